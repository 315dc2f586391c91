// Minimal stand-in for <gflags/gflags.h> (oracle build only).
#ifndef ORACLE_SHIM_GFLAGS_H_
#define ORACLE_SHIM_GFLAGS_H_
#include <string>
#include <stdint.h>
#define DEFINE_bool(n, d, h) bool FLAGS_##n = d
#define DEFINE_int32(n, d, h) int32_t FLAGS_##n = d
#define DEFINE_int64(n, d, h) int64_t FLAGS_##n = d
#define DEFINE_uint64(n, d, h) uint64_t FLAGS_##n = d
#define DEFINE_double(n, d, h) double FLAGS_##n = d
#define DEFINE_string(n, d, h) std::string FLAGS_##n = d
#define DECLARE_bool(n) extern bool FLAGS_##n
#define DECLARE_int32(n) extern int32_t FLAGS_##n
#define DECLARE_int64(n) extern int64_t FLAGS_##n
#define DECLARE_uint64(n) extern uint64_t FLAGS_##n
#define DECLARE_double(n) extern double FLAGS_##n
#define DECLARE_string(n) extern std::string FLAGS_##n
namespace google {
inline unsigned ParseCommandLineFlags(int*, char***, bool) { return 0; }
}
#endif
