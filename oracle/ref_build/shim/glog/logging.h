// Minimal stand-in for <glog/logging.h>, used ONLY to compile the unmodified
// reference sources from /root/reference into the test oracle (oracle/_ref).
// Control/metadata only: cannot change any computed result.
#ifndef ORACLE_SHIM_GLOG_LOGGING_H_
#define ORACLE_SHIM_GLOG_LOGGING_H_
#include <cstdlib>
#include <time.h>
#include <unistd.h>
#include <vector>
#include <iostream>
#include <sstream>
#include <string>

namespace google {
inline void InitGoogleLogging(const char*) {}
inline void InstallFailureSignalHandler() {}
inline void GetExistingTempDirectories(std::vector<std::string>* d) { d->push_back("/tmp"); }
}  // namespace google

namespace oracle_shim {
class LogMessage {
 public:
  LogMessage(const char* file, int line, int sev) : sev_(sev) {
    s_ << (sev >= 3 ? "F " : sev == 2 ? "E " : sev == 1 ? "W " : "I ")
       << file << ":" << line << "] ";
  }
  ~LogMessage() {
    if (sev_ >= 1) std::cerr << s_.str() << std::endl;
    if (sev_ >= 3) abort();
  }
  std::ostream& stream() { return s_; }
 private:
  int sev_;
  std::ostringstream s_;
};
struct Voidify { void operator&(std::ostream&) {} };
struct NullStream : std::ostream { NullStream() : std::ostream(nullptr) {} };
template <typename T> T CheckNotNull(const char* f, int l, const char* n, T&& t) {
  if (t == nullptr) { LogMessage(f, l, 3).stream() << n; }
  return std::forward<T>(t);
}
}  // namespace oracle_shim

#define ORACLE_SEV_INFO 0
#define ORACLE_SEV_WARNING 1
#define ORACLE_SEV_ERROR 2
#define ORACLE_SEV_FATAL 3
#define ORACLE_SEV_DFATAL 2
#define LOG(sev) ::oracle_shim::LogMessage(__FILE__, __LINE__, ORACLE_SEV_##sev).stream()
#define LOG_IF(sev, c) !(c) ? (void)0 : ::oracle_shim::Voidify() & LOG(sev)
#define VLOG_IS_ON(n) false
#define VLOG(n) true ? (void)0 : ::oracle_shim::Voidify() & LOG(INFO)
#define DVLOG(n) VLOG(n)
#define DLOG(sev) true ? (void)0 : ::oracle_shim::Voidify() & LOG(sev)
#define LOG_STRING(sev, vec) LOG(sev)
#define LOG_ASSERT(c) CHECK(c)
#define LOG_EVERY_N(sev, n) LOG(sev)
#define LOG_FIRST_N(sev, n) LOG(sev)
#define CHECK(c) (c) ? (void)0 : ::oracle_shim::Voidify() & LOG(FATAL) << "Check failed: " #c " "
#define ORACLE_CHECK_OP(a, b, op) CHECK((a) op (b))
#define CHECK_EQ(a, b) ORACLE_CHECK_OP(a, b, ==)
#define CHECK_NE(a, b) ORACLE_CHECK_OP(a, b, !=)
#define CHECK_LE(a, b) ORACLE_CHECK_OP(a, b, <=)
#define CHECK_LT(a, b) ORACLE_CHECK_OP(a, b, <)
#define CHECK_GE(a, b) ORACLE_CHECK_OP(a, b, >=)
#define CHECK_GT(a, b) ORACLE_CHECK_OP(a, b, >)
#define CHECK_NOTNULL(v) ::oracle_shim::CheckNotNull(__FILE__, __LINE__, "'" #v "' Must be non NULL", (v))
#define CHECK_STREQ(a, b) CHECK(std::string(a) == std::string(b))
#define PCHECK(c) CHECK(c)
#define ORACLE_DCHECK_SINK(c) true ? (void)0 : ::oracle_shim::Voidify() & LOG(INFO) << (c)
#define DCHECK(c) ORACLE_DCHECK_SINK(c)
#define DCHECK_EQ(a, b) ORACLE_DCHECK_SINK((a) == (b))
#define DCHECK_NE(a, b) ORACLE_DCHECK_SINK((a) != (b))
#define DCHECK_LE(a, b) ORACLE_DCHECK_SINK((a) <= (b))
#define DCHECK_LT(a, b) ORACLE_DCHECK_SINK((a) < (b))
#define DCHECK_GE(a, b) ORACLE_DCHECK_SINK((a) >= (b))
#define DCHECK_GT(a, b) ORACLE_DCHECK_SINK((a) > (b))
#define DCHECK_NOTNULL(v) (v)
#endif
