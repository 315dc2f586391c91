// Force-included before every reference TU (oracle build only).
// libstdc++-13 deletes std::hash<std::_Bit_reference>; the reference's BOOL
// IN-set (comparison_bound_expressions.cc:173) needs it. Values follow
// operators::Hash(bool) (base/infrastructure/operators.h:357-360).
#ifndef ORACLE_SHIM_FIXUPS_H_
#define ORACLE_SHIM_FIXUPS_H_
#include <time.h>
#include <unistd.h>
#ifdef __cplusplus
#include <vector>
#include <functional>
namespace std {
template <> struct hash<_Bit_reference> {
  size_t operator()(_Bit_reference b) const { return bool(b) ? 23 : 34; }
};
}
#endif
#endif
