// device_utils.h -- device-wide primitives shared by sort.cu and join.cu.
#ifndef SSB_CSRC_DEVICE_UTILS_H_
#define SSB_CSRC_DEVICE_UTILS_H_
#include <algorithm>

#include "common.h"

namespace ssb {
// In-place exclusive prefix sum of n u64 values; *d_total (device, optional) receives the sum.
// Synchronises the context stream.
int exclusive_scan_u64(ssb_ctx* ctx, unsigned long long* d_data, unsigned long long n, unsigned long long* d_total);
// Stable LSD radix sort of (key, value) pairs on key bits [begin_bit, end_bit).
int radix_sort_pairs(ssb_ctx* ctx, unsigned long long** keys, long long** vals, unsigned long long** keys_tmp,
                     long long** vals_tmp, unsigned long long n, int begin_bit, int end_bit);
unsigned grid_1d(ssb_ctx* ctx, long long n, int block);

#ifdef __CUDACC__
// ---- single-pass prefix over tiles (decoupled look-back) --------------------------------------
// One 64-bit status word per tile: flag in the top two bits, value below, so flag and value
// travel in one relaxed store / load and no fence is needed.
static constexpr unsigned long long kFlagAgg = 1ull << 62, kFlagPrefix = 2ull << 62, kFlagMask = 3ull << 62;

__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Called by one full warp. Publishes `total` (this tile's count) and returns the sum of the
// counts of all earlier tiles: the warp inspects 32 predecessors per step (one status word per
// lane), waits until every word up to the nearest published prefix is there, and adds them up.
// Tiles must be handed out in launch order (atomic ticket), so a predecessor is always running.
__device__ __forceinline__ unsigned long long tile_prefix_warp(unsigned long long* status, unsigned long long tile,
                                                               unsigned long long total, int lane) {
  if (lane == 0) st_relaxed_u64(&status[tile], (tile == 0 ? kFlagPrefix : kFlagAgg) | total);
  if (tile == 0) return 0;
  unsigned long long excl = 0;
  long long t = static_cast<long long>(tile) - 1;   // this step looks at tiles t, t-1, ..., t-31
  for (;;) {
    const long long idx = t - lane;
    const unsigned long long w = idx >= 0 ? ld_relaxed_u64(&status[idx]) : kFlagPrefix;   // before tile 0: empty prefix
    const unsigned ready = __ballot_sync(0xffffffffu, (w & kFlagMask) != 0);
    const unsigned pref = __ballot_sync(0xffffffffu, (w & kFlagMask) == kFlagPrefix);
    const int fp = pref ? __ffs(pref) - 1 : 32;                       // nearest published prefix
    const unsigned need = fp >= 31 ? 0xffffffffu : ((2u << fp) - 1u);   // lanes 0..fp must be published
    if ((ready & need) != need) continue;                             // poll the same window again
    unsigned long long val = (lane <= fp) ? (w & ~kFlagMask) : 0ull;
    for (int d = 16; d > 0; d >>= 1) val += __shfl_xor_sync(0xffffffffu, val, d);
    excl += val;
    if (fp < 32) break;
    t -= 32;
  }
  if (lane == 0) st_relaxed_u64(&status[tile], kFlagPrefix | (excl + total));
  return excl;
}
#endif
}  // namespace ssb
#endif
