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
}  // namespace ssb
#endif
