// sort.cu -- device exclusive scan, LSD radix sort of (key, row id) pairs, gather.
//
// Replaces SortPermutation / SortTypedColumn (cursor/core/sort.cc:150-322,781-805), which
// std::sort an int64 permutation column by column, by a stable least-significant-digit radix
// sort: key columns are visited from the least to the most significant; each column is
// turned into order-preserving unsigned keys (sign flip, IEEE flip, bitwise NOT for
// DESCENDING) gathered through the current permutation, and sorted 8 bits at a time
// (histogram -> scan -> stable scatter). NULLs sort first for ASCENDING and last for
// DESCENDING (sort.cc:174-238) through one extra 1-bit pass per nullable column.
// Gather replaces ViewCursorWithSelectionVector / ColumnCopierFn
// (cursor/infrastructure/view_cursor.cc:94-115, base/infrastructure/copy_column.cc:200-286).
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.h"
#include "device_utils.h"

namespace ssb {

// ------------------------------------------------------------------ exclusive scan (u64)
enum { kScanThreads = 256, kScanItems = 8, kScanTile = kScanThreads * kScanItems };

__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(const unsigned long long* __restrict__ in,
                                                                    unsigned long long n,
                                                                    unsigned long long* __restrict__ sums) {
  __shared__ unsigned long long warp_sum[kScanThreads / 32];
  const unsigned long long base = static_cast<unsigned long long>(blockIdx.x) * kScanTile;
  unsigned long long s = 0;
  for (int k = 0; k < kScanItems; ++k) {
    const unsigned long long i = base + k * kScanThreads + threadIdx.x;
    if (i < n) s += in[i];
  }
  for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
  if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int w = 0; w < kScanThreads / 32; ++w) t += warp_sum[w];
    sums[blockIdx.x] = t;
  }
}

// Each thread owns kScanItems consecutive elements of the tile.
__global__ void __launch_bounds__(kScanThreads) scan_down_kernel(unsigned long long* __restrict__ data,
                                                                  unsigned long long n,
                                                                  const unsigned long long* __restrict__ block_base) {
  __shared__ unsigned long long warp_sum[kScanThreads / 32];
  const unsigned long long base = static_cast<unsigned long long>(blockIdx.x) * kScanTile +
                                  static_cast<unsigned long long>(threadIdx.x) * kScanItems;
  unsigned long long v[kScanItems];
  unsigned long long s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    v[k] = base + k < n ? data[base + k] : 0ull;
    s += v[k];
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long incl = s;
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned long long y = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += y;
  }
  if (lane == 31) warp_sum[warp] = incl;
  __syncthreads();
  unsigned long long before = block_base ? block_base[blockIdx.x] : 0ull;
  for (int w = 0; w < warp; ++w) before += warp_sum[w];
  unsigned long long run = before + incl - s;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    if (base + k < n) data[base + k] = run;
    run += v[k];
  }
}

int exclusive_scan_u64(ssb_ctx* ctx, unsigned long long* d_data, unsigned long long n, unsigned long long* d_total) {
  if (n == 0) {
    if (d_total) SSB_CUDA(ctx, cudaMemsetAsync(d_total, 0, 8, ctx->stream));
    return 0;
  }
  const unsigned long long nb = (n + kScanTile - 1) / kScanTile;
  unsigned long long* sums = nullptr;
  SSB_CUDA(ctx, cudaMalloc(&sums, (nb + 1) * 8));
  scan_reduce_kernel<<<static_cast<unsigned>(nb), kScanThreads, 0, ctx->stream>>>(d_data, n, sums);
  ++ctx->launches;
  int rc = 0;
  if (nb > 1) {
    rc = exclusive_scan_u64(ctx, sums, nb, d_total);
  } else if (d_total) {
    cudaMemcpyAsync(d_total, sums, 8, cudaMemcpyDeviceToDevice, ctx->stream);
  }
  if (rc == 0) {
    if (nb == 1) cudaMemsetAsync(sums, 0, 8, ctx->stream);
    scan_down_kernel<<<static_cast<unsigned>(nb), kScanThreads, 0, ctx->stream>>>(d_data, n, sums);
    ++ctx->launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) rc = cuda_fail(ctx, e, "scan");
  }
  cudaStreamSynchronize(ctx->stream);
  cudaFree(sums);
  return rc;
}

// ------------------------------------------------------------------ radix sort of pairs
enum { kSortThreads = 256, kSortItemsPerThread = 16, kSortTile = kSortThreads * kSortItemsPerThread,
       kRadixBits = 8, kRadix = 1 << kRadixBits };

// Block digit histogram; hist is digit-major: hist[d * nblocks + block].
__global__ void __launch_bounds__(kSortThreads) radix_hist_kernel(const unsigned long long* __restrict__ keys,
                                                                   unsigned long long n, int shift,
                                                                   unsigned long long* __restrict__ hist,
                                                                   unsigned int nblocks) {
  __shared__ unsigned int cnt[kRadix];
  for (int d = threadIdx.x; d < kRadix; d += kSortThreads) cnt[d] = 0;
  __syncthreads();
  const unsigned long long base = static_cast<unsigned long long>(blockIdx.x) * kSortTile;
  for (int k = 0; k < kSortItemsPerThread; ++k) {
    const unsigned long long i = base + k * kSortThreads + threadIdx.x;
    if (i < n) atomicAdd(&cnt[(keys[i] >> shift) & (kRadix - 1)], 1u);
  }
  __syncthreads();
  for (int d = threadIdx.x; d < kRadix; d += kSortThreads) hist[static_cast<unsigned long long>(d) * nblocks + blockIdx.x] = cnt[d];
}

// Stable scatter. Warp w of the block owns the items [w*512, (w+1)*512) of the tile and walks
// them 32 at a time, so ranks follow the input order.
__global__ void __launch_bounds__(kSortThreads) radix_scatter_kernel(
    const unsigned long long* __restrict__ keys_in, const long long* __restrict__ vals_in,
    unsigned long long* __restrict__ keys_out, long long* __restrict__ vals_out, unsigned long long n, int shift,
    const unsigned long long* __restrict__ hist_scanned, unsigned int nblocks) {
  constexpr int NW = kSortThreads / 32;
  constexpr int PER_WARP = kSortTile / NW;
  __shared__ unsigned int wcnt[NW][kRadix];     // per-warp digit counts, then exclusive bases
  __shared__ unsigned long long gbase[kRadix];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int d = threadIdx.x; d < kRadix * NW; d += kSortThreads) (&wcnt[0][0])[d] = 0;
  __syncthreads();
  const unsigned long long tile0 = static_cast<unsigned long long>(blockIdx.x) * kSortTile + static_cast<unsigned long long>(warp) * PER_WARP;
  unsigned long long k[PER_WARP / 32];
  unsigned short rank[PER_WARP / 32];
  const unsigned lt = (1u << lane) - 1u;
#pragma unroll
  for (int r = 0; r < PER_WARP / 32; ++r) {
    const unsigned long long i = tile0 + r * 32 + lane;
    const bool live = i < n;
    k[r] = live ? keys_in[i] : ~0ull;
    const unsigned d = live ? static_cast<unsigned>((k[r] >> shift) & (kRadix - 1)) : kRadix;   // dead lanes: no digit
    const unsigned active = __ballot_sync(0xffffffffu, live);
    unsigned peers = 0;
    if (live) peers = __match_any_sync(active, d);
    unsigned before = 0;
    if (live) before = wcnt[warp][d];
    __syncwarp();
    if (live) {
      rank[r] = static_cast<unsigned short>(before + __popc(peers & lt));
      if ((__ffs(peers) - 1) == lane) wcnt[warp][d] = before + __popc(peers);
    }
    __syncwarp();
  }
  __syncthreads();
  // exclusive scan over the warps for every digit + global base of (digit, block)
  for (int d = threadIdx.x; d < kRadix; d += kSortThreads) {
    unsigned int run = 0;
    for (int w = 0; w < NW; ++w) { const unsigned int c = wcnt[w][d]; wcnt[w][d] = run; run += c; }
    gbase[d] = hist_scanned[static_cast<unsigned long long>(d) * nblocks + blockIdx.x];
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < PER_WARP / 32; ++r) {
    const unsigned long long i = tile0 + r * 32 + lane;
    if (i < n) {
      const unsigned d = static_cast<unsigned>((k[r] >> shift) & (kRadix - 1));
      const unsigned long long pos = gbase[d] + wcnt[warp][d] + rank[r];
      keys_out[pos] = k[r];
      vals_out[pos] = vals_in[i];
    }
  }
}

// Sorts (keys, vals) by bits [begin_bit, end_bit) of keys; stable. The result ends in
// (*keys, *vals); tmp buffers are swapped in and out as needed.
int radix_sort_pairs(ssb_ctx* ctx, unsigned long long** keys, long long** vals, unsigned long long** keys_tmp,
                     long long** vals_tmp, unsigned long long n, int begin_bit, int end_bit) {
  if (n <= 1) return 0;
  const unsigned int nblocks = static_cast<unsigned int>((n + kSortTile - 1) / kSortTile);
  unsigned long long* hist = nullptr;
  SSB_CUDA(ctx, cudaMalloc(&hist, static_cast<size_t>(nblocks) * kRadix * 8));
  int rc = 0;
  for (int shift = begin_bit; shift < end_bit && rc == 0; shift += kRadixBits) {
    radix_hist_kernel<<<nblocks, kSortThreads, 0, ctx->stream>>>(*keys, n, shift, hist, nblocks);
    ++ctx->launches;
    rc = exclusive_scan_u64(ctx, hist, static_cast<unsigned long long>(nblocks) * kRadix, nullptr);
    if (rc) break;
    radix_scatter_kernel<<<nblocks, kSortThreads, 0, ctx->stream>>>(*keys, *vals, *keys_tmp, *vals_tmp, n, shift, hist, nblocks);
    ++ctx->launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { rc = cuda_fail(ctx, e, "radix sort"); break; }
    std::swap(*keys, *keys_tmp);
    std::swap(*vals, *vals_tmp);
  }
  cudaStreamSynchronize(ctx->stream);
  cudaFree(hist);
  return rc;
}

// ------------------------------------------------------------------ key transforms
__global__ void iota_kernel(long long* p, long long n) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = i;
}

// keys[i] = order-preserving unsigned image of col[perm[i]] (or, with null_pass, 0/1 flags
// that put NULL rows first for ASCENDING and last for DESCENDING).
__global__ void sort_key_kernel(const void* __restrict__ col, const uint32_t* __restrict__ nulls, int phys,
                                int descending, int null_pass, const long long* __restrict__ perm, long long n,
                                unsigned long long* __restrict__ keys) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const long long r = perm[i];
    const bool isn = nulls != nullptr && ((nulls[r >> 5] >> (r & 31)) & 1u);
    if (null_pass) {
      keys[i] = descending ? (isn ? 1ull : 0ull) : (isn ? 0ull : 1ull);
      continue;
    }
    unsigned long long u;
    switch (phys) {
      case T_I32: u = static_cast<uint32_t>(static_cast<const int32_t*>(col)[r]) ^ 0x80000000u; break;
      case T_U32: u = static_cast<const uint32_t*>(col)[r]; break;
      case T_I64: u = static_cast<unsigned long long>(static_cast<const long long*>(col)[r]) ^ 0x8000000000000000ull; break;
      case T_U64: u = static_cast<const unsigned long long*>(col)[r]; break;
      case T_F32: { uint32_t b = static_cast<const uint32_t*>(col)[r]; b = (b & 0x80000000u) ? ~b : (b | 0x80000000u); u = b; } break;
      case T_F64: { unsigned long long b = static_cast<const unsigned long long*>(col)[r]; b = (b >> 63) ? ~b : (b | 0x8000000000000000ull); u = b; } break;
      default: u = static_cast<const uint8_t*>(col)[r] != 0 ? 1ull : 0ull; break;
    }
    if (isn) u = 0;   // value under NULL is garbage: make equal so that later keys decide
    if (descending) {
      const int bits = phys_width(phys) * 8;
      u = (~u) & (bits == 64 ? ~0ull : ((1ull << bits) - 1ull));
    }
    keys[i] = u;
  }
}

// ------------------------------------------------------------------ gather
__global__ void gather_kernel(const void* __restrict__ src, const uint32_t* __restrict__ src_nulls, int width,
                              const long long* __restrict__ idx, long long n, void* __restrict__ dst,
                              uint32_t* __restrict__ dst_nulls) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long n_up = (n + 31) & ~31LL;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_up; i += stride) {
    bool isn = false;
    if (i < n) {
      const long long r = idx[i];
      if (r < 0) {
        isn = true;   // copy_column.cc:112-127: a negative index selects NULL
        if (width == 8) static_cast<unsigned long long*>(dst)[i] = 0;
        else if (width == 4) static_cast<uint32_t*>(dst)[i] = 0;
        else static_cast<uint8_t*>(dst)[i] = 0;
      } else {
        if (width == 8) static_cast<unsigned long long*>(dst)[i] = static_cast<const unsigned long long*>(src)[r];
        else if (width == 4) static_cast<uint32_t*>(dst)[i] = static_cast<const uint32_t*>(src)[r];
        else static_cast<uint8_t*>(dst)[i] = static_cast<const uint8_t*>(src)[r];
        isn = src_nulls != nullptr && ((src_nulls[r >> 5] >> (r & 31)) & 1u);
      }
    }
    if (dst_nulls != nullptr) {
      const uint32_t w = __ballot_sync(0xffffffffu, isn);
      if ((threadIdx.x & 31) == 0) dst_nulls[i >> 5] = w;
    }
  }
}

unsigned grid_1d(ssb_ctx* ctx, long long n, int block) {
  long long g = div_up(n, block);
  const long long cap = static_cast<long long>(ctx->num_sms) * 8;
  return static_cast<unsigned>(g > cap ? cap : (g < 1 ? 1 : g));
}

}  // namespace ssb

using namespace ssb;

extern "C" {

int ssb_gather(ssb_ctx* ctx, const ssb_column* src, const int64_t* d_idx, int64_t n, const ssb_column* dst) {
  if (n <= 0) return 0;
  const int w = width_of(src->dtype);
  if (w == 0 || width_of(dst->dtype) != w) return fail(ctx, SSB_ERROR_INVALID_ARGUMENT_TYPE, "gather: column types differ");
  gather_kernel<<<grid_1d(ctx, n, 256), 256, 0, ctx->stream>>>(src->data, src->nulls, w,
                                                                reinterpret_cast<const long long*>(d_idx), n,
                                                                dst->data, dst->nulls);
  ++ctx->launches;
  SSB_CUDA(ctx, cudaGetLastError());
  return 0;
}

int ssb_sort_permutation(ssb_ctx* ctx, int32_t n_keys, const ssb_column* keys, const int32_t* descending,
                         int64_t rows, int64_t* d_perm) {
  if (rows < 0) return fail(ctx, SSB_ERROR_INVALID_ARGUMENT_VALUE, "negative row count");
  if (rows == 0) return 0;
  TimedRegion timed(ctx);
  long long* perm = reinterpret_cast<long long*>(d_perm);
  iota_kernel<<<grid_1d(ctx, rows, 256), 256, 0, ctx->stream>>>(perm, rows);
  ++ctx->launches;
  if (n_keys == 0 || rows == 1) { SSB_CUDA(ctx, cudaGetLastError()); return 0; }
  unsigned long long *k0 = nullptr, *k1 = nullptr;
  long long* p1 = nullptr;
  SSB_CUDA(ctx, cudaMalloc(&k0, static_cast<size_t>(rows) * 8));
  SSB_CUDA(ctx, cudaMalloc(&k1, static_cast<size_t>(rows) * 8));
  SSB_CUDA(ctx, cudaMalloc(&p1, static_cast<size_t>(rows) * 8));
  long long* pa = perm;
  long long* pb = p1;
  int rc = 0;
  for (int c = n_keys - 1; c >= 0 && rc == 0; --c) {   // least significant key first
    const int phys = phys_of(keys[c].dtype);
    if (phys < 0) { rc = fail(ctx, SSB_ERROR_INVALID_ARGUMENT_TYPE, "unsupported sort key type"); break; }
    const int desc = descending[c] ? 1 : 0;
    sort_key_kernel<<<grid_1d(ctx, rows, 256), 256, 0, ctx->stream>>>(keys[c].data, keys[c].nulls, phys, desc, 0, pa, rows, k0);
    ++ctx->launches;
    rc = radix_sort_pairs(ctx, &k0, &pa, &k1, &pb, static_cast<unsigned long long>(rows), 0, phys_width(phys) * 8);
    if (rc == 0 && keys[c].nulls != nullptr) {
      sort_key_kernel<<<grid_1d(ctx, rows, 256), 256, 0, ctx->stream>>>(keys[c].data, keys[c].nulls, phys, desc, 1, pa, rows, k0);
      ++ctx->launches;
      rc = radix_sort_pairs(ctx, &k0, &pa, &k1, &pb, static_cast<unsigned long long>(rows), 0, 8);
    }
  }
  if (rc == 0 && pa != perm) {
    cudaMemcpyAsync(perm, pa, static_cast<size_t>(rows) * 8, cudaMemcpyDeviceToDevice, ctx->stream);
  }
  cudaStreamSynchronize(ctx->stream);
  cudaFree(k0);
  cudaFree(k1);
  cudaFree(pa == perm ? pb : pa);
  if (rc == 0) SSB_CUDA(ctx, cudaGetLastError());
  return rc;
}

}  // extern "C"
