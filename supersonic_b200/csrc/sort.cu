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
  SSB_CUDA(ctx, tmp_malloc(ctx, &sums, (nb + 1) * 8));
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
  tmp_free(ctx, sums);
  return rc;
}

// ------------------------------------------------------------------ radix sort of pairs
// One-sweep LSD radix sort (8 bits per pass) of (u64 key, i64 value) pairs.
//   1. radix_diff_kernel     OR of key ^ key[0]: digits whose bits never vary are skipped
//   2. radix_hist_kernel     one read of the keys gives the digit histograms of ALL passes
//   3. radix_base_kernel     exclusive scan of each 256-bin histogram = global digit bases
//   4. radix_onesweep_kernel one launch per pass: a CTA takes the next 4096-pair tile (atomic
//      ticket, so a tile never waits for one that has not started), ranks its keys stably with
//      warp match + per-warp digit counters, publishes the tile's digit counts and resolves the
//      counts of all earlier tiles by decoupled look-back (one thread per digit; flag and value
//      share one 64-bit word, so no fence is needed), stages the tile sorted by digit in shared
//      memory and writes it out in coalesced runs. Keys and values are read once and written
//      once per pass (32 B per pair).
enum { kSortThreads = 256, kSortItemsPerThread = 16, kSortTile = kSortThreads * kSortItemsPerThread,
       kRadixBits = 8, kRadix = 1 << kRadixBits, kMaxPasses = 8 };

struct PassList {
  int n;
  int shift[kMaxPasses];
};

__global__ void __launch_bounds__(256) radix_diff_kernel(const unsigned long long* __restrict__ keys, unsigned long long n,
                                                          unsigned long long* __restrict__ out_mask) {
  const unsigned long long k0 = keys[0];
  unsigned long long m = 0;
  const unsigned long long stride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
  unsigned long long i = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n; i += 4 * stride) {
    const unsigned long long a = keys[i], b = keys[i + stride], c = keys[i + 2 * stride], d = keys[i + 3 * stride];
    m |= (a ^ k0) | (b ^ k0) | (c ^ k0) | (d ^ k0);
  }
  for (; i < n; i += stride) m |= keys[i] ^ k0;
  for (int d = 16; d > 0; d >>= 1) m |= __shfl_xor_sync(0xffffffffu, m, d);
  if ((threadIdx.x & 31) == 0 && m) atomicOr(out_mask, m);
}

// ghist[p * 256 + digit] += number of keys with that digit in pass p.
__global__ void __launch_bounds__(256) radix_hist_kernel(const unsigned long long* __restrict__ keys, unsigned long long n,
                                                          const PassList pl, unsigned long long* __restrict__ ghist) {
  __shared__ unsigned int h[kMaxPasses * kRadix];
  for (int j = threadIdx.x; j < pl.n * kRadix; j += blockDim.x) h[j] = 0;
  __syncthreads();
  const unsigned long long stride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
  unsigned long long i = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n; i += 4 * stride) {
    unsigned long long k[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) k[q] = keys[i + q * stride];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      for (int p = 0; p < pl.n; ++p) atomicAdd(&h[p * kRadix + ((k[q] >> pl.shift[p]) & (kRadix - 1))], 1u);
    }
  }
  for (; i < n; i += stride) {
    const unsigned long long k = keys[i];
    for (int p = 0; p < pl.n; ++p) atomicAdd(&h[p * kRadix + ((k >> pl.shift[p]) & (kRadix - 1))], 1u);
  }
  __syncthreads();
  for (int j = threadIdx.x; j < pl.n * kRadix; j += blockDim.x) {
    if (h[j]) atomicAdd(&ghist[j], static_cast<unsigned long long>(h[j]));
  }
}

// One CTA of 256 threads per pass: in-place exclusive scan of the pass's 256 bins.
__global__ void __launch_bounds__(kRadix) radix_base_kernel(unsigned long long* __restrict__ ghist) {
  __shared__ unsigned long long wsum[kRadix / 32];
  unsigned long long* h = ghist + static_cast<size_t>(blockIdx.x) * kRadix;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned long long v = h[threadIdx.x];
  unsigned long long incl = v;
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned long long y = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += y;
  }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  unsigned long long before = 0;
  for (int w = 0; w < warp; ++w) before += wsum[w];
  h[threadIdx.x] = before + incl - v;
}

// aux layout: [0] ticket (u32 in a u64 word), [1 .. 1 + tiles*256) status words of the pass.
// FULL: every tile of the launch is complete (n is a multiple of the tile), no row guards.
struct SweepSmem {
  unsigned long long stage[kSortTile];        // 32 KB: the tile ordered by digit (keys, then values)
  unsigned long long gadj[kRadix];            // global position of local position 0 of a digit run
  unsigned int lstart[kRadix];                // first local position of a digit
  unsigned short wcnt[kSortThreads / 32][kRadix];   // per-warp digit counts, then exclusive bases
  unsigned char dig[kSortTile];               // digit of the staged element
  unsigned int wsum[kSortThreads / 32];
  unsigned int s_tile;
};

template <bool FULL>
__device__ __forceinline__ void onesweep_tile(
    SweepSmem& sm, const unsigned long long tile, const unsigned long long* __restrict__ keys_in,
    const long long* __restrict__ vals_in, unsigned long long* __restrict__ keys_out, long long* __restrict__ vals_out,
    unsigned long long n, int shift, const unsigned long long* __restrict__ gbase, unsigned long long* __restrict__ aux) {
  constexpr int NW = kSortThreads / 32;
  constexpr int PER_WARP = kSortTile / NW;
  constexpr int ROUNDS = PER_WARP / 32;
  constexpr int HALF = ROUNDS / 2;
  constexpr int LOOK = 4;                                // predecessors inspected per look-back step
  unsigned long long (&stage)[kSortTile] = sm.stage;
  unsigned long long (&gadj)[kRadix] = sm.gadj;
  unsigned int (&lstart)[kRadix] = sm.lstart;
  unsigned short (&wcnt)[NW][kRadix] = sm.wcnt;
  unsigned char (&dig)[kSortTile] = sm.dig;
  unsigned int (&wsum)[NW] = sm.wsum;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned long long* status = aux + 1;
  const unsigned long long tile0 = tile * kSortTile;
  const unsigned long long warp0 = tile0 + static_cast<unsigned long long>(warp) * PER_WARP + lane;
  const unsigned long long left = n - tile0;
  const int cnt_tile = (FULL || left >= static_cast<unsigned long long>(kSortTile)) ? kSortTile : static_cast<int>(left);

  unsigned long long k[ROUNDS];
  unsigned short pos[ROUNDS];
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    const unsigned long long i = warp0 + r * 32;
    k[r] = (FULL || i < n) ? keys_in[i] : ~0ull;
  }
  const unsigned lt = (1u << lane) - 1u;
  // ranks: the warp matches run back to back (their latency overlaps), then the per-warp digit
  // counters are walked in row order
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    unsigned peers[HALF];
#pragma unroll
    for (int q = 0; q < HALF; ++q) {
      const int r = h * HALF + q;
      const unsigned d = static_cast<unsigned>((k[r] >> shift) & (kRadix - 1));
      if (FULL) {
        peers[q] = __match_any_sync(0xffffffffu, d);
      } else {
        const bool live = warp0 + r * 32 < n;
        const unsigned active = __ballot_sync(0xffffffffu, live);
        peers[q] = 0;
        if (live) peers[q] = __match_any_sync(active, d);
      }
    }
#pragma unroll
    for (int q = 0; q < HALF; ++q) {
      const int r = h * HALF + q;
      const unsigned d = static_cast<unsigned>((k[r] >> shift) & (kRadix - 1));
      const bool live = FULL || peers[q] != 0u;
      unsigned before = 0;
      if (live) before = wcnt[warp][d];
      __syncwarp();
      if (live) {
        pos[r] = static_cast<unsigned short>(before + __popc(peers[q] & lt));
        if ((__ffs(peers[q]) - 1) == lane) wcnt[warp][d] = static_cast<unsigned short>(before + __popc(peers[q]));
      }
      __syncwarp();
    }
  }
  __syncthreads();
  // thread d owns digit d: scan over the warps, publish the tile's count, scan over the digits
  unsigned int total = 0;
  {
    const int d = tid;
#pragma unroll
    for (int w = 0; w < NW; ++w) { const unsigned int c = wcnt[w][d]; wcnt[w][d] = static_cast<unsigned short>(total); total += c; }
    st_relaxed_u64(&status[tile * kRadix + d], (tile == 0 ? kFlagPrefix : kFlagAgg) | total);
    unsigned int incl = total;
    for (int s = 1; s < 32; s <<= 1) {
      const unsigned int y = __shfl_up_sync(0xffffffffu, incl, s);
      if (lane >= s) incl += y;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    unsigned int before = 0;
    for (int w = 0; w < warp; ++w) before += wsum[w];
    lstart[d] = before + incl - total;
  }
  __syncthreads();
  // stage the keys ordered by digit
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    if (FULL || warp0 + r * 32 < n) {
      const unsigned d = static_cast<unsigned>((k[r] >> shift) & (kRadix - 1));
      const unsigned int q = lstart[d] + wcnt[warp][d] + pos[r];
      pos[r] = static_cast<unsigned short>(q);
      stage[q] = k[r];
      dig[q] = static_cast<unsigned char>(d);
    }
  }
  // the values travel in a second round through the same staging buffer; fetch them now so
  // that their latency overlaps the look-back
  long long v[ROUNDS];
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    const unsigned long long i = warp0 + r * 32;
    v[r] = (FULL || i < n) ? vals_in[i] : 0;
  }
  {
    // decoupled look-back, LOOK predecessors per step: their status words are fetched together
    // and consumed in order up to the first one that is not published yet
    const int d = tid;
    unsigned long long excl = 0;
    if (tile > 0) {
      unsigned long long t = tile;   // predecessors t-1, t-2, ... are still to be accounted for
      for (;;) {
        unsigned long long w[LOOK];
#pragma unroll
        for (int q = 0; q < LOOK; ++q) {
          w[q] = kFlagPrefix;        // before tile 0: an empty prefix
          if (static_cast<unsigned long long>(q) < t) w[q] = ld_relaxed_u64(&status[(t - 1 - q) * kRadix + d]);
        }
        bool done = false;
        int used = 0;
#pragma unroll
        for (int q = 0; q < LOOK; ++q) {
          if (done || used != q) continue;
          const unsigned long long f = w[q] & kFlagMask;
          if (f == 0) continue;      // not published yet: poll again from here
          excl += w[q] & ~kFlagMask;
          ++used;
          if (f == kFlagPrefix) done = true;
        }
        if (done) break;
        t -= static_cast<unsigned long long>(used);
      }
      st_relaxed_u64(&status[tile * kRadix + d], kFlagPrefix | (excl + total));
    }
    gadj[d] = gbase[d] + excl - lstart[d];
  }
  __syncthreads();
  for (int i = tid; i < cnt_tile; i += kSortThreads) keys_out[gadj[dig[i]] + i] = stage[i];
  __syncthreads();
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    if (FULL || warp0 + r * 32 < n) stage[pos[r]] = static_cast<unsigned long long>(v[r]);
  }
  __syncthreads();
  for (int i = tid; i < cnt_tile; i += kSortThreads) vals_out[gadj[dig[i]] + i] = static_cast<long long>(stage[i]);
}

__global__ void __launch_bounds__(kSortThreads, 3) radix_onesweep_kernel(
    const unsigned long long* __restrict__ keys_in, const long long* __restrict__ vals_in,
    unsigned long long* __restrict__ keys_out, long long* __restrict__ vals_out, unsigned long long n, int shift,
    const unsigned long long* __restrict__ gbase, unsigned long long* __restrict__ aux) {
  __shared__ SweepSmem sm;
  const int tid = threadIdx.x;
  if (tid == 0) sm.s_tile = atomicAdd(reinterpret_cast<unsigned int*>(aux), 1u);
  for (int d = tid; d < kRadix * (kSortThreads / 32); d += kSortThreads) (&sm.wcnt[0][0])[d] = 0;
  __syncthreads();
  const unsigned long long tile = sm.s_tile;
  // complete tiles run without row guards (CTA-uniform choice)
  if (n - tile * kSortTile >= static_cast<unsigned long long>(kSortTile)) {
    onesweep_tile<true>(sm, tile, keys_in, vals_in, keys_out, vals_out, n, shift, gbase, aux);
  } else {
    onesweep_tile<false>(sm, tile, keys_in, vals_in, keys_out, vals_out, n, shift, gbase, aux);
  }
}

// Sorts (keys, vals) by bits [begin_bit, end_bit) of keys; stable. The result ends in
// (*keys, *vals); tmp buffers are swapped in and out as needed. One host synchronisation
// (the varying-bits mask decides which passes run at all).
int radix_sort_pairs(ssb_ctx* ctx, unsigned long long** keys, long long** vals, unsigned long long** keys_tmp,
                     long long** vals_tmp, unsigned long long n, int begin_bit, int end_bit) {
  if (n <= 1 || end_bit <= begin_bit) return 0;
  const unsigned long long tiles = (n + kSortTile - 1) / kSortTile;
  // aux: [mask][ghist 8 x 256][ticket][status tiles x 256]
  const size_t hist_words = static_cast<size_t>(kMaxPasses) * kRadix;
  const size_t aux_words = 1 + hist_words + 1 + static_cast<size_t>(tiles) * kRadix;
  unsigned long long* aux = nullptr;
  SSB_CUDA(ctx, tmp_malloc(ctx, &aux, aux_words * 8));
  unsigned long long* d_mask = aux;
  unsigned long long* ghist = aux + 1;
  unsigned long long* pass_aux = aux + 1 + hist_words;
  int rc = 0;
  cudaError_t e = cudaMemsetAsync(aux, 0, (1 + hist_words) * 8, ctx->stream);
  const unsigned grid = grid_1d(ctx, static_cast<long long>((n + 3) / 4), 256);
  if (e == cudaSuccess) {
    radix_diff_kernel<<<grid, 256, 0, ctx->stream>>>(*keys, n, d_mask);
    ++ctx->launches;
    e = cudaMemcpyAsync(ctx->h_count, d_mask, 8, cudaMemcpyDeviceToHost, ctx->stream);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) { tmp_free(ctx, aux); return cuda_fail(ctx, e, "radix sort setup"); }
  const unsigned long long varying = static_cast<unsigned long long>(*ctx->h_count);
  PassList pl;
  pl.n = 0;
  for (int shift = begin_bit; shift < end_bit && pl.n < kMaxPasses; shift += kRadixBits) {
    if ((varying >> shift) & (kRadix - 1)) pl.shift[pl.n++] = shift;
  }
  if (pl.n > 0) {
    radix_hist_kernel<<<grid, 256, 0, ctx->stream>>>(*keys, n, pl, ghist);
    ++ctx->launches;
    radix_base_kernel<<<pl.n, kRadix, 0, ctx->stream>>>(ghist);
    ++ctx->launches;
    for (int p = 0; p < pl.n; ++p) {
      e = cudaMemsetAsync(pass_aux, 0, (1 + static_cast<size_t>(tiles) * kRadix) * 8, ctx->stream);
      if (e != cudaSuccess) break;
      radix_onesweep_kernel<<<static_cast<unsigned>(tiles), kSortThreads, 0, ctx->stream>>>(
          *keys, *vals, *keys_tmp, *vals_tmp, n, pl.shift[p], ghist + static_cast<size_t>(p) * kRadix, pass_aux);
      ++ctx->launches;
      std::swap(*keys, *keys_tmp);
      std::swap(*vals, *vals_tmp);
    }
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) rc = cuda_fail(ctx, e, "radix sort");
  }
  cudaStreamSynchronize(ctx->stream);
  tmp_free(ctx, aux);
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
      // -0.0 and +0.0 compare equal in the reference (sort.cc:151 compares with operator<): one image for both
      case T_F32: { uint32_t b = static_cast<const uint32_t*>(col)[r]; if ((b << 1) == 0u) b = 0u; b = (b & 0x80000000u) ? ~b : (b | 0x80000000u); u = b; } break;
      case T_F64: { unsigned long long b = static_cast<const unsigned long long*>(col)[r]; if ((b << 1) == 0ull) b = 0ull; b = (b >> 63) ? ~b : (b | 0x8000000000000000ull); u = b; } break;
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

// dst[idx[i]] = src[i]: the inverse of gather for index lists without duplicates.
__global__ void scatter_kernel(const void* __restrict__ src, int width, const long long* __restrict__ idx, long long n,
                               void* __restrict__ dst) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const long long r = idx[i];
    if (r < 0) continue;
    if (width == 8) static_cast<unsigned long long*>(dst)[r] = static_cast<const unsigned long long*>(src)[i];
    else if (width == 4) static_cast<uint32_t*>(dst)[r] = static_cast<const uint32_t*>(src)[i];
    else static_cast<uint8_t*>(dst)[r] = static_cast<const uint8_t*>(src)[i];
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

int ssb_scatter(ssb_ctx* ctx, const ssb_column* src, const int64_t* d_idx, int64_t n, const ssb_column* dst) {
  if (n <= 0) return 0;
  const int w = width_of(src->dtype);
  if (w == 0 || width_of(dst->dtype) != w) return fail(ctx, SSB_ERROR_INVALID_ARGUMENT_TYPE, "scatter: column types differ");
  if (src->nulls != nullptr) return fail(ctx, SSB_ERROR_NOT_IMPLEMENTED, "scatter of a column with a null bitmap");
  scatter_kernel<<<grid_1d(ctx, n, 256), 256, 0, ctx->stream>>>(src->data, w, reinterpret_cast<const long long*>(d_idx), n, dst->data);
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
  cudaError_t e = tmp_malloc(ctx, &k0, static_cast<size_t>(rows) * 8);
  if (e == cudaSuccess) e = tmp_malloc(ctx, &k1, static_cast<size_t>(rows) * 8);
  if (e == cudaSuccess) e = tmp_malloc(ctx, &p1, static_cast<size_t>(rows) * 8);
  if (e != cudaSuccess) {   // out of memory: give back what was obtained
    tmp_free(ctx, k0); tmp_free(ctx, k1); tmp_free(ctx, p1);
    return cuda_fail(ctx, e, "sort scratch");
  }
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
  tmp_free(ctx, k0);
  tmp_free(ctx, k1);
  tmp_free(ctx, pa == perm ? pb : pa);
  if (rc == 0) SSB_CUDA(ctx, cudaGetLastError());
  return rc;
}

}  // extern "C"
