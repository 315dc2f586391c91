// strings.cu -- STRING / BINARY columns on the device.
//
// The reference keeps a variable-length cell as a StringPiece into an arena
// (base/infrastructure/types.h:53-68, block.h:259-281, base/memory/arena.h:48) and compares cells
// with memcmp over the common prefix, the shorter one first (utils/strings/stringpiece.h:268-283);
// its hash set hashes the bytes and confirms with operator== (row_hash_set.cc:424-498).
//
// Here a variable-length column is (offsets INT64[rows + 1], bytes) in HBM, and the relational
// kernels (group, join, sort, compare) never touch the bytes: ssb_string_rank turns the column into
// dense ORDER-PRESERVING codes -- code(a) < code(b) iff a < b in the reference's order, equal
// codes iff equal bytes -- so GroupAggregate / HashJoin / Sort / Equal / Less over STRING keys run
// on INT64 columns through the kernels that already exist, and ssb_string_gather materialises the
// bytes of the rows that survive.
//
// Ranking = most-significant-chunk-first refinement: at depth d every string contributes the
// big-endian image of its bytes [8d, 8d + 8) (zero padded) and the number of bytes it still has
// there (0..8, 9 = more follow); rows are sorted (stable one-sweep radix sort of sort.cu) by
// (class so far, chunk, count) and the runs of equal triples become the new classes. Zero padding
// is safe because the count separates "ab" from "ab\0". ceil(max_len / 8) rounds.
#include <cuda_runtime.h>

#include "common.h"
#include "device_utils.h"

namespace ssb {
namespace {

__global__ void __launch_bounds__(256) string_chunk_kernel(const long long* __restrict__ offsets, const unsigned char* __restrict__ bytes,
                                                            long long rows, long long depth, unsigned long long* __restrict__ chunk,
                                                            int* __restrict__ count) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < rows; i += stride) {
    const long long b = offsets[i], e = offsets[i + 1];
    const long long at = b + depth * 8;
    long long left = e - at;
    if (left < 0) left = 0;
    unsigned long long v = 0;
    const int n = left > 8 ? 8 : static_cast<int>(left);
    for (int k = 0; k < n; ++k) v |= static_cast<unsigned long long>(bytes[at + k]) << (56 - 8 * k);
    chunk[i] = v;
    count[i] = left > 8 ? 9 : n;
  }
}

__global__ void fill_zero_kernel(long long* p, long long n) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = 0;
}

// first_rows[c] = perm[starts[c]]
__global__ void representative_kernel(const long long* __restrict__ perm, const long long* __restrict__ starts, long long n,
                                      long long* __restrict__ out) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = perm[starts[i]];
}

// lengths of the selected strings (index < 0: an empty string), written as u64 for the scan; slot n = 0
__global__ void gather_lengths_kernel(const long long* __restrict__ offsets, const long long* __restrict__ idx, long long n,
                                      unsigned long long* __restrict__ out) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i <= n; i += stride) {
    unsigned long long len = 0;
    if (i < n) {
      const long long r = idx != nullptr ? idx[i] : i;
      if (r >= 0) len = static_cast<unsigned long long>(offsets[r + 1] - offsets[r]);
    }
    out[i] = len;
  }
}

// One warp per string: copies bytes [offsets[r], offsets[r + 1]) to out_bytes + out_offsets[i].
__global__ void __launch_bounds__(256) gather_bytes_kernel(const long long* __restrict__ offsets, const unsigned char* __restrict__ bytes,
                                                            const long long* __restrict__ idx, long long n,
                                                            const long long* __restrict__ out_offsets, unsigned char* __restrict__ out_bytes) {
  const int lane = threadIdx.x & 31;
  const long long warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) {
    const long long r = idx != nullptr ? idx[i] : i;
    if (r < 0) continue;
    const long long b = offsets[r], len = offsets[r + 1] - b;
    const unsigned char* src = bytes + b;
    unsigned char* dst = out_bytes + out_offsets[i];
    for (long long k = lane; k < len; k += 32) dst[k] = src[k];
  }
}

__global__ void shift_offsets_kernel(const long long* __restrict__ src, long long n, long long delta, long long* __restrict__ dst) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = src[i] + delta;
}

}  // namespace
}  // namespace ssb

using namespace ssb;

extern "C" {

int ssb_string_rank(ssb_ctx* ctx, const int64_t* d_offsets, const uint8_t* d_bytes, int64_t rows, int64_t max_len, int64_t* d_codes,
                    int64_t* d_first_rows, int64_t* n_distinct) {
  *n_distinct = 0;
  if (rows < 0 || max_len < 0) return fail(ctx, SSB_ERROR_INVALID_ARGUMENT_VALUE, "negative row count or length");
  if (rows == 0) return 0;
  const size_t n = static_cast<size_t>(rows);
  unsigned long long *chunk = nullptr, *chunk_s = nullptr;
  int *count = nullptr, *count_s = nullptr;
  long long *perm = nullptr, *ids_s = nullptr, *cid = nullptr, *starts = nullptr;
  cudaError_t e = tmp_malloc(ctx, &chunk, n * 8);
  if (e == cudaSuccess) e = tmp_malloc(ctx, &chunk_s, n * 8);
  if (e == cudaSuccess) e = tmp_malloc(ctx, &count, n * 4);
  if (e == cudaSuccess) e = tmp_malloc(ctx, &count_s, n * 4);
  if (e == cudaSuccess) e = tmp_malloc(ctx, &perm, n * 8);
  if (e == cudaSuccess) e = tmp_malloc(ctx, &ids_s, n * 8);
  if (e == cudaSuccess) e = tmp_malloc(ctx, &cid, n * 8);
  if (e == cudaSuccess) e = tmp_malloc(ctx, &starts, n * 8);
  auto release = [&]() {
    tmp_free(ctx, chunk); tmp_free(ctx, chunk_s); tmp_free(ctx, count); tmp_free(ctx, count_s);
    tmp_free(ctx, perm); tmp_free(ctx, ids_s); tmp_free(ctx, cid); tmp_free(ctx, starts);
  };
  if (e != cudaSuccess) { release(); return cuda_fail(ctx, e, "string rank scratch"); }
  const unsigned grid = grid_1d(ctx, rows, 256);
  fill_zero_kernel<<<grid, 256, 0, ctx->stream>>>(reinterpret_cast<long long*>(d_codes), rows);
  ++ctx->launches;
  const int64_t rounds = max_len <= 8 ? 1 : (max_len + 7) / 8;
  int rc = 0;
  int64_t clusters = 0;
  for (int64_t d = 0; d < rounds && rc == 0; ++d) {
    string_chunk_kernel<<<grid, 256, 0, ctx->stream>>>(reinterpret_cast<const long long*>(d_offsets), d_bytes, rows, d, chunk, count);
    ++ctx->launches;
    ssb_column keys[3];
    keys[0].data = d_codes; keys[0].nulls = nullptr; keys[0].dtype = SSB_INT64; keys[0].reserved = 0;
    keys[1].data = chunk; keys[1].nulls = nullptr; keys[1].dtype = SSB_UINT64; keys[1].reserved = 0;
    keys[2].data = count; keys[2].nulls = nullptr; keys[2].dtype = SSB_INT32; keys[2].reserved = 0;
    const int32_t asc[3] = {0, 0, 0};
    rc = ssb_sort_permutation(ctx, 3, keys, asc, rows, reinterpret_cast<int64_t*>(perm));
    if (rc) break;
    ssb_column sorted[3] = {keys[0], keys[1], keys[2]};
    sorted[0].data = ids_s; sorted[1].data = chunk_s; sorted[2].data = count_s;
    for (int c = 0; c < 3 && rc == 0; ++c) rc = ssb_gather(ctx, &keys[c], reinterpret_cast<const int64_t*>(perm), rows, &sorted[c]);
    if (rc) break;
    rc = ssb_cluster_ids(ctx, 3, sorted, rows, reinterpret_cast<int64_t*>(cid), reinterpret_cast<int64_t*>(starts), &clusters);
    if (rc) break;
    ssb_column src, dst;
    src.data = cid; src.nulls = nullptr; src.dtype = SSB_INT64; src.reserved = 0;
    dst = src; dst.data = d_codes;
    rc = ssb_scatter(ctx, &src, reinterpret_cast<const int64_t*>(perm), rows, &dst);
  }
  if (rc == 0 && d_first_rows != nullptr) {
    representative_kernel<<<grid_1d(ctx, clusters, 256), 256, 0, ctx->stream>>>(perm, starts, clusters, reinterpret_cast<long long*>(d_first_rows));
    ++ctx->launches;
  }
  if (rc == 0) {
    e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = cuda_fail(ctx, e, "string rank");
  }
  release();
  if (rc == 0) *n_distinct = clusters;
  return rc;
}

int ssb_string_gather_offsets(ssb_ctx* ctx, const int64_t* d_offsets, const int64_t* d_idx, int64_t n, int64_t* d_out_offsets,
                              int64_t* total_bytes) {
  *total_bytes = 0;
  if (n < 0) return fail(ctx, SSB_ERROR_INVALID_ARGUMENT_VALUE, "negative row count");
  gather_lengths_kernel<<<grid_1d(ctx, n + 1, 256), 256, 0, ctx->stream>>>(reinterpret_cast<const long long*>(d_offsets),
                                                                          reinterpret_cast<const long long*>(d_idx), n,
                                                                          reinterpret_cast<unsigned long long*>(d_out_offsets));
  ++ctx->launches;
  unsigned long long* d_total = nullptr;
  cudaError_t e = tmp_malloc(ctx, &d_total, 8);
  if (e != cudaSuccess) return cuda_fail(ctx, e, "string gather scratch");
  int rc = exclusive_scan_u64(ctx, reinterpret_cast<unsigned long long*>(d_out_offsets), static_cast<unsigned long long>(n) + 1, d_total);
  if (rc == 0) {
    e = cudaMemcpyAsync(ctx->h_count, d_total, 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = cuda_fail(ctx, e, "string gather offsets");
    else *total_bytes = *ctx->h_count;
  }
  tmp_free(ctx, d_total);
  return rc;
}

int ssb_string_gather_bytes(ssb_ctx* ctx, const int64_t* d_offsets, const uint8_t* d_bytes, const int64_t* d_idx, int64_t n,
                            const int64_t* d_out_offsets, uint8_t* d_out_bytes) {
  if (n <= 0) return 0;
  long long warps = n;
  long long blocks = (warps * 32 + 255) / 256;
  const long long cap = static_cast<long long>(ctx->num_sms) * 16;
  if (blocks > cap) blocks = cap;
  gather_bytes_kernel<<<static_cast<unsigned>(blocks), 256, 0, ctx->stream>>>(reinterpret_cast<const long long*>(d_offsets), d_bytes,
                                                                             reinterpret_cast<const long long*>(d_idx), n,
                                                                             reinterpret_cast<const long long*>(d_out_offsets), d_out_bytes);
  ++ctx->launches;
  SSB_CUDA(ctx, cudaGetLastError());
  return 0;
}

int ssb_string_shift_offsets(ssb_ctx* ctx, const int64_t* d_src, int64_t n, int64_t delta, int64_t* d_dst) {
  if (n <= 0) return 0;
  shift_offsets_kernel<<<grid_1d(ctx, n, 256), 256, 0, ctx->stream>>>(reinterpret_cast<const long long*>(d_src), n, delta,
                                                                      reinterpret_cast<long long*>(d_dst));
  ++ctx->launches;
  SSB_CUDA(ctx, cudaGetLastError());
  return 0;
}

}  // extern "C"
