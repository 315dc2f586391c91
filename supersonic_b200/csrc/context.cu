// context.cu -- context, device memory, null-vector conversion and the synthetic generator.
#include <cuda_runtime.h>

#include <mutex>
#include <stdio.h>

#include <atomic>

#include "common.h"

namespace ssb {

int fail(ssb_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->last_error = msg;
  return code;
}

int cuda_fail(ssb_ctx* ctx, cudaError_t e, const char* what) {
  std::string msg = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what;
  cudaGetLastError();
  return fail(ctx, e == cudaErrorMemoryAllocation ? SSB_ERROR_MEMORY_EXCEEDED : SSB_ERROR_UNKNOWN, msg);
}

int scratch(ssb_ctx* ctx, size_t bytes, void** out) {
  if (bytes > ctx->scratch_bytes) {
    if (ctx->scratch) {
      SSB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      SSB_CUDA(ctx, cudaFree(ctx->scratch));
      ctx->scratch = nullptr;
      ctx->scratch_bytes = 0;
    }
    size_t want = bytes + bytes / 4 + 4096;
    SSB_CUDA(ctx, cudaMalloc(&ctx->scratch, want));
    ctx->scratch_bytes = want;
  }
  *out = ctx->scratch;
  return 0;
}

// Temporaries. Small ones come from the device's stream-ordered pool (cudaMallocAsync). Large ones (>= 1 MiB) are
// cudaMalloc blocks cached per context: the stream-ordered pool satisfies a mix of large sizes by splitting and
// re-mapping cached memory, which costs milliseconds per call once the free list is fragmented (measured: the sharded
// join's partition 0.3 -> 2.0 ms, build 1.4 -> 3.1 ms, probe 6.0 -> 10.5 ms from its second call on). A block is
// reused only by the context that freed it, so the order of work on the context's one stream keeps reuse safe.
namespace {
std::mutex& tmp_mutex() { static std::mutex m; return m; }
std::vector<ssb_ctx*>& tmp_contexts() { static std::vector<ssb_ctx*> v; return v; }
constexpr size_t kTmpLarge = size_t(1) << 20;
constexpr size_t kTmpCacheLimit = size_t(48) << 30;   // cached (free) bytes per context before the largest blocks go back

// Frees the cached blocks of every context on `device` (lock held by the caller).
void tmp_flush_device(int device) {
  for (ssb_ctx* c : tmp_contexts()) {
    if (c->device != device) continue;
    for (auto& kv : c->tmp_cache) { cudaFree(kv.second); c->tmp_blocks.erase(kv.second); }
    c->tmp_cache.clear();
    c->tmp_cached_bytes = 0;
  }
}
}  // namespace

void tmp_register(ssb_ctx* ctx) {
  std::lock_guard<std::mutex> lock(tmp_mutex());
  ctx->tmp_cached_bytes = 0;
  tmp_contexts().push_back(ctx);
}
void tmp_unregister(ssb_ctx* ctx) {
  std::lock_guard<std::mutex> lock(tmp_mutex());
  for (auto& kv : ctx->tmp_cache) cudaFree(kv.second);
  ctx->tmp_cache.clear();
  ctx->tmp_blocks.clear();
  ctx->tmp_cached_bytes = 0;
  std::vector<ssb_ctx*>& v = tmp_contexts();
  for (size_t i = 0; i < v.size(); ++i) if (v[i] == ctx) { v.erase(v.begin() + i); break; }
}
void tmp_release_cached(int device) {
  std::lock_guard<std::mutex> lock(tmp_mutex());
  tmp_flush_device(device);
}

cudaError_t tmp_malloc_bytes(ssb_ctx* ctx, void** out, size_t bytes) {
  *out = nullptr;
  if (bytes >= kTmpLarge) {
    const size_t granule = size_t(2) << 20;
    const size_t want = (bytes + granule - 1) / granule * granule;
    std::lock_guard<std::mutex> lock(tmp_mutex());
    std::multimap<size_t, void*>::iterator it = ctx->tmp_cache.lower_bound(want);
    if (it != ctx->tmp_cache.end() && it->first <= want + want / 8) {   // a close fit: big blocks are not cut up
      *out = it->second;
      ctx->tmp_cached_bytes -= it->first;
      ctx->tmp_cache.erase(it);
      return cudaSuccess;
    }
    cudaError_t e = cudaMalloc(out, want);
    if (e == cudaErrorMemoryAllocation) {
      cudaGetLastError();
      cudaStreamSynchronize(ctx->stream);
      tmp_flush_device(ctx->device);
      cudaMemPool_t pool;
      if (cudaDeviceGetDefaultMemPool(&pool, ctx->device) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
      e = cudaMalloc(out, want);
    }
    if (e == cudaSuccess) ctx->tmp_blocks[*out] = want;
    return e;
  }
  cudaError_t e = cudaMallocAsync(out, bytes ? bytes : 8, ctx->stream);
  if (e == cudaErrorMemoryAllocation) {
    // give cached blocks back and try once more
    cudaGetLastError();
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, ctx->device) == cudaSuccess) {
      cudaStreamSynchronize(ctx->stream);
      cudaMemPoolTrimTo(pool, 0);
    }
    tmp_release_cached(ctx->device);
    e = cudaMallocAsync(out, bytes ? bytes : 8, ctx->stream);
  }
  return e;
}

void tmp_free(ssb_ctx* ctx, void* ptr) {
  if (!ptr) return;
  {
    std::lock_guard<std::mutex> lock(tmp_mutex());
    std::unordered_map<void*, size_t>::iterator it = ctx->tmp_blocks.find(ptr);
    if (it == ctx->tmp_blocks.end()) {
      // a block of another context of this process: pending work of the freeing stream must not outlive it
      for (ssb_ctx* other : tmp_contexts()) {
        if (other == ctx) continue;
        std::unordered_map<void*, size_t>::iterator jt = other->tmp_blocks.find(ptr);
        if (jt == other->tmp_blocks.end()) continue;
        cudaStreamSynchronize(ctx->stream);
        other->tmp_cache.insert(std::make_pair(jt->second, ptr));
        other->tmp_cached_bytes += jt->second;
        return;
      }
    }
    if (it != ctx->tmp_blocks.end()) {
      ctx->tmp_cache.insert(std::make_pair(it->second, ptr));
      ctx->tmp_cached_bytes += it->second;
      while (ctx->tmp_cached_bytes > kTmpCacheLimit && !ctx->tmp_cache.empty()) {   // the largest blocks go first
        std::multimap<size_t, void*>::iterator last = --ctx->tmp_cache.end();
        cudaStreamSynchronize(ctx->stream);
        cudaFree(last->second);
        ctx->tmp_blocks.erase(last->second);
        ctx->tmp_cached_bytes -= last->first;
        ctx->tmp_cache.erase(last);
      }
      return;
    }
  }
  cudaFreeAsync(ptr, ctx->stream);
}

// bool per row -> bitmap: each warp packs 32 rows with one ballot.
__global__ void pack_nulls_kernel(const uint8_t* __restrict__ bools, long long rows,
                                  uint32_t* __restrict__ bitmap) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long rows_up = (rows + 31) & ~31LL;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < rows_up; i += stride) {
    const bool v = i < rows && bools[i] != 0;
    const uint32_t w = __ballot_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0) bitmap[i >> 5] = w;
  }
}

__global__ void unpack_nulls_kernel(const uint32_t* __restrict__ bitmap, long long rows,
                                    uint8_t* __restrict__ bools) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < rows; i += stride) {
    bools[i] = (bitmap[i >> 5] >> (i & 31)) & 1u;
  }
}

__host__ __device__ inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

__host__ __device__ inline uint64_t gen_value(uint64_t seed, uint64_t stream, uint64_t row, int kind,
                                              int64_t lo, uint64_t span) {
  const uint64_t u = splitmix64((seed ^ (stream * 0x9E3779B97F4A7C15ull)) + row);
  switch (kind) {
    case 0: return static_cast<uint64_t>(lo) + (span ? (u & (span - 1)) : u);
    case 1: return static_cast<uint64_t>(lo) + (u % span);
    case 2: { double d = static_cast<double>(u >> 44) * (1.0 / 1024.0); uint64_t b; memcpy(&b, &d, 8); return b; }
    case 4: return (row * static_cast<uint64_t>(lo)) % span;   // a permutation of [0, span) when gcd(lo, span) = 1
    case 5: { double d = static_cast<double>(lo + static_cast<int64_t>(u % span)) * (1.0 / 16.0); uint64_t b; memcpy(&b, &d, 8); return b; }
    default: { double d = static_cast<double>(u >> 11) * (1.0 / 9007199254740992.0); uint64_t b; memcpy(&b, &d, 8); return b; }
  }
}

__global__ void generate_kernel(uint64_t* __restrict__ out, long long rows, long long first_row,
                                uint64_t seed, uint64_t stream, int kind, int64_t lo, uint64_t span) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < rows; i += stride) {
    out[i] = gen_value(seed, stream, static_cast<uint64_t>(first_row + i), kind, lo, span);
  }
}

}  // namespace ssb

using namespace ssb;

extern "C" {

int ssb_abi_version(void) { return SSB_ABI_VERSION; }

int ssb_ctx_create(int device, ssb_ctx** out) {
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0 || device < 0 || device >= count) {
    cudaGetLastError();
    return SSB_ERROR_UNKNOWN;   // no CUDA device: there is no fallback
  }
  ssb_ctx* ctx = new ssb_ctx;
  ctx->device = device;
  ctx->launches = 0;
  ctx->timing = false;
  ctx->ev_valid = false;
  ctx->scratch = nullptr;
  ctx->scratch_bytes = 0;
  if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return SSB_ERROR_UNKNOWN; }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  ctx->num_sms = prop.multiProcessorCount;
  ctx->smem_optin = prop.sharedMemPerBlockOptin;
  ctx->smem_per_sm = prop.sharedMemPerMultiprocessor;
  ctx->smem_reserved = prop.reservedSharedMemPerBlock;
  if (prop.major < 10) {
    fprintf(stderr, "libssb200: device %d is sm_%d%d; this library carries sm_100a code only\n",
            device, prop.major, prop.minor);
    delete ctx;
    return SSB_ERROR_NOT_IMPLEMENTED;
  }
  cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
  tmp_register(ctx);
  {
    // keep freed temporaries cached in the pool (default: released at the next synchronisation)
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      uint64_t keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
  }
  cudaEventCreate(&ctx->ev0);
  cudaEventCreate(&ctx->ev1);
  cudaEventCreate(&ctx->tm0);
  cudaEventCreate(&ctx->tm1);
  cudaMalloc(&ctx->d_fail, sizeof(int32_t));
  cudaMalloc(&ctx->d_count, sizeof(int64_t));
  cudaMemset(ctx->d_fail, 0, sizeof(int32_t));
  cudaMallocHost(&ctx->h_count, sizeof(int64_t));
  cudaMallocHost(&ctx->h_fail, sizeof(int32_t));
  if (cudaGetLastError() != cudaSuccess) { delete ctx; return SSB_ERROR_UNKNOWN; }
  *out = ctx;
  return 0;
}

void ssb_ctx_destroy(ssb_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  tmp_unregister(ctx);
  if (ctx->scratch) cudaFree(ctx->scratch);
  cudaFree(ctx->d_fail);
  cudaFree(ctx->d_count);
  cudaFreeHost(ctx->h_count);
  cudaFreeHost(ctx->h_fail);
  cudaEventDestroy(ctx->ev0);
  cudaEventDestroy(ctx->ev1);
  cudaEventDestroy(ctx->tm0);
  cudaEventDestroy(ctx->tm1);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* ssb_last_error(const ssb_ctx* ctx) { return ctx ? ctx->last_error.c_str() : "no context"; }
void* ssb_ctx_stream(ssb_ctx* ctx) { return ctx->stream; }
int64_t ssb_ctx_launch_count(const ssb_ctx* ctx) { return ctx->launches; }

int ssb_ctx_sync(ssb_ctx* ctx) {
  SSB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

int ssb_ctx_enable_timing(ssb_ctx* ctx, int enable) {
  ctx->timing = enable != 0;
  ctx->ev_valid = false;
  return 0;
}

int ssb_ctx_last_kernel_ms(ssb_ctx* ctx, float* ms) {
  if (!ctx->timing || !ctx->ev_valid) return fail(ctx, SSB_ERROR_INVALID_ARGUMENT_VALUE, "no timed region recorded");
  SSB_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
  SSB_CUDA(ctx, cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
  return 0;
}

int ssb_ctx_timer_start(ssb_ctx* ctx) {
  SSB_CUDA(ctx, cudaEventRecord(ctx->tm0, ctx->stream));
  return 0;
}
int ssb_ctx_timer_stop(ssb_ctx* ctx, float* ms) {
  SSB_CUDA(ctx, cudaEventRecord(ctx->tm1, ctx->stream));
  SSB_CUDA(ctx, cudaEventSynchronize(ctx->tm1));
  SSB_CUDA(ctx, cudaEventElapsedTime(ms, ctx->tm0, ctx->tm1));
  return 0;
}

int ssb_malloc(ssb_ctx* ctx, size_t bytes, void** out) {
  *out = nullptr;
  SSB_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaError_t e = cudaMalloc(out, bytes ? bytes : 1);
  if (e == cudaErrorMemoryAllocation) {
    cudaGetLastError();
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, ctx->device) == cudaSuccess) {
      cudaStreamSynchronize(ctx->stream);
      cudaMemPoolTrimTo(pool, 0);
    }
    tmp_release_cached(ctx->device);
    e = cudaMalloc(out, bytes ? bytes : 1);
  }
  SSB_CUDA(ctx, e);
  return 0;
}
int ssb_free(ssb_ctx* ctx, void* ptr) {
  if (ptr) SSB_CUDA(ctx, cudaFree(ptr));
  return 0;
}
int ssb_malloc_host(ssb_ctx* ctx, size_t bytes, void** out) {
  *out = nullptr;
  SSB_CUDA(ctx, cudaMallocHost(out, bytes ? bytes : 1));
  return 0;
}
int ssb_free_host(ssb_ctx* ctx, void* ptr) {
  if (ptr) SSB_CUDA(ctx, cudaFreeHost(ptr));
  return 0;
}
static std::atomic<unsigned long long> g_h2d_bytes(0), g_d2h_bytes(0);

int ssb_memcpy_h2d(ssb_ctx* ctx, void* dst, const void* src, size_t bytes) {
  if (bytes) SSB_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  g_h2d_bytes.fetch_add(bytes, std::memory_order_relaxed);
  return 0;
}
int ssb_memcpy_d2h(ssb_ctx* ctx, void* dst, const void* src, size_t bytes) {
  if (bytes) SSB_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  g_d2h_bytes.fetch_add(bytes, std::memory_order_relaxed);
  return 0;
}
void ssb_transfer_bytes(uint64_t* h2d, uint64_t* d2h) {
  if (h2d) *h2d = g_h2d_bytes.load(std::memory_order_relaxed);
  if (d2h) *d2h = g_d2h_bytes.load(std::memory_order_relaxed);
}
int ssb_memcpy_d2d(ssb_ctx* ctx, void* dst, const void* src, size_t bytes) {
  if (bytes) SSB_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  return 0;
}
int ssb_memset(ssb_ctx* ctx, void* dst, int value, size_t bytes) {
  if (bytes) SSB_CUDA(ctx, cudaMemsetAsync(dst, value, bytes, ctx->stream));
  return 0;
}

int ssb_pointer_is_device(const void* ptr) {
  if (ptr == nullptr) return 0;
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, ptr) != cudaSuccess) { cudaGetLastError(); return 0; }
  return attr.type == cudaMemoryTypeDevice ? 1 : 0;
}

static unsigned grid_for(ssb_ctx* ctx, int64_t n, int block) {
  int64_t g = div_up(n, block);
  const int64_t cap = static_cast<int64_t>(ctx->num_sms) * 8;
  if (g > cap) g = cap;
  return static_cast<unsigned>(g < 1 ? 1 : g);
}

int ssb_nulls_pack(ssb_ctx* ctx, const uint8_t* d_bools, int64_t rows, uint32_t* d_bitmap) {
  if (rows <= 0) return 0;
  pack_nulls_kernel<<<grid_for(ctx, rows, 256), 256, 0, ctx->stream>>>(d_bools, rows, d_bitmap);
  ++ctx->launches;
  SSB_CUDA(ctx, cudaGetLastError());
  return 0;
}

int ssb_nulls_unpack(ssb_ctx* ctx, const uint32_t* d_bitmap, int64_t rows, uint8_t* d_bools) {
  if (rows <= 0) return 0;
  unpack_nulls_kernel<<<grid_for(ctx, rows, 256), 256, 0, ctx->stream>>>(d_bitmap, rows, d_bools);
  ++ctx->launches;
  SSB_CUDA(ctx, cudaGetLastError());
  return 0;
}

int ssb_generate(ssb_ctx* ctx, void* d_out, int64_t rows, int64_t first_row, uint64_t seed,
                 uint64_t stream, int kind, int64_t lo, uint64_t span) {
  if (rows <= 0) return 0;
  if ((kind == 1 || kind == 4 || kind == 5) && span == 0) return fail(ctx, SSB_ERROR_INVALID_ARGUMENT_VALUE, "this kind needs a span");
  if (kind == 0 && (span & (span - 1))) return fail(ctx, SSB_ERROR_INVALID_ARGUMENT_VALUE, "kind 0 needs a power-of-two span");
  generate_kernel<<<grid_for(ctx, rows, 256), 256, 0, ctx->stream>>>(
      static_cast<uint64_t*>(d_out), rows, first_row, seed, stream, kind, lo, span);
  ++ctx->launches;
  SSB_CUDA(ctx, cudaGetLastError());
  return 0;
}

void ssb_generate_host(void* out, int64_t rows, int64_t first_row, uint64_t seed, uint64_t stream,
                       int kind, int64_t lo, uint64_t span) {
  uint64_t* o = static_cast<uint64_t*>(out);
  for (int64_t i = 0; i < rows; ++i) {
    o[i] = gen_value(seed, stream, static_cast<uint64_t>(first_row + i), kind, lo, span);
  }
}

}  // extern "C"
