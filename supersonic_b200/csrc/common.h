// common.h -- context object and helpers shared by the translation units of libssb200.so.
#ifndef SSB_CSRC_COMMON_H_
#define SSB_CSRC_COMMON_H_

#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/supersonic_b200.h"
#include "ops.h"

struct ssb_ctx {
  int device;
  cudaStream_t stream;
  int num_sms;
  size_t smem_optin;          // max dynamic shared memory per block
  size_t smem_per_sm, smem_reserved;
  std::string last_error;
  int64_t launches;
  bool timing;
  cudaEvent_t ev0, ev1;
  cudaEvent_t tm0, tm1;       // user stopwatch
  bool ev_valid;
  // reusable device scratch (grown on demand)
  void* scratch;
  size_t scratch_bytes;
  int32_t* d_fail;            // failure flag of signaling ops
  int64_t* d_count;           // one int64 result slot
  int64_t* h_count;           // pinned mirror
  int32_t* h_fail;            // pinned mirror
  // large temporaries (tmp_malloc): freed blocks are kept per context and handed out again by size, so that an
  // operator called in a loop neither pays cudaMalloc nor makes the driver's pool re-map memory every time
  std::multimap<size_t, void*> tmp_cache;            // free blocks by size
  std::unordered_map<void*, size_t> tmp_blocks;      // every block obtained through the cache (in use or free)
  size_t tmp_cached_bytes;
};

namespace ssb {

int fail(ssb_ctx* ctx, int code, const std::string& msg);
int cuda_fail(ssb_ctx* ctx, cudaError_t e, const char* what);
// Returns the context scratch buffer, at least `bytes` large (contents undefined).
int scratch(ssb_ctx* ctx, size_t bytes, void** out);

// Stream-ordered temporaries from the device's memory pool (cudaMallocAsync with the release
// threshold lifted in ssb_ctx_create): repeated operator calls reuse their scratch instead of
// paying a cudaMalloc / cudaFree pair (and its implicit device synchronisation) every time.
cudaError_t tmp_malloc_bytes(ssb_ctx* ctx, void** out, size_t bytes);
void tmp_free(ssb_ctx* ctx, void* ptr);
// Gives the cached large temporaries of every context on `device` back to the driver (out-of-memory paths).
void tmp_release_cached(int device);
template <class T>
inline cudaError_t tmp_malloc(ssb_ctx* ctx, T** out, size_t bytes) {
  return tmp_malloc_bytes(ctx, reinterpret_cast<void**>(out), bytes);
}

#define SSB_CUDA(ctx, call)                                          \
  do {                                                               \
    cudaError_t e_ = (call);                                         \
    if (e_ != cudaSuccess) return ::ssb::cuda_fail((ctx), e_, #call); \
  } while (0)

struct TimedRegion {
  explicit TimedRegion(ssb_ctx* c) : ctx(c) {
    if (ctx->timing) cudaEventRecord(ctx->ev0, ctx->stream);
  }
  ~TimedRegion() {
    if (ctx->timing) { cudaEventRecord(ctx->ev1, ctx->stream); ctx->ev_valid = true; }
  }
  ssb_ctx* ctx;
};

// SSB_* dtype -> physical type, or -1.
inline int phys_of(int dtype) {
  switch (dtype) {
    case SSB_INT32: case SSB_DATE: case SSB_ENUM: return T_I32;
    case SSB_INT64: case SSB_DATETIME: return T_I64;
    case SSB_UINT32: return T_U32;
    case SSB_UINT64: return T_U64;
    case SSB_FLOAT: return T_F32;
    case SSB_DOUBLE: return T_F64;
    case SSB_BOOL: return T_B8;
    default: return -1;
  }
}
inline int width_of(int dtype) { int p = phys_of(dtype); return p < 0 ? 0 : phys_width(p); }

inline int64_t div_up(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace ssb
#endif  // SSB_CSRC_COMMON_H_
