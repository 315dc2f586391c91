// comm.cu -- the multi-GPU seam of the C ABI: one process per GPU, a
// communicator per context, ragged all-to-all / all-gather of column slices, and the exchange
// step of the row-range sharded operators built on them.
//
// The reference has no counterpart (cursor/core/* runs on one thread of one host); the operators
// that need an exchange are the ones whose CPU form keeps global state: GroupAggregate
// (aggregate_groups.cc:332-433: one hash set for the whole input) and HashJoin (hash_join.cc:
// 406-517: one index over the whole rhs). Row-range shards of Compute / Filter / Project need none.
//
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy the process already holds, e.g.
// PyTorch's, or the system library), so single-GPU users of libssb200.so never load it. All
// transfers are enqueued on the context's stream: kernels and collectives stay ordered without
// host synchronisation; only row counts travel through the host (NCCL's send / receive sizes are
// host values).
#include <dlfcn.h>
#include <nccl.h>
#include <stdio.h>
#include <string.h>
#include <unistd.h>

#include <string>
#include <vector>

#include "comm.h"
#include "common.h"

namespace ssb {
namespace {

struct Nccl {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string error;
};

Nccl* nccl() {
  static Nccl* n = [] {
    Nccl* x = new Nccl;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* name : names) {
      x->lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (x->lib) break;
    }
    if (!x->lib) { x->error = std::string("cannot load NCCL: ") + dlerror(); return x; }
#define SSB_NCCL_SYM(field, sym)                                                     \
  x->field = reinterpret_cast<decltype(x->field)>(dlsym(x->lib, sym));               \
  if (!x->field && x->error.empty()) x->error = std::string("NCCL symbol missing: ") + sym;
    SSB_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    SSB_NCCL_SYM(CommInitRank, "ncclCommInitRank")
    SSB_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    SSB_NCCL_SYM(GroupStart, "ncclGroupStart")
    SSB_NCCL_SYM(GroupEnd, "ncclGroupEnd")
    SSB_NCCL_SYM(Send, "ncclSend")
    SSB_NCCL_SYM(Recv, "ncclRecv")
    SSB_NCCL_SYM(AllGather, "ncclAllGather")
    SSB_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef SSB_NCCL_SYM
    return x;
  }();
  return n;
}

int nccl_fail(ssb_ctx* ctx, ncclResult_t r, const char* what) {
  Nccl* n = nccl();
  return fail(ctx, SSB_ERROR_UNKNOWN, std::string(what) + ": " + (n->GetErrorString ? n->GetErrorString(r) : "NCCL error"));
}

#define SSB_NCCL(ctx, call)                                      \
  do {                                                           \
    ncclResult_t r_ = (call);                                    \
    if (r_ != ncclSuccess) return nccl_fail((ctx), r_, #call);   \
  } while (0)

}  // namespace
}  // namespace ssb

using namespace ssb;

struct ssb_comm {
  ssb_ctx* ctx;
  ncclComm_t comm;
  int world, rank;
  int64_t* d_counts;   // [2 * world * kCountSlots] staging of row counts
  int64_t* h_counts;   // pinned mirror
};

namespace ssb {

int comm_world(const ssb_comm* c) { return c->world; }
int comm_rank(const ssb_comm* c) { return c->rank; }
ssb_ctx* comm_ctx(const ssb_comm* c) { return c->ctx; }

enum { kCountSlots = 8 };

// Every rank contributes `n` (<= kCountSlots) int64 values; h_all[r * n + i] = value i of rank r.
int comm_all_gather_counts(ssb_comm* c, const int64_t* h_mine, int n, int64_t* h_all) {
  ssb_ctx* ctx = c->ctx;
  if (n < 1 || n > kCountSlots) return fail(ctx, SSB_ERROR_INVALID_ARGUMENT_VALUE, "too many counts");
  Nccl* nc = nccl();
  int64_t* h_send = c->h_counts;
  int64_t* h_recv = c->h_counts + kCountSlots;
  for (int i = 0; i < n; ++i) h_send[i] = h_mine[i];
  int64_t* d_send = c->d_counts;
  int64_t* d_recv = c->d_counts + kCountSlots;
  SSB_CUDA(ctx, cudaMemcpyAsync(d_send, h_send, sizeof(int64_t) * n, cudaMemcpyHostToDevice, ctx->stream));
  SSB_NCCL(ctx, nc->AllGather(d_send, d_recv, static_cast<size_t>(n), ncclInt64, c->comm, ctx->stream));
  SSB_CUDA(ctx, cudaMemcpyAsync(h_recv, d_recv, sizeof(int64_t) * n * c->world, cudaMemcpyDeviceToHost, ctx->stream));
  SSB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < n * c->world; ++i) h_all[i] = h_recv[i];
  return 0;
}

// h_recv[r] = the value rank r holds for this rank in its h_send[this rank].
int comm_exchange_counts(ssb_comm* c, const int64_t* h_send, int64_t* h_recv) {
  ssb_ctx* ctx = c->ctx;
  Nccl* nc = nccl();
  const int W = c->world;
  int64_t* hs = c->h_counts;
  int64_t* hr = c->h_counts + kCountSlots * (1 + W);   // behind the all-gather area
  int64_t* ds = c->d_counts;
  int64_t* dr = c->d_counts + kCountSlots * (1 + W);
  for (int r = 0; r < W; ++r) hs[r] = h_send[r];
  SSB_CUDA(ctx, cudaMemcpyAsync(ds, hs, sizeof(int64_t) * W, cudaMemcpyHostToDevice, ctx->stream));
  SSB_NCCL(ctx, nc->GroupStart());
  for (int r = 0; r < W; ++r) {
    SSB_NCCL(ctx, nc->Send(ds + r, 1, ncclInt64, r, c->comm, ctx->stream));
    SSB_NCCL(ctx, nc->Recv(dr + r, 1, ncclInt64, r, c->comm, ctx->stream));
  }
  SSB_NCCL(ctx, nc->GroupEnd());
  SSB_CUDA(ctx, cudaMemcpyAsync(hr, dr, sizeof(int64_t) * W, cudaMemcpyDeviceToHost, ctx->stream));
  SSB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (int r = 0; r < W; ++r) h_recv[r] = hr[r];
  return 0;
}

// One fused exchange of several columns: column i is cut into `world` consecutive slices of
// send_rows[r] elements (width[i] bytes each); slice r goes to rank r, and the slices received
// from ranks 0..world-1 are laid out consecutively (recv_rows[r] elements each) in recv[i]. All
// columns travel in ONE NCCL group (one launch, every peer pair busy at once). Asynchronous.
int comm_all_to_all_v(ssb_comm* c, int n_cols, const void* const* send, void* const* recv, const int32_t* width,
                      const int64_t* send_rows, const int64_t* recv_rows) {
  ssb_ctx* ctx = c->ctx;
  Nccl* nc = nccl();
  const int W = c->world;
  SSB_NCCL(ctx, nc->GroupStart());
  for (int i = 0; i < n_cols; ++i) {
    size_t so = 0, ro = 0;
    for (int r = 0; r < W; ++r) {
      const size_t sb = static_cast<size_t>(send_rows[r]) * width[i], rb = static_cast<size_t>(recv_rows[r]) * width[i];
      if (r == c->rank) {
        // the own slice stays on the device
        if (sb) SSB_CUDA(ctx, cudaMemcpyAsync(static_cast<char*>(recv[i]) + ro, static_cast<const char*>(send[i]) + so, sb,
                                              cudaMemcpyDeviceToDevice, ctx->stream));
      } else {
        if (sb) SSB_NCCL(ctx, nc->Send(static_cast<const char*>(send[i]) + so, sb, ncclInt8, r, c->comm, ctx->stream));
        if (rb) SSB_NCCL(ctx, nc->Recv(static_cast<char*>(recv[i]) + ro, rb, ncclInt8, r, c->comm, ctx->stream));
      }
      so += sb;
      ro += rb;
    }
  }
  SSB_NCCL(ctx, nc->GroupEnd());
  return 0;
}

// Ragged all-gather of several columns: every rank contributes my_rows elements per column and
// receives all ranks' contributions in rank order (all_rows[r] elements from rank r). Asynchronous.
int comm_all_gather_v(ssb_comm* c, int n_cols, const void* const* send, void* const* recv, const int32_t* width,
                      const int64_t* all_rows) {
  ssb_ctx* ctx = c->ctx;
  Nccl* nc = nccl();
  const int W = c->world;
  SSB_NCCL(ctx, nc->GroupStart());
  for (int i = 0; i < n_cols; ++i) {
    size_t ro = 0;
    const size_t sb = static_cast<size_t>(all_rows[c->rank]) * width[i];
    for (int r = 0; r < W; ++r) {
      const size_t rb = static_cast<size_t>(all_rows[r]) * width[i];
      if (r == c->rank) {
        if (sb) SSB_CUDA(ctx, cudaMemcpyAsync(static_cast<char*>(recv[i]) + ro, send[i], sb, cudaMemcpyDeviceToDevice, ctx->stream));
      } else {
        if (sb) SSB_NCCL(ctx, nc->Send(send[i], sb, ncclInt8, r, c->comm, ctx->stream));
        if (rb) SSB_NCCL(ctx, nc->Recv(static_cast<char*>(recv[i]) + ro, rb, ncclInt8, r, c->comm, ctx->stream));
      }
      ro += rb;
    }
  }
  SSB_NCCL(ctx, nc->GroupEnd());
  return 0;
}

}  // namespace ssb

extern "C" {

int ssb_comm_unique_id(uint8_t* id) {
  Nccl* nc = nccl();
  if (!nc->error.empty()) return SSB_ERROR_UNKNOWN;
  ncclUniqueId u;
  static_assert(sizeof(ncclUniqueId) == SSB_COMM_ID_BYTES, "NCCL unique id size");
  if (nc->GetUniqueId(&u) != ncclSuccess) return SSB_ERROR_UNKNOWN;
  memcpy(id, &u, sizeof(u));
  return 0;
}

int ssb_comm_create(ssb_ctx* ctx, const uint8_t* id, int32_t world, int32_t rank, ssb_comm** out) {
  *out = nullptr;
  Nccl* nc = nccl();
  if (!nc->error.empty()) return fail(ctx, SSB_ERROR_UNKNOWN, nc->error);
  if (world < 1 || world > 64 || rank < 0 || rank >= world) return fail(ctx, SSB_ERROR_INVALID_ARGUMENT_VALUE, "communicator: bad world / rank");
  SSB_CUDA(ctx, cudaSetDevice(ctx->device));
  ncclUniqueId u;
  memcpy(&u, id, sizeof(u));
  ssb_comm* c = new ssb_comm;
  c->ctx = ctx;
  c->world = world;
  c->rank = rank;
  c->comm = nullptr;
  c->d_counts = nullptr;
  c->h_counts = nullptr;
  ncclResult_t r = nc->CommInitRank(&c->comm, world, u, rank);
  if (r != ncclSuccess) { delete c; return nccl_fail(ctx, r, "ncclCommInitRank"); }
  const size_t n = static_cast<size_t>(kCountSlots) * (1 + world) * 2 + 2 * world;
  cudaError_t e = cudaMalloc(&c->d_counts, n * sizeof(int64_t));
  if (e == cudaSuccess) e = cudaMallocHost(&c->h_counts, n * sizeof(int64_t));
  if (e != cudaSuccess) { ssb_comm_destroy(c); return cuda_fail(ctx, e, "communicator buffers"); }
  *out = c;
  return 0;
}

int ssb_comm_create_file(ssb_ctx* ctx, const char* path, int32_t world, int32_t rank, int32_t timeout_s, ssb_comm** out) {
  *out = nullptr;
  uint8_t id[SSB_COMM_ID_BYTES];
  const std::string tmp = std::string(path) + ".tmp";
  if (rank == 0) {
    if (ssb_comm_unique_id(id) != 0) return fail(ctx, SSB_ERROR_UNKNOWN, nccl()->error.empty() ? "ncclGetUniqueId failed" : nccl()->error);
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f || fwrite(id, 1, sizeof(id), f) != sizeof(id)) { if (f) fclose(f); return fail(ctx, SSB_ERROR_UNKNOWN, "cannot write the rendezvous file"); }
    fclose(f);
    if (rename(tmp.c_str(), path) != 0) return fail(ctx, SSB_ERROR_UNKNOWN, "cannot publish the rendezvous file");
  } else {
    bool got = false;
    for (int waited = 0; waited < timeout_s * 20 && !got; ++waited) {
      FILE* f = fopen(path, "rb");
      if (f) {
        got = fread(id, 1, sizeof(id), f) == sizeof(id);
        fclose(f);
      }
      if (!got) usleep(50000);
    }
    if (!got) return fail(ctx, SSB_ERROR_UNKNOWN, "timed out waiting for the rendezvous file");
  }
  return ssb_comm_create(ctx, id, world, rank, out);
}

void ssb_comm_destroy(ssb_comm* c) {
  if (!c) return;
  if (c->comm) nccl()->CommDestroy(c->comm);
  if (c->d_counts) cudaFree(c->d_counts);
  if (c->h_counts) cudaFreeHost(c->h_counts);
  delete c;
}

int32_t ssb_comm_rank(const ssb_comm* c) { return c->rank; }
int32_t ssb_comm_size(const ssb_comm* c) { return c->world; }

int ssb_comm_exchange_counts(ssb_comm* c, const int64_t* h_send, int64_t* h_recv) {
  return comm_exchange_counts(c, h_send, h_recv);
}

int ssb_comm_all_to_all(ssb_comm* c, int32_t n_cols, const void* const* send, void* const* recv, const int32_t* width,
                        const int64_t* send_rows, const int64_t* recv_rows) {
  if (n_cols < 0) return fail(c->ctx, SSB_ERROR_INVALID_ARGUMENT_VALUE, "negative column count");
  return comm_all_to_all_v(c, n_cols, send, recv, width, send_rows, recv_rows);
}

}  // extern "C"
