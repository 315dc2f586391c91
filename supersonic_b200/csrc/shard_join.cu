// shard_join.cu -- HashJoin over row-range shards, one process per GPU, behind the C ABI.
//
// The reference builds ONE index over the whole rhs and streams the lhs through it on one thread
// (cursor/core/hash_join.cc:406-517, 707-831). With both sides sharded by row range the index is split by key
// hash instead of being rebuilt W times:
//   1. every rank hash-partitions its build rows into W parts (ssb_partition_rows: one stable 8-bit sweep) and
//      packs key + payload columns in part order (one gather per column);
//   2. ONE grouped ncclSend/ncclRecv exchange moves part r of every rank to rank r (key and payload columns in
//      the same NCCL group); rank r then owns every build row of hash part r, in global insertion order
//      (chunks arrive in rank order, the partition is stable);
//   3. rank r builds the table of its part (a compact table: 16-byte slots at load 0.6);
//   4. the W tables and the received payload columns are all-gathered (a second grouped exchange), and every
//      rank attaches them as one index (ssb_join_attach_parts): a probe hashes its key, picks the part, and
//      walks that part's table.
// The probe side never moves: the output is in lhs order, and for each lhs row in rhs insertion order, exactly
// as hash_join.cc:793-831 produces it. Build work per rank is 1/W of the whole; what crosses NVLink is the
// build side once (all-to-all) plus the tables once (all-gather) -- the right trade when the build side is the
// small one, which is what the reference's own guidance asks of callers (hash_join.h:35-69: "the right hand
// side is the index").
#include <cuda_runtime.h>

#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <chrono>
#include <vector>

#include "comm.h"
#include "common.h"

using namespace ssb;

namespace {
// SSB200_DEBUG_SHARD_JOIN=1: wall-clock per phase (each mark synchronises the stream), printed by rank 0.
struct PhaseClock {
  ssb_ctx* ctx;
  bool on;
  int rank;
  std::chrono::steady_clock::time_point last;
  PhaseClock(ssb_ctx* c, int r) : ctx(c), on(getenv("SSB200_DEBUG_SHARD_JOIN") != nullptr), rank(r) {
    if (on) { cudaStreamSynchronize(ctx->stream); last = std::chrono::steady_clock::now(); }
  }
  void mark(const char* what) {
    if (!on) return;
    cudaStreamSynchronize(ctx->stream);
    const std::chrono::steady_clock::time_point now = std::chrono::steady_clock::now();
    if (rank == 0) fprintf(stderr, "[ssb200] shard join %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - last).count());
    last = now;
  }
};
}  // namespace

struct ssb_shard_join {
  ssb_comm* comm;
  ssb_ctx* ctx;
  ssb_join* attached;                 // index over the gathered tables (or, dense keys: over the gathered key column)
  void* keys;                         // dense keys: all ranks' key columns, end to end (global rhs row order)
  void* tables;                       // all ranks' slot arrays, end to end
  std::vector<void*> payload;         // all ranks' received payload columns, end to end (global rhs row order)
  std::vector<int32_t> payload_types;
  int64_t total_rows;                 // rhs rows with a non-NULL key, over all ranks
};

extern "C" {

void ssb_shard_join_destroy(ssb_shard_join* j) {
  if (j == nullptr) return;
  if (j->attached) ssb_join_destroy(j->attached);
  cudaStreamSynchronize(j->ctx->stream);
  tmp_free(j->ctx, j->tables);
  tmp_free(j->ctx, j->keys);
  for (size_t i = 0; i < j->payload.size(); ++i) tmp_free(j->ctx, j->payload[i]);
  delete j;
}

int ssb_shard_join_build(ssb_comm* comm, const ssb_column* key, int32_t n_payload, const ssb_column* payload, int64_t rows,
                         ssb_shard_join** out) {
  *out = nullptr;
  ssb_ctx* ctx = comm_ctx(comm);
  const int W = comm_world(comm), rank = comm_rank(comm);
  if (rows < 0 || n_payload < 0) return fail(ctx, SSB_ERROR_INVALID_ARGUMENT_VALUE, "negative row or column count");
  if (W + 1 > 256) return fail(ctx, SSB_ERROR_NOT_IMPLEMENTED, "more than 255 ranks");
  const int kw = width_of(key->dtype);
  if (kw == 0) return fail(ctx, SSB_ERROR_INVALID_ARGUMENT_TYPE, "unsupported join key type");
  for (int i = 0; i < n_payload; ++i) {
    if (width_of(payload[i].dtype) == 0) return fail(ctx, SSB_ERROR_INVALID_ARGUMENT_TYPE, "unsupported payload column type");
    if (payload[i].nulls != nullptr) {
      return fail(ctx, SSB_ERROR_NOT_IMPLEMENTED, "nullable payload columns: send the is_null vector as a BOOL payload column");
    }
  }
  const int n_cols = 1 + n_payload;
  std::vector<int32_t> width(n_cols);
  width[0] = kw;
  for (int i = 0; i < n_payload; ++i) width[1 + i] = width_of(payload[i].dtype);

  // ---- 0. dense integer keys (a surrogate primary key: values spanning at most 4 x the rows of ALL ranks, no NULL
  // bitmap): nothing is partitioned and no table travels. The key and payload columns are all-gathered as they are
  // (rank order = global insertion order) and every rank builds the direct index over the gathered keys
  // (JoinTable::dense_rows: one scatter of 4-byte row numbers, no hashing) -- half the bytes of the partitioned form
  // over NVLink (no tables) and a build that is one pass, so repeating it on every rank costs less than exchanging it.
  PhaseClock clock(ctx, rank);
  {   // collective: every rank takes part in the decision, whatever its own shard looks like
    long long range[3] = {0, 0, 0};
    int eligible = 0;
    if (key->nulls == nullptr) {
      if (int rc0 = join_key_range(ctx, key, rows, range, &eligible)) eligible = 0;   // a failed range pass: the general form
    }
    std::vector<int64_t> all4(4 * W, 0);
    const int64_t mine4[4] = {range[0], range[1], range[2], eligible};
    if (int rc0 = comm_all_gather_counts(comm, mine4, 4, all4.data())) return rc0;
    long long lo = INT64_MAX, hi = INT64_MIN, count = 0;
    bool all_eligible = true;
    std::vector<int64_t> rows_of(W);
    for (int r = 0; r < W; ++r) {
      all_eligible = all_eligible && all4[4 * r + 3] != 0;
      rows_of[r] = all4[4 * r + 2];
      if (all4[4 * r + 2] > 0) { lo = std::min<long long>(lo, all4[4 * r]); hi = std::max<long long>(hi, all4[4 * r + 1]); }
      count += all4[4 * r + 2];
    }
    clock.mark("key range");
    if (all_eligible && count >= 4096 && join_dense_fits(lo, hi, count, count)) {   // the same decision on every rank
      ssb_shard_join* dj = new ssb_shard_join;
      dj->comm = comm; dj->ctx = ctx; dj->attached = nullptr; dj->tables = nullptr; dj->keys = nullptr; dj->total_rows = count;
      int rc0 = 0;
      cudaError_t e0 = tmp_malloc_bytes(ctx, &dj->keys, static_cast<size_t>(count) * kw + 64);
      if (e0 != cudaSuccess) rc0 = cuda_fail(ctx, e0, "sharded join keys");
      dj->payload.assign(n_payload, nullptr);
      dj->payload_types.resize(n_payload);
      std::vector<const void*> src(n_cols);
      std::vector<void*> dst(n_cols);
      src[0] = key->data; dst[0] = dj->keys;
      for (int i = 0; i < n_payload && rc0 == 0; ++i) {
        dj->payload_types[i] = payload[i].dtype;
        e0 = tmp_malloc_bytes(ctx, &dj->payload[i], static_cast<size_t>(count) * width[1 + i] + 64);
        if (e0 != cudaSuccess) rc0 = cuda_fail(ctx, e0, "sharded join payload");
        src[1 + i] = payload[i].data; dst[1 + i] = dj->payload[i];
      }
      if (rc0 == 0) rc0 = comm_all_gather_v(comm, n_cols, src.data(), dst.data(), width.data(), rows_of.data());
      clock.mark("all-gather keys + payload");
      if (rc0 == 0) {
        ssb_column k = *key;
        k.data = dj->keys;
        k.nulls = nullptr;
        rc0 = ssb_join_build(ctx, 1, &k, count, SSB_KEYS_UNIQUE, &dj->attached);
      }
      clock.mark("dense index");
      if (rc0 != 0) { ssb_shard_join_destroy(dj); return rc0; }
      *out = dj;
      return 0;
    }
  }
  // ---- 1. partition: parts 0..W-1 by key hash, part W = rows with a NULL key (they never match and stay behind)
  long long* d_perm = nullptr;
  cudaError_t e = tmp_malloc(ctx, &d_perm, static_cast<size_t>(rows) * 8 + 64);
  if (e != cudaSuccess) return cuda_fail(ctx, e, "sharded join scratch");
  std::vector<int64_t> part_rows(W + 1, 0), send_rows(W, 0), recv_rows(W, 0);
  std::vector<void*> send(n_cols, nullptr), recv(n_cols, nullptr);
  ssb_join* local = nullptr;
  ssb_shard_join* j = nullptr;
  auto cleanup = [&]() {
    cudaStreamSynchronize(ctx->stream);
    tmp_free(ctx, d_perm);
    for (int i = 0; i < n_cols; ++i) tmp_free(ctx, send[i]);
    for (int i = 0; i < n_cols; ++i) tmp_free(ctx, recv[i]);
    if (local) ssb_join_destroy(local);
  };
  int rc = ssb_partition_rows(ctx, 1, key, rows, W, W, reinterpret_cast<int64_t*>(d_perm), part_rows.data());
  if (rc) { cleanup(); return rc; }
  clock.mark("partition");
  int64_t sent = 0;
  for (int r = 0; r < W; ++r) { send_rows[r] = part_rows[r]; sent += part_rows[r]; }
  // ---- 2. pack and exchange
  for (int i = 0; i < n_cols && rc == 0; ++i) {
    e = tmp_malloc_bytes(ctx, &send[i], static_cast<size_t>(sent) * width[i] + 64);
    if (e != cudaSuccess) { rc = cuda_fail(ctx, e, "sharded join send buffer"); break; }
    ssb_column src = i == 0 ? *key : payload[i - 1];
    src.nulls = nullptr;   // the rows that travel have a key; payload columns are NOT NULL
    ssb_column dst = src;
    dst.data = send[i];
    if (sent > 0) rc = ssb_gather(ctx, &src, reinterpret_cast<const int64_t*>(d_perm), sent, &dst);
  }
  clock.mark("pack");
  if (rc == 0) rc = comm_exchange_counts(comm, send_rows.data(), recv_rows.data());
  int64_t received = 0;
  for (int r = 0; r < W; ++r) received += recv_rows[r];
  for (int i = 0; i < n_cols && rc == 0; ++i) {
    e = tmp_malloc_bytes(ctx, &recv[i], static_cast<size_t>(received) * width[i] + 64);
    if (e != cudaSuccess) rc = cuda_fail(ctx, e, "sharded join receive buffer");
  }
  if (rc == 0) rc = comm_all_to_all_v(comm, n_cols, send.data(), recv.data(), width.data(), send_rows.data(), recv_rows.data());
  clock.mark("exchange (all-to-all)");
  // ---- 3. the table of this rank's part
  const void* d_slots = nullptr;
  int64_t capacity = 0;
  if (rc == 0) {
    ssb_column k = *key;
    k.data = recv[0];
    k.nulls = nullptr;
    rc = ssb_join_build(ctx, 1, &k, received, SSB_KEYS_UNIQUE | SSB_KEYS_COMPACT_TABLE, &local);
  }
  if (rc == 0) rc = ssb_join_table(local, &d_slots, &capacity);
  clock.mark("build");
  // ---- 4. all-gather the tables and the payload columns
  std::vector<int64_t> all(2 * W, 0);
  if (rc == 0) {
    const int64_t mine[2] = {received, capacity};
    rc = comm_all_gather_counts(comm, mine, 2, all.data());
  }
  if (rc != 0) { cleanup(); return rc; }
  std::vector<int64_t> rows_of(W), cap_of(W), row_offset(W), cap_offset(W);
  int64_t total_rows = 0, total_cap = 0;
  for (int r = 0; r < W; ++r) {
    rows_of[r] = all[2 * r];
    cap_of[r] = all[2 * r + 1];
    row_offset[r] = total_rows;
    cap_offset[r] = total_cap;
    total_rows += rows_of[r];
    total_cap += cap_of[r];
  }
  j = new ssb_shard_join;
  j->comm = comm;
  j->ctx = ctx;
  j->attached = nullptr;
  j->tables = nullptr;
  j->keys = nullptr;
  j->total_rows = total_rows;
  e = tmp_malloc_bytes(ctx, &j->tables, static_cast<size_t>(total_cap) * 16 + 64);
  if (e != cudaSuccess) rc = cuda_fail(ctx, e, "sharded join tables");
  if (rc == 0) {
    const void* s1[1] = {d_slots};
    void* r1[1] = {j->tables};
    const int32_t w1[1] = {16};
    rc = comm_all_gather_v(comm, 1, s1, r1, w1, cap_of.data());
  }
  if (rc == 0 && n_payload > 0) {
    j->payload.assign(n_payload, nullptr);
    j->payload_types.resize(n_payload);
    for (int i = 0; i < n_payload && rc == 0; ++i) {
      j->payload_types[i] = payload[i].dtype;
      e = tmp_malloc_bytes(ctx, &j->payload[i], static_cast<size_t>(total_rows) * width[1 + i] + 64);
      if (e != cudaSuccess) rc = cuda_fail(ctx, e, "sharded join payload");
    }
    if (rc == 0) rc = comm_all_gather_v(comm, n_payload, recv.data() + 1, j->payload.data(), width.data() + 1, rows_of.data());
  }
  clock.mark("all-gather tables + payload");
  if (rc == 0) {
    std::vector<const void*> slot_ptrs(W);
    for (int r = 0; r < W; ++r) slot_ptrs[r] = static_cast<const char*>(j->tables) + static_cast<size_t>(cap_offset[r]) * 16;
    rc = ssb_join_attach_parts(ctx, key->dtype, W, slot_ptrs.data(), cap_of.data(), row_offset.data(), &j->attached);
  }
  // the exchange reads the send / receive buffers and the local table: wait for it, then let them go
  cleanup();
  clock.mark("attach + cleanup");
  if (rc != 0) { ssb_shard_join_destroy(j); return rc; }
  *out = j;
  return 0;
}

int ssb_shard_join_probe(ssb_shard_join* j, const ssb_column* key, int64_t rows, int32_t join_type, int64_t* n_pairs,
                         const int64_t** d_lhs_rows, const int64_t** d_rhs_rows) {
  PhaseClock clock(j->ctx, comm_rank(j->comm));
  const int rc = ssb_join_probe(j->attached, key, rows, join_type, n_pairs, d_lhs_rows, d_rhs_rows);
  clock.mark("probe");
  return rc;
}

int ssb_shard_join_probe_materialize(ssb_shard_join* j, const ssb_column* keys, int64_t rows, int32_t join_type, int32_t n_lhs,
                                     const ssb_column* lhs_cols, int32_t n_rhs, const int32_t* rhs_payload, const ssb_column* out_cols,
                                     uint8_t* d_matched, int64_t* n_rows) {
  std::vector<ssb_column> rhs(n_rhs > 0 ? n_rhs : 1);
  for (int i = 0; i < n_rhs; ++i) {
    if (int rc = ssb_shard_join_payload(j, rhs_payload[i], &rhs[i], nullptr)) return rc;
  }
  PhaseClock clock(j->ctx, comm_rank(j->comm));
  const int rc = ssb_join_probe_materialize(j->attached, keys, rows, join_type, n_lhs, lhs_cols, n_rhs, rhs.data(), out_cols, d_matched, n_rows);
  clock.mark("probe + materialise");
  return rc;
}

int ssb_shard_join_form(const ssb_shard_join* j) { return j->keys != nullptr ? 1 : 0; }

int ssb_shard_join_payload(const ssb_shard_join* j, int32_t i, ssb_column* out, int64_t* rows) {
  if (i < 0 || i >= static_cast<int32_t>(j->payload.size())) return fail(j->ctx, SSB_ERROR_INVALID_ARGUMENT_VALUE, "no such payload column");
  out->data = j->payload[i];
  out->nulls = nullptr;
  out->dtype = j->payload_types[i];
  out->reserved = 0;
  if (rows) *rows = j->total_rows;
  return 0;
}

}  // extern "C"
