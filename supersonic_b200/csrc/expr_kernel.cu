// expr_kernel.cu -- the fused Compute / Project / Filter kernel for sm_100a.
//
// Replaces, in one launch over a whole shard, the reference's per-1024-row-block chain
//   ViewCursor::Next -> BoundExpressionTree::Evaluate -> VectorBinaryPrimitive loops
//   -> FilterCursor::PrepareInputRowIds -> SelectiveViewCopier gather
// (cursor/core/compute.cc:49-56, cursor/core/filter.cc:96-230,
//  expression/vector/vector_primitives.h:70-353, base/infrastructure/copy_column.cc:200-286).
//
// Shape of the kernel (HBM-bound; no tensor cores because nothing here is a contraction):
//  * persistent CTAs; tile t (NT*R rows) belongs to CTA t % gridDim.x, so the CTAs walk the
//    table in lock-step "waves" of gridDim.x tiles
//  * input column tiles are staged into shared memory by TMA bulk copies (cp.async.bulk, one
//    elected thread, mbarrier complete_tx), `stages` tiles in flight per CTA: HBM latency is
//    hidden by bytes in flight, not by occupancy
//  * the expression program (bytecode in the constant bank) is interpreted warp-uniformly.
//    A thread owns R rows whose values live in registers (the accumulator); shared memory
//    slots hold operands only. Hot (op, type) pairs on NOT NULL operands are pre-decoded
//    superinstructions (operand address = one add); the rest runs through ops.h's alu()
//  * Filter: the predicate runs first; warp ballots + a scan over the warps give every kept
//    row its position inside the tile, outputs are stored already compacted into a staging
//    buffer, and the tile's kept count is published. One iteration later the CTA reads the
//    counts of its whole wave with one coalesced load (wave-synchronous prefix: no serial
//    look-back chain), and copies the staged rows out with fully coalesced stores. Output
//    order = input order, one pass over the data.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <algorithm>
#include <stdlib.h>

#include "common.h"
#include "group_device.h"
#include "program.h"

namespace ssb {

// ------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// TMA 1-D bulk copy global -> shared, completion counted on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// streaming (evict-first) global stores: every output byte is written once
__device__ __forceinline__ void st_cs_u64(void* p, uint64_t v) {
  asm volatile("st.global.cs.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_cs_u32(void* p, uint32_t v) {
  asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

static constexpr unsigned long long kValid = 1ull << 63;

// Consumer-only barrier (named barrier 1): the producer warp does not take part.
template <int NT>
__device__ __forceinline__ void bar_consumers() {
  asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
}

// NT consumer threads evaluate the program; one extra producer warp issues the TMA fills, so
// that no consumer warp carries the descriptor loop on its critical path.
//
// Instruction budget. The kernel is bound by instruction issue, not by bytes (ncu: 12 resident
// warps, "wait" stalls), so everything that runs once per tile is kept off the constant bank
// and off 64-bit arithmetic: the program is expanded once per CTA into a shared-memory table of
// pre-resolved entries (one LDS.128 per instruction, the next one prefetched while the current
// one runs), tile / stage / buffer indices are running 32-bit counters, and the wave scan uses
// loop-invariant lane masks and REDUX.
struct __align__(16) TabEntry {
  uint32_t code_flags;   // Code | flags << 16
  uint32_t x;            // address of slot a | low half of the immediate | output offset
  uint32_t y;            // address of slot b (left operand of fused forms)
  uint32_t z;            // high half of the immediate | index a
};
static constexpr uint32_t kCodeEnd = 0xffffu;

// ---- aggregation sink: CTA-local state ----------------------------------------------------------
// The CTA learns its (few) groups on the fly: the first row of a group goes through the global
// table (find_slot_*), claims the next local entry and publishes its key values; later rows find
// the entry by comparing their keys with the published ones (held in registers for the tile).
// Every thread owns a private accumulator per (local group, aggregate) -- [group][aggregate][thread]:
// bank = thread, no conflicts, no atomics in the row loop.
struct SinkShared {
  unsigned long long* t_acc;
  unsigned long long* l_key;
  unsigned int* l_knull;
  unsigned int* l_slot;
  unsigned int* l_ready;
  unsigned int* l_fp;
  unsigned char* t_seen;
};
__device__ __forceinline__ uint32_t sink_fp32(unsigned long long k0, unsigned long long k1, unsigned int knull) {
  const unsigned long long h = (k0 + knull) * 0x9E3779B97F4A7C15ull + k1 * 0xC2B2AE3D27D4EB4Full;
  return static_cast<uint32_t>(h >> 32) ^ static_cast<uint32_t>(h);
}
// Layout of the sink area (must match sink_bytes_per_thread / sink_fixed_bytes on the host).
__device__ __forceinline__ SinkShared sink_layout(unsigned char* base, int SG, int SA, int NT) {
  SinkShared sk;
  sk.t_acc = reinterpret_cast<unsigned long long*>(base);                     // [(SG + 1) * SA][NT], block SG = trash
  sk.l_key = sk.t_acc + (SG + 1) * SA * NT;                                   // [kTinyGroups][2] key values of the local entries
  sk.l_knull = reinterpret_cast<unsigned int*>(sk.l_key + kTinyGroups * 2);   // [kTinyGroups] bit c: key column c is NULL
  sk.l_slot = sk.l_knull + kTinyGroups;                                       // [kTinyGroups] global slot + 1, 0 = free
  sk.l_ready = sk.l_slot + kTinyGroups;                                       // [1] bit e: entry e is published
  sk.l_fp = sk.l_ready + 4;                                                   // [kTinyGroups] 32-bit fingerprint of the entry's key
  sk.t_seen = reinterpret_cast<unsigned char*>(sk.l_fp + kTinyGroups);        // [(SG + 1) * SA][NT]
  return sk;
}
// Identity of a thread's partial. Sums start from the value that addition leaves exact (-0.0 for
// floating point: -0.0 + x == x for every x, +0.0 would turn a sum of -0.0 into +0.0).
__device__ __forceinline__ unsigned long long sink_identity(const AggDev& ag) {
  if (ag.fn == SSB_AGG_SUM && ag.out_phys == T_F64) return 0x8000000000000000ull;
  if (ag.fn == SSB_AGG_SUM && ag.out_phys == T_F32) return 0x80000000ull;
  return identity_dev(ag);
}
__device__ __forceinline__ long long sink_global_slot(const GroupParams& gp, unsigned long long k0, unsigned long long k1, unsigned int knull) {
  if (gp.packed) return find_slot_packed_kv(gp, (knull & 1u) != 0, k0);
  unsigned long long kv[kMaxKeys];
  kv[0] = k0; kv[1] = k1;
  return find_slot_generic_kv(gp, kv, knull);
}
// Cold path of the group lookup (first sight of a key in this CTA). Returns the local entry, -1
// when the CTA has no free entry (the row goes to the global table), -2 when the global table is
// full (the row is deferred).
__device__ __noinline__ int sink_insert(const GroupParams& gp, unsigned char* sink_base, int SA, int NT, int n_local,
                                        unsigned long long k0, unsigned long long k1, unsigned int knull) {
  const SinkShared sk = sink_layout(sink_base, n_local, SA, NT);
  const long long slot = sink_global_slot(gp, k0, k1, knull);
  if (slot < 0) return -2;
  const unsigned int want = static_cast<unsigned int>(slot) + 1u;
  for (int e = 0; e < n_local && e < kTinyGroups; ++e) {
    const unsigned int old = atomicCAS(&sk.l_slot[e], 0u, want);
    if (old == 0u) {
      sk.l_key[2 * e] = k0;
      sk.l_key[2 * e + 1] = k1;
      sk.l_knull[e] = knull;
      sk.l_fp[e] = sink_fp32(k0, k1, knull);
      __threadfence_block();
      atomicOr(sk.l_ready, 1u << e);
      return e;
    }
    if (old == want) return e;
  }
  return -1;
}

// Cold path: a row whose group has no local entry in this CTA goes to the global table.
__device__ __noinline__ void sink_apply_global(const GroupParams& gp, int a, unsigned long long k0, unsigned long long k1,
                                               unsigned int knull, unsigned long long v, bool count_star) {
  const long long slot = sink_global_slot(gp, k0, k1, knull);
  const AggDev& ag = gp.agg[a];
  if (count_star) { atomicAdd(&ag.acc[static_cast<unsigned long long>(slot) * ag.stride], 1ull); return; }
  apply(ag, slot, v, 1ull);
  if (ag.seen != nullptr) ag.seen[slot] = 1u;
}

// SINK: the outputs (group-by keys, then aggregate inputs) feed the aggregation table of group.cu
// directly -- per-thread accumulators in shared memory for the CTA's first `sink_groups` groups,
// the global table beyond -- instead of being staged, compacted and copied out. Filter then
// needs no compaction at all (the predicate only masks rows), so the wave scan, the staging
// buffers and the copy-out disappear from the kernel.
template <int NT, int R, bool SINK>
__device__ __forceinline__ void expr_kernel_body(const ExprParams& p) {
  constexpr int TILE = NT * R;
  constexpr int NW = NT / 32;
  constexpr int WROWS = 32 * R;          // rows owned by one warp: [warp * WROWS, (warp + 1) * WROWS)
  extern __shared__ __align__(1024) unsigned char smem[];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool producer = warp == NW;
  const int row_first = warp * WROWS + lane;   // thread's row k is row_first + 32 * k
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.off_bar);
  uint32_t* warp_cnt = reinterpret_cast<uint32_t*>(smem + p.off_scan);   // [2][16] kept rows per warp
  uint2* red = reinterpret_cast<uint2*>(warp_cnt + 32);                  // [2][16] {before, sum} per warp
  uint32_t* s_meta = reinterpret_cast<uint32_t*>(red + 32);              // [0..3] kept rows of the staged tiles
  uint32_t* nullw_base = reinterpret_cast<uint32_t*>(smem + p.off_nullw);
  TabEntry* itab = reinterpret_cast<TabEntry*>(smem + p.off_itab);       // [stages][n_insn + 1]

  const int G = static_cast<int>(gridDim.x);
  const int bid = static_cast<int>(blockIdx.x);
  const int num_tiles = static_cast<int>(p.num_tiles);
  const int n_my = (num_tiles - bid + G - 1) / G;
  const int S = p.stages;
  const int n_tab = p.n_insn + 1;

  if (tid == 0 && p.use_tma) {
    for (int s = 0; s < S; ++s) mbar_init(&bars[s], 1);
    fence_barrier_init();
  }
  // Expand the program: one entry per (stage, instruction) with absolute shared-memory offsets.
  for (int e = tid; e < S * n_tab; e += NT + 32) {
    const int s = e / n_tab, pc = e - s * n_tab;
    TabEntry t;
    if (pc == p.n_insn) {
      t.code_flags = kCodeEnd; t.x = t.y = t.z = 0;
    } else {
      const Insn& in = p.insn[pc];
      const uint32_t so = static_cast<uint32_t>(s) * p.stage_bytes;
      t.code_flags = static_cast<uint32_t>(in.code) | (static_cast<uint32_t>(in.flags) << 16);
      t.x = (in.off_a & 0x7fffffffu) + ((in.off_a >> 31) ? so : 0u);
      t.y = (in.off_b & 0x7fffffffu) + ((in.off_b >> 31) ? so : 0u);
      t.z = static_cast<uint32_t>(static_cast<int32_t>(in.a));
      const uint32_t c = in.code;
      if (c >= C_MAD_I64 && c < C_MAD_END) {
        const uint32_t oc = p.insn_c[pc];
        t.z = (oc & 0x7fffffffu) + ((oc >> 31) ? so : 0u);
      }
      const bool bin_imm = c >= C_BIN_BASE && c < C_BIN_END && ((c - C_BIN_BASE) & 1u);
      if (c == C_LOADK || bin_imm) {
        const u64 v = p.imm[in.a];
        t.x = static_cast<uint32_t>(v);
        t.z = static_cast<uint32_t>(v >> 32);
      } else if (c == C_OUT8 || c == C_OUT4) {
        t.x = p.out_off[in.a];
      }
    }
    itab[e] = t;
  }
  // ---- aggregation sink state (thread-private words are indexed [..][thread]); layout = sink_layout() below
  const GroupParams* gp = static_cast<const GroupParams*>(p.sink_gp);
  const int SG = p.sink_groups, SA = p.sink_n_aggs, SNK = p.sink_n_keys;
  const SinkShared sk = sink_layout(smem + p.sink_off, SG, SA, NT);
  const uint32_t sink_block = static_cast<uint32_t>(SA) * NT * 8u;                            // bytes of one group's accumulators
  const uint32_t sink_trash = static_cast<uint32_t>(SG) * sink_block;
  if constexpr (SINK) {
    if (tid < NT) {
      for (int i = tid; i < (SG + 1) * SA * NT; i += NT) {
        sk.t_acc[i] = sink_identity(gp->agg[(i / NT) % SA]);
        sk.t_seen[i] = 0;
      }
      if (tid < kTinyGroups) sk.l_slot[tid] = 0u;
      if (tid == 0) *sk.l_ready = 0u;
    }
  }
  __syncthreads();

  auto slot_data = [&](int slot, int stage) -> unsigned char* {
    return slot < p.n_in ? smem + p.off_data + stage * p.stage_bytes + p.in_off[slot]
                         : smem + p.off_tmp + (slot - p.n_in) * (TILE * 8);
  };
  auto slot_nullw = [&](int slot, int stage) -> uint32_t* {
    if (slot < p.n_in) {
      const int row = p.in_nullw[slot];
      return row < 0 ? nullptr : nullw_base + (stage * p.stage_nullw + row) * (TILE / 32);
    }
    return nullw_base + (p.stages * p.stage_nullw + (slot - p.n_in)) * (TILE / 32);
  };

  // Issues the TMA fill of `stage` with tile `tile` (full tiles only).
  auto issue = [&](int tile, int stage) {
    const long long row0 = static_cast<long long>(tile) * TILE;
    if (!p.use_tma || p.rows - row0 < TILE) return;
    uint32_t bytes = p.stage_tx_bytes;
    for (int i = 0; i < p.n_in; ++i) {
      if (p.in_nullw[i] >= 0 && p.in_nulls[i] != nullptr) bytes += TILE / 8;
    }
    fence_proxy_async();
    mbar_expect_tx(&bars[stage], bytes);
    for (int i = 0; i < p.n_in; ++i) {
      const int w = p.in_width[i];
      tma_load_1d(smem + p.off_data + stage * p.stage_bytes + p.in_off[i],
                  static_cast<const unsigned char*>(p.in_data[i]) + row0 * w, TILE * w, &bars[stage]);
      if (p.in_nullw[i] >= 0 && p.in_nulls[i] != nullptr) {
        tma_load_1d(slot_nullw(i, stage), p.in_nulls[i] + row0 / 32, TILE / 8, &bars[stage]);
      }
    }
  };

  const int kdefer_ = p.defer;
  const int n_obuf = kdefer_ + 1;

  if (producer) {
    // ---- producer warp: keeps `S` tiles in flight; refills a stage as soon as the consumers
    // release it (the __syncthreads that ends every iteration).
    if (lane == 0) {
      for (int s = 0; s < S && s < n_my; ++s) issue(bid + s * G, s);
    }
    int stage = 0;
    for (int it = 0; it < n_my + kdefer_; ++it) {
      __syncthreads();
      if (lane == 0 && it + S < n_my) issue(bid + (it + S) * G, stage);
      if (++stage == S) stage = 0;
    }
    return;
  }

  const uint32_t all = (R >= 32) ? 0xffffffffu : ((1u << R) - 1u);
  const uint32_t lt = (1u << lane) - 1u;
  uint32_t fail = 0;

  // Wave-synchronous prefix, shared by all consumer threads: at the top of an iteration every
  // thread fetches a few of the kept-row counts the wave of kdefer_ tiles ago published (together
  // one coalesced read of gridDim.x words); the L2 round trip overlaps with the evaluation of the
  // current tile; the partial sums are combined across the warps through the barrier that ends
  // the evaluation anyway. Which of a thread's words exist (full wave / last wave) and which
  // precede this CTA's tile never changes: three lane masks computed once.
  constexpr int kPrefetch = (768 + NT - 1) / NT;   // status words per thread: covers gridDim.x <= 768
  const int n_waves = (num_tiles + G - 1) / G;
  const int last_count = num_tiles - (n_waves - 1) * G;   // tiles in the last wave
  uint32_t full_mask = 0, last_mask = 0, before_mask = 0;
#pragma unroll
  for (int q = 0; q < kPrefetch; ++q) {
    const int j = tid + q * NT;
    full_mask |= (j < G ? 1u : 0u) << q;
    last_mask |= (j < last_count ? 1u : 0u) << q;
    before_mask |= (j < bid ? 1u : 0u) << q;
  }
  const unsigned long long* wave_ptr = p.tile_status + tid;   // status words of the wave scanned next
  long long wave_base = 0;   // kept rows of all earlier waves (every thread keeps a copy)

  int stage = 0;             // it % S
  uint32_t parity = 0;       // (it / S) & 1
  int ob = 0;                // it % n_obuf: staging buffer written by this iteration
  for (int it = 0; it < n_my + kdefer_; ++it) {
    unsigned long long pre[kPrefetch];
    const bool scan_wave = !SINK && p.has_pred && it >= kdefer_;
    const unsigned long long* wp = wave_ptr;
    if (scan_wave) {
      const uint32_t m = (it - kdefer_ == n_waves - 1) ? last_mask : full_mask;
#pragma unroll
      for (int q = 0; q < kPrefetch; ++q) {
        pre[q] = kValid;
        if (q * NT < G && ((m >> q) & 1u)) pre[q] = ld_relaxed(wp + q * NT);   // first test is CTA-uniform
      }
      wave_ptr += G;
    }
    auto finish_wave = [&]() {
      unsigned long long andv = kValid;
#pragma unroll
      for (int q = 0; q < kPrefetch; ++q) andv &= pre[q];
      if (!(andv & kValid) && !p.debug_nowait) {   // a CTA of the wave is behind: wait for it
#pragma unroll
        for (int q = 0; q < kPrefetch; ++q) {
          while (!(pre[q] & kValid)) pre[q] = ld_relaxed(wp + q * NT);
        }
      }
      uint32_t before = 0, sum = 0;
#pragma unroll
      for (int q = 0; q < kPrefetch; ++q) {
        const uint32_t c = static_cast<uint32_t>(pre[q]);
        sum += c;
        before += ((before_mask >> q) & 1u) ? c : 0u;
      }
      before = __reduce_add_sync(0xffffffffu, before);
      sum = __reduce_add_sync(0xffffffffu, sum);
      if (lane == 0) red[(it & 1) * 16 + warp] = make_uint2(before, sum);
    };
    // ======================================================== evaluate tile `it`
    if (it < n_my) {
      const int tile = bid + it * G;
      const long long row0 = static_cast<long long>(tile) * TILE;
      const int n = static_cast<int>(p.rows - row0 < TILE ? p.rows - row0 : TILE);
      const bool via_tma = p.use_tma && n == TILE;
      unsigned char* obuf = smem + p.off_out + ob * p.out_bytes;

      if (via_tma) {
        mbar_wait(&bars[stage], parity);
      } else {
        // plain-load path: last (partial) tile, or columns not 16-byte aligned
        for (int i = 0; i < p.n_in; ++i) {
          const int w = p.in_width[i];
          unsigned char* dst = smem + p.off_data + stage * p.stage_bytes + p.in_off[i];
          const unsigned char* src = static_cast<const unsigned char*>(p.in_data[i]) + row0 * w;
          if (w == 8) {
            for (int r = tid; r < n; r += NT) reinterpret_cast<u64*>(dst)[r] = reinterpret_cast<const u64*>(src)[r];
          } else if (w == 4) {
            for (int r = tid; r < n; r += NT) reinterpret_cast<uint32_t*>(dst)[r] = reinterpret_cast<const uint32_t*>(src)[r];
          } else {
            for (int r = tid; r < n; r += NT) dst[r] = src[r];
          }
        }
        bar_consumers<NT>();
      }
      // Null words of inputs declared nullable: TMA delivered them when the column has a
      // bitmap; otherwise (no bitmap in this run, or the plain-load path) fill them here.
      if (via_tma ? p.fill_nullw_tma : p.fill_nullw_plain) {
        for (int i = 0; i < p.n_in; ++i) {
          if (p.in_nullw[i] < 0) continue;
          const bool have = p.in_nulls[i] != nullptr;
          if (via_tma && have) continue;
          uint32_t* wdst = slot_nullw(i, stage);
          for (int wi = tid; wi < TILE / 32; wi += NT) {
            uint32_t wv = 0;
            if (have && wi * 32 < n) wv = p.in_nulls[i][row0 / 32 + wi];
            wdst[wi] = wv;
          }
        }
        bar_consumers<NT>();
      }
      // Without a predicate no barrier separates this tile's staging writes from the copy-out
      // of the tile evaluated two iterations ago (same buffer): add one.
      if (!SINK && !p.has_pred) bar_consumers<NT>();

      uint32_t live = all;
      if (n != TILE) {
        live = 0;
#pragma unroll
        for (int k = 0; k < R; ++k) live |= (row_first + 32 * k < n ? 1u : 0u) << k;
      }

      // ---- the accumulator machine
      u64 acc[R];
      uint32_t accn = 0, pass = live;
      int pos[R];
#pragma unroll
      for (int k = 0; k < R; ++k) { acc[k] = 0; pos[k] = row_first + 32 * k; }

      // ---- aggregation sink: runs once per tile, after the program (below)

      // +0 acc (op) slot, +1 acc (op) imm, +2 slot (op) slot, +3 slot (op) imm; EXPR sees x, y
#define SSB_BIN4(CODE0, T, EXPR)                                                         \
  case (CODE0): {                                                                        \
    const T* ps = reinterpret_cast<const T*>(pa);                                        \
    _Pragma("unroll") for (int k = 0; k < R; ++k) {                                      \
      const T x = Codec<T>::dec(acc[k]);                                                 \
      const T y = ps[row_first + 32 * k];                                                \
      acc[k] = EXPR;                                                                     \
    }                                                                                    \
  } break;                                                                               \
  case (CODE0) + 1: {                                                                    \
    const T y = Codec<T>::dec(immv);                                                     \
    _Pragma("unroll") for (int k = 0; k < R; ++k) {                                      \
      const T x = Codec<T>::dec(acc[k]);                                                 \
      acc[k] = EXPR;                                                                     \
    }                                                                                    \
  } break;                                                                               \
  case (CODE0) + 2: {                                                                    \
    const T* ps = reinterpret_cast<const T*>(pa);                                        \
    const T* pl = reinterpret_cast<const T*>(pb);                                        \
    _Pragma("unroll") for (int k = 0; k < R; ++k) {                                      \
      const T x = pl[row_first + 32 * k];                                                \
      const T y = ps[row_first + 32 * k];                                                \
      acc[k] = EXPR;                                                                     \
    }                                                                                    \
    accn = 0;                                                                            \
  } break;                                                                               \
  case (CODE0) + 3: {                                                                    \
    const T* pl = reinterpret_cast<const T*>(pb);                                        \
    const T y = Codec<T>::dec(immv);                                                     \
    _Pragma("unroll") for (int k = 0; k < R; ++k) {                                      \
      const T x = pl[row_first + 32 * k];                                                \
      acc[k] = EXPR;                                                                     \
    }                                                                                    \
    accn = 0;                                                                            \
  } break;
      // comparisons collect one bit per row; materialised as 0/1, or fed to the compaction
#define SSB_CMP4(CODE0, T, COND)                                                         \
  case (CODE0): {                                                                        \
    const T* ps = reinterpret_cast<const T*>(pa);                                        \
    _Pragma("unroll") for (int k = 0; k < R; ++k) {                                      \
      const T x = Codec<T>::dec(acc[k]);                                                 \
      const T y = ps[row_first + 32 * k];                                                \
      bits |= ((COND) ? 1u : 0u) << k;                                                   \
    }                                                                                    \
    post = 1;                                                                            \
  } break;                                                                               \
  case (CODE0) + 1: {                                                                    \
    const T y = Codec<T>::dec(immv);                                                     \
    _Pragma("unroll") for (int k = 0; k < R; ++k) {                                      \
      const T x = Codec<T>::dec(acc[k]);                                                 \
      bits |= ((COND) ? 1u : 0u) << k;                                                   \
    }                                                                                    \
    post = 1;                                                                            \
  } break;                                                                               \
  case (CODE0) + 2: {                                                                    \
    const T* ps = reinterpret_cast<const T*>(pa);                                        \
    const T* pl = reinterpret_cast<const T*>(pb);                                        \
    _Pragma("unroll") for (int k = 0; k < R; ++k) {                                      \
      const T x = pl[row_first + 32 * k];                                                \
      const T y = ps[row_first + 32 * k];                                                \
      bits |= ((COND) ? 1u : 0u) << k;                                                   \
    }                                                                                    \
    accn = 0;                                                                            \
    post = 1;                                                                            \
  } break;                                                                               \
  case (CODE0) + 3: {                                                                    \
    const T* pl = reinterpret_cast<const T*>(pb);                                        \
    const T y = Codec<T>::dec(immv);                                                     \
    _Pragma("unroll") for (int k = 0; k < R; ++k) {                                      \
      const T x = pl[row_first + 32 * k];                                                \
      bits |= ((COND) ? 1u : 0u) << k;                                                   \
    }                                                                                    \
    accn = 0;                                                                            \
    post = 1;                                                                            \
  } break;
      // multiply-add: +0 acc * slot a + slot b, +1 slot c * slot a + slot b
#define SSB_MAD2(CODE0, T, EXPR)                                                         \
  case (CODE0): {                                                                        \
    const T* ps = reinterpret_cast<const T*>(pa);                                        \
    const T* pz = reinterpret_cast<const T*>(pb);                                        \
    _Pragma("unroll") for (int k = 0; k < R; ++k) {                                      \
      const T x = Codec<T>::dec(acc[k]);                                                 \
      const T y = ps[row_first + 32 * k];                                                \
      const T z = pz[row_first + 32 * k];                                                \
      acc[k] = Codec<T>::enc(EXPR);                                                      \
    }                                                                                    \
  } break;                                                                               \
  case (CODE0) + 1: {                                                                    \
    const T* ps = reinterpret_cast<const T*>(pa);                                        \
    const T* pz = reinterpret_cast<const T*>(pb);                                        \
    const T* pl = reinterpret_cast<const T*>(smem + cur.z);                              \
    _Pragma("unroll") for (int k = 0; k < R; ++k) {                                      \
      const T x = pl[row_first + 32 * k];                                                \
      const T y = ps[row_first + 32 * k];                                                \
      const T z = pz[row_first + 32 * k];                                                \
      acc[k] = Codec<T>::enc(EXPR);                                                      \
    }                                                                                    \
    accn = 0;                                                                            \
  } break;
#define SSB_ENC(T, V) Codec<T>::enc(V)
#define SSB_BIN_TYPE(BASE, T, ADD, SUB, SUBR, MUL)                                       \
  SSB_BIN4((BASE) + 4 * B_ADD, T, SSB_ENC(T, ADD))                                       \
  SSB_BIN4((BASE) + 4 * B_SUB, T, SSB_ENC(T, SUB))                                       \
  SSB_BIN4((BASE) + 4 * B_SUBR, T, SSB_ENC(T, SUBR))                                     \
  SSB_BIN4((BASE) + 4 * B_MUL, T, SSB_ENC(T, MUL))                                       \
  SSB_CMP4((BASE) + 4 * B_LT, T, x < y)                                                  \
  SSB_CMP4((BASE) + 4 * B_GT, T, y < x)                                                  \
  SSB_CMP4((BASE) + 4 * B_EQ, T, x == y)

      const TabEntry* tab = itab + stage * n_tab;
      TabEntry next = tab[0];
      for (int pc = 0;; ++pc) {
        const TabEntry cur = next;
        const uint32_t code = cur.code_flags & 0xffffu;
        if (code == kCodeEnd) break;
        next = tab[pc + 1];   // in flight while this instruction runs
        if (code != C_GENERIC) {
          // pre-decoded fast path: operands cannot be NULL, addresses are resolved
          const unsigned char* pa = smem + cur.x;
          const unsigned char* pb = smem + cur.y;
          const u64 immv = static_cast<u64>(cur.x) | (static_cast<u64>(cur.z) << 32);
          uint32_t bits = 0;
          int post = 0;   // 1: comparison mask in `bits`; 2: K_PRED on the accumulator
          switch (code) {
            case C_LOAD8: {
              const u64* ps = reinterpret_cast<const u64*>(pa);
#pragma unroll
              for (int k = 0; k < R; ++k) acc[k] = ps[row_first + 32 * k];
              accn = 0;
            } break;
            case C_LOAD4: {
              const uint32_t* ps = reinterpret_cast<const uint32_t*>(pa);
#pragma unroll
              for (int k = 0; k < R; ++k) acc[k] = ps[row_first + 32 * k];
              accn = 0;
            } break;
            case C_LOADK: {
#pragma unroll
              for (int k = 0; k < R; ++k) acc[k] = immv;
              accn = 0;
            } break;
            SSB_BIN_TYPE(C_BIN_I64, int64_t, Arith<int64_t>::add(x, y), Arith<int64_t>::sub(x, y),
                         Arith<int64_t>::sub(y, x), Arith<int64_t>::mul(x, y))
            SSB_BIN_TYPE(C_BIN_F64, double, x + y, x - y, y - x, x * y)
            SSB_BIN_TYPE(C_BIN_I32, int32_t, Arith<int32_t>::add(x, y), Arith<int32_t>::sub(x, y),
                         Arith<int32_t>::sub(y, x), Arith<int32_t>::mul(x, y))
            SSB_MAD2(C_MAD_I64, int64_t, Arith<int64_t>::add(Arith<int64_t>::mul(x, y), z))
            SSB_MAD2(C_MAD_F64, double, __dadd_rn(__dmul_rn(x, y), z))   // two roundings, as the separate ops
            SSB_MAD2(C_MAD_I32, int32_t, Arith<int32_t>::add(Arith<int32_t>::mul(x, y), z))
            case C_AND3_S:
            case C_OR3_S: {
              // rhs is never NULL here; acc may be
              uint32_t av = 0, bv = 0;
#pragma unroll
              for (int k = 0; k < R; ++k) {
                av |= static_cast<uint32_t>(acc[k] & 1u) << k;
                bv |= static_cast<uint32_t>(pa[row_first + 32 * k] != 0) << k;
              }
              uint32_t val, nul;
              if (code == C_OR3_S) { val = (av & ~accn) | bv; nul = accn & ~bv; }
              else { val = av & ~accn & bv; nul = accn & bv; }
#pragma unroll
              for (int k = 0; k < R; ++k) acc[k] = (val >> k) & 1u;
              accn = nul & all;
            } break;
            case C_PRED: {
#pragma unroll
              for (int k = 0; k < R; ++k) bits |= static_cast<uint32_t>(acc[k] & 1u) << k;
              post = 2;
            } break;
            case C_STORE8: {
              u64* dst = reinterpret_cast<u64*>(smem + cur.x);
#pragma unroll
              for (int k = 0; k < R; ++k) dst[row_first + 32 * k] = acc[k];
            } break;
            case C_STORE4: {
              uint32_t* dst = reinterpret_cast<uint32_t*>(smem + cur.x);
#pragma unroll
              for (int k = 0; k < R; ++k) dst[row_first + 32 * k] = static_cast<uint32_t>(acc[k]);
            } break;
            case C_OUT8: {
              u64* dst = reinterpret_cast<u64*>(obuf + cur.x);
#pragma unroll
              for (int k = 0; k < R; ++k) if ((pass >> k) & 1u) dst[pos[k]] = acc[k];
            } break;
            case C_OUT4: {
              uint32_t* dst = reinterpret_cast<uint32_t*>(obuf + cur.x);
#pragma unroll
              for (int k = 0; k < R; ++k) if ((pass >> k) & 1u) dst[pos[k]] = static_cast<uint32_t>(acc[k]);
            } break;
            default: break;
          }
          if (post) {
            if (post == 1) {
              if ((cur.code_flags >> 16) & F_NEGATE) bits ^= all;
              if (!((cur.code_flags >> 16) & F_THEN_PRED)) {
#pragma unroll
                for (int k = 0; k < R; ++k) acc[k] = (bits >> k) & 1u;
                continue;
              }
            }
            if constexpr (SINK) {
              pass = bits & ~accn & live;   // the predicate only masks rows: nothing is compacted
            } else {
              pass = bits & ~accn & live;
              // in-tile compaction: the warp's rows are contiguous, so positions inside the warp
              // come from R ballots; one small scan over the warps gives the warp bases
              uint32_t run = 0;
#pragma unroll
              for (int k = 0; k < R; ++k) {
                const uint32_t m = __ballot_sync(0xffffffffu, (pass >> k) & 1u);
                pos[k] = static_cast<int>(run + __popc(m & lt));
                run += __popc(m);
              }
              uint32_t* wc = warp_cnt + 16 * (it & 1);   // double buffered: one barrier per tile
              if (lane == 0) wc[warp] = run;
              bar_consumers<NT>();
              uint32_t before = 0, total = 0;
#pragma unroll
              for (int w = 0; w < NW; ++w) {
                const uint32_t c = wc[w];
                if (w < warp) before += c;
                total += c;
              }
#pragma unroll
              for (int k = 0; k < R; ++k) pos[k] += static_cast<int>(before);
              if (tid == 0) {
                s_meta[ob] = total;
                st_relaxed(&p.tile_status[tile], kValid | total);   // this tile's share of its wave
              }
            }
          }
          continue;
        }
        // ---- generic path: NULL-carrying operands, narrow or mixed types, rare ops.
        // Works on 4 rows at a time so that its register footprint does not grow with R.
        const Insn& in = p.insn[pc];
        auto fetch4 = [&](int c4, int idx, bool is_imm, bool null_const, bool nullable, int width, u64 (&v)[4], uint32_t& nn) {
          if (is_imm) {
#pragma unroll
            for (int q = 0; q < 4; ++q) v[q] = p.imm[idx];
            nn = null_const ? 15u : 0u;
            return;
          }
          const unsigned char* base = slot_data(idx, stage);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int r = row_first + 32 * (c4 * 4 + q);
            v[q] = width == 8 ? reinterpret_cast<const u64*>(base)[r]
                              : (width == 4 ? static_cast<u64>(reinterpret_cast<const uint32_t*>(base)[r]) : static_cast<u64>(base[r]));
          }
          nn = 0;
          if (nullable) {
            const uint32_t* w = slot_nullw(idx, stage);
            if (w != nullptr) {
#pragma unroll
              for (int q = 0; q < 4; ++q) nn |= ((w[warp * R + c4 * 4 + q] >> lane) & 1u) << q;
            }
          }
        };
        switch (in.kind) {
          case K_LOAD:
          case K_ALU1:
          case K_ALU2:
          case K_ALU3: {
#pragma unroll
            for (int c4 = 0; c4 < R / 4; ++c4) {
              u64 a4[4], r1[4], r2[4];
              uint32_t an = (accn >> (4 * c4)) & 15u, n1 = 0, n2 = 0;
#pragma unroll
              for (int q = 0; q < 4; ++q) { a4[q] = acc[c4 * 4 + q]; r1[q] = 0; r2[q] = 0; }
              if (in.kind != K_ALU1) fetch4(c4, in.a, in.flags & F_RHS_IMM, in.flags & F_RHS_NULLK, in.rhs_nullable & 1, in.rw, r1, n1);
              // rhs2 has the width of rhs (the branches of IF), except a guard, which is a BOOL
              if (in.kind == K_ALU3) fetch4(c4, in.b, in.flags & F_RHS2_IMM, in.rhs_nullable & 4, in.rhs_nullable & 2,
                                            (in.flags & F_GUARDED) ? 1 : in.rw, r2, n2);
              if (in.kind == K_LOAD) {
#pragma unroll
                for (int q = 0; q < 4; ++q) a4[q] = r1[q];
                an = n1;
              } else {
                uint32_t f4 = 0;
                alu<4>(in, a4, an, r1, n1, r2, n2, (live >> (4 * c4)) & 15u, f4);
                fail |= f4;
              }
#pragma unroll
              for (int q = 0; q < 4; ++q) acc[c4 * 4 + q] = a4[q];
              accn = (accn & ~(15u << (4 * c4))) | ((an & 15u) << (4 * c4));
            }
          } break;
          case K_STORE: {
            unsigned char* base = slot_data(in.a, stage);
            if (in.rw == 8) {
#pragma unroll
              for (int k = 0; k < R; ++k) reinterpret_cast<u64*>(base)[row_first + 32 * k] = acc[k];
            } else if (in.rw == 4) {
#pragma unroll
              for (int k = 0; k < R; ++k) reinterpret_cast<uint32_t*>(base)[row_first + 32 * k] = static_cast<uint32_t>(acc[k]);
            } else {
#pragma unroll
              for (int k = 0; k < R; ++k) base[row_first + 32 * k] = static_cast<unsigned char>(acc[k]);
            }
            if (in.rhs_nullable & 1) {
              uint32_t* w = slot_nullw(in.a, stage);
#pragma unroll
              for (int k = 0; k < R; ++k) {
                const uint32_t b = __ballot_sync(0xffffffffu, (accn >> k) & 1u);
                if (lane == 0) w[warp * R + k] = b;
              }
              __syncwarp();   // read back by the lanes of this warp only
            }
          } break;
          case K_OUT: {
            const int j = in.a;
            unsigned char* dst = obuf + p.out_off[j];
            if (in.rw == 8) {
#pragma unroll
              for (int k = 0; k < R; ++k) if ((pass >> k) & 1u) reinterpret_cast<u64*>(dst)[pos[k]] = acc[k];
            } else if (in.rw == 4) {
#pragma unroll
              for (int k = 0; k < R; ++k) if ((pass >> k) & 1u) reinterpret_cast<uint32_t*>(dst)[pos[k]] = static_cast<uint32_t>(acc[k]);
            } else {
#pragma unroll
              for (int k = 0; k < R; ++k) if ((pass >> k) & 1u) dst[pos[k]] = static_cast<unsigned char>(acc[k]);
            }
            if (p.out_null_off[j] != 0xffffffffu) {
              unsigned char* nb = obuf + p.out_null_off[j];
#pragma unroll
              for (int k = 0; k < R; ++k) if ((pass >> k) & 1u) nb[pos[k]] = static_cast<unsigned char>((accn >> k) & 1u);
            }
          } break;
          default: break;
        }
      }
#undef SSB_BIN4
#undef SSB_CMP4
#undef SSB_MAD2
#undef SSB_BIN_TYPE
#undef SSB_ENC
      if constexpr (SINK) {
        // ================================================== the aggregation sink of this tile
        // Every output of the program sits in a shared-memory slot (an input column of this stage or
        // a temporary). Keys first: the group of each of the thread's R rows, found by comparing the
        // keys against the CTA's published entries held in registers (branch-free; the first row of a
        // group, or a row of a group without a local entry, takes the cold path).
        auto src_load = [&](int j, u64 (&v)[R]) -> uint32_t {   // values of output j for the thread's rows; returns the NULL bits
          const uint32_t o = p.sink_src_off[j];
          const unsigned char* base = smem + (o & 0x7fffffffu) + ((o >> 31) ? static_cast<uint32_t>(stage) * p.stage_bytes : 0u);
          const int w = p.sink_src_w[j];
          if (w == 8) {
#pragma unroll
            for (int k = 0; k < R; ++k) v[k] = reinterpret_cast<const u64*>(base)[row_first + 32 * k];
          } else if (w == 4) {
#pragma unroll
            for (int k = 0; k < R; ++k) v[k] = reinterpret_cast<const uint32_t*>(base)[row_first + 32 * k];
          } else {
#pragma unroll
            for (int k = 0; k < R; ++k) v[k] = base[row_first + 32 * k];
          }
          uint32_t nn = 0;
          if (p.sink_src_nullable[j]) {
            const uint32_t* wds = slot_nullw(p.sink_src_slot[j], stage);
            if (wds != nullptr) {
#pragma unroll
              for (int k = 0; k < R; ++k) nn |= ((wds[warp * R + k] >> lane) & 1u) << k;
            }
          }
          return nn;
        };
        u64 kv0[R], kv1[R];
        uint32_t kn0 = 0, kn1 = 0;
#pragma unroll
        for (int k = 0; k < R; ++k) { kv0[k] = 0; kv1[k] = 0; }
        if (SNK > 0) kn0 = src_load(0, kv0);
        if (SNK > 1) kn1 = src_load(1, kv1);
        if (kn0 | kn1) {
#pragma unroll
          for (int k = 0; k < R; ++k) { if ((kn0 >> k) & 1u) kv0[k] = 0; if ((kn1 >> k) & 1u) kv1[k] = 0; }
        }
        // 32-bit fingerprints select the candidate entry (two instructions per row and entry); the key
        // values confirm it. A confirmed mismatch is a miss (cold path, which is exact).
        const uint32_t ready = *reinterpret_cast<volatile unsigned int*>(sk.l_ready);
        uint32_t fpr[R];
#pragma unroll
        for (int k = 0; k < R; ++k) {
          const uint32_t kn = ((kn0 >> k) & 1u) | (((kn1 >> k) & 1u) << 1);
          fpr[k] = sink_fp32(kv0[k], kv1[k], kn);
        }
        int gsel[R];
#pragma unroll
        for (int k = 0; k < R; ++k) gsel[k] = -1;
#pragma unroll
        for (int e = 0; e < kTinyGroups; ++e) {
          if (!((ready >> e) & 1u)) continue;   // CTA-uniform
          const uint32_t ef = sk.l_fp[e];
#pragma unroll
          for (int k = 0; k < R; ++k) if (fpr[k] == ef) gsel[k] = e;
        }
#pragma unroll
        for (int k = 0; k < R; ++k) {
          const int e = gsel[k] < 0 ? 0 : gsel[k];
          const uint32_t kn = ((kn0 >> k) & 1u) | (((kn1 >> k) & 1u) << 1);
          const bool same = sk.l_key[2 * e] == kv0[k] && sk.l_key[2 * e + 1] == kv1[k] && sk.l_knull[e] == kn;
          if (!same) gsel[k] = -1;
        }
        uint32_t goff[R];
        uint32_t ovf = 0;
        uint32_t miss = 0;
#pragma unroll
        for (int k = 0; k < R; ++k) {
          const bool on = (pass >> k) & 1u;
          goff[k] = (on && gsel[k] >= 0) ? static_cast<uint32_t>(gsel[k]) * sink_block : sink_trash;
          miss |= (on && gsel[k] < 0 ? 1u : 0u) << k;
        }
        if (miss) {   // cold: first sight of a group in this CTA, or no local entry left for it
#pragma unroll
          for (int k = 0; k < R; ++k) {
            if (!((miss >> k) & 1u)) continue;
            const uint32_t kn = ((kn0 >> k) & 1u) | (((kn1 >> k) & 1u) << 1);
            const int g = sink_insert(*gp, smem + p.sink_off, SA, NT, SG, kv0[k], kv1[k], kn);
            if (g >= 0) {
              goff[k] = static_cast<uint32_t>(g) * sink_block;
            } else if (g == -1) {
              ovf |= 1u << k;
            } else {   // the global table is full: the host grows it and replays this row
              const unsigned long long d = atomicAdd(gp->n_deferred, 1ull);
              gp->deferred[d] = row0 + row_first + 32 * k;
              pass &= ~(1u << k);
            }
          }
        }
        unsigned long long* const acc_tid = sk.t_acc + tid;
        // COUNT(*) aggregates count the row
        for (uint32_t m = p.sink_count_star; m; m &= m - 1) {
          const int a = __ffs(m) - 1;
          unsigned char* const base = reinterpret_cast<unsigned char*>(acc_tid + a * NT);
#pragma unroll
          for (int k = 0; k < R; ++k) *reinterpret_cast<unsigned long long*>(base + goff[k]) += 1ull;
          if (ovf) {
#pragma unroll
            for (int k = 0; k < R; ++k) {
              if (!((ovf >> k) & 1u)) continue;
              const uint32_t kn = ((kn0 >> k) & 1u) | (((kn1 >> k) & 1u) << 1);
              sink_apply_global(*gp, a, kv0[k], kv1[k], kn, 0ull, true);
            }
          }
        }
        // value outputs: output j feeds the aggregates in sink_out_aggs[j]
        for (int j = SNK; j < p.n_out; ++j) {
          if (p.sink_out_aggs[j] == 0u) continue;
          u64 v[R];
          const uint32_t nulls = src_load(j, v);
          const uint32_t valid = pass & ~nulls;
          uint32_t off[R];   // rows without a value (predicate, NULL, overflow) accumulate into the trash block
#pragma unroll
          for (int k = 0; k < R; ++k) off[k] = ((valid >> k) & 1u) ? goff[k] : sink_trash;
          for (uint32_t m = p.sink_out_aggs[j]; m; m &= m - 1) {
            const int a = __ffs(m) - 1;
            const uint32_t code = (p.sink_pad >> (2 * a)) & 3u;
            unsigned char* const base = reinterpret_cast<unsigned char*>(acc_tid + a * NT);
            if (code != TA_OTHER && code != TA_COUNT && ((p.sink_seen_mask >> a) & 1u)) {   // nullable input: all-NULL groups stay NULL
#pragma unroll
              for (int k = 0; k < R; ++k) sk.t_seen[(reinterpret_cast<unsigned long long*>(base + off[k]) - sk.t_acc)] = 1;
            }
            if (code == TA_SUM_F64) {
#pragma unroll
              for (int k = 0; k < R; ++k) {
                unsigned long long* q = reinterpret_cast<unsigned long long*>(base + off[k]);
                *q = Codec<double>::enc(Codec<double>::dec(*q) + Codec<double>::dec(v[k]));
              }
            } else if (code == TA_SUM_U64) {
#pragma unroll
              for (int k = 0; k < R; ++k) *reinterpret_cast<unsigned long long*>(base + off[k]) += v[k];
            } else if (code == TA_COUNT) {
#pragma unroll
              for (int k = 0; k < R; ++k) *reinterpret_cast<unsigned long long*>(base + off[k]) += 1ull;
            } else {
              const AggDev& ag = gp->agg[a];
#pragma unroll
              for (int k = 0; k < R; ++k) {
                unsigned long long* q = reinterpret_cast<unsigned long long*>(base + off[k]);
                unsigned char* sn = sk.t_seen + (q - sk.t_acc);
                *q = *sn ? combine(ag, *q, v[k]) : v[k];   // the first value is taken as it is (NaN, -0.0)
                *sn = 1;
              }
            }
            if (ovf & valid) {   // groups beyond the CTA's local entries: the global table
#pragma unroll
              for (int k = 0; k < R; ++k) {
                if (!(((ovf & valid) >> k) & 1u)) continue;
                const uint32_t kn = ((kn0 >> k) & 1u) | (((kn1 >> k) & 1u) << 1);
                sink_apply_global(*gp, a, kv0[k], kv1[k], kn, v[k], false);
              }
            }
          }
        }
      }
    }
    if (scan_wave) finish_wave();
    __syncthreads();   // all threads: the stage and the temporaries are free; the wave sums are published

    // ======================================================== write out tile `it - kdefer_`
    if (!SINK && it >= kdefer_) {
      const int ob_out = (ob + 1 == n_obuf) ? 0 : ob + 1;   // (it - kdefer_) % n_obuf
      const int tile = bid + (it - kdefer_) * G;
      const unsigned char* obuf = smem + p.off_out + ob_out * p.out_bytes;
      int total;
      long long base;
      if (p.has_pred) {
        uint32_t before = 0, sum = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
          const uint2 v = red[(it & 1) * 16 + w];
          before += v.x;
          sum += v.y;
        }
        base = wave_base + before;
        wave_base += sum;
        total = static_cast<int>(s_meta[ob_out]);
        if (tile == num_tiles - 1 && tid == 0 && p.d_out_rows != nullptr) *p.d_out_rows = base + total;
      } else {
        base = static_cast<long long>(tile) * TILE;
        total = static_cast<int>(p.rows - base < TILE ? p.rows - base : TILE);
      }
      for (int j = 0; j < p.n_out; ++j) {
        const unsigned char* src = obuf + p.out_off[j];
        const int w = p.out_width[j];
        if (w == 8) {
          u64* d = static_cast<u64*>(p.out_data[j]) + base;
          const u64* s = reinterpret_cast<const u64*>(src);
          if (total == TILE) {   // all loads first, then all stores
            u64 v[R];
#pragma unroll
            for (int k = 0; k < R; ++k) v[k] = s[tid + k * NT];
#pragma unroll
            for (int k = 0; k < R; ++k) st_cs_u64(d + tid + k * NT, v[k]);
          } else {
#pragma unroll
            for (int k = 0; k < R; ++k) {
              if (k * NT >= total) break;   // CTA-uniform
              const int i = tid + k * NT;
              if (i < total) st_cs_u64(d + i, s[i]);
            }
          }
        } else if (w == 4) {
          uint32_t* d = static_cast<uint32_t*>(p.out_data[j]) + base;
          const uint32_t* s = reinterpret_cast<const uint32_t*>(src);
          if (total == TILE) {
            uint32_t v[R];
#pragma unroll
            for (int k = 0; k < R; ++k) v[k] = s[tid + k * NT];
#pragma unroll
            for (int k = 0; k < R; ++k) st_cs_u32(d + tid + k * NT, v[k]);
          } else {
#pragma unroll
            for (int k = 0; k < R; ++k) {
              if (k * NT >= total) break;   // CTA-uniform
              const int i = tid + k * NT;
              if (i < total) st_cs_u32(d + i, s[i]);
            }
          }
        } else {
          unsigned char* d = static_cast<unsigned char*>(p.out_data[j]) + base;
          for (int i = tid; i < total; i += NT) d[i] = src[i];
        }
        if (p.out_null_off[j] != 0xffffffffu && p.out_nulls[j] != nullptr) {
          // output bitmap words: interior words belong to this tile alone, the first and last may
          // be shared with the neighbouring tiles (the bitmap was zeroed before the launch)
          const unsigned char* nb = obuf + p.out_null_off[j];
          const long long first = base & ~31LL;
          const long long end = base + total;
          for (long long g = first + tid; g < ((end + 31) & ~31LL); g += NT) {
            const bool bit = g >= base && g < end && nb[g - base] != 0;
            const uint32_t word = __ballot_sync(0xffffffffu, bit);
            if (lane == 0 && word != 0u) {
              if (g >= base && g + 32 <= end) p.out_nulls[j][g >> 5] = word;
              else atomicOr(&p.out_nulls[j][g >> 5], word);
            }
          }
        }
      }
      // No barrier needed here: the next write into this staging buffer happens two evaluations
      // later, behind at least one __syncthreads of the next iteration.
    }
    if (++stage == S) { stage = 0; parity ^= 1u; }
    if (++ob == n_obuf) ob = 0;
  }
  if constexpr (SINK) {
    // flush: the threads' partials of every (local group, aggregate) are combined by one warp and
    // applied to the global table once per CTA. Accumulators of the SUM / COUNT codes start from an
    // identity that adding leaves exact (0, -0.0), so untouched partials need no bookkeeping; the
    // others (MIN / MAX, narrow sums) carry a seen byte per partial.
    bar_consumers<NT>();
    for (int ga = warp; ga < SG * SA; ga += NW) {
      const int g = ga / SA, a = ga - g * SA;
      if (sk.l_slot[g] == 0u) continue;
      const AggDev& ag = gp->agg[a];
      const uint32_t code = (p.sink_pad >> (2 * a)) & 3u;
      const bool track = code == TA_OTHER || ((p.sink_seen_mask >> a) & 1u);   // partials carry a seen byte
      unsigned long long acc2 = code == TA_OTHER ? 0ull : sink_identity(ag);
      bool has = code != TA_OTHER;
      unsigned int any = 0;
      for (int t = lane; t < NT; t += 32) {
        const unsigned long long x = sk.t_acc[ga * NT + t];
        const unsigned int sn = track ? sk.t_seen[ga * NT + t] : 1u;
        any |= sn;
        if (code == TA_COUNT || code == TA_SUM_U64) acc2 += x;
        else if (code == TA_SUM_F64) acc2 = Codec<double>::enc(Codec<double>::dec(acc2) + Codec<double>::dec(x));
        else if (sn) { acc2 = has ? combine(ag, acc2, x) : x; has = true; }
      }
      for (int d = 16; d > 0; d >>= 1) {
        const unsigned long long ov = __shfl_xor_sync(0xffffffffu, acc2, d);
        const bool oh = __shfl_xor_sync(0xffffffffu, has ? 1 : 0, d) != 0;
        any |= __shfl_xor_sync(0xffffffffu, any, d);
        if (code == TA_COUNT || code == TA_SUM_U64) acc2 += ov;
        else if (code == TA_SUM_F64) acc2 = Codec<double>::enc(Codec<double>::dec(acc2) + Codec<double>::dec(ov));
        else if (oh) { acc2 = has ? combine(ag, acc2, ov) : ov; has = true; }
      }
      if (lane != 0) continue;
      const long long slot = static_cast<long long>(sk.l_slot[g] - 1u);
      if (ag.fn == SSB_AGG_COUNT) { if (acc2) atomicAdd(&ag.acc[static_cast<unsigned long long>(slot) * ag.stride], acc2); continue; }
      if (!has || !any) continue;   // no value of this aggregate reached the group in this CTA (all NULL)
      if (ag.seen != nullptr) ag.seen[slot] = 1u;
      apply(ag, slot, acc2, 0ull);
    }
    if (fail && p.d_fail != nullptr) atomicOr(p.d_fail, 1);
    return;
  }
  if (!p.has_pred && p.d_out_rows != nullptr && blockIdx.x == 0 && tid == 0) *p.d_out_rows = p.rows;
  if (fail && p.d_fail != nullptr) atomicOr(p.d_fail, 1);
}


template <int NT, int R>
__global__ void __launch_bounds__(NT + 32) expr_kernel(const __grid_constant__ ExprParams p) {
  expr_kernel_body<NT, R, false>(p);
}
// The sink instantiation carries more live state through the dispatch loop (group ids of the
// thread's rows, the predicate mask, accumulator addresses): it is allowed the registers of two
// resident CTAs per SM instead of being squeezed to the default (which spilled that state).
template <int NT, int R>
__global__ void __launch_bounds__(NT + 32, (NT >= 192 ? 1 : 2)) expr_sink_kernel(const __grid_constant__ ExprParams p) {
  expr_kernel_body<NT, R, true>(p);
}

// ------------------------------------------------------------------ host side
// Kernel variants: threads per CTA x rows per thread. More rows per thread amortise the
// interpreter's dispatch; the tile (and with it the shared-memory stage) grows with both.
struct Variant {
  int threads, rows_per_thread;
  void (*kernel)(const ExprParams);
  void (*sink_kernel)(const ExprParams);   // aggregation-sink instantiation, or nullptr
};
static const Variant kVariants[] = {
    {256, 4, expr_kernel<256, 4>, expr_sink_kernel<256, 4>},
    {128, 8, expr_kernel<128, 8>, expr_sink_kernel<128, 8>},
    {256, 8, expr_kernel<256, 8>, nullptr},
    {128, 16, expr_kernel<128, 16>, nullptr},
    {64, 16, expr_kernel<64, 16>, nullptr},
    {128, 4, expr_kernel<128, 4>, expr_sink_kernel<128, 4>},
    {64, 8, expr_kernel<64, 8>, expr_sink_kernel<64, 8>},
    {96, 8, expr_kernel<96, 8>, expr_sink_kernel<96, 8>},
    {96, 4, expr_kernel<96, 4>, expr_sink_kernel<96, 4>},
};
static const int kNumVariants = sizeof(kVariants) / sizeof(kVariants[0]);
static const int kDefaultVariant = 7;   // 96 consumer threads x 8 rows = 768-row tiles, three CTAs per SM

static int launch_program(ssb_program* sp, const ssb_column* inputs, int64_t rows,
                          const ssb_column* outputs, int64_t* d_out_rows) {
  if (!sp->parts.empty()) {
    // column groups: same predicate, same kept rows, each group writes its own output columns;
    // one timed region around all of them
    ssb_ctx* c = sp->ctx;
    const bool timing = c->timing;
    if (timing) cudaEventRecord(c->ev0, c->stream);
    c->timing = false;
    int rc = 0;
    for (size_t q = 0; q < sp->parts.size() && rc == 0; ++q) {
      const ssb_program::Part& part = sp->parts[q];
      std::vector<ssb_column> in(part.inputs.size() ? part.inputs.size() : 1);
      for (size_t i = 0; i < part.inputs.size(); ++i) in[i] = inputs[part.inputs[i]];
      rc = launch_program(part.prog, in.data(), rows, outputs + part.first_out, d_out_rows);
    }
    c->timing = timing;
    if (timing) { cudaEventRecord(c->ev1, c->stream); c->ev_valid = true; }
    return rc;
  }
  ssb_ctx* ctx = sp->ctx;
  Program& prog = sp->prog;
  ExprParams p = prog.params;
  const Variant& var = kVariants[prog.variant];
  if (rows < 0) return fail(ctx, SSB_ERROR_INVALID_ARGUMENT_VALUE, "negative row count");
  if (rows == 0) {
    if (d_out_rows) SSB_CUDA(ctx, cudaMemsetAsync(d_out_rows, 0, sizeof(int64_t), ctx->stream));
    return 0;
  }
  bool aligned = true;
  for (int i = 0; i < p.n_in; ++i) {
    if (inputs[i].data == nullptr) return fail(ctx, SSB_ERROR_INVALID_ARGUMENT_VALUE, "input column without data");
    if (phys_of(inputs[i].dtype) != phys_of(prog.input_types[i])) return fail(ctx, SSB_ERROR_INVALID_ARGUMENT_TYPE, "input column type differs from the compiled program");
    if (inputs[i].nulls != nullptr && !p.in_nullable[i]) return fail(ctx, SSB_ERROR_INVALID_ARGUMENT_VALUE, "input column carries NULLs but was declared NOT_NULLABLE");
    p.in_data[i] = inputs[i].data;
    p.in_nulls[i] = inputs[i].nulls;
    if (reinterpret_cast<uintptr_t>(inputs[i].data) & 15) aligned = false;
    if (inputs[i].nulls && (reinterpret_cast<uintptr_t>(inputs[i].nulls) & 15)) aligned = false;
  }
  for (int j = 0; j < p.n_out; ++j) {
    if (outputs[j].data == nullptr) return fail(ctx, SSB_ERROR_INVALID_ARGUMENT_VALUE, "output column without data");
    if (p.out_nullable[j] && outputs[j].nulls == nullptr) return fail(ctx, SSB_ERROR_INVALID_ARGUMENT_VALUE, "nullable output column without a null bitmap");
    p.out_data[j] = outputs[j].data;
    p.out_nulls[j] = p.out_nullable[j] ? outputs[j].nulls : nullptr;
    // every nullable output bitmap is zeroed: only non-zero words are written by the kernel
    if (p.out_nulls[j]) {
      SSB_CUDA(ctx, cudaMemsetAsync(p.out_nulls[j], 0, static_cast<size_t>(div_up(rows, 32) + 1) * 4, ctx->stream));
    }
  }
  p.fill_nullw_tma = p.fill_nullw_plain = 0;
  for (int i = 0; i < p.n_in; ++i) {
    if (p.in_nullw[i] < 0) continue;
    p.fill_nullw_plain = 1;
    if (p.in_nulls[i] == nullptr) p.fill_nullw_tma = 1;
  }
  p.rows = rows;
  p.num_tiles = div_up(rows, p.tile);
  if (p.num_tiles > 0x7fff0000LL) return fail(ctx, SSB_ERROR_NOT_IMPLEMENTED, "more than 2^31 tiles in one launch");
  p.use_tma = aligned ? 1 : 0;
  p.d_out_rows = d_out_rows;
  p.d_fail = prog.has_signaling ? ctx->d_fail : nullptr;
  p.debug_nowait = getenv("SSB200_DEBUG_NOWAIT") ? 1 : 0;   // experiment only: breaks output positions
  if (p.has_pred) {
    void* st = nullptr;
    if (int rc = scratch(ctx, static_cast<size_t>(p.num_tiles) * 8, &st)) return rc;
    p.tile_status = static_cast<unsigned long long*>(st);
    SSB_CUDA(ctx, cudaMemsetAsync(st, 0, static_cast<size_t>(p.num_tiles) * 8, ctx->stream));
  }
  long long grid = static_cast<long long>(ctx->num_sms) * sp->max_ctas_per_sm;
  if (p.has_pred && grid > 768) grid = 768;   // the wave scan reads at most 768 status words
  if (grid > p.num_tiles) grid = p.num_tiles;
  TimedRegion timed(ctx);
  if (p.has_pred) {
    // Filter: the CTAs of a wave read each other's kept-row counts (the wave-synchronous prefix spins
    // on status words of CTAs with higher indices), so the whole grid must be resident at once. The
    // grid is sized from the occupancy query, but only a cooperative launch makes co-residency a
    // guarantee when other kernels (the second streaming lane, another context) share the device:
    // the runtime then starts the grid only when all of it fits, or fails the launch.
    void* args[] = {&p};
    SSB_CUDA(ctx, cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(var.kernel), dim3(static_cast<unsigned>(grid)),
                                              dim3(var.threads + 32), args, prog.smem_bytes, ctx->stream));
  } else {
    var.kernel<<<static_cast<unsigned>(grid), var.threads + 32, prog.smem_bytes, ctx->stream>>>(p);
  }
  ++ctx->launches;
  SSB_CUDA(ctx, cudaGetLastError());
  return 0;
}


// ---- aggregation sink: host side -------------------------------------------------------------
// Shared memory of the sink per consumer thread / fixed part (must match the pointer arithmetic at
// the top of the kernel).
static uint32_t sink_bytes_per_thread(int n_keys, int n_aggs, int groups, int rows_per_thread) {
  (void)n_keys; (void)rows_per_thread;
  return static_cast<uint32_t>((groups + 1) * n_aggs * 9);   // accumulator + seen byte per (group incl. trash, aggregate)
}
static uint32_t sink_fixed_bytes() {
  return static_cast<uint32_t>(kTinyGroups * 2 * 8 + kTinyGroups * 4 * 3 + 16 + 64);
}

// Compiles (once per shape) the twin of `base` whose outputs feed the aggregation sink.
int sink_program_for(ssb_program* base, int n_keys, int n_aggs, int groups, ssb_program** out) {
  ssb_ctx* ctx = base->ctx;
  *out = nullptr;
  const uint32_t key = static_cast<uint32_t>(n_keys) | (static_cast<uint32_t>(n_aggs) << 8) | (static_cast<uint32_t>(groups) << 16);
  if (base->sink != nullptr && base->sink_key == key) { *out = base->sink; return 0; }
  delete base->sink;
  base->sink = nullptr;
  const Program& bp = base->prog;
  struct Try { int variant, ctas; };
  // 512-row tiles of 64 consumer threads x 8 rows first: eight rows per thread amortise the
  // interpreter's dispatch, the small tile leaves room for two or three resident CTAs next to the
  // per-thread accumulators (the kernel is bound by instruction issue and latency, not by bytes)
  // Measured on the Q1 shape (200M rows, profiles/r2d_q1_sink_variants.txt): the kernel is bound by instruction
  // issue with one consumer warp per scheduler, so consumer warps per SM decide. One CTA of 256 consumer threads x
  // 4 rows (8 consumer warps, the whole SM's shared memory) 9.06 ms; two CTAs of 64 x 8 (4 warps, half the
  // dispatch cost per row) 11.05 ms; 128 x 8 in one CTA 10.8 ms; two CTAs of 96 x 8 14.4 ms.
  Try tries[] = {{0, 1}, {6, 3}, {6, 2}, {7, 2}, {8, 3}, {8, 2}, {6, 1}, {7, 1}, {8, 1}};
  if (const char* ev = getenv("SSB200_SINK_VARIANT")) {   // experiments: pin the tile variant / resident CTAs
    const int v = atoi(ev), c = getenv("SSB200_SINK_CTAS") ? atoi(getenv("SSB200_SINK_CTAS")) : 2;
    if (v >= 0 && v < kNumVariants && kVariants[v].sink_kernel != nullptr && c >= 1) {
      for (size_t i = 0; i < sizeof(tries) / sizeof(tries[0]); ++i) { tries[i].variant = v; tries[i].ctas = c; }
    }
  }
  ssb_program* best = nullptr;
  long long best_score = -1;
  std::string err;
  int rc = SSB_ERROR_NOT_IMPLEMENTED;
  for (size_t i = 0; i < sizeof(tries) / sizeof(tries[0]); ++i) {
    const Variant& var = kVariants[tries[i].variant];
    const int c = tries[i].ctas;
    const uint32_t budget = static_cast<uint32_t>((ctx->smem_per_sm - c * ctx->smem_reserved) / c);
    ssb_program* cand = new ssb_program;
    cand->ctx = ctx;
    std::string e2;
    const int r2 = compile_program(bp.nodes.data(), static_cast<int32_t>(bp.nodes.size()), static_cast<int32_t>(bp.input_types.size()),
                                   bp.input_types.data(), bp.input_nullable.data(), bp.outputs.data(),
                                   static_cast<int32_t>(bp.outputs.size()), bp.predicate, var.threads * var.rows_per_thread,
                                   budget, static_cast<uint32_t>(ctx->smem_optin), &cand->prog, &e2,
                                   sink_bytes_per_thread(n_keys, n_aggs, groups, var.rows_per_thread), sink_fixed_bytes(), var.threads);
    if (r2 != 0) { delete cand; rc = r2; err = e2; continue; }
    cand->prog.variant = tries[i].variant;
    const int stages = cand->prog.params.stages;
    const bool fits = cand->prog.smem_bytes <= budget;
    const int resident = fits ? c : static_cast<int>(ctx->smem_per_sm / (cand->prog.smem_bytes + ctx->smem_reserved));
    // bytes in flight per SM (three stages are enough), more resident CTAs (= consumer warps) on a tie
    const long long score = static_cast<long long>(stages >= 2 ? (stages > 3 ? 3 : stages) : 0) * (resident < 1 ? 1 : resident) *
                                cand->prog.params.tile * 16 + (resident < 1 ? 1 : resident);
    // the tries are in order of preference: the first one with two stages and two resident CTAs is taken
    if (stages >= 2 && resident >= c) { delete best; best = cand; break; }
    if (score > best_score) { delete best; best = cand; best_score = score; } else { delete cand; }
  }
  if (best == nullptr) return fail(ctx, rc, err);
  if (getenv("SSB200_DEBUG_PLAN")) {
    fprintf(stderr, "[ssb200] sink plan: variant %d (tile %d), %d stages, %u bytes of shared memory per CTA\n", best->prog.variant,
            best->prog.params.tile, best->prog.params.stages, best->prog.smem_bytes);
  }
  const Variant& var = kVariants[best->prog.variant];
  cudaError_t e = cudaFuncSetAttribute(var.sink_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(ctx->smem_optin));
  int occ = 0;
  if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, var.sink_kernel, var.threads + 32, best->prog.smem_bytes);
  if (e != cudaSuccess || occ < 1) { delete best; return cuda_fail(ctx, e, "occupancy(expr_kernel sink)"); }
  best->max_ctas_per_sm = occ;
  base->sink = best;
  base->sink_key = key;
  *out = best;
  return 0;
}

// Launches the sink kernel of `sp` (from sink_program_for) over `rows` rows. d_gp: the table's
// GroupParams in device memory; out_aggs[j]: aggregates fed by output j; count_star: COUNT(*) mask.
int launch_program_sink(ssb_program* sp, const ssb_column* inputs, int64_t rows, const void* d_gp, int n_keys,
                        int n_aggs, int groups, const uint32_t* out_aggs, uint32_t count_star, uint32_t pad_codes,
                        uint32_t seen_mask) {
  ssb_ctx* ctx = sp->ctx;
  Program& prog = sp->prog;
  ExprParams p = prog.params;
  const Variant& var = kVariants[prog.variant];
  if (rows <= 0) return 0;
  bool aligned = true;
  for (int i = 0; i < p.n_in; ++i) {
    p.in_data[i] = inputs[i].data;
    p.in_nulls[i] = inputs[i].nulls;
    if (reinterpret_cast<uintptr_t>(inputs[i].data) & 15) aligned = false;
    if (inputs[i].nulls && (reinterpret_cast<uintptr_t>(inputs[i].nulls) & 15)) aligned = false;
  }
  p.fill_nullw_tma = p.fill_nullw_plain = 0;
  for (int i = 0; i < p.n_in; ++i) {
    if (p.in_nullw[i] < 0) continue;
    p.fill_nullw_plain = 1;
    if (p.in_nulls[i] == nullptr) p.fill_nullw_tma = 1;
  }
  p.rows = rows;
  p.num_tiles = div_up(rows, p.tile);
  if (p.num_tiles > 0x7fff0000LL) return fail(ctx, SSB_ERROR_NOT_IMPLEMENTED, "more than 2^31 tiles in one launch");
  p.use_tma = aligned ? 1 : 0;
  p.d_out_rows = nullptr;
  p.tile_status = nullptr;
  p.d_fail = prog.has_signaling ? ctx->d_fail : nullptr;
  p.debug_nowait = 0;
  p.sink_gp = d_gp;
  p.sink_n_keys = n_keys;
  p.sink_n_aggs = n_aggs;
  p.sink_groups = groups;
  for (int j = 0; j < kMaxOut; ++j) p.sink_out_aggs[j] = j < p.n_out ? out_aggs[j] : 0u;
  p.sink_count_star = count_star;
  p.sink_pad = pad_codes;
  p.sink_seen_mask = seen_mask;
  p.sink_keys_not_null = 1;
  for (int j = 0; j < n_keys; ++j) if (p.out_nullable[j]) p.sink_keys_not_null = 0;
  long long grid = static_cast<long long>(ctx->num_sms) * sp->max_ctas_per_sm;
  if (grid > p.num_tiles) grid = p.num_tiles;
  var.sink_kernel<<<static_cast<unsigned>(grid), var.threads + 32, prog.smem_bytes, ctx->stream>>>(p);
  ++ctx->launches;
  SSB_CUDA(ctx, cudaGetLastError());
  return 0;
}

}  // namespace ssb

using namespace ssb;

extern "C" {

// One kernel for the whole program (plan search over tile variants and resident CTAs).
static int create_single_program(ssb_ctx* ctx, const ssb_expr_node* nodes, int32_t n_nodes,
                                 int32_t n_inputs, const int32_t* input_types,
                                 const int32_t* input_nullable, const int32_t* outputs,
                                 int32_t n_outputs, int32_t predicate, ssb_program** out) {
  *out = nullptr;
  int variant = kDefaultVariant;
  // Narrow Filter plans (at most 16 input bytes per row) are bound by instruction issue, not by
  // bytes: 2048-row tiles of 128 consumer threads x 16 rows amortise the per-tile work better
  // (measured, 400M rows: Filter(d<K) -> a 172 -> 229 G rows/s; the four-input C2 plan prefers the
  // 768-row default: 148 vs 105 G rows/s).
  int in_bytes = 0;
  for (int i = 0; i < n_inputs; ++i) in_bytes += width_of(input_types[i]);
  bool pinned_variant = false;
  if (predicate >= 0 && in_bytes <= 16) { variant = 3; pinned_variant = true; }
  if (const char* env = getenv("SSB200_EXPR_VARIANT")) {
    pinned_variant = true;
    const int v = atoi(env);
    if (v >= 0 && v < kNumVariants) variant = v;
  }
  // Two resident CTAs per SM is the design point; wide plans fall back to fewer stages, then
  // to the smallest tile.
  // Resident CTAs per SM (measured, profiles/r1_summary.md): Filter is insensitive between 3 and
  // 4 and prefers deeper staging; predicate-free programs gain 24 % from the fourth CTA.
  int ctas = predicate >= 0 ? 3 : 4;
  if (const char* env = getenv("SSB200_EXPR_CTAS")) { const int c = atoi(env); if (c >= 1 && c <= 8) ctas = c; }
  // Plan search. The default (768-row tiles, `ctas` resident CTAs) is kept whenever it leaves
  // three input stages in flight. Wide plans (many input / output columns: the Q1 shape has
  // seven of each) would otherwise end with one CTA per SM and a single stage, i.e. loads,
  // evaluation and stores in lock step; for them smaller tiles and fewer CTAs are tried and the
  // plan with the most bytes in flight per SM (stages >= 2 first) wins.
  struct Try { int variant, ctas; };
  std::vector<Try> tries;
  tries.push_back({variant, ctas});
  if (!pinned_variant) {
    const int half = 8;   // 96 x 4 = 384-row tiles
    for (int c = ctas; c >= 1; --c) { if (c != ctas) tries.push_back({variant, c}); tries.push_back({half, c}); }
    tries.push_back({0, 1});
  }
  ssb_program* sp = nullptr;
  std::string err;
  int rc = 0;
  long long best_score = -1;
  for (size_t i = 0; i < tries.size(); ++i) {
    const Variant& var = kVariants[tries[i].variant];
    const int c = tries[i].ctas;
    const uint32_t budget = static_cast<uint32_t>((ctx->smem_per_sm - c * ctx->smem_reserved) / c);
    ssb_program* cand = new ssb_program;
    cand->ctx = ctx;
    std::string e2;
    const int r2 = compile_program(nodes, n_nodes, n_inputs, input_types, input_nullable, outputs, n_outputs, predicate,
                                   var.threads * var.rows_per_thread, budget, static_cast<uint32_t>(ctx->smem_optin),
                                   &cand->prog, &e2);
    if (r2 != 0) {
      delete cand;
      if (sp == nullptr) { rc = r2; err = e2; }
      if (r2 != SSB_ERROR_NOT_IMPLEMENTED) break;   // a bind error: no other tile will change it
      continue;
    }
    cand->prog.variant = tries[i].variant;
    const int stages = cand->prog.params.stages;
    const bool fits = cand->prog.smem_bytes <= budget;
    const int resident = fits ? c : static_cast<int>((ctx->smem_per_sm) / (cand->prog.smem_bytes + ctx->smem_reserved));
    // bytes of input in flight per SM, with a heavy penalty for a single stage
    // rows in flight per SM; a fourth stage adds nothing (three already hide the HBM latency), so it
    // must not buy out resident CTAs. Equal rows in flight: more resident CTAs (more warps) win.
    long long score = static_cast<long long>(stages >= 2 ? (stages > 3 ? 3 : stages) : 0) * (resident < 1 ? 1 : resident) *
                          cand->prog.params.tile * 16 + (resident < 1 ? 1 : resident);
    if (i == 0 && stages >= 3 && fits) score = 1LL << 40;   // the measured default
    if (score > best_score) {
      delete sp;
      sp = cand;
      best_score = score;
      rc = 0;
    } else {
      delete cand;
    }
    if (best_score >= (1LL << 40)) break;
  }
  if (rc) return fail(ctx, rc, err);
  const Variant& var = kVariants[sp->prog.variant];
  cudaError_t e = cudaFuncSetAttribute(var.kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(ctx->smem_optin));
  if (e != cudaSuccess) { delete sp; return cuda_fail(ctx, e, "cudaFuncSetAttribute(expr_kernel)"); }
  int occ = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, var.kernel, var.threads + 32, sp->prog.smem_bytes);
  if (e != cudaSuccess || occ < 1) { delete sp; return cuda_fail(ctx, e, "occupancy(expr_kernel)"); }
  sp->max_ctas_per_sm = occ;
  *out = sp;
  return 0;
}

// The part of a program that computes outputs [first, first + count): only the nodes those outputs
// and the predicate reach, over only the input columns they read.
static int create_part(ssb_ctx* ctx, const Program& whole, int first, int count, ssb_program::Part* part) {
  const int n = static_cast<int>(whole.nodes.size());
  std::vector<char> used(n, 0);
  std::vector<int> stack;
  for (int j = first; j < first + count; ++j) stack.push_back(whole.outputs[j]);
  if (whole.predicate >= 0) stack.push_back(whole.predicate);
  while (!stack.empty()) {
    const int i = stack.back();
    stack.pop_back();
    if (used[i]) continue;
    used[i] = 1;
    if (whole.nodes[i].op == SSB_OP_INPUT) continue;
    for (int a = 0; a < 3; ++a) if (whole.nodes[i].arg[a] >= 0 && whole.nodes[i].op != SSB_OP_CONST) stack.push_back(whole.nodes[i].arg[a]);
  }
  std::vector<int> new_index(n, -1), input_slot(whole.input_types.size(), -1);
  std::vector<ssb_expr_node> nodes;
  std::vector<int32_t> types, nullable;
  part->inputs.clear();
  for (int i = 0; i < n; ++i) {
    if (!used[i]) continue;
    ssb_expr_node nd = whole.nodes[i];
    if (nd.op == SSB_OP_INPUT) {
      const int col = nd.arg[0];
      if (input_slot[col] < 0) {
        input_slot[col] = static_cast<int>(part->inputs.size());
        part->inputs.push_back(col);
        types.push_back(whole.input_types[col]);
        nullable.push_back(whole.input_nullable[col]);
      }
      nd.arg[0] = input_slot[col];
    } else if (nd.op != SSB_OP_CONST) {
      for (int a = 0; a < 3; ++a) if (nd.arg[a] >= 0) nd.arg[a] = new_index[nd.arg[a]];
    }
    new_index[i] = static_cast<int>(nodes.size());
    nodes.push_back(nd);
  }
  std::vector<int32_t> outs;
  for (int j = first; j < first + count; ++j) outs.push_back(new_index[whole.outputs[j]]);
  int32_t dummy = 0;
  part->first_out = first;
  part->n_out = count;
  return create_single_program(ctx, nodes.data(), static_cast<int32_t>(nodes.size()), static_cast<int32_t>(types.size()),
                               types.empty() ? &dummy : types.data(), nullable.empty() ? &dummy : nullable.data(), outs.data(),
                               count, whole.predicate >= 0 ? new_index[whole.predicate] : -1, &part->prog);
}

int ssb_program_create(ssb_ctx* ctx, const ssb_expr_node* nodes, int32_t n_nodes,
                       int32_t n_inputs, const int32_t* input_types,
                       const int32_t* input_nullable, const int32_t* outputs,
                       int32_t n_outputs, int32_t predicate, ssb_program** out) {
  ssb_program* sp = nullptr;
  if (int rc = create_single_program(ctx, nodes, n_nodes, n_inputs, input_types, input_nullable, outputs, n_outputs, predicate, &sp)) {
    *out = nullptr;
    return rc;
  }
  *out = sp;
  // Column groups for wide plans: a plan whose single kernel keeps fewer than three input stages in
  // flight or fewer than six consumer warps per SM (C2 variant B, nine outputs: one CTA per SM,
  // measured 1.0 TB/s) is evaluated as k kernels that share the predicate; every group reads the
  // predicate's inputs again, which costs less than running the whole plan starved (2.2 TB/s).
  static const int split_env = getenv("SSB200_EXPR_SPLIT") ? atoi(getenv("SSB200_EXPR_SPLIT")) : -1;
  auto starved = [](const ssb_program* q) {   // fewer than three stages in flight, or fewer than six consumer warps per SM
    const Variant& v = kVariants[q->prog.variant];
    return q->prog.params.stages < 3 || q->max_ctas_per_sm * (v.threads / 32) < 6;
  };
  const bool poor = starved(sp);
  if (split_env == 0 || n_outputs < 2 || (!poor && split_env < 2)) return 0;
  // group counts to try: about three outputs per group first (measured best on variant B), then more, then fewer
  std::vector<int> ks;
  if (split_env >= 2) {
    ks.push_back(split_env);
  } else {
    const int k0 = std::max(2, std::min(4, (n_outputs + 2) / 3));
    for (int k = k0; k <= 4; ++k) ks.push_back(k);
    for (int k = k0 - 1; k >= 2; --k) ks.push_back(k);
  }
  for (size_t ki = 0; ki < ks.size(); ++ki) {
    const int k = ks[ki];
    if (k > n_outputs) continue;
    std::vector<ssb_program::Part> parts(k);
    bool ok = true;
    for (int q = 0; q < k && ok; ++q) {
      const int first = n_outputs * q / k, last = n_outputs * (q + 1) / k;
      parts[q].prog = nullptr;
      if (create_part(ctx, sp->prog, first, last - first, &parts[q]) != 0) { ok = false; break; }
      if (starved(parts[q].prog)) ok = split_env >= 2;
    }
    if (getenv("SSB200_DEBUG_PLAN")) {
      fprintf(stderr, "[ssb200] plan: %d outputs, single kernel: tile %d, %d stages, %d CTAs/SM; %d column groups %s:", n_outputs,
              sp->prog.params.tile, sp->prog.params.stages, sp->max_ctas_per_sm, k, ok ? "accepted" : "rejected");
      for (int q = 0; q < k; ++q) {
        if (parts[q].prog) fprintf(stderr, " [%d outs, %zu ins, tile %d, %d stages, %d CTAs/SM]", parts[q].n_out, parts[q].inputs.size(),
                                   parts[q].prog->prog.params.tile, parts[q].prog->prog.params.stages, parts[q].prog->max_ctas_per_sm);
      }
      fprintf(stderr, "\n");
    }
    if (ok) { sp->parts = parts; return 0; }
    for (int q = 0; q < k; ++q) delete parts[q].prog;
  }
  return 0;
}

void ssb_program_destroy(ssb_program* prog) { delete prog; }
int32_t ssb_program_output_type(const ssb_program* prog, int32_t j) { return prog->prog.out_types[j]; }
int32_t ssb_program_output_nullable(const ssb_program* prog, int32_t j) { return prog->prog.out_nullable[j]; }
int32_t ssb_program_bytes_per_input_row(const ssb_program* prog) { return prog->prog.bytes_in_row; }
int32_t ssb_program_bytes_per_output_row(const ssb_program* prog) { return prog->prog.bytes_out_row; }

int ssb_program_run(ssb_program* prog, const ssb_column* inputs, int64_t rows,
                    const ssb_column* outputs, int64_t* d_out_rows) {
  return launch_program(prog, inputs, rows, outputs, d_out_rows);
}

int ssb_program_check_failure(ssb_program* prog) {
  ssb_ctx* ctx = prog->ctx;
  if (!prog->prog.has_signaling) return 0;
  SSB_CUDA(ctx, cudaMemcpyAsync(ctx->h_fail, ctx->d_fail, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  SSB_CUDA(ctx, cudaMemsetAsync(ctx->d_fail, 0, sizeof(int32_t), ctx->stream));
  SSB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (*ctx->h_fail) return fail(ctx, SSB_ERROR_EVALUATION_ERROR, "evaluation error (division by zero in a signaling expression)");
  return 0;
}

int ssb_program_run_sync(ssb_program* prog, const ssb_column* inputs, int64_t rows,
                         const ssb_column* outputs, int64_t* out_rows) {
  ssb_ctx* ctx = prog->ctx;
  if (int rc = launch_program(prog, inputs, rows, outputs, ctx->d_count)) return rc;
  SSB_CUDA(ctx, cudaMemcpyAsync(ctx->h_count, ctx->d_count, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
  SSB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (out_rows) *out_rows = *ctx->h_count;
  return ssb_program_check_failure(prog);
}

}  // extern "C"
