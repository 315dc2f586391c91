// jit_rows.h -- Filter -> Compute -> GroupAggregate as ONE kernel specialised for one plan.
//
// Compiled at run time (NVRTC -> sm_100a cubin, csrc/jit.cu) behind ssb_group_update_program: the
// reference evaluates this chain as GroupAggregateCursor pulling blocks from FilterCursor /
// ComputeCursor (cursor/core/aggregate_groups.cc:332-433, filter.cc:96-230, compute.cc:49-56),
// one column primitive per expression node and block. Here the bound expression program, the
// column types and the aggregate list are compile-time constants of the translation unit that
// includes this file (the generated prelude defines `Spec` and the three X-macro lists below), so
// the accumulator machine of ops.h unfolds into straight-line code on registers: the same alu()
// the interpreting kernels call gives the values, only the dispatch is gone.
//
//   SSB_JIT_INPUTS(X)   X(c, PHYS, NULLABLE)                       one line per input column
//   SSB_JIT_PROGRAM(X)  X(KIND, MOP, T1, T2, FLAGS, RHSN, A, B, IMMA, IMMB)   one line per instruction
//   SSB_JIT_AGGS(X)     X(a, FN, IN_PHYS, OUT_PHYS, OUT, PAD)      one line per aggregate (OUT < 0: COUNT(*))
//
// Rows are read straight from the columns in HBM (coalesced, R rows per thread in flight); rows
// failing the predicate are skipped; accumulation is the few-groups scheme of group.cu: per-thread
// accumulators of up to Spec::G CTA-local groups in shared memory, the global table as overflow,
// rows the table cannot place are deferred to the host's replay loop (feed_slice).
#ifndef SSB_CSRC_JIT_ROWS_H_
#define SSB_CSRC_JIT_ROWS_H_

#include "group_device.h"

namespace ssb {

enum { kJitMaxIn = 12 };

struct JitRun {
  const void* in_data[kJitMaxIn];
  const uint32_t* in_nulls[kJitMaxIn];
  int32_t* d_fail;
};

#if defined(__CUDACC_RTC__)

__host__ __device__ constexpr int jit_max1(int v) { return v < 1 ? 1 : v; }

struct JitRow {
  enum { R = Spec::R };
  u64 acc[R];
  uint32_t accn, pass, live, fail;
  u64 sv[jit_max1(Spec::N_SLOT)][R];
  uint32_t sn[jit_max1(Spec::N_SLOT)];
  u64 ov[jit_max1(Spec::N_OUT)][R];
  uint32_t on[jit_max1(Spec::N_OUT)];
};

// The input values of the R rows one thread evaluates per step; the next step's are loaded before this step's
// are evaluated, so every warp keeps its loads in flight while it computes.
struct JitIn {
  enum { R = Spec::R };
  long long rows[R];
  uint32_t live;
  u64 v[jit_max1(Spec::N_IN)][R];
  uint32_t n[jit_max1(Spec::N_IN)];
};

template <int C, int PHYS, int NULLABLE>
__device__ __forceinline__ void jit_load(JitIn& in, const JitRun& run) {
  constexpr int R = Spec::R;
  uint32_t nn = 0;
#pragma unroll
  for (int j = 0; j < R; ++j) {
    u64 v = 0;
    if (in.rows[j] >= 0) {
      bool isn = false;
      if (NULLABLE) isn = bit_at(run.in_nulls[C], in.rows[j]);
      if (isn) nn |= 1u << j;
      else v = load_raw(run.in_data[C], PHYS, in.rows[j]);
    }
    in.v[C][j] = v;
  }
  in.n[C] = nn;
}

__device__ __forceinline__ void jit_fetch(JitIn& in, const GroupParams& p, const JitRun& run, long long base) {
  constexpr int R = Spec::R, T = Spec::T;
  in.live = 0;
#pragma unroll
  for (int j = 0; j < R; ++j) {
    const long long i = base + static_cast<long long>(j) * T;
    in.rows[j] = i < p.rows ? (p.row_index ? p.row_index[i] : i) : -1;
    if (in.rows[j] >= 0) in.live |= 1u << j;
  }
#define SSB_JIT_X_LOAD(C, PHYS, NULLABLE) jit_load<C, PHYS, NULLABLE>(in, run);
  SSB_JIT_INPUTS(SSB_JIT_X_LOAD)
#undef SSB_JIT_X_LOAD
}

// One instruction of the program; every field is a template constant, so alu() folds to the one operation.
template <int KIND, int MOP, int T1, int T2, int FLAGS, int RHSN, int A, int B, unsigned long long IMMA, unsigned long long IMMB>
__device__ __forceinline__ void jit_step(JitRow& s) {
  constexpr int R = Spec::R;
  constexpr uint32_t all = (1u << R) - 1u;
  u64 r1[R], r2[R];
  uint32_t n1 = 0, n2 = 0;
#pragma unroll
  for (int j = 0; j < R; ++j) { r1[j] = 0; r2[j] = 0; }
  if (KIND == K_LOAD || KIND == K_ALU2 || KIND == K_ALU3) {
    if (FLAGS & F_RHS_IMM) {
#pragma unroll
      for (int j = 0; j < R; ++j) r1[j] = IMMA;
      n1 = (FLAGS & F_RHS_NULLK) ? all : 0u;
    } else {
#pragma unroll
      for (int j = 0; j < R; ++j) r1[j] = s.sv[A][j];
      n1 = s.sn[A];
    }
  }
  if (KIND == K_ALU3) {
    if (FLAGS & F_RHS2_IMM) {
#pragma unroll
      for (int j = 0; j < R; ++j) r2[j] = IMMB;
      n2 = (RHSN & 4) ? all : 0u;
    } else {
#pragma unroll
      for (int j = 0; j < R; ++j) r2[j] = s.sv[B][j];
      n2 = s.sn[B];
    }
  }
  if (KIND == K_LOAD) {
#pragma unroll
    for (int j = 0; j < R; ++j) s.acc[j] = r1[j];
    s.accn = n1;
  } else if (KIND == K_STORE) {
#pragma unroll
    for (int j = 0; j < R; ++j) s.sv[A][j] = s.acc[j];
    s.sn[A] = s.accn;
  } else if (KIND == K_ALU1 || KIND == K_ALU2 || KIND == K_ALU3) {
    Insn in;
    in.kind = KIND; in.mop = MOP; in.t = T1; in.t2 = T2; in.flags = FLAGS; in.rhs_nullable = RHSN; in.rw = 0;
    in.a = A; in.b = B; in.code = 0; in.pad2 = 0; in.off_a = 0; in.off_b = 0;
    alu<R>(in, s.acc, s.accn, r1, n1, r2, n2, s.live, s.fail);
  } else if (KIND == K_PRED) {
    uint32_t t = 0;
#pragma unroll
    for (int j = 0; j < R; ++j) t |= static_cast<uint32_t>(s.acc[j] & 1u) << j;
    s.pass = t & ~s.accn & s.live;
  } else if (KIND == K_OUT) {
#pragma unroll
    for (int j = 0; j < R; ++j) s.ov[A][j] = s.acc[j];
    s.on[A] = s.accn;
  }
}

// Cold paths, kept out of line so that their registers do not count against the row loop.
__device__ __noinline__ void jit_overflow(const GroupParams& p, int a, long long slot, unsigned long long v) {
  const AggDev& ag = p.agg[a];
  apply(ag, slot, v, 1ull);
  if (ag.seen != nullptr) ag.seen[slot] = 1u;
}
struct JitLocal {   // Spec::G == 0 (many groups: every row goes to the global table) keeps one unused entry
  unsigned long long l_key[jit_max1(Spec::G)][jit_max1(Spec::NK)];
  unsigned int l_fp[jit_max1(Spec::G)];   // 32-bit fingerprints: a hit is confirmed against l_key / l_knull
  unsigned int l_knull[jit_max1(Spec::G)];
  unsigned int l_slot[jit_max1(Spec::G)];
  unsigned int l_ready;
};
// The row's group is not among the CTA-local entries this thread knows: finds (or inserts) its slot in the global
// table and claims a local entry for it. Returns the local entry or -1 (*slot < 0: the row must be deferred).
struct JitKey { unsigned long long v[jit_max1(Spec::NK)]; };   // by value: the key stays in registers at the call
// Returns the slot in the high word ((slot + 1) << 8, 0 = defer the row) and the local entry + 1 in the low byte (0 = none).
__device__ __noinline__ unsigned long long jit_claim(const GroupParams& p, JitLocal& L, const JitKey key, unsigned int knull,
                                                     unsigned int fp) {
  constexpr int G = Spec::G, NK = Spec::NK;
  const unsigned long long* kv = key.v;
  const long long slot = p.packed ? find_slot_packed_kv(p, (knull & 1u) != 0, kv[0]) : find_slot_generic_kv(p, kv, knull);
  if (slot < 0) return 0ull;
  const unsigned int want = static_cast<unsigned int>(slot) + 1u;
  int g = -1;
  for (int e = 0; e < G && g < 0; ++e) {
    const unsigned int old = atomicCAS(&L.l_slot[e], 0u, want);
    if (old == 0u) {
      for (int c = 0; c < NK; ++c) L.l_key[e][c] = kv[c];
      L.l_knull[e] = knull;
      L.l_fp[e] = fp;
      __threadfence_block();
      atomicOr(&L.l_ready, 1u << e);
      g = e;
    } else if (old == want) {
      g = e;
    }
  }
  return (static_cast<unsigned long long>(slot + 1) << 8) | static_cast<unsigned long long>(g + 1);
}

// One aggregate of one row: into the thread's accumulator of local group g, or (g < 0) the global table.
// Returns the aggregate's bit when the local accumulator took a value (the "seen" mask of the flush).
template <int AI, int FN, int IN_PHYS, int OUT_PHYS, int OUT, int PAD>
__device__ __forceinline__ uint32_t jit_accumulate(const GroupParams& p, const JitRow& s, int j, int g, long long slot,
                                                   unsigned long long* t_acc, int tid) {
  constexpr int T = Spec::T;
  unsigned long long v = 0;
  if (OUT >= 0) {
    if ((s.on[OUT < 0 ? 0 : OUT] >> j) & 1u) return 0u;   // NULL input: no contribution
    v = s.ov[OUT < 0 ? 0 : OUT][j];
    if (FN != SSB_AGG_COUNT && IN_PHYS != OUT_PHYS) v = convert_value(v, IN_PHYS, OUT_PHYS);
  }
  if (g < 0) {   // more groups than local entries in this CTA: straight to the global table
    jit_overflow(p, AI, slot, v);
    return 0u;
  }
  unsigned long long* accp = &t_acc[(g * Spec::A + AI) * T + tid];
  if (PAD == TA_COUNT) { *accp += 1ull; return 0u; }
  if (PAD == TA_SUM_F64) { *accp = Codec<double>::enc(Codec<double>::dec(*accp) + Codec<double>::dec(v)); }
  else if (PAD == TA_SUM_U64) { *accp += v; }
  else {
    AggDev ag;
    ag.fn = FN; ag.out_phys = OUT_PHYS;
    *accp = combine(ag, *accp, v);
  }
  return 1u << AI;
}

// Spec::G == 0: one aggregate of one row straight into the global table (atomics), the function and type folded.
template <int AI, int FN, int IN_PHYS, int OUT_PHYS, int OUT>
__device__ __forceinline__ void jit_apply_direct(const GroupParams& p, const JitRow& s, int j, long long slot) {
  unsigned long long v = 0;
  if (OUT >= 0) {
    if ((s.on[OUT < 0 ? 0 : OUT] >> j) & 1u) return;   // NULL input: no contribution
    v = s.ov[OUT < 0 ? 0 : OUT][j];
    if (FN != SSB_AGG_COUNT && IN_PHYS != OUT_PHYS) v = convert_value(v, IN_PHYS, OUT_PHYS);
  }
  AggDev ag = p.agg[AI];
  ag.fn = FN;
  ag.out_phys = OUT_PHYS;
  apply(ag, slot, v, 1ull);
  if (ag.seen != nullptr) ag.seen[slot] = 1u;
}

extern "C" __global__ void __launch_bounds__(Spec::T, Spec::MIN_CTAS) ssb_jit_rows(const __grid_constant__ GroupParams p,
                                                                                    const __grid_constant__ JitRun run) {
  constexpr int T = Spec::T, R = Spec::R, G = Spec::G, A = Spec::A, NK = Spec::NK;
  extern __shared__ unsigned long long dyn[];
  unsigned long long* t_acc = dyn;                                       // [G * A][T]
  unsigned int* t_seen = reinterpret_cast<unsigned int*>(t_acc + G * A * T);   // [G][T]
  __shared__ JitLocal L;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < G * A * T; i += T) t_acc[i] = identity_dev(p.agg[(i / T) % A]);
  for (int i = tid; i < G * T; i += T) t_seen[i] = 0u;
  if (tid < G) L.l_slot[tid] = 0u;
  if (tid == 0) L.l_ready = 0u;
  __syncthreads();
  unsigned int my_ready = 0;
  unsigned int my_fp[jit_max1(G)];
#pragma unroll
  for (int e = 0; e < jit_max1(G); ++e) my_fp[e] = 0;
  uint32_t fail = 0;
  const long long stride = static_cast<long long>(gridDim.x) * T * R;
  // Spec::PREFETCH = how many steps ahead the inputs are loaded (0: this step's loads are issued at the end of the
  // previous step, nothing overlaps the evaluation; 1 / 2: one / two steps' loads stay in flight under it)
  JitIn next, next2;
  long long base = static_cast<long long>(blockIdx.x) * T * R + tid;
  if (base < p.rows) jit_fetch(next, p, run, base);
  if (Spec::PREFETCH >= 2 && base + stride < p.rows) jit_fetch(next2, p, run, base + stride);
  for (; base < p.rows; base += stride) {
    JitRow s;
    long long rows_[R];
    s.live = next.live;
#pragma unroll
    for (int j = 0; j < R; ++j) rows_[j] = next.rows[j];
#pragma unroll
    for (int c = 0; c < Spec::N_IN; ++c) {
      s.sn[c] = next.n[c];
#pragma unroll
      for (int j = 0; j < R; ++j) s.sv[c][j] = next.v[c][j];
    }
    if (Spec::PREFETCH >= 2) {
      next = next2;
      if (base + 2 * stride < p.rows) jit_fetch(next2, p, run, base + 2 * stride);
    } else if (Spec::PREFETCH == 1) {
      if (base + stride < p.rows) jit_fetch(next, p, run, base + stride);
    }
#pragma unroll
    for (int c = Spec::N_IN; c < jit_max1(Spec::N_SLOT); ++c) {
      s.sn[c] = 0;
#pragma unroll
      for (int j = 0; j < R; ++j) s.sv[c][j] = 0;
    }
#pragma unroll
    for (int o = 0; o < jit_max1(Spec::N_OUT); ++o) {
      s.on[o] = 0;
#pragma unroll
      for (int j = 0; j < R; ++j) s.ov[o][j] = 0;
    }
    s.accn = 0; s.pass = s.live; s.fail = 0;
#pragma unroll
    for (int j = 0; j < R; ++j) s.acc[j] = 0;
#define SSB_JIT_X_STEP(KIND, MOP, T1, T2, FLAGS, RHSN, A_, B_, IMMA, IMMB) jit_step<KIND, MOP, T1, T2, FLAGS, RHSN, A_, B_, IMMA, IMMB>(s);
    SSB_JIT_PROGRAM(SSB_JIT_X_STEP)
#undef SSB_JIT_X_STEP
    fail |= s.fail;
    const unsigned int ready_now = *reinterpret_cast<volatile unsigned int*>(&L.l_ready);
    if (ready_now != my_ready) {
      my_ready = ready_now;
#pragma unroll
      for (int e = 0; e < G; ++e) if ((my_ready >> e) & 1u) my_fp[e] = *reinterpret_cast<volatile unsigned int*>(&L.l_fp[e]);
    }
#pragma unroll
    for (int j = 0; j < R; ++j) {
      if (!((s.pass >> j) & 1u)) continue;
      JitKey key;
      unsigned long long (&kv)[jit_max1(NK)] = key.v;
      unsigned int knull = 0;
      kv[0] = 0;
#pragma unroll
      for (int c = 0; c < NK; ++c) {
        kv[c] = 0;
        if ((s.on[c] >> j) & 1u) knull |= 1u << c; else kv[c] = s.ov[c][j];
      }
      if (G == 0) {   // many groups: no CTA-local entries, the row's slot in the global table takes the values
        const long long direct = p.packed ? find_slot_packed_kv(p, (knull & 1u) != 0, kv[0]) : find_slot_generic_kv(p, kv, knull);
        if (direct < 0) {
          const unsigned long long d = atomicAdd(p.n_deferred, 1ull);
          p.deferred[d] = rows_[j];
          continue;
        }
#define SSB_JIT_X_DIRECT(AI, FN, IN_PHYS, OUT_PHYS, OUT, PAD) jit_apply_direct<AI, FN, IN_PHYS, OUT_PHYS, OUT>(p, s, j, direct);
        SSB_JIT_AGGS(SSB_JIT_X_DIRECT)
#undef SSB_JIT_X_DIRECT
        continue;
      }
      // a cheap 32-bit fingerprint (the keys are compared in full after a hit): one multiply-add per key half
      unsigned int fp = 0x9E3779B9u + knull;
#pragma unroll
      for (int c = 0; c < NK; ++c) {
        fp = (fp ^ static_cast<unsigned int>(kv[c])) * 0x85EBCA77u + static_cast<unsigned int>(kv[c] >> 32) * 0xC2B2AE3Du;
      }
      int g = -1;
#pragma unroll
      for (int e = 0; e < G; ++e) if (my_fp[e] == fp) g = e;
      if (g >= 0 && !((my_ready >> g) & 1u)) g = -1;   // an entry this thread has not seen published yet
      if (g >= 0) {
        bool same = L.l_knull[g] == knull;
#pragma unroll
        for (int c = 0; c < NK; ++c) same = same && L.l_key[g][c] == kv[c];
        if (!same) g = -1;
      }
      long long slot = -1;
      if (g < 0) {
        const unsigned long long claimed = jit_claim(p, L, key, knull, fp);
        slot = static_cast<long long>(claimed >> 8) - 1;
        g = static_cast<int>(claimed & 0xffull) - 1;
        if (slot < 0) {
          const unsigned long long d = atomicAdd(p.n_deferred, 1ull);
          p.deferred[d] = rows_[j];
          continue;
        }
      }
      uint32_t seen = 0;
#define SSB_JIT_X_AGG(AI, FN, IN_PHYS, OUT_PHYS, OUT, PAD) seen |= jit_accumulate<AI, FN, IN_PHYS, OUT_PHYS, OUT, PAD>(p, s, j, g, slot, t_acc, tid);
      SSB_JIT_AGGS(SSB_JIT_X_AGG)
#undef SSB_JIT_X_AGG
      if (seen) t_seen[g * T + tid] |= seen;
    }
    if (!Spec::PREFETCH && base + stride < p.rows) jit_fetch(next, p, run, base + stride);
  }
  if (fail && run.d_fail != nullptr) atomicOr(run.d_fail, 1);
  __syncthreads();
  for (int ga = warp; ga < G * A; ga += T / 32) {
    const int g = ga / A, a = ga - g * A;
    if (L.l_slot[g] == 0u) continue;
    const AggDev& ag = p.agg[a];
    unsigned long long acc2 = 0;
    bool has = false;
    for (int t = lane; t < T; t += 32) {
      const unsigned long long x = t_acc[ga * T + t];
      if (ag.fn == SSB_AGG_COUNT) { acc2 += x; }
      else if ((t_seen[g * T + t] >> a) & 1u) { acc2 = has ? combine(ag, acc2, x) : x; has = true; }
    }
    for (int d = 16; d > 0; d >>= 1) {
      const unsigned long long ov = __shfl_xor_sync(0xffffffffu, acc2, d);
      const bool oh = __shfl_xor_sync(0xffffffffu, has ? 1 : 0, d) != 0;
      if (ag.fn == SSB_AGG_COUNT) acc2 += ov;
      else if (oh) { acc2 = has ? combine(ag, acc2, ov) : ov; has = true; }
    }
    if (lane != 0) continue;
    const long long slot = static_cast<long long>(L.l_slot[g] - 1u);
    if (ag.fn == SSB_AGG_COUNT) { if (acc2) atomicAdd(&ag.acc[static_cast<unsigned long long>(slot) * ag.stride], acc2); continue; }
    if (!has) continue;
    if (ag.seen != nullptr) ag.seen[slot] = 1u;
    apply(ag, slot, acc2, 0ull);
  }
}

#endif  // __CUDACC_RTC__

}  // namespace ssb
#endif  // SSB_CSRC_JIT_ROWS_H_
