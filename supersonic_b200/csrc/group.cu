// group.cu -- GroupAggregate / ScalarAggregate on the GPU.
//
// Replaces the reference's chain (per 1024-row block, one thread)
//   RowHashSetImpl::InsertUnique  (cursor/infrastructure/row_hash_set.cc:458-518)
//   Aggregator::UpdateAggregations -> ColumnAggregatorImpl::UpdateAggregation
//                                    (cursor/core/aggregator.cc:206-221, column_aggregator.cc:108-226)
// by an open-addressed hash table in HBM (L2-resident for the headline sizes):
//   * slot claim by atomicCAS on a 64-bit key word (single-column keys) or on a state word
//     with the key columns stored beside it (multi-column keys)
//   * accumulators are 8-byte words per (slot, aggregate) updated with L2 atomics
//     (atomicAdd on u64 / double, atomicMin/Max, CAS loops for floating MIN/MAX)
//   * the common shapes have their own kernels: one packed key with plain aggregates (four rows
//     per thread in flight), a handful of groups (per-thread accumulators in shared memory, no
//     atomics in the row loop), up to 256 groups (CTA-private shared-memory table)
//   * rows that find no free slot within the probe limit are appended to a deferred list; the
//     host grows the table (re-inserting the old slots) and replays only those rows, so every
//     row is accumulated exactly once
// NULL is a group key of its own (row_hash_set.cc:81-90); an aggregate whose inputs were all
// NULL is NULL (column_aggregator.cc:108-125); COUNT never is.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <vector>

#include "comm.h"
#include "device_utils.h"
#include "common.h"
#include "group_device.h"
#include "jit.h"
#include "jit_rows.h"
#include "program.h"

namespace ssb {

__global__ void __launch_bounds__(256) group_update_kernel(const __grid_constant__ GroupParams p) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < p.rows; i += stride) {
    const long long row = p.row_index ? p.row_index[i] : i;
    const long long slot = p.packed ? find_slot_packed(p, row) : find_slot_generic(p, row);
    if (slot < 0) {
      const unsigned long long d = atomicAdd(p.n_deferred, 1ull);
      p.deferred[d] = row;
      continue;
    }
    for (int a = 0; a < p.n_aggs; ++a) {
      const AggDev& ag = p.agg[a];
      if (ag.in_phys >= 0 && bit_at(ag.in_nulls, row)) continue;
      if (ag.fn == SSB_AGG_FIRST || ag.fn == SSB_AGG_LAST) {
        // the first / last non-NULL input in row order (column_aggregator.cc:108-166 walks the
        // rows in order): the launch only elects the row, first_last_resolve_kernel fetches it
        if (p.merge) {   // partial tables hold one row per group: plain stores, `dst` rows come first
          const unsigned long long v = load_raw(ag.in_data, ag.in_phys, row);
          if (ag.fn == SSB_AGG_LAST || ag.seen[slot] == 0u) ag.acc[static_cast<unsigned long long>(slot) * ag.stride] = v;
          ag.seen[slot] = 1u;
        } else if (ag.fn == SSB_AGG_FIRST) {
          atomicMin(&ag.cand[slot], static_cast<unsigned long long>(row));
        } else {
          atomicMax(&ag.cand[slot], static_cast<unsigned long long>(row) + 1ull);
        }
        continue;
      }
      unsigned long long v = 0, cnt = 1;
      if (ag.in_phys >= 0) {
        v = load_raw(ag.in_data, ag.in_phys, row);
        if (ag.fn == SSB_AGG_COUNT) { cnt = p.merge ? v : 1ull; }
        else v = convert_value(v, ag.in_phys, ag.out_phys);
      }
      apply(ag, slot, v, cnt);
      if (ag.seen != nullptr) ag.seen[slot] = 1u;
    }
  }
}

// FIRST / LAST: after every launch, each slot takes the value of the row the launch elected
// (FIRST only while the group has no value yet: earlier launches hold earlier rows).
__global__ void first_last_resolve_kernel(const __grid_constant__ GroupParams p, unsigned long long total_slots) {
  const unsigned long long stride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
  for (unsigned long long s = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; s < total_slots; s += stride) {
    for (int a = 0; a < p.n_aggs; ++a) {
      const AggDev& ag = p.agg[a];
      if (ag.cand == nullptr) continue;
      const unsigned long long c = ag.cand[s];
      if (ag.fn == SSB_AGG_FIRST) {
        if (c == ~0ull) continue;
        ag.cand[s] = ~0ull;
        if (ag.seen[s] != 0u) continue;
        ag.acc[s * ag.stride] = convert_value(load_raw(ag.in_data, ag.in_phys, static_cast<long long>(c)), ag.in_phys, ag.out_phys);
        ag.seen[s] = 1u;
      } else {
        if (c == 0ull) continue;
        ag.cand[s] = 0ull;
        ag.acc[s * ag.stride] = convert_value(load_raw(ag.in_data, ag.in_phys, static_cast<long long>(c - 1ull)), ag.in_phys, ag.out_phys);
        ag.seen[s] = 1u;
      }
    }
  }
}

// ---- fast path: one 8-byte NOT NULL key, 8-byte NOT NULL inputs ------------------------------
// The kernel above interprets the aggregate list per row (about 370 instructions per row for
// the C3 shape: ncu smsp__inst_executed, profiles/r1b_summary.md) and is bound by instruction
// issue. This one serves the common shape -- single packed key, every aggregate COUNT or an
// 8-byte SUM/MIN/MAX whose input type equals its result type, no NULL bitmaps, no replay --
// with four rows per thread: the four key loads, then the four first probes, then the value
// loads and reductions are issued back to back, so the dependent L2 round trips of one row
// overlap those of the others. Same table, same hash, same claim protocol as above.
__device__ __forceinline__ long long probe_packed_from(const GroupParams& p, unsigned long long key, unsigned long long s,
                                                       unsigned long long cur) {
  const unsigned long long mask = p.capacity - 1;
  for (int probe = 0; probe < kProbeLimit; ++probe) {
    if (cur == key) return static_cast<long long>(s);
    if (cur == kEmptyKey) {
      const unsigned long long old = atomicCAS(&p.slot_key[s], kEmptyKey, key);
      if (old == kEmptyKey) { atomicAdd(p.n_groups, 1ull); return static_cast<long long>(s); }
      if (old == key) return static_cast<long long>(s);
    }
    s = (s + 1) & mask;
    cur = p.slot_key[s];
  }
  return -1;
}

// PLAIN: every column is 8 bytes wide without a NULL bitmap (plain 8-byte loads, no bit tests).
// Otherwise 4-byte columns and NULL bitmaps are handled in place: a NULL key takes the special
// slot, a NULL input does not contribute (the result is NULL while no input was seen).
template <bool PLAIN>
__global__ void __launch_bounds__(256, 4) group_update_fast_kernel(const __grid_constant__ GroupParams p) {
  constexpr int R = 4;
  constexpr int TILE = 256 * R;
  const unsigned long long* __restrict__ keys = static_cast<const unsigned long long*>(p.key_data[0]);
  const unsigned long long mask = p.capacity - 1;
  const long long tiles = (p.rows + TILE - 1) / TILE;
  for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
    const long long r0 = t * TILE + threadIdx.x;
    unsigned long long key[R], cur[R];
    long long slot[R];
    unsigned int knull = 0;   // bit j: the key of row j is NULL
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const long long row = r0 + j * 256;
      key[j] = kEmptyKey;
      if (row < p.rows) {
        if (PLAIN) key[j] = keys[row];
        else if (bit_at(p.key_nulls[0], row)) knull |= 1u << j;
        else key[j] = load_raw(p.key_data[0], p.key_phys[0], row);
      }
    }
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const unsigned long long s = mix64(key[j]) & mask;
      slot[j] = static_cast<long long>(s);
      cur[j] = p.slot_key[s];   // first probe of all four rows in flight together
    }
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const long long row = r0 + j * 256;
      if (row >= p.rows) { slot[j] = -2; continue; }
      if (key[j] == kEmptyKey) slot[j] = find_slot_packed(p, row);       // NULL key / the one key that needs the special slot
      else if (cur[j] != key[j]) slot[j] = probe_packed_from(p, key[j], static_cast<unsigned long long>(slot[j]), cur[j]);
      if (slot[j] == -1) {
        const unsigned long long d = atomicAdd(p.n_deferred, 1ull);
        p.deferred[d] = row;
      }
    }
    for (int a = 0; a < p.n_aggs; ++a) {
      const AggDev& ag = p.agg[a];
      unsigned int valid = 0;   // bit j: row j contributes to this aggregate
#pragma unroll
      for (int j = 0; j < R; ++j) {
        if (slot[j] >= 0 && (PLAIN || ag.in_phys < 0 || !bit_at(ag.in_nulls, r0 + j * 256))) valid |= 1u << j;
      }
      if (ag.fn == SSB_AGG_COUNT) {
#pragma unroll
        for (int j = 0; j < R; ++j) if ((valid >> j) & 1u) atomicAdd(&ag.acc[slot[j]], 1ull);
        continue;
      }
      unsigned long long v[R];
#pragma unroll
      for (int j = 0; j < R; ++j) {
        v[j] = 0ull;
        if ((valid >> j) & 1u) {
          v[j] = PLAIN ? static_cast<const unsigned long long*>(ag.in_data)[r0 + j * 256] : load_raw(ag.in_data, ag.in_phys, r0 + j * 256);
        }
      }
      if (ag.fn == SSB_AGG_SUM && ag.out_phys == T_F64) {
#pragma unroll
        for (int j = 0; j < R; ++j) if ((valid >> j) & 1u) atomicAdd(reinterpret_cast<double*>(&ag.acc[slot[j]]), Codec<double>::dec(v[j]));
      } else if (ag.fn == SSB_AGG_SUM && (ag.out_phys == T_I64 || ag.out_phys == T_U64)) {   // wrapping add
#pragma unroll
        for (int j = 0; j < R; ++j) if ((valid >> j) & 1u) atomicAdd(&ag.acc[slot[j]], v[j]);
      } else {
#pragma unroll
        for (int j = 0; j < R; ++j) if ((valid >> j) & 1u) apply(ag, slot[j], v[j], 1ull);
      }
      if (ag.seen != nullptr) {
#pragma unroll
        for (int j = 0; j < R; ++j) if ((valid >> j) & 1u) ag.seen[slot[j]] = 1u;
      }
    }
  }
}

// ---- few groups: CTA-private accumulators in shared memory ---------------------------------
// With few distinct keys every row of the table hits the same handful of L2 words; the kernel
// above would serialise on them. Here each CTA keeps its own small table in shared memory
// (keyed by the slot index of the global table, so keys of any shape work), accumulates with
// shared-memory atomics, and flushes one partial per (group, aggregate) at the end.
enum { kLocalSlots = 512 };

__global__ void __launch_bounds__(256) group_update_smem_kernel(const __grid_constant__ GroupParams p) {
  __shared__ unsigned int l_slot[kLocalSlots];                          // global slot + 1, 0 = empty
  __shared__ unsigned int l_seen[kLocalSlots];                          // bit a: aggregate a saw a value
  __shared__ unsigned long long l_acc[kLocalSlots * kLocalMaxAggs];
  for (int i = threadIdx.x; i < kLocalSlots; i += blockDim.x) { l_slot[i] = 0u; l_seen[i] = 0u; }
  for (int i = threadIdx.x; i < kLocalSlots * kLocalMaxAggs; i += blockDim.x) {
    const int a = i % kLocalMaxAggs;
    unsigned long long id = 0;
    if (a < p.n_aggs) {
      const AggDev& ag = p.agg[a];
      if (ag.fn == SSB_AGG_MIN) id = ag.out_phys == T_F64 ? Codec<double>::enc(__longlong_as_double(0x7ff0000000000000LL))
                                   : ag.out_phys == T_F32 ? Codec<float>::enc(__uint_as_float(0x7f800000u))
                                   : (ag.out_phys == T_I64 || ag.out_phys == T_I32) ? static_cast<unsigned long long>(INT64_MAX) : ~0ull;
      if (ag.fn == SSB_AGG_MAX) id = ag.out_phys == T_F64 ? Codec<double>::enc(__longlong_as_double(0xfff0000000000000LL))
                                   : ag.out_phys == T_F32 ? Codec<float>::enc(__uint_as_float(0xff800000u))
                                   : (ag.out_phys == T_I64 || ag.out_phys == T_I32) ? static_cast<unsigned long long>(INT64_MIN) : 0ull;
    }
    l_acc[i] = id;
  }
  __syncthreads();
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < p.rows; i += stride) {
    const long long row = p.row_index ? p.row_index[i] : i;
    const long long slot = p.packed ? find_slot_packed(p, row) : find_slot_generic(p, row);
    if (slot < 0) {
      const unsigned long long d = atomicAdd(p.n_deferred, 1ull);
      p.deferred[d] = row;
      continue;
    }
    // local entry of this global slot
    const unsigned int want = static_cast<unsigned int>(slot) + 1u;
    unsigned int h = (static_cast<unsigned int>(slot) * 2654435761u) & (kLocalSlots - 1);
    int e = -1;
    for (int probe = 0; probe < 16; ++probe) {
      unsigned int cur = l_slot[h];
      if (cur == 0u) cur = atomicCAS(&l_slot[h], 0u, want);
      if (cur == 0u || cur == want) { e = static_cast<int>(h); break; }
      h = (h + 1) & (kLocalSlots - 1);
    }
    for (int a = 0; a < p.n_aggs; ++a) {
      const AggDev& ag = p.agg[a];
      if (ag.in_phys >= 0 && bit_at(ag.in_nulls, row)) continue;
      unsigned long long v = 0, cnt = 1;
      if (ag.in_phys >= 0) {
        v = load_raw(ag.in_data, ag.in_phys, row);
        if (ag.fn == SSB_AGG_COUNT) { cnt = p.merge ? v : 1ull; }
        else v = convert_value(v, ag.in_phys, ag.out_phys);
      }
      if (e >= 0) {
        AggDev local = ag;
        local.acc = &l_acc[a];
        local.stride = kLocalMaxAggs;
        apply(local, e, v, cnt);
        if (ag.seen != nullptr) atomicOr(&l_seen[e], 1u << a);
      } else {   // the CTA's table is full: go to the global table directly
        apply(ag, slot, v, cnt);
        if (ag.seen != nullptr) ag.seen[slot] = 1u;
      }
    }
  }
  __syncthreads();
  // flush: one atomic per (group, aggregate) and CTA
  for (int e = threadIdx.x; e < kLocalSlots; e += blockDim.x) {
    const unsigned int ls = l_slot[e];
    if (ls == 0u) continue;
    const long long slot = static_cast<long long>(ls - 1u);
    for (int a = 0; a < p.n_aggs; ++a) {
      const AggDev& ag = p.agg[a];
      const unsigned long long v = l_acc[e * kLocalMaxAggs + a];
      unsigned long long* dst = &ag.acc[static_cast<unsigned long long>(slot) * ag.stride];
      if (ag.fn == SSB_AGG_COUNT) { if (v) atomicAdd(dst, v); continue; }
      if (ag.seen != nullptr) {
        if (!((l_seen[e] >> a) & 1u)) continue;   // only NULL inputs in this CTA
        ag.seen[slot] = 1u;
      }
      if (ag.fn == SSB_AGG_SUM && ag.out_phys != T_F64 && ag.out_phys != T_F32) atomicAdd(dst, v);   // wrapped 64-bit partial
      else apply(ag, slot, v, 0ull);
    }
  }
}

// ---- a handful of groups: per-thread accumulators, no atomics in the row loop ------------------
// With <= 8 groups (the Q1 shape: 6) every row of a warp hits the same few accumulators, and even
// shared-memory atomics serialise (a DOUBLE add is a CAS loop there). Here every thread owns a
// private accumulator per (group, aggregate) in shared memory ([group][aggregate][thread]: bank =
// thread, no conflicts), a row costs one load-combine-store per aggregate, and the threads'
// partials are tree-combined once at the end: one global update per (group, aggregate) and CTA.
// The CTA learns its groups on the fly: the first row of a group goes through the global table
// (find_slot_*), claims the next local entry by CAS on the slot id and publishes the key values;
// later rows find the group by comparing their key with the published ones (broadcast reads).

// AggDev::pad carries a pre-decoded accumulate code for this kernel (set by the host).

// FAST: every key and input column is 8 bytes wide without a NULL bitmap, no merge, no replay
// list -- the loads are plain 8-byte loads and no per-aggregate NULL bookkeeping is needed.
// K2: at most two key columns (the unrolled key loops shrink from eight to two iterations).
template <bool FAST, bool K2>
__global__ void __launch_bounds__(kTinyThreads) group_update_tiny_kernel(const __grid_constant__ GroupParams p) {
  constexpr int T = kTinyThreads;
  constexpr int KMAX = K2 ? 2 : kMaxKeys;
  extern __shared__ unsigned long long t_acc[];                  // [kTinyGroups * n_aggs][T]
  __shared__ unsigned int t_seen[kTinyGroups][T];                // bit a: this thread saw a value of aggregate a
  __shared__ unsigned long long l_key[kTinyGroups][kMaxKeys];
  __shared__ unsigned long long l_fp[kTinyGroups];               // fingerprint of the entry's key
  __shared__ unsigned int l_knull[kTinyGroups];
  __shared__ unsigned int l_slot[kTinyGroups];                   // global slot + 1, 0 = free; claimed in order
  __shared__ unsigned int l_ready;                               // bit e: key values of entry e are published
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int A = p.n_aggs;
  const int NK = p.n_keys;
  for (int i = tid; i < kTinyGroups * A * T; i += T) t_acc[i] = identity_dev(p.agg[(i / T) % A]);
  for (int i = tid; i < kTinyGroups * T; i += T) (&t_seen[0][0])[i] = 0u;
  if (tid < kTinyGroups) l_slot[tid] = 0u;
  if (tid == 0) l_ready = 0u;
  __syncthreads();
  // register copy of the published fingerprints, refreshed when the ready mask changes
  unsigned int my_ready = 0;
  unsigned long long my_fp[kTinyGroups];
#pragma unroll
  for (int e = 0; e < kTinyGroups; ++e) my_fp[e] = 0;
  // R rows per thread and iteration; all key and value loads of the R rows are issued before the
  // first one is used, so a thread pays one HBM round trip per iteration, not two per row
  constexpr int R = FAST ? 4 : 2;
  const long long stride = static_cast<long long>(gridDim.x) * T * R;
  for (long long base = static_cast<long long>(blockIdx.x) * T * R + tid; base < p.rows; base += stride) {
    long long rows_[R];
    unsigned long long kv[R][KMAX], vv[R][kLocalMaxAggs];
    unsigned int knull[R], vnull[R];
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const long long i = base + static_cast<long long>(j) * T;
      rows_[j] = i < p.rows ? ((!FAST && p.row_index) ? p.row_index[i] : i) : -1;
    }
#pragma unroll
    for (int j = 0; j < R; ++j) {
      knull[j] = 0;
      vnull[j] = 0;
#pragma unroll
      for (int c = 0; c < KMAX; ++c) {
        kv[j][c] = 0;
        if (c < NK && rows_[j] >= 0) {
          if (FAST) kv[j][c] = static_cast<const unsigned long long*>(p.key_data[c])[rows_[j]];
          else if (bit_at(p.key_nulls[c], rows_[j])) knull[j] |= 1u << c;
          else kv[j][c] = load_raw(p.key_data[c], p.key_phys[c], rows_[j]);
        }
      }
#pragma unroll
      for (int a = 0; a < kLocalMaxAggs; ++a) {
        vv[j][a] = 0;
        if (a < A && p.agg[a].in_phys >= 0 && rows_[j] >= 0) {
          if (FAST) vv[j][a] = static_cast<const unsigned long long*>(p.agg[a].in_data)[rows_[j]];
          else if (bit_at(p.agg[a].in_nulls, rows_[j])) vnull[j] |= 1u << a;
          else vv[j][a] = load_raw(p.agg[a].in_data, p.agg[a].in_phys, rows_[j]);
        }
      }
    }
    const unsigned int ready_now = *reinterpret_cast<volatile unsigned int*>(&l_ready);
    if (ready_now != my_ready) {
      my_ready = ready_now;
#pragma unroll
      for (int e = 0; e < kTinyGroups; ++e) if ((my_ready >> e) & 1u) my_fp[e] = *reinterpret_cast<volatile unsigned long long*>(&l_fp[e]);
    }
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const long long row = rows_[j];
      if (row < 0) continue;
      unsigned long long fp;
      if (FAST && K2) {
        // NOT NULL keys, at most two columns: one key is its own (exact) fingerprint
        fp = NK > 1 ? kv[j][0] * 0x9E3779B97F4A7C15ull + kv[j][1] : kv[j][0];
      } else {
        fp = 0x9E3779B97F4A7C15ull + knull[j];
#pragma unroll
        for (int c = 0; c < KMAX; ++c) if (c < NK) fp = (fp ^ kv[j][c]) * 0xff51afd7ed558ccdULL + c;
      }
      int g = -1;
#pragma unroll
      for (int e = 0; e < kTinyGroups; ++e) if (((my_ready >> e) & 1u) && my_fp[e] == fp) g = e;
      if (g >= 0 && !(FAST && K2 && NK <= 1)) {
        // equal fingerprints: confirm on the key values (a different key restarts below)
        bool same = l_knull[g] == knull[j];
#pragma unroll
        for (int c = 0; c < KMAX; ++c) if (c < NK) same = same && l_key[g][c] == kv[j][c];
        if (!same) g = -1;
      }
      long long slot = -1;
      if (g < 0) {
        // first sight of this group in the CTA (or its entry is still being published)
        slot = p.packed ? find_slot_packed(p, row) : find_slot_generic(p, row);
        if (slot < 0) {
          const unsigned long long d = atomicAdd(p.n_deferred, 1ull);
          p.deferred[d] = row;
          continue;
        }
        const unsigned int want = static_cast<unsigned int>(slot) + 1u;
        for (int e = 0; e < kTinyGroups && g < 0; ++e) {
          const unsigned int old = atomicCAS(&l_slot[e], 0u, want);
          if (old == 0u) {
#pragma unroll
            for (int c = 0; c < KMAX; ++c) if (c < NK) l_key[e][c] = kv[j][c];
            l_knull[e] = knull[j];
            l_fp[e] = fp;
            __threadfence_block();
            atomicOr(&l_ready, 1u << e);
            g = e;
          } else if (old == want) {
            g = e;
          }
        }
      }
      if (g >= 0 && FAST) t_seen[g][tid] = 0xffffffffu;   // NOT NULL inputs: every aggregate saw this row
      unsigned long long* const acc_base = t_acc + (g < 0 ? 0 : g) * A * T + tid;
#pragma unroll
      for (int a = 0; a < kLocalMaxAggs; ++a) {
        if (a >= A) break;
        const AggDev& ag = p.agg[a];
        if (!FAST && ((vnull[j] >> a) & 1u)) continue;
        unsigned long long v = vv[j][a], cnt = 1;
        if (!FAST && ag.in_phys >= 0) {
          if (ag.fn == SSB_AGG_COUNT) { cnt = p.merge ? v : 1ull; }
          else if (ag.in_phys != ag.out_phys) v = convert_value(v, ag.in_phys, ag.out_phys);
        }
        if (g < 0) {   // more groups than local entries in this CTA: straight to the global table
          apply(ag, slot, v, cnt);
          if (ag.seen != nullptr) ag.seen[slot] = 1u;
          continue;
        }
        unsigned long long* acc = acc_base + a * T;
        switch (ag.pad) {
          case TA_COUNT: *acc += cnt; break;
          case TA_SUM_F64: *acc = Codec<double>::enc(Codec<double>::dec(*acc) + Codec<double>::dec(v)); break;
          case TA_SUM_U64: *acc += v; break;
          default: *acc = combine(ag, *acc, v); break;
        }
        if (!FAST && ag.pad != TA_COUNT) t_seen[g][tid] |= 1u << a;
      }
    }
  }
  __syncthreads();
  // flush: warp w combines the T partials of the (group, aggregate) pairs w, w + T/32, ...
  for (int ga = warp; ga < kTinyGroups * A; ga += T / 32) {
    const int g = ga / A, a = ga - g * A;
    if (l_slot[g] == 0u) continue;
    const AggDev& ag = p.agg[a];
    unsigned long long acc = 0;
    bool has = false;
    for (int t = lane; t < T; t += 32) {
      const unsigned long long x = t_acc[ga * T + t];
      if (ag.fn == SSB_AGG_COUNT) { acc += x; }
      else if ((t_seen[g][t] >> a) & 1u) { acc = has ? combine(ag, acc, x) : x; has = true; }
    }
    for (int d = 16; d > 0; d >>= 1) {
      const unsigned long long ov = __shfl_xor_sync(0xffffffffu, acc, d);
      const bool oh = __shfl_xor_sync(0xffffffffu, has ? 1 : 0, d) != 0;
      if (ag.fn == SSB_AGG_COUNT) acc += ov;
      else if (oh) { acc = has ? combine(ag, acc, ov) : ov; has = true; }
    }
    if (lane != 0) continue;
    const long long slot = static_cast<long long>(l_slot[g] - 1u);
    if (ag.fn == SSB_AGG_COUNT) { if (acc) atomicAdd(&ag.acc[static_cast<unsigned long long>(slot) * ag.stride], acc); continue; }
    if (!has) continue;
    if (ag.seen != nullptr) ag.seen[slot] = 1u;
    apply(ag, slot, acc, 0ull);
  }
}

// ---- fused Filter -> Compute -> GroupAggregate ------------------------------------------------
// The row-wise child of a GroupAggregate (cursor/core/aggregate_groups.cc:332-433 pulling blocks
// from ComputeCursor / FilterCursor, compute.cc:49-56, filter.cc:96-230) is evaluated inside the
// aggregation kernel: nothing is materialised between the two operators, the plan reads its
// input columns once (the Q1 shape: 56 B per row instead of 56 + 55 written + 55 read again).
// Every thread runs the expression program of its own R rows on the accumulator machine of
// ops.h (the program is warp-uniform, so the dispatch does not diverge); operand slots, i.e.
// the row's input values and temporaries, are thread-private words in shared memory
// ([slot][row][thread]: conflict-free). The outputs of the program are the group-by key columns
// followed by the aggregate inputs; rows failing the predicate are skipped; accumulation is the
// tiny-group scheme above (private accumulators per thread, global table as overflow).
enum { kRowMaxIn = 12, kRowMaxOut = 15, kRowMaxInsn = 64, kRowMaxImm = 16 };

struct RowProg {
  int32_t n_insn, n_in, n_tmp, n_out;
  Insn insn[kRowMaxInsn];
  unsigned long long imm[kRowMaxImm];
  const void* in_data[kRowMaxIn];
  const uint32_t* in_nulls[kRowMaxIn];
  int32_t in_phys[kRowMaxIn];
  int32_t agg_out[kMaxAggs];      // program output feeding aggregate a, -1: COUNT(*)
  int32_t has_pred;
  int32_t* d_fail;
};

__global__ void __launch_bounds__(kTinyThreads) group_update_rows_kernel(const __grid_constant__ GroupParams p,
                                                                          const __grid_constant__ RowProg rp) {
  constexpr int T = kTinyThreads;
  constexpr int R = 2;
  extern __shared__ unsigned long long dyn[];
  const int A = p.n_aggs;
  const int NK = p.n_keys;
  const int n_slots = rp.n_in + rp.n_tmp;
  unsigned long long* t_acc = dyn;                                         // [kTinyGroups * A][T]
  unsigned long long* s_val = t_acc + kTinyGroups * A * T;                  // [n_slots][R][T]
  unsigned long long* s_out = s_val + n_slots * R * T;                      // [n_out][R][T]
  unsigned int* s_nul = reinterpret_cast<unsigned int*>(s_out + rp.n_out * R * T);   // [n_slots][T], bit r
  unsigned int* s_outn = s_nul + n_slots * T;                               // [n_out][T]
  __shared__ unsigned int t_seen[kTinyGroups][T];
  __shared__ unsigned long long l_key[kTinyGroups][kMaxKeys];
  __shared__ unsigned long long l_fp[kTinyGroups];
  __shared__ unsigned int l_knull[kTinyGroups];
  __shared__ unsigned int l_slot[kTinyGroups];
  __shared__ unsigned int l_ready;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < kTinyGroups * A * T; i += T) t_acc[i] = identity_dev(p.agg[(i / T) % A]);
  for (int i = tid; i < kTinyGroups * T; i += T) (&t_seen[0][0])[i] = 0u;
  if (tid < kTinyGroups) l_slot[tid] = 0u;
  if (tid == 0) l_ready = 0u;
  __syncthreads();
  unsigned int my_ready = 0;
  unsigned long long my_fp[kTinyGroups];
#pragma unroll
  for (int e = 0; e < kTinyGroups; ++e) my_fp[e] = 0;
  uint32_t fail = 0;
  const long long stride = static_cast<long long>(gridDim.x) * T * R;
  for (long long base = static_cast<long long>(blockIdx.x) * T * R + tid; base < p.rows; base += stride) {
    long long rows_[R];
    uint32_t live = 0;
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const long long i = base + static_cast<long long>(j) * T;
      rows_[j] = i < p.rows ? (p.row_index ? p.row_index[i] : i) : -1;
      if (rows_[j] >= 0) live |= 1u << j;
    }
    // ---- all input loads of the R rows first, then into the thread's operand slots
    {
      unsigned long long iv[kRowMaxIn][R];
      uint32_t inul[kRowMaxIn];
#pragma unroll
      for (int c = 0; c < kRowMaxIn; ++c) {
        inul[c] = 0;
#pragma unroll
        for (int j = 0; j < R; ++j) {
          iv[c][j] = 0;
          if (c < rp.n_in && rows_[j] >= 0) {
            if (bit_at(rp.in_nulls[c], rows_[j])) inul[c] |= 1u << j; else iv[c][j] = load_raw(rp.in_data[c], rp.in_phys[c], rows_[j]);
          }
        }
      }
#pragma unroll
      for (int c = 0; c < kRowMaxIn; ++c) {
        if (c < rp.n_in) {
#pragma unroll
          for (int j = 0; j < R; ++j) s_val[(c * R + j) * T + tid] = iv[c][j];
          s_nul[c * T + tid] = inul[c];
        }
      }
    }
    // ---- the row program (K_LOAD / K_STORE / K_ALU* / K_PRED / K_OUT)
    u64 acc[R];
    uint32_t accn = 0, pass = live;
#pragma unroll
    for (int j = 0; j < R; ++j) acc[j] = 0;
    for (int pc = 0; pc < rp.n_insn; ++pc) {
      const Insn& in = rp.insn[pc];
      const uint32_t code = in.code;
      if (code != C_GENERIC) {
        // pre-decoded cases (operands cannot be NULL): the hot (op, type) pairs of ops.h
        if (code >= C_BIN_BASE && code < C_BIN_END) {
          const int rel = static_cast<int>(code) - C_BIN_BASE;
          const int type_base = rel / 28, op = (rel % 28) / 4;
          u64 y[R];
          if (rel & 1) {
#pragma unroll
            for (int j = 0; j < R; ++j) y[j] = rp.imm[in.a];
          } else {
#pragma unroll
            for (int j = 0; j < R; ++j) y[j] = s_val[(in.a * R + j) * T + tid];
          }
#define SSB_ROW_BIN(TYPE)                                                                   \
  _Pragma("unroll") for (int j = 0; j < R; ++j) {                                           \
    const TYPE a_ = Codec<TYPE>::dec(acc[j]), b_ = Codec<TYPE>::dec(y[j]);                  \
    u64 r_;                                                                                 \
    switch (op) {                                                                           \
      case B_ADD: r_ = Codec<TYPE>::enc(Arith<TYPE>::add(a_, b_)); break;                   \
      case B_SUB: r_ = Codec<TYPE>::enc(Arith<TYPE>::sub(a_, b_)); break;                   \
      case B_SUBR: r_ = Codec<TYPE>::enc(Arith<TYPE>::sub(b_, a_)); break;                  \
      case B_MUL: r_ = Codec<TYPE>::enc(Arith<TYPE>::mul(a_, b_)); break;                   \
      case B_LT: r_ = ((a_ < b_) != neg) ? 1u : 0u; break;                                  \
      case B_GT: r_ = ((b_ < a_) != neg) ? 1u : 0u; break;                                  \
      default: r_ = ((a_ == b_) != neg) ? 1u : 0u; break;                                   \
    }                                                                                       \
    acc[j] = r_;                                                                            \
  }
          const bool neg = (in.flags & F_NEGATE) != 0;
          if (type_base == 0) { SSB_ROW_BIN(int64_t) } else if (type_base == 1) { SSB_ROW_BIN(double) } else { SSB_ROW_BIN(int32_t) }
#undef SSB_ROW_BIN
          continue;
        }
        if (code == C_LOAD8 || code == C_LOAD4) {
#pragma unroll
          for (int j = 0; j < R; ++j) acc[j] = s_val[(in.a * R + j) * T + tid];
          accn = 0;
          continue;
        }
        if (code == C_LOADK) {
#pragma unroll
          for (int j = 0; j < R; ++j) acc[j] = rp.imm[in.a];
          accn = 0;
          continue;
        }
        if (code == C_OUT8 || code == C_OUT4) {
#pragma unroll
          for (int j = 0; j < R; ++j) s_out[(in.a * R + j) * T + tid] = acc[j];
          s_outn[in.a * T + tid] = 0u;
          continue;
        }
        if (code == C_PRED) {
          uint32_t t = 0;
#pragma unroll
          for (int j = 0; j < R; ++j) t |= static_cast<uint32_t>(acc[j] & 1u) << j;
          pass = t & ~accn & live;
          continue;
        }
        // C_AND3_S / C_OR3_S and anything else: the generic path below computes the same
      }
      u64 r1[R], r2[R];
      uint32_t n1 = 0, n2 = 0;
#pragma unroll
      for (int j = 0; j < R; ++j) { r1[j] = 0; r2[j] = 0; }
      if (in.kind == K_LOAD || in.kind == K_ALU2 || in.kind == K_ALU3) {
        if (in.flags & F_RHS_IMM) {
#pragma unroll
          for (int j = 0; j < R; ++j) r1[j] = rp.imm[in.a];
          n1 = (in.flags & F_RHS_NULLK) ? 3u : 0u;
        } else {
#pragma unroll
          for (int j = 0; j < R; ++j) r1[j] = s_val[(in.a * R + j) * T + tid];
          n1 = s_nul[in.a * T + tid];
        }
      }
      if (in.kind == K_ALU3) {
        if (in.flags & F_RHS2_IMM) {
#pragma unroll
          for (int j = 0; j < R; ++j) r2[j] = rp.imm[in.b];
          n2 = (in.rhs_nullable & 4) ? 3u : 0u;
        } else {
#pragma unroll
          for (int j = 0; j < R; ++j) r2[j] = s_val[(in.b * R + j) * T + tid];
          n2 = s_nul[in.b * T + tid];
        }
      }
      switch (in.kind) {
        case K_LOAD:
#pragma unroll
          for (int j = 0; j < R; ++j) acc[j] = r1[j];
          accn = n1;
          break;
        case K_STORE:
#pragma unroll
          for (int j = 0; j < R; ++j) s_val[(in.a * R + j) * T + tid] = acc[j];
          s_nul[in.a * T + tid] = accn;
          break;
        case K_ALU1:
        case K_ALU2:
        case K_ALU3:
          alu<R>(in, acc, accn, r1, n1, r2, n2, live, fail);
          break;
        case K_PRED: {
          uint32_t t = 0;
#pragma unroll
          for (int j = 0; j < R; ++j) t |= static_cast<uint32_t>(acc[j] & 1u) << j;
          pass = t & ~accn & live;
        } break;
        case K_OUT:
#pragma unroll
          for (int j = 0; j < R; ++j) s_out[(in.a * R + j) * T + tid] = acc[j];
          s_outn[in.a * T + tid] = accn;
          break;
        default: break;
      }
    }
    const unsigned int ready_now = *reinterpret_cast<volatile unsigned int*>(&l_ready);
    if (ready_now != my_ready) {
      my_ready = ready_now;
#pragma unroll
      for (int e = 0; e < kTinyGroups; ++e) if ((my_ready >> e) & 1u) my_fp[e] = *reinterpret_cast<volatile unsigned long long*>(&l_fp[e]);
    }
#pragma unroll
    for (int j = 0; j < R; ++j) {
      if (!((pass >> j) & 1u)) continue;
      const long long row = rows_[j];
      unsigned long long kv[kMaxKeys];
      unsigned int knull = 0;
#pragma unroll
      for (int c = 0; c < kMaxKeys; ++c) {
        kv[c] = 0;
        if (c < NK) {
          if ((s_outn[c * T + tid] >> j) & 1u) knull |= 1u << c; else kv[c] = s_out[(c * R + j) * T + tid];
        }
      }
      unsigned long long fp = 0x9E3779B97F4A7C15ull + knull;
#pragma unroll
      for (int c = 0; c < kMaxKeys; ++c) if (c < NK) fp = (fp ^ kv[c]) * 0xff51afd7ed558ccdULL + c;
      int g = -1;
#pragma unroll
      for (int e = 0; e < kTinyGroups; ++e) if (((my_ready >> e) & 1u) && my_fp[e] == fp) g = e;
      if (g >= 0) {
        bool same = l_knull[g] == knull;
#pragma unroll
        for (int c = 0; c < kMaxKeys; ++c) if (c < NK) same = same && l_key[g][c] == kv[c];
        if (!same) g = -1;
      }
      long long slot = -1;
      if (g < 0) {
        slot = p.packed ? find_slot_packed_kv(p, (knull & 1u) != 0, kv[0]) : find_slot_generic_kv(p, kv, knull);
        if (slot < 0) {
          const unsigned long long d = atomicAdd(p.n_deferred, 1ull);
          p.deferred[d] = row;
          continue;
        }
        const unsigned int want = static_cast<unsigned int>(slot) + 1u;
        for (int e = 0; e < kTinyGroups && g < 0; ++e) {
          const unsigned int old = atomicCAS(&l_slot[e], 0u, want);
          if (old == 0u) {
#pragma unroll
            for (int c = 0; c < kMaxKeys; ++c) if (c < NK) l_key[e][c] = kv[c];
            l_knull[e] = knull;
            l_fp[e] = fp;
            __threadfence_block();
            atomicOr(&l_ready, 1u << e);
            g = e;
          } else if (old == want) {
            g = e;
          }
        }
      }
#pragma unroll
      for (int a = 0; a < kLocalMaxAggs; ++a) {
        if (a >= A) break;
        const AggDev& ag = p.agg[a];
        const int o = rp.agg_out[a];
        unsigned long long v = 0;
        if (o >= 0) {
          if ((s_outn[o * T + tid] >> j) & 1u) continue;   // NULL input: no contribution
          v = s_out[(o * R + j) * T + tid];
          if (ag.fn != SSB_AGG_COUNT && ag.in_phys != ag.out_phys) v = convert_value(v, ag.in_phys, ag.out_phys);
        }
        if (g < 0) {   // more groups than local entries in this CTA: straight to the global table
          apply(ag, slot, v, 1ull);
          if (ag.seen != nullptr) ag.seen[slot] = 1u;
          continue;
        }
        unsigned long long* accp = &t_acc[(g * A + a) * T + tid];
        switch (ag.pad) {
          case TA_COUNT: *accp += 1ull; break;
          case TA_SUM_F64: *accp = Codec<double>::enc(Codec<double>::dec(*accp) + Codec<double>::dec(v)); break;
          case TA_SUM_U64: *accp += v; break;
          default: *accp = combine(ag, *accp, v); break;
        }
        if (ag.pad != TA_COUNT) t_seen[g][tid] |= 1u << a;
      }
    }
  }
  if (fail && rp.d_fail != nullptr) atomicOr(rp.d_fail, 1);
  __syncthreads();
  for (int ga = warp; ga < kTinyGroups * A; ga += T / 32) {
    const int g = ga / A, a = ga - g * A;
    if (l_slot[g] == 0u) continue;
    const AggDev& ag = p.agg[a];
    unsigned long long acc2 = 0;
    bool has = false;
    for (int t = lane; t < T; t += 32) {
      const unsigned long long x = t_acc[ga * T + t];
      if (ag.fn == SSB_AGG_COUNT) { acc2 += x; }
      else if ((t_seen[g][t] >> a) & 1u) { acc2 = has ? combine(ag, acc2, x) : x; has = true; }
    }
    for (int d = 16; d > 0; d >>= 1) {
      const unsigned long long ov = __shfl_xor_sync(0xffffffffu, acc2, d);
      const bool oh = __shfl_xor_sync(0xffffffffu, has ? 1 : 0, d) != 0;
      if (ag.fn == SSB_AGG_COUNT) acc2 += ov;
      else if (oh) { acc2 = has ? combine(ag, acc2, ov) : ov; has = true; }
    }
    if (lane != 0) continue;
    const long long slot = static_cast<long long>(l_slot[g] - 1u);
    if (ag.fn == SSB_AGG_COUNT) { if (acc2) atomicAdd(&ag.acc[static_cast<unsigned long long>(slot) * ag.stride], acc2); continue; }
    if (!has) continue;
    if (ag.seen != nullptr) ag.seen[slot] = 1u;
    apply(ag, slot, acc2, 0ull);
  }
}

// ---- finalize: dense result columns ---------------------------------------------------------
struct FinalizeParams {
  GroupParams g;
  unsigned long long total_slots;      // capacity + 2
  unsigned long long* block_counts;    // [grid]
  void* key_out[kMaxKeys];
  uint32_t* key_out_nulls[kMaxKeys];
  void* agg_out[kMaxAggs];
  uint32_t* agg_out_nulls[kMaxAggs];
};

__device__ __forceinline__ bool slot_used(const GroupParams& p, unsigned long long s) {
  if (p.n_keys == 0) return s == 0;
  if (p.packed) {
    if (s < p.capacity) return p.slot_key[s * p.stride] != kEmptyKey;
    return p.slot_state[s - p.capacity] != 0u;
  }
  return s < p.capacity && p.slot_state[s] == 2u;
}

__global__ void __launch_bounds__(256) group_count_kernel(const __grid_constant__ FinalizeParams f) {
  __shared__ unsigned int warp_sum[8];
  const unsigned long long per_block = (f.total_slots + gridDim.x - 1) / gridDim.x;
  const unsigned long long begin = per_block * blockIdx.x;
  unsigned long long end = begin + per_block;
  if (end > f.total_slots) end = f.total_slots;
  unsigned int c = 0;
  for (unsigned long long s = begin + threadIdx.x; s < end; s += blockDim.x) c += slot_used(f.g, s) ? 1u : 0u;
  for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
  if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int w = 0; w < 8; ++w) t += warp_sum[w];
    f.block_counts[blockIdx.x] = t;
  }
}

__device__ __forceinline__ void store_typed(void* base, int phys, unsigned long long pos, unsigned long long v) {
  switch (phys_width(phys)) {
    case 8: static_cast<unsigned long long*>(base)[pos] = v; break;
    case 4: static_cast<uint32_t*>(base)[pos] = static_cast<uint32_t>(v); break;
    default: static_cast<uint8_t*>(base)[pos] = static_cast<uint8_t>(v); break;
  }
}

// One block per slot range; positions inside the block come from a block-wide scan so that
// the output order is the slot order (deterministic for a given table size).
__global__ void __launch_bounds__(256) group_emit_kernel(const __grid_constant__ FinalizeParams f) {
  __shared__ unsigned long long s_base;
  __shared__ unsigned int warp_off[8];
  const GroupParams& p = f.g;
  if (threadIdx.x == 0) {
    unsigned long long b = 0;
    for (unsigned int i = 0; i < blockIdx.x; ++i) b += f.block_counts[i];
    s_base = b;
  }
  __syncthreads();
  const unsigned long long per_block = (f.total_slots + gridDim.x - 1) / gridDim.x;
  const unsigned long long begin = per_block * blockIdx.x;
  unsigned long long end = begin + per_block;
  if (end > f.total_slots) end = f.total_slots;
  unsigned long long running = s_base;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (unsigned long long s0 = begin; s0 < end; s0 += blockDim.x) {
    const unsigned long long s = s0 + threadIdx.x;
    const bool used = s < end && slot_used(p, s);
    const unsigned m = __ballot_sync(0xffffffffu, used);
    if (lane == 0) warp_off[warp] = __popc(m);
    __syncthreads();
    unsigned int before = 0, total = 0;
    for (int w = 0; w < 8; ++w) { if (w < warp) before += warp_off[w]; total += warp_off[w]; }
    if (used) {
      const unsigned long long pos = running + before + __popc(m & ((1u << lane) - 1u));
      // keys
      for (int c = 0; c < p.n_keys; ++c) {
        unsigned long long kv = 0;
        bool kn = false;
        if (p.packed) {
          if (s < p.capacity) kv = p.slot_key[s * p.stride];
          else if (s == p.capacity) kv = kEmptyKey;
          else kn = true;
        } else {
          kv = p.key_store[c][s];
          kn = (p.key_store_null[s] >> c) & 1u;
        }
        store_typed(f.key_out[c], p.key_phys[c], pos, kv);
        if (kn && f.key_out_nulls[c] != nullptr) atomicOr(&f.key_out_nulls[c][pos >> 5], 1u << (pos & 31));
      }
      for (int a = 0; a < p.n_aggs; ++a) {
        const AggDev& ag = p.agg[a];
        store_typed(f.agg_out[a], ag.out_phys, pos, ag.acc[s * ag.stride]);
        const bool isnull = ag.fn != SSB_AGG_COUNT && ag.seen != nullptr && ag.seen[s] == 0u;
        if (isnull && f.agg_out_nulls[a] != nullptr) atomicOr(&f.agg_out_nulls[a][pos >> 5], 1u << (pos & 31));
      }
    }
    running += total;
    __syncthreads();
  }
}

// Initialises the slot records: key word = EMPTY, accumulators = their identities.
struct RecordInit { unsigned long long word[16]; };
__global__ void fill_records_kernel(unsigned long long* rec, unsigned long long n_slots, unsigned int stride, RecordInit init) {
  const unsigned long long n = n_slots * stride;
  const unsigned long long step = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
  for (unsigned long long i = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += step) rec[i] = init.word[i & (stride - 1)];
}

// ---- dense integer keys ----------------------------------------------------------------------------------------
// The general table costs three L2 accesses per row for the C3 shape (read the slot's key word, RED the sum, RED the
// count), and the L2's request rate -- not HBM -- bounds the kernel. When the keys seen so far fill a narrow integer
// range densely, slot = key - lo needs no key word: two REDs per row. Rows outside the range go to the general table
// (deferred list); at the end the touched slots are merged into the general table as partial aggregates, so every
// later step (growth, merge between ranks, finalize) is unchanged. SUM and COUNT aggregates over NOT NULL 8-byte
// columns, one NOT NULL 8-byte integer key.
struct DenseParams {
  const unsigned long long* keys;
  long long rows;
  unsigned long long lo, range;
  int32_t n_aggs;
  int32_t code[kMaxAggs];                  // 0 COUNT, 1 SUM f64, 2 SUM 64-bit integer
  const unsigned long long* in[kMaxAggs];
  unsigned long long* acc[kMaxAggs];
  unsigned long long* hits;                // or NULL
  long long* deferred;
  unsigned long long* n_deferred;
};
__global__ void __launch_bounds__(256, 4) group_update_dense_kernel(const __grid_constant__ DenseParams p) {
  constexpr int R = 4;
  constexpr int TILE = 256 * R;
  const long long tiles = (p.rows + TILE - 1) / TILE;
  for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
    const long long r0 = t * TILE + threadIdx.x;
    unsigned long long slot[R];
    bool in_range[R];
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const long long row = r0 + j * 256;
      slot[j] = row < p.rows ? p.keys[row] - p.lo : ~0ull;
      in_range[j] = slot[j] < p.range;
    }
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const long long row = r0 + j * 256;
      if (row < p.rows && !in_range[j]) {
        const unsigned long long d = atomicAdd(p.n_deferred, 1ull);
        p.deferred[d] = row;
      }
    }
    for (int a = 0; a < p.n_aggs; ++a) {
      if (p.code[a] == 0) {
#pragma unroll
        for (int j = 0; j < R; ++j) if (in_range[j]) atomicAdd(&p.acc[a][slot[j]], 1ull);
        continue;
      }
      unsigned long long v[R];
#pragma unroll
      for (int j = 0; j < R; ++j) v[j] = in_range[j] ? p.in[a][r0 + j * 256] : 0ull;
      if (p.code[a] == 1) {
#pragma unroll
        for (int j = 0; j < R; ++j) if (in_range[j]) atomicAdd(reinterpret_cast<double*>(&p.acc[a][slot[j]]), Codec<double>::dec(v[j]));
      } else {
#pragma unroll
        for (int j = 0; j < R; ++j) if (in_range[j]) atomicAdd(&p.acc[a][slot[j]], v[j]);
      }
    }
    if (p.hits != nullptr) {
#pragma unroll
      for (int j = 0; j < R; ++j) if (in_range[j]) atomicAdd(&p.hits[slot[j]], 1ull);
    }
  }
}
// min / max of the first rows' keys as signed or unsigned 64-bit values: out[0] = min, out[1] = max
__global__ void dense_minmax_kernel(const unsigned long long* __restrict__ keys, long long rows, int is_signed, unsigned long long* out) {
  const unsigned long long flip = is_signed ? 0x8000000000000000ull : 0ull;   // order-preserving image
  unsigned long long lo = ~0ull, hi = 0ull;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < rows; i += stride) {
    const unsigned long long k = keys[i] ^ flip;
    lo = k < lo ? k : lo;
    hi = k > hi ? k : hi;
  }
  for (int d = 16; d > 0; d >>= 1) {
    const unsigned long long l2 = __shfl_xor_sync(0xffffffffu, lo, d), h2 = __shfl_xor_sync(0xffffffffu, hi, d);
    lo = l2 < lo ? l2 : lo;
    hi = h2 > hi ? h2 : hi;
  }
  if ((threadIdx.x & 31) == 0) { atomicMin(&out[0], lo); atomicMax(&out[1], hi); }
}
// The touched slots as rows of partial aggregates: key = lo + slot, one value per aggregate; *n counts them.
struct DenseOut { unsigned long long* key; unsigned long long* agg[8]; unsigned long long* n; };
__global__ void dense_flush_kernel(DenseParams p, const unsigned long long* __restrict__ hits, DenseOut out) {
  const unsigned long long stride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
  for (unsigned long long s = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; s < p.range; s += stride) {
    if (hits[s] == 0) continue;
    const unsigned long long o = atomicAdd(out.n, 1ull);
    out.key[o] = p.lo + s;
    for (int a = 0; a < p.n_aggs; ++a) out.agg[a][o] = p.acc[a][s];
  }
}

// Fills a u64 array with a value (accumulator identities, empty keys).
// ---- AggregateClusters (cursor/core/aggregate_clusters.cc:67-125,233-300): rows with equal keys that are
// CONSECUTIVE in the input form a cluster; a key that comes back later starts a new one. flag[i] = row i
// starts a cluster: its key differs from row i - 1 (NULL equals NULL, values by operator==, so a NaN differs
// from everything and -0.0 equals +0.0, as operators::Equal does).
struct ClusterKeys {
  int32_t n_keys;
  int32_t phys[kMaxKeys];
  const void* data[kMaxKeys];
  const uint32_t* nulls[kMaxKeys];
};
__global__ void __launch_bounds__(256) cluster_flag_kernel(ClusterKeys k, long long rows, unsigned long long* __restrict__ flag) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < rows; i += stride) {
    bool differs = false;
    if (i > 0) {
      for (int c = 0; c < k.n_keys && !differs; ++c) {
        const bool na = bit_at(k.nulls[c], i - 1), nb = bit_at(k.nulls[c], i);
        if (na || nb) { differs = na != nb; continue; }
        const unsigned long long a = load_raw(k.data[c], k.phys[c], i - 1), b = load_raw(k.data[c], k.phys[c], i);
        if (k.phys[c] == T_F64) differs = !(Codec<double>::dec(a) == Codec<double>::dec(b));
        else if (k.phys[c] == T_F32) differs = !(Codec<float>::dec(a) == Codec<float>::dec(b));
        else differs = a != b;
      }
    }
    flag[i] = (i == 0 || differs) ? 1ull : 0ull;
  }
}
// scanned[i] = number of cluster starts before row i (exclusive scan of flag): ids[i] = scanned[i] + flag[i] - 1;
// the start row of every cluster is recorded.
__global__ void __launch_bounds__(256) cluster_ids_kernel(const unsigned long long* __restrict__ scanned, ClusterKeys k, long long rows,
                                                           long long* __restrict__ ids, long long* __restrict__ starts) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < rows; i += stride) {
    const bool start = i == 0 || scanned[i + 1] != scanned[i];
    const long long id = static_cast<long long>(scanned[i + 1]) - 1;   // inclusive count - 1
    ids[i] = id;
    if (start) starts[id] = i;
  }
}

// Floating-point key columns: is some non-NULL key a NaN?
__global__ void nan_key_kernel(const void* data, const uint32_t* nulls, int is_f64, long long rows, int* found) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  bool hit = false;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < rows; i += stride) {
    if (nulls != nullptr && ((nulls[i >> 5] >> (i & 31)) & 1u)) continue;
    if (is_f64) { const double v = static_cast<const double*>(data)[i]; hit = hit || v != v; }
    else { const float v = static_cast<const float*>(data)[i]; hit = hit || v != v; }
  }
  if (hit) *found = 1;
}

__global__ void iota_rows_kernel(long long* p, long long n) {
  for (long long i = threadIdx.x; i < n; i += blockDim.x) p[i] = i;
}
__global__ void fill_u64_kernel(unsigned long long* p, unsigned long long n, unsigned long long v) {
  const unsigned long long stride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
  for (unsigned long long i = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = v;
}

}  // namespace ssb

using namespace ssb;

struct ssb_group {
  ssb_ctx* ctx;
  int n_keys, n_aggs;
  std::vector<int32_t> key_types, key_nullable;
  std::vector<ssb_agg_spec> aggs;
  bool packed;
  unsigned long long capacity;
  // table storage
  unsigned long long* rec;        // slot records
  unsigned long long stride;
  unsigned long long* slot_key;
  uint32_t* slot_state;
  unsigned long long* key_store[kMaxKeys];
  uint32_t* key_store_null;
  unsigned long long* acc[kMaxAggs];
  uint32_t* seen[kMaxAggs];
  unsigned long long* cand[kMaxAggs];   // FIRST / LAST candidates
  bool has_first_last;
  unsigned long long* counters;   // [0] n_groups, [1] n_deferred
  unsigned long long* h_counters; // pinned
  long long* deferred;
  size_t deferred_cap;
  long long rows_seen;
  bool merged_any;
  // result storage
  void* key_out[kMaxKeys];
  uint32_t* key_out_nulls[kMaxKeys];
  void* agg_out[kMaxAggs];
  uint32_t* agg_out_nulls[kMaxAggs];
  unsigned long long* block_counts;
  long long out_capacity;
  // dense-key accumulators (group_update_dense_kernel): slot = key - dense_lo, no key word to read
  int dense_state;                        // 0 undecided, 1 active, -1 not applicable to this input
  long long dense_lo;
  unsigned long long dense_range;
  unsigned long long* dense_acc[kMaxAggs];
  unsigned long long* dense_hits;         // rows per slot when no aggregate counts every row
  int dense_hits_agg;                     // the aggregate whose accumulator counts every row, or -1
};

namespace ssb {

static unsigned long long identity_of(const ssb_agg_spec& a) {
  const int phys = phys_of(a.out_type);
  if (a.fn == SSB_AGG_MIN) {
    switch (phys) {
      case T_F64: return Codec<double>::enc(__builtin_huge_val());
      case T_F32: return Codec<float>::enc(__builtin_huge_valf());
      case T_I64: case T_I32: return static_cast<unsigned long long>(INT64_MAX);
      default: return ~0ull;
    }
  }
  if (a.fn == SSB_AGG_MAX) {
    switch (phys) {
      case T_F64: return Codec<double>::enc(-__builtin_huge_val());
      case T_F32: return Codec<float>::enc(-__builtin_huge_valf());
      case T_I64: case T_I32: return static_cast<unsigned long long>(INT64_MIN);
      default: return 0ull;
    }
  }
  return 0ull;
}

static void free_table(ssb_group* g) {
  ssb_ctx* ctx = g->ctx;
  tmp_free(ctx, g->rec); g->rec = nullptr; g->slot_key = nullptr;
  tmp_free(ctx, g->slot_state); g->slot_state = nullptr;
  for (int c = 0; c < kMaxKeys; ++c) { tmp_free(ctx, g->key_store[c]); g->key_store[c] = nullptr; }
  tmp_free(ctx, g->key_store_null); g->key_store_null = nullptr;
  for (int a = 0; a < kMaxAggs; ++a) {
    g->acc[a] = nullptr;
    tmp_free(ctx, g->seen[a]); g->seen[a] = nullptr;
    tmp_free(ctx, g->cand[a]); g->cand[a] = nullptr;
  }
}

static unsigned fill_grid(ssb_ctx* ctx, unsigned long long n) {
  unsigned long long g = (n + 255) / 256;
  const unsigned long long cap = static_cast<unsigned long long>(ctx->num_sms) * 8;
  return static_cast<unsigned>(g > cap ? cap : (g < 1 ? 1 : g));
}

static int alloc_table(ssb_group* g, unsigned long long capacity) {
  ssb_ctx* ctx = g->ctx;
  g->capacity = capacity;
  const unsigned long long total = capacity + 2;
  // Structure of arrays: [key words][accumulator 0][accumulator 1]... Measured on B200 the
  // 48 MB SoA table of the C3 shape (1M groups) stays L2 resident and runs 30 % faster than
  // 32-byte slot records (64 MB), although a row then touches three sectors instead of one.
  const unsigned long long stride = 1;
  g->stride = stride;
  const unsigned long long arrays = 1 + static_cast<unsigned long long>(g->n_aggs);
  SSB_CUDA(ctx, tmp_malloc(ctx, &g->rec, total * arrays * 8));
  SSB_CUDA(ctx, cudaMemsetAsync(g->rec, 0xff, total * 8, ctx->stream));   // keys = EMPTY
  g->slot_key = g->rec;
  for (int a = 0; a < g->n_aggs; ++a) {
    g->acc[a] = g->rec + (1 + a) * total;
    const unsigned long long id = identity_of(g->aggs[a]);
    if (id == 0) {
      SSB_CUDA(ctx, cudaMemsetAsync(g->acc[a], 0, total * 8, ctx->stream));
    } else {
      fill_u64_kernel<<<fill_grid(ctx, total), 256, 0, ctx->stream>>>(g->acc[a], total, id);
      ++ctx->launches;
    }
  }
  if (g->packed) {
    SSB_CUDA(ctx, tmp_malloc(ctx, &g->slot_state, 2 * 4));
    SSB_CUDA(ctx, cudaMemsetAsync(g->slot_state, 0, 2 * 4, ctx->stream));
  } else {
    SSB_CUDA(ctx, tmp_malloc(ctx, &g->slot_state, total * 4));
    SSB_CUDA(ctx, cudaMemsetAsync(g->slot_state, 0, total * 4, ctx->stream));
    for (int c = 0; c < g->n_keys; ++c) SSB_CUDA(ctx, tmp_malloc(ctx, &g->key_store[c], total * 8));
    SSB_CUDA(ctx, tmp_malloc(ctx, &g->key_store_null, total * 4));
  }
  for (int a = 0; a < g->n_aggs; ++a) {
    // `seen` decides NULL-ness of SUM/MIN/MAX results; only inputs that can be NULL need it
    // (a ScalarAggregate over an empty input is NULL as well: aggregate_scalar.cc:40-90)
    const bool first_last = g->aggs[a].fn == SSB_AGG_FIRST || g->aggs[a].fn == SSB_AGG_LAST;
    if (g->aggs[a].fn != SSB_AGG_COUNT && (g->aggs[a].in_nullable || g->n_keys == 0 || first_last)) {
      SSB_CUDA(ctx, tmp_malloc(ctx, &g->seen[a], total * 4));
      SSB_CUDA(ctx, cudaMemsetAsync(g->seen[a], 0, total * 4, ctx->stream));
    }
    if (first_last) {
      SSB_CUDA(ctx, tmp_malloc(ctx, &g->cand[a], total * 8));
      SSB_CUDA(ctx, cudaMemsetAsync(g->cand[a], g->aggs[a].fn == SSB_AGG_FIRST ? 0xff : 0, total * 8, ctx->stream));
    }
  }
  SSB_CUDA(ctx, cudaGetLastError());
  return 0;
}

static void fill_table_params(const ssb_group* g, GroupParams* p) {
  memset(p, 0, sizeof(*p));
  p->n_keys = g->n_keys;
  p->n_aggs = g->n_aggs;
  p->packed = g->packed ? 1 : 0;
  for (int c = 0; c < g->n_keys; ++c) {
    p->key_phys[c] = phys_of(g->key_types[c]);
    p->key_store[c] = g->key_store[c];
  }
  p->capacity = g->capacity;
  p->slot_key = g->slot_key;
  p->stride = g->stride;
  p->slot_state = g->slot_state;
  p->key_store_null = g->key_store_null;
  p->n_groups = &g->counters[0];
  p->n_deferred = &g->counters[1];
  for (int a = 0; a < g->n_aggs; ++a) {
    p->agg[a].fn = g->aggs[a].fn;
    p->agg[a].in_phys = g->aggs[a].input < 0 ? -1 : phys_of(g->aggs[a].in_type);
    p->agg[a].out_phys = phys_of(g->aggs[a].out_type);
    p->agg[a].in_nullable = g->aggs[a].in_nullable;
    p->agg[a].acc = g->acc[a];
    p->agg[a].stride = static_cast<uint32_t>(g->stride);
    p->agg[a].seen = g->seen[a];
    p->agg[a].cand = g->cand[a];
  }
}

static int read_counters(ssb_group* g) {
  ssb_ctx* ctx = g->ctx;
  SSB_CUDA(ctx, cudaMemcpyAsync(g->h_counters, g->counters, 16, cudaMemcpyDeviceToHost, ctx->stream));
  SSB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

static unsigned update_grid(ssb_ctx* ctx, long long rows) {
  long long g = div_up(rows, 256);
  const long long cap = static_cast<long long>(ctx->num_sms) * 8;
  return static_cast<unsigned>(g > cap ? cap : (g < 1 ? 1 : g));
}

static int feed(ssb_group* g, const ssb_column* keys, const ssb_column* values, int64_t rows, bool merge,
                bool internal);

// Moves every group of the current table into a table of `new_capacity` slots.
static int grow_table(ssb_group* g, unsigned long long new_capacity) {
  ssb_ctx* ctx = g->ctx;
  int64_t n = 0;
  std::vector<ssb_column> keys(g->n_keys ? g->n_keys : 1), aggs(g->n_aggs ? g->n_aggs : 1);
  if (int rc = ssb_group_finalize(g, &n, keys.data(), aggs.data())) return rc;
  // the dense copy must outlive the rebuild: detach it from the group
  void* key_out[kMaxKeys]; uint32_t* key_nulls[kMaxKeys]; void* agg_out[kMaxAggs]; uint32_t* agg_nulls[kMaxAggs];
  for (int c = 0; c < kMaxKeys; ++c) { key_out[c] = g->key_out[c]; key_nulls[c] = g->key_out_nulls[c]; g->key_out[c] = nullptr; g->key_out_nulls[c] = nullptr; }
  for (int a = 0; a < kMaxAggs; ++a) { agg_out[a] = g->agg_out[a]; agg_nulls[a] = g->agg_out_nulls[a]; g->agg_out[a] = nullptr; g->agg_out_nulls[a] = nullptr; }
  g->out_capacity = 0;
  // finalize reports NOT-NULL keys / COUNT columns without bitmap; that is what merge expects
  free_table(g);
  int rc = alloc_table(g, new_capacity);
  if (rc == 0) {
    cudaMemsetAsync(g->counters, 0, 16, ctx->stream);
    g->h_counters[0] = g->h_counters[1] = 0;
    if (n > 0 && g->n_keys > 0) rc = feed(g, keys.data(), aggs.data(), n, true, true);
  }
  cudaStreamSynchronize(ctx->stream);
  for (int c = 0; c < kMaxKeys; ++c) { tmp_free(ctx, key_out[c]); tmp_free(ctx, key_nulls[c]); }
  for (int a = 0; a < kMaxAggs; ++a) { tmp_free(ctx, agg_out[a]); tmp_free(ctx, agg_nulls[a]); }
  return rc;
}

// One slice: launch, then grow-and-replay until no row is deferred.
// How to run a slice through the aggregation sink of expr_kernel (first round; replays of
// deferred rows use the per-thread row evaluator).
// The plan-specialised form of group_update_rows_kernel (csrc/jit.cu, csrc/jit_rows.h).
struct JitLaunch {
  JitKernel k;
  JitRowsShape shape;
  JitRun run;
};

struct SinkLaunch {
  ssb_program* twin;
  const ssb_column* inputs;
  int groups;
  uint32_t out_aggs[16];
  uint32_t count_star;
};

// replay0 (a tmp_malloc'ed list of n_replay0 row numbers, ownership taken): only those rows of the slice are fed.
static int feed_slice(ssb_group* g, const ssb_column* keys, const ssb_column* values, long long rows, bool merge,
                      const RowProg* fused = nullptr, size_t fused_smem = 0, const SinkLaunch* sink = nullptr,
                      long long* replay0 = nullptr, long long n_replay0 = 0, const JitLaunch* jit = nullptr) {
  ssb_ctx* ctx = g->ctx;
  long long remaining = replay0 != nullptr ? n_replay0 : rows;
  long long* replay = replay0;
  int rc = 0;
  for (int round = 0; round < 48 && remaining > 0; ++round) {
    if (g->deferred_cap < static_cast<size_t>(remaining)) {
      tmp_free(ctx, g->deferred);
      g->deferred = nullptr;
      g->deferred_cap = 0;
      cudaError_t e = tmp_malloc(ctx, &g->deferred, static_cast<size_t>(remaining) * 8);
      if (e != cudaSuccess) { rc = cuda_fail(ctx, e, "deferred row list"); break; }
      g->deferred_cap = static_cast<size_t>(remaining);
    }
    GroupParams p;
    fill_table_params(g, &p);
    for (int c = 0; c < g->n_keys; ++c) {
      p.key_data[c] = fused ? nullptr : keys[c].data;
      p.key_nulls[c] = fused ? nullptr : keys[c].nulls;
    }
    for (int a = 0; a < g->n_aggs; ++a) {
      if (merge) {
        p.agg[a].in_phys = phys_of(g->aggs[a].out_type);   // partial results carry the output type
        p.agg[a].in_data = values[a].data;
        p.agg[a].in_nulls = values[a].nulls;
      } else if (g->aggs[a].input >= 0 && fused == nullptr) {
        p.agg[a].in_data = values[g->aggs[a].input].data;
        p.agg[a].in_nulls = values[g->aggs[a].input].nulls;
      }
    }
    p.merge = merge ? 1 : 0;
    p.rows = remaining;
    p.row_index = replay;
    p.deferred = g->deferred;
    // Few groups after the first megarow (or a scalar aggregate): combine inside the warp.
    cudaMemsetAsync(&g->counters[1], 0, 8, ctx->stream);
    // few groups so far (and few enough aggregates): CTA-private shared-memory tables
    const bool few = !g->has_first_last && g->n_aggs <= kLocalMaxAggs && (g->n_keys == 0 || (g->rows_seen >= kProbeRowsFirst && g->h_counters[0] <= 256));
    // single packed key, COUNT or aggregates whose input type equals the result type, no replay;
    // `plain` additionally: 8-byte columns without NULL bitmaps
    static const bool tiny_enabled = getenv("SSB200_GROUP_TINY") == nullptr || atoi(getenv("SSB200_GROUP_TINY")) != 0;
    static const bool fast_enabled = getenv("SSB200_GROUP_FAST") == nullptr || atoi(getenv("SSB200_GROUP_FAST")) != 0;
    bool fast = fused == nullptr && fast_enabled && !g->has_first_last && !few && !merge && replay == nullptr && g->packed && g->n_keys == 1 && g->stride == 1;
    bool plain = fast && phys_width(p.key_phys[0]) == 8 && p.key_nulls[0] == nullptr;
    for (int a = 0; fast && a < g->n_aggs; ++a) {
      const AggDev& ag = p.agg[a];
      if (ag.fn == SSB_AGG_COUNT) { if (ag.in_phys >= 0 && ag.in_nulls != nullptr) plain = false; continue; }
      if (ag.in_phys != ag.out_phys) fast = false;
      if (ag.in_nulls != nullptr || phys_width(ag.in_phys) != 8) plain = false;
    }
    if (fused != nullptr) {
      // Filter -> Compute -> GroupAggregate in one kernel: the keys and aggregate inputs are outputs
      // of the row program, not columns in memory
      for (int a = 0; a < g->n_aggs; ++a) {
        p.agg[a].in_phys = fused->agg_out[a] < 0 ? -1 : phys_of(g->aggs[a].in_type);
        p.agg[a].in_data = nullptr;
        p.agg[a].in_nulls = nullptr;
      }
    }
    GroupParams* d_params = nullptr;
    if (sink != nullptr && replay == nullptr) {
      // the tile-wide expression kernel with the aggregation as its sink (expr_kernel<.., SINK>)
      cudaError_t e2 = tmp_malloc(ctx, &d_params, sizeof(GroupParams));
      if (e2 == cudaSuccess) e2 = cudaMemcpyAsync(d_params, &p, sizeof(GroupParams), cudaMemcpyHostToDevice, ctx->stream);
      if (e2 != cudaSuccess) { rc = cuda_fail(ctx, e2, "sink parameters"); break; }
      uint32_t pad_codes = 0, seen_mask = 0;
      for (int a = 0; a < g->n_aggs; ++a) {
        const AggDev& ag = p.agg[a];
        const uint32_t code = ag.fn == SSB_AGG_COUNT ? TA_COUNT
                              : (ag.fn == SSB_AGG_SUM && ag.out_phys == T_F64) ? TA_SUM_F64
                              : (ag.fn == SSB_AGG_SUM && (ag.out_phys == T_I64 || ag.out_phys == T_U64)) ? TA_SUM_U64 : TA_OTHER;
        pad_codes |= code << (2 * a);
        if (ag.seen != nullptr) seen_mask |= 1u << a;
      }
      rc = launch_program_sink(sink->twin, sink->inputs, remaining, d_params, g->n_keys, g->n_aggs, sink->groups,
                               sink->out_aggs, sink->count_star, pad_codes, seen_mask);
      --ctx->launches;   // counted below
      if (rc) { tmp_free(ctx, d_params); break; }
    } else if (fused != nullptr && jit != nullptr) {
      const size_t smem = jit_rows_smem(jit->shape);
      const void* fn = jit->k.kernel;
      cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
      int per_sm = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, jit->shape.threads, smem) != cudaSuccess || per_sm < 1) { cudaGetLastError(); per_sm = 1; }
      const long long rows_per_cta = static_cast<long long>(jit->shape.threads) * jit->shape.rows_per_thread;
      long long ctas = static_cast<long long>(ctx->num_sms) * per_sm;
      if (ctas > div_up(remaining, rows_per_cta)) ctas = div_up(remaining, rows_per_cta);
      JitRun run = jit->run;
      void* args[] = {&p, &run};
      cudaLaunchKernel(fn, dim3(static_cast<unsigned>(ctas)), dim3(static_cast<unsigned>(jit->shape.threads)), args, smem, ctx->stream);
      jit_note_launch();
    } else if (fused != nullptr) {
      cudaFuncSetAttribute(group_update_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(fused_smem));
      long long per_sm = static_cast<long long>((ctx->smem_per_sm - 8192) / (fused_smem + 8192));
      if (per_sm < 1) per_sm = 1;
      if (per_sm > 8) per_sm = 8;
      long long ctas = static_cast<long long>(ctx->num_sms) * per_sm;
      if (ctas > div_up(remaining, kTinyThreads * 2)) ctas = div_up(remaining, kTinyThreads * 2);
      group_update_rows_kernel<<<static_cast<unsigned>(ctas), kTinyThreads, fused_smem, ctx->stream>>>(p, *fused);
    } else if (fast) {
      long long ctas = static_cast<long long>(ctx->num_sms) * 4;
      if (ctas > div_up(remaining, 1024)) ctas = div_up(remaining, 1024);
      if (plain) group_update_fast_kernel<true><<<static_cast<unsigned>(ctas), 256, 0, ctx->stream>>>(p);
      else group_update_fast_kernel<false><<<static_cast<unsigned>(ctas), 256, 0, ctx->stream>>>(p);
    } else if (few && tiny_enabled && g->n_aggs >= 1 && (g->n_keys == 0 || g->h_counters[0] <= kTinyGroups)) {
      const size_t smem = static_cast<size_t>(kTinyGroups) * g->n_aggs * kTinyThreads * 8;
      bool tfast = !merge && replay == nullptr;
      for (int c = 0; tfast && c < g->n_keys; ++c) tfast = phys_width(p.key_phys[c]) == 8 && p.key_nulls[c] == nullptr;
      for (int a = 0; a < g->n_aggs; ++a) {
        AggDev& ag = p.agg[a];
        ag.pad = ag.fn == SSB_AGG_COUNT ? TA_COUNT
                 : (ag.fn == SSB_AGG_SUM && ag.out_phys == T_F64) ? TA_SUM_F64
                 : (ag.fn == SSB_AGG_SUM && (ag.out_phys == T_I64 || ag.out_phys == T_U64)) ? TA_SUM_U64 : TA_OTHER;
        if (ag.in_phys >= 0 && (ag.in_nulls != nullptr || phys_width(ag.in_phys) != 8 || (ag.fn != SSB_AGG_COUNT && ag.in_phys != ag.out_phys))) tfast = false;
      }
      auto kernel = g->n_keys <= 2 ? (tfast ? group_update_tiny_kernel<true, true> : group_update_tiny_kernel<false, true>)
                                   : (tfast ? group_update_tiny_kernel<true, false> : group_update_tiny_kernel<false, false>);
      cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
      long long per_sm = static_cast<long long>((ctx->smem_per_sm - 8192) / (smem + 8192));
      if (per_sm < 1) per_sm = 1;
      if (per_sm > 8) per_sm = 8;
      long long ctas = static_cast<long long>(ctx->num_sms) * per_sm;
      const long long rows_per_cta = kTinyThreads * (tfast ? 4 : 2);
      if (ctas > div_up(remaining, rows_per_cta)) ctas = div_up(remaining, rows_per_cta);
      kernel<<<static_cast<unsigned>(ctas), kTinyThreads, smem, ctx->stream>>>(p);
    } else if (few) {
      long long ctas = static_cast<long long>(ctx->num_sms) * 4;
      if (ctas > div_up(remaining, 256)) ctas = div_up(remaining, 256);
      group_update_smem_kernel<<<static_cast<unsigned>(ctas), 256, 0, ctx->stream>>>(p);
    } else {
      group_update_kernel<<<update_grid(ctx, remaining), 256, 0, ctx->stream>>>(p);
    }
    ++ctx->launches;
    if (g->has_first_last && !merge) {
      first_last_resolve_kernel<<<fill_grid(ctx, g->capacity + 2), 256, 0, ctx->stream>>>(p, g->capacity + 2);
      ++ctx->launches;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { tmp_free(ctx, d_params); rc = cuda_fail(ctx, e, "group_update_kernel"); break; }
    if ((rc = read_counters(g))) { tmp_free(ctx, d_params); break; }
    tmp_free(ctx, d_params);
    const unsigned long long n_def = g->h_counters[1];
    if (n_def == 0) { remaining = 0; break; }
    long long* next = nullptr;
    e = tmp_malloc(ctx, &next, n_def * 8);
    if (e != cudaSuccess) { rc = cuda_fail(ctx, e, "replay list"); break; }
    cudaMemcpyAsync(next, g->deferred, n_def * 8, cudaMemcpyDeviceToDevice, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    tmp_free(ctx, replay);
    replay = next;
    unsigned long long want = g->capacity * 2;
    while (want < 4 * (g->h_counters[0] + n_def)) want *= 2;
    if ((rc = grow_table(g, want))) break;
    remaining = static_cast<long long>(n_def);
  }
  if (rc == 0 && remaining > 0) rc = fail(ctx, SSB_ERROR_UNKNOWN, "group-by table did not converge");
  cudaStreamSynchronize(ctx->stream);
  tmp_free(ctx, replay);
  return rc;
}

// The reference's hash set asks operator== after the hash (row_hash_set.cc:487-498), so every row with a NaN key
// is a group of its own; this table compares bit images and would put equal NaN payloads into one group.
// Refused (ERROR_NOT_IMPLEMENTED) rather than answered differently; the check reads float key columns only.
static int refuse_nan_keys(ssb_group* g, const ssb_column* keys, int64_t rows) {
  ssb_ctx* ctx = g->ctx;
  for (int c = 0; c < g->n_keys && rows > 0; ++c) {
    const int ph = phys_of(g->key_types[c]);
    if (ph != T_F32 && ph != T_F64) continue;
    SSB_CUDA(ctx, cudaMemsetAsync(ctx->d_fail, 0, sizeof(int32_t), ctx->stream));
    nan_key_kernel<<<update_grid(ctx, rows), 256, 0, ctx->stream>>>(keys[c].data, keys[c].nulls, ph == T_F64 ? 1 : 0, rows, ctx->d_fail);
    ++ctx->launches;
    SSB_CUDA(ctx, cudaMemcpyAsync(ctx->h_fail, ctx->d_fail, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    SSB_CUDA(ctx, cudaMemsetAsync(ctx->d_fail, 0, sizeof(int32_t), ctx->stream));
    SSB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (*ctx->h_fail) {
      return fail(ctx, SSB_ERROR_NOT_IMPLEMENTED, "NaN in a floating-point group-by key: the reference makes every such row a group of its own "
                                                   "(row_hash_set.cc:487-498); not supported on the B200 path");
    }
  }
  return 0;
}

// ---- dense integer keys: host side ------------------------------------------------------------------------------
static bool dense_shape(const ssb_group* g, const ssb_column* keys, const ssb_column* values) {
  static const bool enabled = getenv("SSB200_GROUP_DENSE") == nullptr || atoi(getenv("SSB200_GROUP_DENSE")) != 0;
  if (!enabled || g->n_keys != 1 || g->has_first_last || g->n_aggs < 1 || g->n_aggs > 8) return false;
  const int kp = phys_of(g->key_types[0]);
  if ((kp != T_I64 && kp != T_U64) || keys[0].nulls != nullptr || phys_of(keys[0].dtype) != kp) return false;
  for (int a = 0; a < g->n_aggs; ++a) {
    const ssb_agg_spec& sp = g->aggs[a];
    if (sp.fn == SSB_AGG_COUNT) {
      if (sp.input >= 0 && values[sp.input].nulls != nullptr) return false;
      if (phys_width(phys_of(sp.out_type)) != 8) return false;
      continue;
    }
    if (sp.fn != SSB_AGG_SUM || sp.input < 0) return false;
    const int ip = phys_of(sp.in_type), op = phys_of(sp.out_type);
    if (ip != op || (op != T_F64 && op != T_I64 && op != T_U64)) return false;
    if (values[sp.input].nulls != nullptr || phys_of(values[sp.input].dtype) != ip) return false;
  }
  return true;
}

// After the first rows went through the general table: do their keys fill a narrow range densely?
static int dense_decide(ssb_group* g, const ssb_column* keys, long long rows) {
  ssb_ctx* ctx = g->ctx;
  g->dense_state = -1;
  const unsigned long long groups = g->h_counters[0];
  if (groups < 65536 || rows < 1) return 0;   // few groups: the shared-memory kernels and the L2-resident table are fine
  unsigned long long* mm = nullptr;
  cudaError_t e = tmp_malloc(ctx, &mm, 16);
  if (e != cudaSuccess) return cuda_fail(ctx, e, "dense key scratch");
  const unsigned long long init[2] = {~0ull, 0ull};
  cudaMemcpyAsync(mm, init, 16, cudaMemcpyHostToDevice, ctx->stream);
  const long long sample = rows < (1ll << 20) ? rows : (1ll << 20);
  const int is_signed = phys_of(g->key_types[0]) == T_I64 ? 1 : 0;
  dense_minmax_kernel<<<update_grid(ctx, sample), 256, 0, ctx->stream>>>(static_cast<const unsigned long long*>(keys[0].data), sample, is_signed, mm);
  ++ctx->launches;
  unsigned long long h[2];
  e = cudaMemcpyAsync(h, mm, 16, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  tmp_free(ctx, mm);
  if (e != cudaSuccess) return cuda_fail(ctx, e, "dense key range");
  const unsigned long long span = h[1] - h[0] + 1;          // in the order-preserving image
  if (h[1] < h[0] || span == 0 || span > (1ull << 25)) return 0;
  const unsigned long long margin = span / 8 + 4096;
  const unsigned long long flip = is_signed ? 0x8000000000000000ull : 0ull;
  unsigned long long lo_img = h[0] > margin ? h[0] - margin : 0ull;
  unsigned long long hi_img = h[1] + margin < h[1] ? ~0ull : h[1] + margin;
  const unsigned long long range = hi_img - lo_img + 1;
  if (range == 0 || range > (1ull << 25) || range > 64 * groups) return 0;   // sparse: most slots would stay empty
  bool need_hits = true;
  g->dense_hits_agg = -1;
  for (int a = 0; a < g->n_aggs; ++a) {
    if (g->aggs[a].fn == SSB_AGG_COUNT) { g->dense_hits_agg = a; need_hits = false; break; }   // counts every row (inputs are NOT NULL)
  }
  e = cudaSuccess;
  for (int a = 0; a < g->n_aggs && e == cudaSuccess; ++a) {
    e = tmp_malloc(ctx, &g->dense_acc[a], range * 8);
    if (e == cudaSuccess) e = cudaMemsetAsync(g->dense_acc[a], 0, range * 8, ctx->stream);
  }
  if (e == cudaSuccess && need_hits) {
    e = tmp_malloc(ctx, &g->dense_hits, range * 8);
    if (e == cudaSuccess) e = cudaMemsetAsync(g->dense_hits, 0, range * 8, ctx->stream);
  }
  if (e != cudaSuccess) {   // no memory for it: stay with the general table
    cudaGetLastError();
    for (int a = 0; a < kMaxAggs; ++a) { tmp_free(ctx, g->dense_acc[a]); g->dense_acc[a] = nullptr; }
    tmp_free(ctx, g->dense_hits);
    g->dense_hits = nullptr;
    return 0;
  }
  g->dense_lo = static_cast<long long>(lo_img ^ flip);   // back from the image: wrapping subtraction key - lo is the slot
  g->dense_range = range;
  g->dense_state = 1;
  if (getenv("SSB200_DEBUG_PLAN")) fprintf(stderr, "[ssb200] group-by: dense keys, %llu slots from %lld\n", range, g->dense_lo);
  return 0;
}

static void dense_fill(const ssb_group* g, const ssb_column* keys, const ssb_column* values, long long rows, DenseParams* p) {
  memset(p, 0, sizeof(*p));
  p->keys = keys ? static_cast<const unsigned long long*>(keys[0].data) : nullptr;
  p->rows = rows;
  p->lo = static_cast<unsigned long long>(g->dense_lo);
  p->range = g->dense_range;
  p->n_aggs = g->n_aggs;
  for (int a = 0; a < g->n_aggs; ++a) {
    const ssb_agg_spec& sp = g->aggs[a];
    p->code[a] = sp.fn == SSB_AGG_COUNT ? 0 : (phys_of(sp.out_type) == T_F64 ? 1 : 2);
    p->in[a] = (values && sp.fn != SSB_AGG_COUNT) ? static_cast<const unsigned long long*>(values[sp.input].data) : nullptr;
    p->acc[a] = g->dense_acc[a];
  }
  p->hits = g->dense_hits;
  p->deferred = g->deferred;
  p->n_deferred = &g->counters[1];
}

static int dense_feed(ssb_group* g, const ssb_column* keys, const ssb_column* values, long long rows) {
  ssb_ctx* ctx = g->ctx;
  if (g->deferred_cap < static_cast<size_t>(rows)) {
    tmp_free(ctx, g->deferred);
    g->deferred = nullptr;
    g->deferred_cap = 0;
    cudaError_t e = tmp_malloc(ctx, &g->deferred, static_cast<size_t>(rows) * 8);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "deferred row list");
    g->deferred_cap = static_cast<size_t>(rows);
  }
  DenseParams p;
  dense_fill(g, keys, values, rows, &p);
  cudaMemsetAsync(&g->counters[1], 0, 8, ctx->stream);
  long long ctas = static_cast<long long>(ctx->num_sms) * 4;
  if (ctas > div_up(rows, 1024)) ctas = div_up(rows, 1024);
  group_update_dense_kernel<<<static_cast<unsigned>(ctas), 256, 0, ctx->stream>>>(p);
  ++ctx->launches;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(ctx, e, "group_update_dense_kernel");
  if (int rc = read_counters(g)) return rc;
  const unsigned long long n_def = g->h_counters[1];
  if (n_def == 0) return 0;
  // keys outside the range: the general table takes those rows
  long long* list = nullptr;
  e = tmp_malloc(ctx, &list, n_def * 8);
  if (e != cudaSuccess) return cuda_fail(ctx, e, "replay list");
  cudaMemcpyAsync(list, g->deferred, n_def * 8, cudaMemcpyDeviceToDevice, ctx->stream);
  cudaStreamSynchronize(ctx->stream);
  return feed_slice(g, keys, values, rows, false, nullptr, 0, nullptr, list, static_cast<long long>(n_def));
}

// Merges the touched dense slots into the general table (as partial aggregates) and clears them.
static int dense_flush(ssb_group* g) {
  if (g->dense_state != 1) return 0;
  ssb_ctx* ctx = g->ctx;
  const unsigned long long range = g->dense_range;
  DenseParams p;
  dense_fill(g, nullptr, nullptr, 0, &p);
  DenseOut out;
  memset(&out, 0, sizeof(out));
  cudaError_t e = tmp_malloc(ctx, &out.key, range * 8);
  for (int a = 0; a < g->n_aggs && e == cudaSuccess; ++a) e = tmp_malloc(ctx, &out.agg[a], range * 8);
  if (e == cudaSuccess) e = tmp_malloc(ctx, &out.n, 8);
  auto release = [&]() {
    tmp_free(ctx, out.key);
    for (int a = 0; a < 8; ++a) tmp_free(ctx, out.agg[a]);
    tmp_free(ctx, out.n);
  };
  if (e != cudaSuccess) { release(); return cuda_fail(ctx, e, "dense flush scratch"); }
  cudaMemsetAsync(out.n, 0, 8, ctx->stream);
  const unsigned long long* hits = g->dense_hits != nullptr ? g->dense_hits : g->dense_acc[g->dense_hits_agg];
  dense_flush_kernel<<<update_grid(ctx, static_cast<long long>(range)), 256, 0, ctx->stream>>>(p, hits, out);
  ++ctx->launches;
  e = cudaMemcpyAsync(ctx->h_count, out.n, 8, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) { release(); return cuda_fail(ctx, e, "dense flush"); }
  const long long touched = *ctx->h_count;
  int rc = 0;
  if (touched > 0) {
    ssb_column kc, ac[8];
    kc.data = out.key; kc.nulls = nullptr; kc.dtype = g->key_types[0]; kc.reserved = 0;
    for (int a = 0; a < g->n_aggs; ++a) { ac[a].data = out.agg[a]; ac[a].nulls = nullptr; ac[a].dtype = g->aggs[a].out_type; ac[a].reserved = 0; }
    g->dense_state = -2;   // the merge below must take the general path
    rc = feed(g, &kc, ac, touched, true, true);
    g->dense_state = 1;
  }
  for (int a = 0; a < g->n_aggs; ++a) cudaMemsetAsync(g->dense_acc[a], 0, range * 8, ctx->stream);
  if (g->dense_hits != nullptr) cudaMemsetAsync(g->dense_hits, 0, range * 8, ctx->stream);
  cudaStreamSynchronize(ctx->stream);
  release();
  return rc;
}

static int feed(ssb_group* g, const ssb_column* keys, const ssb_column* values, int64_t rows, bool merge,
                bool internal) {
  ssb_ctx* ctx = g->ctx;
  if (rows < 0) return fail(ctx, SSB_ERROR_INVALID_ARGUMENT_VALUE, "negative row count");
  if (int rc = refuse_nan_keys(g, keys, rows)) return rc;
  int n_values = 0;
  for (int a = 0; a < g->n_aggs; ++a) if (g->aggs[a].input >= n_values) n_values = g->aggs[a].input + 1;
  if (merge) n_values = g->n_aggs;
  long long offset = 0;
  while (offset < rows) {
    long long n = rows - offset;
    // The first megarow runs alone so that the group count seen so far can pick the strategy.
    if (!internal && g->rows_seen < kProbeRowsFirst && n > kProbeRowsFirst) n = kProbeRowsFirst;
    // Slices bound the deferred-row list (one entry per row of a slice in the worst case).
    if (n > kSliceRows) n = kSliceRows;
    ssb_column k2[kMaxKeys], v2[kMaxAggs];
    for (int c = 0; c < g->n_keys; ++c) {
      k2[c] = keys[c];
      k2[c].data = static_cast<char*>(keys[c].data) + static_cast<size_t>(offset) * width_of(keys[c].dtype);
      if (keys[c].nulls) k2[c].nulls = keys[c].nulls + offset / 32;   // offset is a multiple of 32
    }
    for (int v = 0; v < n_values; ++v) {
      v2[v] = values[v];
      v2[v].data = static_cast<char*>(values[v].data) + static_cast<size_t>(offset) * width_of(values[v].dtype);
      if (values[v].nulls) v2[v].nulls = values[v].nulls + offset / 32;
    }
    // dense integer keys (decided once, after the first rows have gone through the general table)
    if (!merge && g->dense_state >= 0 && g->rows_seen >= kProbeRowsFirst && dense_shape(g, k2, v2)) {
      if (g->dense_state == 0) { if (int rc = dense_decide(g, k2, n)) return rc; }
      if (g->dense_state == 1) {
        if (int rc = dense_feed(g, k2, v2, n)) return rc;
        offset += n;
        if (!internal) g->rows_seen += n;
        continue;
      }
    }
    if (int rc = feed_slice(g, k2, v2, n, merge)) return rc;
    offset += n;
    if (!internal) g->rows_seen += n;
  }
  return 0;
}

}  // namespace ssb

extern "C" {

int ssb_group_create(ssb_ctx* ctx, int32_t n_keys, const int32_t* key_types, const int32_t* key_nullable,
                     int32_t n_aggs, const ssb_agg_spec* aggs, int64_t expected_groups, ssb_group** out) {
  *out = nullptr;
  if (n_keys < 0 || n_keys > kMaxKeys) return fail(ctx, SSB_ERROR_NOT_IMPLEMENTED, "too many group-by key columns");
  if (n_aggs < 0 || n_aggs > kMaxAggs) return fail(ctx, SSB_ERROR_NOT_IMPLEMENTED, "too many aggregates");
  for (int c = 0; c < n_keys; ++c) {
    if (phys_of(key_types[c]) < 0) return fail(ctx, SSB_ERROR_INVALID_ARGUMENT_TYPE, "unsupported group-by key type");
  }
  for (int a = 0; a < n_aggs; ++a) {
    const ssb_agg_spec& s = aggs[a];
    if (s.fn != SSB_AGG_SUM && s.fn != SSB_AGG_MIN && s.fn != SSB_AGG_MAX && s.fn != SSB_AGG_COUNT &&
        s.fn != SSB_AGG_FIRST && s.fn != SSB_AGG_LAST) return fail(ctx, SSB_ERROR_NOT_IMPLEMENTED, "unknown aggregate function");
    if (phys_of(s.out_type) < 0 || (s.input >= 0 && phys_of(s.in_type) < 0)) return fail(ctx, SSB_ERROR_INVALID_ARGUMENT_TYPE, "unsupported aggregate type");
    if (s.fn != SSB_AGG_COUNT && s.input < 0) return fail(ctx, SSB_ERROR_INVALID_ARGUMENT_VALUE, "aggregate without input");
    if (s.fn == SSB_AGG_SUM && phys_of(s.out_type) == T_B8) return fail(ctx, SSB_ERROR_INVALID_ARGUMENT_TYPE, "SUM of BOOL");
    if (s.fn == SSB_AGG_COUNT && phys_width(phys_of(s.out_type)) < 4) return fail(ctx, SSB_ERROR_INVALID_ARGUMENT_TYPE, "COUNT needs an integer result");
  }
  ssb_group* g = new ssb_group();
  memset(static_cast<void*>(&g->rec), 0, sizeof(ssb_group) - offsetof(ssb_group, rec));
  g->ctx = ctx;
  g->n_keys = n_keys;
  g->n_aggs = n_aggs;
  g->key_types.assign(key_types, key_types + n_keys);
  g->key_nullable.assign(key_nullable, key_nullable + n_keys);
  g->aggs.assign(aggs, aggs + n_aggs);
  g->packed = n_keys <= 1;
  for (int a = 0; a < n_aggs; ++a) if (aggs[a].fn == SSB_AGG_FIRST || aggs[a].fn == SSB_AGG_LAST) g->has_first_last = true;
  unsigned long long cap = 1ull << 16;
  if (n_keys == 0) cap = 2;
  const unsigned long long want = expected_groups > 0 ? static_cast<unsigned long long>(expected_groups) * 2 : (1ull << 20);
  while (n_keys > 0 && cap < want) cap *= 2;
  cudaError_t e = tmp_malloc(ctx, &g->counters, 16);
  if (e == cudaSuccess) e = cudaMallocHost(&g->h_counters, 16);
  if (e != cudaSuccess) { delete g; return cuda_fail(ctx, e, "group counters"); }
  cudaMemsetAsync(g->counters, 0, 16, ctx->stream);
  g->h_counters[0] = g->h_counters[1] = 0;
  if (int rc = alloc_table(g, cap)) { ssb_group_destroy(g); return rc; }
  *out = g;
  return 0;
}

void ssb_group_destroy(ssb_group* g) {
  if (!g) return;
  ssb_ctx* ctx = g->ctx;
  cudaStreamSynchronize(ctx->stream);
  free_table(g);
  tmp_free(ctx, g->counters);
  cudaFreeHost(g->h_counters);
  tmp_free(ctx, g->deferred);
  tmp_free(ctx, g->block_counts);
  for (int a = 0; a < kMaxAggs; ++a) tmp_free(ctx, g->dense_acc[a]);
  tmp_free(ctx, g->dense_hits);
  for (int c = 0; c < kMaxKeys; ++c) { tmp_free(ctx, g->key_out[c]); tmp_free(ctx, g->key_out_nulls[c]); }
  for (int a = 0; a < kMaxAggs; ++a) { tmp_free(ctx, g->agg_out[a]); tmp_free(ctx, g->agg_out_nulls[a]); }
  delete g;
}

int ssb_group_update(ssb_group* g, const ssb_column* keys, const ssb_column* values, int64_t rows) {
  TimedRegion timed(g->ctx);
  return feed(g, keys, values, rows, false, false);
}

int ssb_group_update_program(ssb_group* g, ssb_program* sp, const ssb_column* inputs, int64_t rows) {
  ssb_ctx* ctx = g->ctx;
  if (rows < 0) return fail(ctx, SSB_ERROR_INVALID_ARGUMENT_VALUE, "negative row count");
  const Program& prog = sp->prog;
  const int n_in = static_cast<int>(prog.input_types.size());
  const int n_out = static_cast<int>(prog.outputs.size());
  int n_values = 0;
  for (int a = 0; a < g->n_aggs; ++a) if (g->aggs[a].input >= n_values) n_values = g->aggs[a].input + 1;
  if (n_out != g->n_keys + n_values) {
    return fail(ctx, SSB_ERROR_INVALID_ARGUMENT_VALUE, "program outputs must be the key columns followed by the aggregate inputs");
  }
  for (int c = 0; c < g->n_keys; ++c) {
    if (phys_of(prog.out_types[c]) != phys_of(g->key_types[c])) return fail(ctx, SSB_ERROR_INVALID_ARGUMENT_TYPE, "key column type differs from the program output");
    if (prog.out_nullable[c] && !g->key_nullable[c]) return fail(ctx, SSB_ERROR_INVALID_ARGUMENT_VALUE, "nullable program output bound to a NOT NULL key");
  }
  for (int a = 0; a < g->n_aggs; ++a) {
    if (g->aggs[a].input < 0) continue;
    const int o = g->n_keys + g->aggs[a].input;
    if (phys_of(prog.out_types[o]) != phys_of(g->aggs[a].in_type)) return fail(ctx, SSB_ERROR_INVALID_ARGUMENT_TYPE, "aggregate input type differs from the program output");
    if (prog.out_nullable[o] && !g->aggs[a].in_nullable && g->aggs[a].fn != SSB_AGG_COUNT) {
      return fail(ctx, SSB_ERROR_INVALID_ARGUMENT_VALUE, "nullable program output bound to an aggregate declared NOT NULL");
    }
  }
  for (int i = 0; i < n_in; ++i) {
    if (inputs[i].data == nullptr && rows > 0) return fail(ctx, SSB_ERROR_INVALID_ARGUMENT_VALUE, "input column without data");
    if (phys_of(inputs[i].dtype) != phys_of(prog.input_types[i])) return fail(ctx, SSB_ERROR_INVALID_ARGUMENT_TYPE, "input column type differs from the compiled program");
    if (inputs[i].nulls != nullptr && !prog.input_nullable[i]) return fail(ctx, SSB_ERROR_INVALID_ARGUMENT_VALUE, "input column carries NULLs but was declared NOT_NULLABLE");
  }
  TimedRegion timed(ctx);
  // ---- the fused form, when the program fits the row evaluator
  RowProg rp;
  memset(&rp, 0, sizeof(rp));
  const int n_tmp = prog.params.n_tmp;
  const int A = g->n_aggs;
  const size_t T = kTinyThreads, R = 2;
  const size_t smem = (static_cast<size_t>(kTinyGroups) * A * T + static_cast<size_t>(n_in + n_tmp) * R * T +
                       static_cast<size_t>(n_out) * R * T) * 8 + static_cast<size_t>(n_in + n_tmp + n_out) * T * 4;
  // Measured (profiles/r1_summary.md): interpreting the row program per thread costs ~1300
  // instructions per row (6 G rows/s on the Q1 shape) against 11 G rows/s for materialising the
  // slice with the tile-wide expression kernel and aggregating it. The per-thread evaluator is
  // therefore opt-in (SSB200_GROUP_FUSED=1, kept as the replay engine of a future sink inside
  // expr_kernel); the default is the sliced two-kernel form below, whose scratch is bounded by
  // the slice, not by the table.
  static const bool fused_enabled = getenv("SSB200_GROUP_FUSED") != nullptr && atoi(getenv("SSB200_GROUP_FUSED")) != 0;
  const bool rows_feasible = !g->has_first_last && A >= 1 && A <= kLocalMaxAggs && n_in <= kRowMaxIn && n_out <= kRowMaxOut &&
                             static_cast<int>(prog.generic.size()) <= kRowMaxInsn && smem + 8192 <= ctx->smem_optin;
  bool float_key = false;   // NaN keys are refused (refuse_nan_keys), which needs the key columns in memory
  for (int c = 0; c < g->n_keys; ++c) { const int ph = phys_of(g->key_types[c]); float_key = float_key || ph == T_F32 || ph == T_F64; }
  const bool feasible = fused_enabled && rows_feasible && !float_key;
  // The aggregation sink inside expr_kernel (tile-wide superinstructions, no output staging, no
  // scratch): the default whenever the plan fits it. Round 1's form (384-row tiles, four rows per
  // thread, a fingerprint loop per row) cost 23 warp instructions per row and lost to the two
  // kernels; this one runs 512-row tiles at eight rows per thread with a hashed group lookup.
  // SSB200_GROUP_SINK=0 selects the two-kernel form (A/B runs, tests).
  const char* sink_env = getenv("SSB200_GROUP_SINK");
  const bool sink_enabled = sink_env == nullptr || atoi(sink_env) != 0;
  bool sink_ok = sink_enabled && rows_feasible && g->n_keys <= 2 && !float_key;   // the row evaluator replays rows the sink had to defer
  for (int a = 0; a < A; ++a) {   // the sink accumulates without conversions
    if (g->aggs[a].fn != SSB_AGG_COUNT && phys_of(g->aggs[a].in_type) != phys_of(g->aggs[a].out_type)) sink_ok = false;
  }
  const char* jit_env = getenv("SSB200_GROUP_JIT");
  const int jit_mode = jit_env == nullptr ? -1 : atoi(jit_env);   // -1: by size, 0: never, 1: always
  const char* jit_rows_env = getenv("SSB200_JIT_MIN_ROWS");
  const long long jit_min_rows = jit_rows_env != nullptr ? atoll(jit_rows_env) : (1LL << 26);
  const char* jit_many_env = getenv("SSB200_JIT_MANY_GROUPS");
  const long long jit_many_groups = jit_many_env != nullptr ? atoll(jit_many_env) : 256;
  bool jit_ok = rows_feasible && !float_key && n_in <= kJitMaxIn;
  if (rows_feasible) {
    rp.n_insn = static_cast<int32_t>(prog.generic.size());
    rp.n_in = n_in; rp.n_tmp = n_tmp; rp.n_out = n_out;
    for (int i = 0; i < rp.n_insn; ++i) rp.insn[i] = prog.generic[i];
    for (int i = 0; i < kRowMaxImm; ++i) rp.imm[i] = prog.params.imm[i];
    for (int i = 0; i < n_in; ++i) rp.in_phys[i] = phys_of(prog.input_types[i]);
    for (int a = 0; a < kMaxAggs; ++a) rp.agg_out[a] = (a < A && g->aggs[a].input >= 0) ? g->n_keys + g->aggs[a].input : -1;
    rp.has_pred = prog.predicate >= 0 ? 1 : 0;
    rp.d_fail = prog.has_signaling ? ctx->d_fail : nullptr;
  }
  int rc = 0;
  long long offset = 0;
  // scratch of the unfused path (allocated on first use)
  std::vector<ssb_column> outs(n_out ? n_out : 1);
  long long outs_rows = 0;
  auto free_outs = [&]() {
    for (int j = 0; j < n_out; ++j) { tmp_free(ctx, outs[j].data); tmp_free(ctx, outs[j].nulls); outs[j].data = nullptr; outs[j].nulls = nullptr; }
  };
  for (int j = 0; j < n_out; ++j) { outs[j].data = nullptr; outs[j].nulls = nullptr; outs[j].dtype = prog.out_types[j]; outs[j].reserved = 0; }
  while (offset < rows && rc == 0) {
    long long n = rows - offset;
    if (g->rows_seen < kProbeRowsFirst && n > kProbeRowsFirst) n = kProbeRowsFirst;
    if (n > kSliceRows) n = kSliceRows;
    std::vector<ssb_column> in2(n_in ? n_in : 1);
    for (int i = 0; i < n_in; ++i) {
      in2[i] = inputs[i];
      in2[i].data = static_cast<char*>(inputs[i].data) + static_cast<size_t>(offset) * width_of(inputs[i].dtype);
      if (inputs[i].nulls) in2[i].nulls = inputs[i].nulls + offset / 32;   // offset is a multiple of 32
    }
    // few groups (known after the first rows): the expression kernel aggregates its own outputs
    bool use_sink = sink_ok && g->rows_seen >= kProbeRowsFirst && (g->n_keys == 0 || g->h_counters[0] <= kTinyGroups);
    SinkLaunch sl;
    if (use_sink) {
      memset(&sl, 0, sizeof(sl));
      sl.groups = g->n_keys == 0 ? 1 : static_cast<int>(g->h_counters[0] < 1 ? 1 : g->h_counters[0]);
      const int src = sink_program_for(sp, g->n_keys, A, sl.groups, &sl.twin);
      if (src != 0) { sink_ok = use_sink = false; }   // too wide for the sink: the two-kernel form below
    }
    // large inputs: the kernel compiled for this plan (csrc/jit.cu). SSB200_GROUP_JIT=1 forces it for every slice
    // (tests), 0 disables it; by default a call with at least SSB200_JIT_MIN_ROWS rows (64M) pays the compilation.
    JitLaunch jl;
    bool use_jit = false;
    // few groups: CTA-local accumulators for exactly the groups seen so far; many groups (>= SSB200_JIT_MANY_GROUPS,
    // 256): no local entries, every row goes to the global table. Atomics on a few hot groups serialise -- measured per
    // 200M rows against the materialising form (profiles/r2m_jit_many_groups.txt): 300 groups 5.9 vs 8.6 ms, 64 groups
    // 12.4 vs 6.9 ms -- so the range in between keeps the materialising form
    const bool groups_known = g->rows_seen >= kProbeRowsFirst;
    const bool few_groups = g->n_keys == 0 || (groups_known ? g->h_counters[0] <= kTinyGroups : jit_mode == 1);
    const bool many_groups = g->n_keys > 0 && groups_known && static_cast<long long>(g->h_counters[0]) >= jit_many_groups;
    if (jit_mode != 0 && jit_ok && (jit_mode == 1 || rows >= jit_min_rows) && (few_groups || many_groups)) {
      memset(&jl, 0, sizeof(jl));
      jl.shape.n_keys = g->n_keys; jl.shape.n_aggs = A;
      jl.shape.groups = g->n_keys == 0 ? 1 : many_groups ? 0 : (groups_known ? static_cast<int>(g->h_counters[0] < 1 ? 1 : g->h_counters[0]) : kTinyGroups);
      jit_rows_tune(&jl.shape);
      for (int a = 0; a < A; ++a) {
        jl.shape.fn[a] = g->aggs[a].fn;
        jl.shape.in_phys[a] = g->aggs[a].input < 0 ? -1 : phys_of(g->aggs[a].in_type);
        jl.shape.out_phys[a] = phys_of(g->aggs[a].out_type);
        jl.shape.out[a] = rp.agg_out[a];
      }
      std::string jerr;
      const std::string src = jit_rows_source(prog, jl.shape, &jerr);
      if (!src.empty() && jit_rows_smem(jl.shape) + 4096 <= ctx->smem_optin && jit_get_kernel(ctx, src, "ssb_jit_rows", &jl.k) == 0) {
        use_jit = true;
        use_sink = false;
      } else {
        jit_ok = false;   // this call keeps the interpreting kernels
      }
    }
    // opt-in: every thread interprets the row program for its own rows (see the note above)
    const bool fuse = !use_sink && feasible && (g->n_keys == 0 || g->rows_seen < kProbeRowsFirst || g->h_counters[0] <= kTinyGroups);
    if (use_jit) {
      for (int i = 0; i < n_in; ++i) { jl.run.in_data[i] = in2[i].data; jl.run.in_nulls[i] = in2[i].nulls; }
      jl.run.d_fail = rp.d_fail;
      rc = feed_slice(g, nullptr, nullptr, n, false, &rp, smem, nullptr, nullptr, 0, &jl);
    } else if (use_sink) {
      for (int i = 0; i < n_in; ++i) { rp.in_data[i] = in2[i].data; rp.in_nulls[i] = in2[i].nulls; }
      sl.inputs = in2.data();
      for (int a = 0; a < A; ++a) {
        if (g->aggs[a].input < 0) sl.count_star |= 1u << a;
        else sl.out_aggs[g->n_keys + g->aggs[a].input] |= 1u << a;
      }
      rc = feed_slice(g, nullptr, nullptr, n, false, &rp, smem, &sl);
    } else if (fuse) {
      for (int i = 0; i < n_in; ++i) { rp.in_data[i] = in2[i].data; rp.in_nulls[i] = in2[i].nulls; }
      rc = feed_slice(g, nullptr, nullptr, n, false, &rp, smem);
    } else {
      // many groups: materialise the slice with the fused Compute / Filter kernel, then aggregate it
      if (outs_rows < n) {
        free_outs();
        cudaError_t e = cudaSuccess;
        for (int j = 0; j < n_out && e == cudaSuccess; ++j) {
          e = tmp_malloc(ctx, &outs[j].data, static_cast<size_t>(n) * 8 + 256);
          if (e == cudaSuccess && prog.out_nullable[j]) e = tmp_malloc(ctx, &outs[j].nulls, static_cast<size_t>(n / 32 + 2) * 4 + 256);
        }
        if (e != cudaSuccess) { rc = cuda_fail(ctx, e, "fused aggregate scratch"); break; }
        outs_rows = n;
      }
      rc = ssb_program_run(sp, in2.data(), n, outs.data(), ctx->d_count);
      if (rc) break;
      cudaError_t e = cudaMemcpyAsync(ctx->h_count, ctx->d_count, 8, cudaMemcpyDeviceToHost, ctx->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
      if (e != cudaSuccess) { rc = cuda_fail(ctx, e, "fused aggregate"); break; }
      const long long kept = *ctx->h_count;
      rc = feed(g, outs.data(), outs.data() + g->n_keys, kept, false, true);
    }
    offset += n;
    g->rows_seen += n;
  }
  cudaStreamSynchronize(ctx->stream);
  free_outs();
  if (rc == 0) rc = ssb_program_check_failure(sp);
  return rc;
}

int ssb_group_merge(ssb_group* dst, int64_t n_groups, const ssb_column* key_cols, const ssb_column* agg_cols) {
  if (int rc = dense_flush(dst)) return rc;
  dst->merged_any = true;
  return feed(dst, key_cols, agg_cols, n_groups, true, false);
}

int ssb_group_finalize(ssb_group* g, int64_t* n_groups, ssb_column* key_out, ssb_column* agg_out) {
  ssb_ctx* ctx = g->ctx;
  if (int rc = dense_flush(g)) return rc;
  if (int rc = read_counters(g)) return rc;
  long long n = g->n_keys == 0 ? 1 : static_cast<long long>(g->h_counters[0]);
  if (g->n_keys == 0 && g->rows_seen == 0 && !g->merged_any) n = 1;   // ScalarAggregate: exactly one row
  const long long cap = n > 0 ? n : 1;
  if (g->out_capacity < cap) {
    for (int c = 0; c < kMaxKeys; ++c) { tmp_free(ctx, g->key_out[c]); tmp_free(ctx, g->key_out_nulls[c]); g->key_out[c] = nullptr; g->key_out_nulls[c] = nullptr; }
    for (int a = 0; a < kMaxAggs; ++a) { tmp_free(ctx, g->agg_out[a]); tmp_free(ctx, g->agg_out_nulls[a]); g->agg_out[a] = nullptr; g->agg_out_nulls[a] = nullptr; }
    for (int c = 0; c < g->n_keys; ++c) {
      SSB_CUDA(ctx, tmp_malloc(ctx, &g->key_out[c], static_cast<size_t>(cap) * 8 + 128));
      SSB_CUDA(ctx, tmp_malloc(ctx, &g->key_out_nulls[c], static_cast<size_t>(cap / 32 + 2) * 4 + 128));
    }
    for (int a = 0; a < g->n_aggs; ++a) {
      SSB_CUDA(ctx, tmp_malloc(ctx, &g->agg_out[a], static_cast<size_t>(cap) * 8 + 128));
      SSB_CUDA(ctx, tmp_malloc(ctx, &g->agg_out_nulls[a], static_cast<size_t>(cap / 32 + 2) * 4 + 128));
    }
    g->out_capacity = cap;
  }
  for (int c = 0; c < g->n_keys; ++c) SSB_CUDA(ctx, cudaMemsetAsync(g->key_out_nulls[c], 0, static_cast<size_t>(cap / 32 + 2) * 4, ctx->stream));
  for (int a = 0; a < g->n_aggs; ++a) SSB_CUDA(ctx, cudaMemsetAsync(g->agg_out_nulls[a], 0, static_cast<size_t>(cap / 32 + 2) * 4, ctx->stream));
  FinalizeParams f;
  memset(&f, 0, sizeof(f));
  fill_table_params(g, &f.g);
  f.total_slots = g->capacity + 2;
  unsigned grid = static_cast<unsigned>(ctx->num_sms) * 4;
  if (grid > f.total_slots / 256 + 1) grid = static_cast<unsigned>(f.total_slots / 256 + 1);
  if (!g->block_counts) SSB_CUDA(ctx, tmp_malloc(ctx, &g->block_counts, static_cast<size_t>(ctx->num_sms) * 4 * 8));
  f.block_counts = g->block_counts;
  for (int c = 0; c < g->n_keys; ++c) { f.key_out[c] = g->key_out[c]; f.key_out_nulls[c] = g->key_out_nulls[c]; }
  for (int a = 0; a < g->n_aggs; ++a) { f.agg_out[a] = g->agg_out[a]; f.agg_out_nulls[a] = g->agg_out_nulls[a]; }
  group_count_kernel<<<grid, 256, 0, ctx->stream>>>(f);
  group_emit_kernel<<<grid, 256, 0, ctx->stream>>>(f);
  ctx->launches += 2;
  SSB_CUDA(ctx, cudaGetLastError());
  SSB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (int c = 0; c < g->n_keys; ++c) {
    key_out[c].data = g->key_out[c];
    key_out[c].nulls = g->key_nullable[c] ? g->key_out_nulls[c] : nullptr;
    key_out[c].dtype = g->key_types[c];
    key_out[c].reserved = 0;
  }
  for (int a = 0; a < g->n_aggs; ++a) {
    agg_out[a].data = g->agg_out[a];
    agg_out[a].nulls = g->aggs[a].fn == SSB_AGG_COUNT ? nullptr : g->agg_out_nulls[a];
    agg_out[a].dtype = g->aggs[a].out_type;
    agg_out[a].reserved = 0;
  }
  *n_groups = n;
  return 0;
}

int ssb_cluster_ids(ssb_ctx* ctx, int32_t n_keys, const ssb_column* keys, int64_t rows, int64_t* d_ids, int64_t* d_starts,
                    int64_t* n_clusters) {
  *n_clusters = 0;
  if (rows < 0) return fail(ctx, SSB_ERROR_INVALID_ARGUMENT_VALUE, "negative row count");
  if (n_keys < 0 || n_keys > kMaxKeys) return fail(ctx, SSB_ERROR_NOT_IMPLEMENTED, "too many clustering key columns");
  if (rows == 0) return 0;
  ClusterKeys k;
  memset(&k, 0, sizeof(k));
  k.n_keys = n_keys;
  for (int c = 0; c < n_keys; ++c) {
    k.phys[c] = phys_of(keys[c].dtype);
    if (k.phys[c] < 0) return fail(ctx, SSB_ERROR_INVALID_ARGUMENT_TYPE, "unsupported clustering key type");
    k.data[c] = keys[c].data;
    k.nulls[c] = keys[c].nulls;
  }
  unsigned long long* flag = nullptr;
  unsigned long long* d_total = nullptr;
  cudaError_t e = tmp_malloc(ctx, &flag, static_cast<size_t>(rows + 1) * 8);
  if (e == cudaSuccess) e = tmp_malloc(ctx, &d_total, 8);
  if (e != cudaSuccess) { tmp_free(ctx, flag); tmp_free(ctx, d_total); return cuda_fail(ctx, e, "cluster scratch"); }
  cluster_flag_kernel<<<update_grid(ctx, rows), 256, 0, ctx->stream>>>(k, rows, flag);
  ++ctx->launches;
  cudaMemsetAsync(flag + rows, 0, 8, ctx->stream);
  int rc = exclusive_scan_u64(ctx, flag, static_cast<unsigned long long>(rows) + 1, d_total);
  if (rc == 0) {
    cluster_ids_kernel<<<update_grid(ctx, rows), 256, 0, ctx->stream>>>(flag, k, rows, reinterpret_cast<long long*>(d_ids),
                                                                       reinterpret_cast<long long*>(d_starts));
    ++ctx->launches;
    e = cudaMemcpyAsync(ctx->h_count, d_total, 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = cuda_fail(ctx, e, "cluster ids");
    else *n_clusters = *ctx->h_count;
  }
  tmp_free(ctx, flag);
  tmp_free(ctx, d_total);
  return rc;
}

// ---- the exchange step of the row-range sharded aggregate ------------------------------------
// Reduce-scatter by key: the dense partial table is hash-partitioned over the ranks, every rank
// merges the partial rows of the keys it owns (about n_groups rows in total, whatever the number of
// ranks) and the merged ranges are all-gathered. is_null flags travel as one byte per row.
int ssb_shard_group_merge(ssb_comm* comm, ssb_group* g, int64_t* n_groups, ssb_column* key_out, ssb_column* agg_out) {
  ssb_ctx* ctx = g->ctx;
  if (comm_ctx(comm) != ctx) return fail(ctx, SSB_ERROR_INVALID_ARGUMENT_VALUE, "the communicator and the table live on different contexts");
  if (g->has_first_last) return fail(ctx, SSB_ERROR_NOT_IMPLEMENTED, "FIRST / LAST over sharded input (the row order across shards is not carried)");
  const int W = comm_world(comm), rank = comm_rank(comm);
  int64_t n = 0;
  if (int rc = ssb_group_finalize(g, &n, key_out, agg_out)) return rc;
  if (W == 1) { *n_groups = n; return 0; }
  const int NK = g->n_keys, NA = g->n_aggs, NC = NK + NA;
  // ---- 1. rows per owner: hash partition of the partial groups (a scalar aggregate's single row goes to rank 0)
  std::vector<int64_t> send_rows(W, 0), recv_rows(W, 0);
  long long* perm = nullptr;
  SSB_CUDA(ctx, tmp_malloc(ctx, &perm, static_cast<size_t>(n > 0 ? n : 1) * 8));
  int rc = 0;
  if (NK == 0) {
    send_rows[0] = n;
    iota_rows_kernel<<<1, 32, 0, ctx->stream>>>(perm, n);
  } else {
    rc = ssb_partition_rows(ctx, NK, key_out, n, W, 0, reinterpret_cast<int64_t*>(perm), send_rows.data());
  }
  // every column travels as values (8 bytes per row: the result containers) + one is_null byte per row
  struct Col { void* vals; uint8_t* nulls; void* rvals; uint8_t* rnulls; const ssb_column* src; };
  std::vector<Col> cols(NC);
  for (int i = 0; i < NC; ++i) { cols[i].vals = nullptr; cols[i].nulls = nullptr; cols[i].rvals = nullptr; cols[i].rnulls = nullptr; cols[i].src = i < NK ? &key_out[i] : &agg_out[i - NK]; }
  uint8_t* bytes = nullptr;
  auto cleanup = [&]() {
    for (int i = 0; i < NC; ++i) { tmp_free(ctx, cols[i].vals); tmp_free(ctx, cols[i].nulls); tmp_free(ctx, cols[i].rvals); tmp_free(ctx, cols[i].rnulls); }
    tmp_free(ctx, perm);
    tmp_free(ctx, bytes);
  };
  const size_t cap = static_cast<size_t>(n > 0 ? n : 1);
  if (rc == 0 && tmp_malloc(ctx, &bytes, cap + 64) != cudaSuccess) rc = fail(ctx, SSB_ERROR_MEMORY_EXCEEDED, "sharded aggregate scratch");
  for (int i = 0; i < NC && rc == 0; ++i) {
    const int w = width_of(cols[i].src->dtype);
    if (tmp_malloc_bytes(ctx, &cols[i].vals, cap * 8 + 64) != cudaSuccess || tmp_malloc(ctx, &cols[i].nulls, cap + 64) != cudaSuccess) {
      rc = fail(ctx, SSB_ERROR_MEMORY_EXCEEDED, "sharded aggregate send buffers");
      break;
    }
    if (n == 0) continue;
    ssb_column src = *cols[i].src, dst;
    src.nulls = nullptr;
    dst.data = cols[i].vals; dst.nulls = nullptr; dst.dtype = src.dtype; dst.reserved = 0;
    rc = ssb_gather(ctx, &src, reinterpret_cast<const int64_t*>(perm), n, &dst);
    if (rc) break;
    if (cols[i].src->nulls != nullptr) {
      rc = ssb_nulls_unpack(ctx, cols[i].src->nulls, n, bytes);
      ssb_column bs, bd;
      bs.data = bytes; bs.nulls = nullptr; bs.dtype = SSB_BOOL; bs.reserved = 0;
      bd.data = cols[i].nulls; bd.nulls = nullptr; bd.dtype = SSB_BOOL; bd.reserved = 0;
      if (rc == 0) rc = ssb_gather(ctx, &bs, reinterpret_cast<const int64_t*>(perm), n, &bd);
    } else {
      cudaMemsetAsync(cols[i].nulls, 0, cap, ctx->stream);
    }
    (void)w;
  }
  if (rc) { cleanup(); return rc; }
  // ---- 2. exchange: counts, then all columns in one grouped NCCL operation
  if ((rc = comm_exchange_counts(comm, send_rows.data(), recv_rows.data()))) { cleanup(); return rc; }
  int64_t got = 0;
  for (int r = 0; r < W; ++r) got += recv_rows[r];
  const size_t rcap = static_cast<size_t>(got > 0 ? got : 1);
  std::vector<const void*> sp;
  std::vector<void*> rp;
  std::vector<int32_t> widths;
  for (int i = 0; i < NC && rc == 0; ++i) {
    if (tmp_malloc_bytes(ctx, &cols[i].rvals, rcap * 8 + 64) != cudaSuccess || tmp_malloc(ctx, &cols[i].rnulls, rcap + 64) != cudaSuccess) {
      rc = fail(ctx, SSB_ERROR_MEMORY_EXCEEDED, "sharded aggregate receive buffers");
      break;
    }
    const int w = width_of(cols[i].src->dtype);
    sp.push_back(cols[i].vals); rp.push_back(cols[i].rvals); widths.push_back(w);
    sp.push_back(cols[i].nulls); rp.push_back(cols[i].rnulls); widths.push_back(1);
  }
  if (rc == 0) rc = comm_all_to_all_v(comm, static_cast<int>(sp.size()), sp.data(), rp.data(), widths.data(), send_rows.data(), recv_rows.data());
  if (rc) { cleanup(); return rc; }
  // ---- 3. merge the received partial rows of the keys this rank owns
  ssb_group* m = nullptr;
  rc = ssb_group_create(ctx, NK, g->key_types.data(), g->key_nullable.data(), NA, g->aggs.data(), got > 0 ? got : 1, &m);
  std::vector<ssb_column> mk(NK ? NK : 1), ma(NA ? NA : 1);
  std::vector<uint32_t*> bitmaps;
  for (int i = 0; i < NC && rc == 0; ++i) {
    ssb_column& c = i < NK ? mk[i] : ma[i - NK];
    c.data = cols[i].rvals; c.dtype = cols[i].src->dtype; c.reserved = 0; c.nulls = nullptr;
    if (cols[i].src->nulls != nullptr && got > 0) {
      uint32_t* bm = nullptr;
      if (tmp_malloc(ctx, &bm, (rcap / 32 + 2) * 4 + 64) != cudaSuccess) { rc = fail(ctx, SSB_ERROR_MEMORY_EXCEEDED, "sharded aggregate bitmaps"); break; }
      bitmaps.push_back(bm);
      rc = ssb_nulls_pack(ctx, cols[i].rnulls, got, bm);
      c.nulls = bm;
    }
  }
  if (rc == 0 && got > 0) rc = ssb_group_merge(m, got, mk.data(), ma.data());
  int64_t mine = 0;
  std::vector<ssb_column> fk(NK ? NK : 1), fa(NA ? NA : 1);
  if (rc == 0) {
    if (got > 0 || (NK == 0 && rank == 0)) rc = ssb_group_finalize(m, &mine, fk.data(), fa.data());
    if (NK == 0 && rank != 0) mine = 0;   // the scalar row lives on rank 0
  }
  for (size_t i = 0; i < bitmaps.size(); ++i) tmp_free(ctx, bitmaps[i]);
  // ---- 4. all-gather the merged ranges: every rank ends with the whole result
  std::vector<int64_t> all_rows(W, 0);
  if (rc == 0) rc = comm_all_gather_counts(comm, &mine, 1, all_rows.data());
  int64_t total = 0;
  for (int r = 0; r < W; ++r) total += all_rows[r];
  if (rc == 0) {
    // result storage of `g` (as ssb_group_finalize lays it out), grown to the global group count
    const long long need = total > 0 ? total : 1;
    if (g->out_capacity < need) {
      for (int c = 0; c < kMaxKeys; ++c) { tmp_free(ctx, g->key_out[c]); tmp_free(ctx, g->key_out_nulls[c]); g->key_out[c] = nullptr; g->key_out_nulls[c] = nullptr; }
      for (int a = 0; a < kMaxAggs; ++a) { tmp_free(ctx, g->agg_out[a]); tmp_free(ctx, g->agg_out_nulls[a]); g->agg_out[a] = nullptr; g->agg_out_nulls[a] = nullptr; }
      cudaError_t e = cudaSuccess;
      for (int c = 0; c < NK && e == cudaSuccess; ++c) {
        e = tmp_malloc_bytes(ctx, &g->key_out[c], static_cast<size_t>(need) * 8 + 128);
        if (e == cudaSuccess) e = tmp_malloc(ctx, &g->key_out_nulls[c], static_cast<size_t>(need / 32 + 2) * 4 + 128);
      }
      for (int a = 0; a < NA && e == cudaSuccess; ++a) {
        e = tmp_malloc_bytes(ctx, &g->agg_out[a], static_cast<size_t>(need) * 8 + 128);
        if (e == cudaSuccess) e = tmp_malloc(ctx, &g->agg_out_nulls[a], static_cast<size_t>(need / 32 + 2) * 4 + 128);
      }
      if (e != cudaSuccess) rc = cuda_fail(ctx, e, "sharded aggregate result");
      g->out_capacity = need;
    }
  }
  std::vector<uint8_t*> gathered_nulls(NC, nullptr), my_nulls(NC, nullptr);
  if (rc == 0) {
    sp.clear(); rp.clear(); widths.clear();
    for (int i = 0; i < NC && rc == 0; ++i) {
      const ssb_column& f = i < NK ? fk[i] : fa[i - NK];
      const int w = width_of(cols[i].src->dtype);
      void* dst = i < NK ? g->key_out[i] : g->agg_out[i - NK];
      sp.push_back(mine > 0 ? f.data : dst); rp.push_back(dst); widths.push_back(w);
      if (cols[i].src->nulls != nullptr) {
        if (tmp_malloc(ctx, &gathered_nulls[i], static_cast<size_t>(total > 0 ? total : 1) + 64) != cudaSuccess ||
            tmp_malloc(ctx, &my_nulls[i], static_cast<size_t>(mine > 0 ? mine : 1) + 64) != cudaSuccess) {
          rc = fail(ctx, SSB_ERROR_MEMORY_EXCEEDED, "sharded aggregate is_null bytes");
          break;
        }
        if (mine > 0) {
          if (f.nulls != nullptr) rc = ssb_nulls_unpack(ctx, f.nulls, mine, my_nulls[i]);
          else cudaMemsetAsync(my_nulls[i], 0, static_cast<size_t>(mine), ctx->stream);
        }
        sp.push_back(my_nulls[i]); rp.push_back(gathered_nulls[i]); widths.push_back(1);
      }
    }
    if (rc == 0) rc = comm_all_gather_v(comm, static_cast<int>(sp.size()), sp.data(), rp.data(), widths.data(), all_rows.data());
    for (int i = 0; i < NC && rc == 0; ++i) {
      if (gathered_nulls[i] == nullptr || total == 0) continue;
      rc = ssb_nulls_pack(ctx, gathered_nulls[i], total, i < NK ? g->key_out_nulls[i] : g->agg_out_nulls[i - NK]);
    }
  }
  cudaStreamSynchronize(ctx->stream);
  for (int i = 0; i < NC; ++i) { tmp_free(ctx, gathered_nulls[i]); tmp_free(ctx, my_nulls[i]); }
  if (m) ssb_group_destroy(m);
  cleanup();
  if (rc) return rc;
  for (int c = 0; c < NK; ++c) {
    key_out[c].data = g->key_out[c];
    key_out[c].nulls = g->key_nullable[c] ? g->key_out_nulls[c] : nullptr;
    key_out[c].dtype = g->key_types[c];
    key_out[c].reserved = 0;
  }
  for (int a = 0; a < NA; ++a) {
    agg_out[a].data = g->agg_out[a];
    agg_out[a].nulls = g->aggs[a].fn == SSB_AGG_COUNT ? nullptr : g->agg_out_nulls[a];
    agg_out[a].dtype = g->aggs[a].out_type;
    agg_out[a].reserved = 0;
  }
  *n_groups = total;
  return 0;
}

}  // extern "C"
