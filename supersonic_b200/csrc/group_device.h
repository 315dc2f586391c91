// group_device.h -- the device side of the aggregation table (structures, slot lookup, accumulate
// steps), shared by group.cu and by the aggregation sink of expr_kernel.cu.
#ifndef SSB_CSRC_GROUP_DEVICE_H_
#define SSB_CSRC_GROUP_DEVICE_H_

#if defined(__CUDACC_RTC__)
#include "jit_rt.h"   // run-time compilation (csrc/jit.cu): device code only
#include "ops.h"
#else
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.h"
#endif

namespace ssb {

enum { kMaxKeys = 8, kMaxAggs = 16, kProbeLimit = 96 };
static constexpr long long kSliceRows = 1LL << 26;   // rows per launch of the update kernels (multiple of 32)
static constexpr long long kProbeRowsFirst = 1LL << 17;   // the first slice runs alone: its group count picks the kernel
static constexpr unsigned long long kEmptyKey = ~0ull;

struct AggDev {
  int32_t fn;            // SSB_AGG_*
  int32_t in_phys;       // physical type of the input column (-1: COUNT(*))
  int32_t out_phys;      // physical type of the accumulator / result
  int32_t in_nullable;
  const void* in_data;
  const uint32_t* in_nulls;
  unsigned long long* acc;   // accumulator of slot s = acc[s * stride] (records: key word, then one word per aggregate)
  uint32_t stride;
  uint32_t pad;
  uint32_t* seen;            // [capacity + 2] or NULL (input not nullable and not merging)
  unsigned long long* cand;  // FIRST / LAST: candidate row of the current launch per slot (FIRST: smallest row,
                             // ~0 = none; LAST: largest row + 1, 0 = none); resolved after every launch
};

struct GroupParams {
  int32_t n_keys, n_aggs;
  int32_t packed;            // single key column stored in the slot word
  int32_t merge;             // inputs are partial aggregates (COUNT adds its input)
  int32_t reserved;
  int32_t key_phys[kMaxKeys];
  const void* key_data[kMaxKeys];
  const uint32_t* key_nulls[kMaxKeys];
  // table
  unsigned long long capacity;       // power of two; special slots: capacity (EMPTY key), capacity+1 (NULL key)
  unsigned long long* slot_key;      // packed: key word of slot s = slot_key[s * stride]; generic: unused
  unsigned long long stride;         // 8-byte words per slot record (power of two)
  uint32_t* slot_state;              // generic: 0 empty, 1 being written, 2 ready; packed: special-slot flags
  unsigned long long* key_store[kMaxKeys];   // generic: stored key values per column [capacity]
  uint32_t* key_store_null;          // generic: bit c set = key column c is NULL [capacity]
  unsigned long long* n_groups;
  AggDev agg[kMaxAggs];
  // rows
  long long rows;
  const long long* row_index;        // replay of deferred rows, or NULL
  long long* deferred;
  unsigned long long* n_deferred;
};

__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}

__device__ __forceinline__ unsigned long long load_raw(const void* base, int phys, long long i) {
  switch (phys_width(phys)) {
    case 8: return static_cast<const unsigned long long*>(base)[i];
    case 4: return static_cast<const uint32_t*>(base)[i];
    default: return static_cast<const uint8_t*>(base)[i];
  }
}
__device__ __forceinline__ bool bit_at(const uint32_t* bm, long long i) {
  return bm != nullptr && ((bm[i >> 5] >> (i & 31)) & 1u);
}

// value of physical type `from` converted (C++ conversion) into a container of type `to`
__device__ __forceinline__ unsigned long long convert_value(unsigned long long v, int from, int to) {
  if (from == to) return v;
  Insn in;
  in.kind = K_ALU1; in.mop = M_CAST; in.t = static_cast<uint8_t>(from); in.t2 = static_cast<uint8_t>(to);
  in.flags = 0; in.rhs_nullable = 0; in.rw = 0; in.a = 0; in.b = 0; in.code = 0; in.pad2 = 0; in.off_a = 0; in.off_b = 0;
  u64 acc[1] = {v}, rhs[1] = {0}, rhs2[1] = {0};
  uint32_t n = 0, fail = 0;
  alu<1>(in, acc, n, rhs, 0u, rhs2, 0u, 1u, fail);
  return acc[0];
}

// ---- slot lookup ---------------------------------------------------------------------------
// Returns the slot of the row's key, inserting it if new; -1 when the probe limit is hit.
__device__ __forceinline__ long long find_slot_packed_kv(const GroupParams& p, bool isnull, unsigned long long key) {
  if (p.n_keys == 0) return 0;
  if (isnull) {
    if (atomicExch(&p.slot_state[1], 1u) == 0u) atomicAdd(p.n_groups, 1ull);
    return static_cast<long long>(p.capacity + 1);
  }
  if (key == kEmptyKey) {
    if (atomicExch(&p.slot_state[0], 1u) == 0u) atomicAdd(p.n_groups, 1ull);
    return static_cast<long long>(p.capacity);
  }
  const unsigned long long mask = p.capacity - 1;
  unsigned long long s = mix64(key) & mask;
  for (int probe = 0; probe < kProbeLimit; ++probe) {
    unsigned long long cur = p.slot_key[s * p.stride];
    if (cur == key) return static_cast<long long>(s);
    if (cur == kEmptyKey) {
      const unsigned long long old = atomicCAS(&p.slot_key[s * p.stride], kEmptyKey, key);
      if (old == kEmptyKey) { atomicAdd(p.n_groups, 1ull); return static_cast<long long>(s); }
      if (old == key) return static_cast<long long>(s);
    }
    s = (s + 1) & mask;
  }
  return -1;
}
__device__ __forceinline__ long long find_slot_packed(const GroupParams& p, long long row) {
  if (p.n_keys == 0) return 0;
  const bool isn = bit_at(p.key_nulls[0], row);
  return find_slot_packed_kv(p, isn, isn ? 0ull : load_raw(p.key_data[0], p.key_phys[0], row));
}

// kv[c] = raw value of key column c (0 where NULL), knull bit c = column c is NULL.
__device__ __forceinline__ long long find_slot_generic_kv(const GroupParams& p, const unsigned long long* kv, uint32_t knull) {
  unsigned long long h = 0x9E3779B97F4A7C15ull;
  for (int c = 0; c < p.n_keys; ++c) {
    const bool isn = (knull >> c) & 1u;
    h = mix64(h ^ (kv[c] + (isn ? 0xdeadbabeull : 0ull))) + c;
  }
  const unsigned long long mask = p.capacity - 1;
  unsigned long long s = h & mask;
  for (int probe = 0; probe < kProbeLimit; ++probe) {
    uint32_t st = *reinterpret_cast<volatile uint32_t*>(&p.slot_state[s]);
    if (st == 0u) {
      const uint32_t old = atomicCAS(&p.slot_state[s], 0u, 1u);
      if (old == 0u) {
        for (int c = 0; c < p.n_keys; ++c) p.key_store[c][s] = kv[c];
        p.key_store_null[s] = knull;
        __threadfence();
        *reinterpret_cast<volatile uint32_t*>(&p.slot_state[s]) = 2u;
        atomicAdd(p.n_groups, 1ull);
        return static_cast<long long>(s);
      }
      st = old;
    }
    while (st == 1u) st = *reinterpret_cast<volatile uint32_t*>(&p.slot_state[s]);
    __threadfence();
    bool same = *reinterpret_cast<volatile uint32_t*>(&p.key_store_null[s]) == knull;
    for (int c = 0; same && c < p.n_keys; ++c) {
      same = *reinterpret_cast<volatile unsigned long long*>(&p.key_store[c][s]) == kv[c];
    }
    if (same) return static_cast<long long>(s);
    s = (s + 1) & mask;
  }
  return -1;
}
__device__ __forceinline__ long long find_slot_generic(const GroupParams& p, long long row) {
  unsigned long long kv[kMaxKeys];
  uint32_t knull = 0;
  for (int c = 0; c < p.n_keys; ++c) {
    const bool isn = bit_at(p.key_nulls[c], row);
    kv[c] = isn ? 0ull : load_raw(p.key_data[c], p.key_phys[c], row);
    if (isn) knull |= 1u << c;
  }
  return find_slot_generic_kv(p, kv, knull);
}

// ---- accumulation -------------------------------------------------------------------------
__device__ __forceinline__ void atomic_min_max_f64(unsigned long long* addr, double v, bool is_min) {
  unsigned long long old = *addr;
  for (;;) {
    const double cur = __longlong_as_double(static_cast<long long>(old));
    if (is_min ? !(v < cur) : !(cur < v)) return;
    const unsigned long long prev = atomicCAS(addr, old, static_cast<unsigned long long>(__double_as_longlong(v)));
    if (prev == old) return;
    old = prev;
  }
}
__device__ __forceinline__ void atomic_add_f32(unsigned long long* addr, float v) {
  atomicAdd(reinterpret_cast<float*>(addr), v);   // low 32 bits of the container (little endian)
}
__device__ __forceinline__ void atomic_min_max_f32(unsigned long long* addr, float v, bool is_min) {
  unsigned int* a = reinterpret_cast<unsigned int*>(addr);
  unsigned int old = *a;
  for (;;) {
    const float cur = __uint_as_float(old);
    if (is_min ? !(v < cur) : !(cur < v)) return;
    const unsigned int prev = atomicCAS(a, old, __float_as_uint(v));
    if (prev == old) return;
    old = prev;
  }
}

// Applies one (already converted) value to the accumulator of `slot`.
__device__ __forceinline__ void apply(const AggDev& a, long long slot, unsigned long long v, unsigned long long count) {
  unsigned long long* dst = &a.acc[static_cast<unsigned long long>(slot) * a.stride];
  switch (a.fn) {
    case SSB_AGG_COUNT: atomicAdd(dst, count); break;
    case SSB_AGG_SUM:
      if (a.out_phys == T_F64) atomicAdd(reinterpret_cast<double*>(dst), Codec<double>::dec(v));
      else if (a.out_phys == T_F32) atomic_add_f32(dst, Codec<float>::dec(v));
      else atomicAdd(dst, a.out_phys == T_I32 ? static_cast<unsigned long long>(static_cast<long long>(Codec<int32_t>::dec(v))) : v);
      break;
    case SSB_AGG_MIN:
    case SSB_AGG_MAX: {
      const bool is_min = a.fn == SSB_AGG_MIN;
      switch (a.out_phys) {
        case T_F64: atomic_min_max_f64(dst, Codec<double>::dec(v), is_min); break;
        case T_F32: atomic_min_max_f32(dst, Codec<float>::dec(v), is_min); break;
        case T_I64: { long long x = Codec<int64_t>::dec(v); if (is_min) atomicMin(reinterpret_cast<long long*>(dst), x); else atomicMax(reinterpret_cast<long long*>(dst), x); } break;
        case T_I32: { long long x = Codec<int32_t>::dec(v); if (is_min) atomicMin(reinterpret_cast<long long*>(dst), x); else atomicMax(reinterpret_cast<long long*>(dst), x); } break;
        default: if (is_min) atomicMin(dst, v); else atomicMax(dst, v); break;   // U32 / U64 / B8 containers
      }
    } break;
    default: break;
  }
}

// Combines two partial values of one aggregate (warp pre-aggregation).
__device__ __forceinline__ unsigned long long combine(const AggDev& a, unsigned long long x, unsigned long long y) {
  switch (a.fn) {
    case SSB_AGG_SUM:
      if (a.out_phys == T_F64) return Codec<double>::enc(Codec<double>::dec(x) + Codec<double>::dec(y));
      if (a.out_phys == T_F32) return Codec<float>::enc(Codec<float>::dec(x) + Codec<float>::dec(y));
      if (a.out_phys == T_I32) return Codec<int32_t>::enc(Arith<int32_t>::add(Codec<int32_t>::dec(x), Codec<int32_t>::dec(y)));
      return x + y;
    case SSB_AGG_MIN:
    case SSB_AGG_MAX: {
      const bool is_min = a.fn == SSB_AGG_MIN;
      bool take_y;
      switch (a.out_phys) {
        case T_F64: take_y = is_min ? Codec<double>::dec(y) < Codec<double>::dec(x) : Codec<double>::dec(x) < Codec<double>::dec(y); break;
        case T_F32: take_y = is_min ? Codec<float>::dec(y) < Codec<float>::dec(x) : Codec<float>::dec(x) < Codec<float>::dec(y); break;
        case T_I64: take_y = is_min ? Codec<int64_t>::dec(y) < Codec<int64_t>::dec(x) : Codec<int64_t>::dec(x) < Codec<int64_t>::dec(y); break;
        case T_I32: take_y = is_min ? Codec<int32_t>::dec(y) < Codec<int32_t>::dec(x) : Codec<int32_t>::dec(x) < Codec<int32_t>::dec(y); break;
        default: take_y = is_min ? y < x : x < y; break;
      }
      return take_y ? y : x;
    }
    default: return x + y;
  }
}


// ---- few groups: shared constants of the per-thread-accumulator kernels -----------------------
enum { kTinyGroups = 8, kTinyThreads = 128, kLocalMaxAggs = 7 };
// AggDev::pad carries a pre-decoded accumulate code (set by the host).
enum { TA_COUNT = 0, TA_SUM_F64 = 1, TA_SUM_U64 = 2, TA_OTHER = 3 };

__device__ __forceinline__ unsigned long long identity_dev(const AggDev& ag) {
  if (ag.fn == SSB_AGG_MIN) {
    return ag.out_phys == T_F64 ? Codec<double>::enc(__longlong_as_double(0x7ff0000000000000LL))
         : ag.out_phys == T_F32 ? Codec<float>::enc(__uint_as_float(0x7f800000u))
         : (ag.out_phys == T_I64 || ag.out_phys == T_I32) ? static_cast<unsigned long long>(INT64_MAX) : ~0ull;
  }
  if (ag.fn == SSB_AGG_MAX) {
    return ag.out_phys == T_F64 ? Codec<double>::enc(__longlong_as_double(0xfff0000000000000LL))
         : ag.out_phys == T_F32 ? Codec<float>::enc(__uint_as_float(0xff800000u))
         : (ag.out_phys == T_I64 || ag.out_phys == T_I32) ? static_cast<unsigned long long>(INT64_MIN) : 0ull;
  }
  return 0ull;
}


}  // namespace ssb
#endif  // SSB_CSRC_GROUP_DEVICE_H_
