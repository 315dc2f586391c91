// join.cu -- hash join build and probe on the GPU.
//
// Replaces HashIndexOnMaterializedCursor::{MaterializeInputAndBuildIndex, MultiLookup} and
// ResultCursor (cursor/core/hash_join.cc:604-625,707-831) over RowHashSet / RowHashMultiSet
// (cursor/infrastructure/row_hash_set.cc:424-608):
//   build   open-addressed table of {key word, first build row}; a slot is claimed with
//           atomicCAS on its row field, duplicates of a key keep the smallest row as the head.
//           NOT_UNIQUE keys: the build rows are additionally grouped per slot with a stable
//           radix sort of (slot, row), giving per key a contiguous run in insertion order.
//   probe   pass 1 looks every lhs row up and writes its match count, an exclusive scan turns
//           counts into output offsets, pass 2 writes the (lhs row, rhs row) pairs. The output
//           is therefore in lhs order and, per lhs row, in build insertion order, exactly the
//           order the reference emits (hash_join.cc:793-831).
// Rows with a NULL in any key column never match (hash_join.cc:67-76,616-617,755-756).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <vector>

#include "comm.h"
#include "common.h"
#include "device_utils.h"

namespace ssb {

enum { kJoinMaxKeys = 8 };

struct JoinKeys {
  int32_t n_keys;
  int32_t phys[kJoinMaxKeys];
  const void* data[kJoinMaxKeys];
  const uint32_t* nulls[kJoinMaxKeys];
};

struct JoinTable {
  unsigned long long capacity;    // power of two
  // slot s = 16 bytes {first key column's raw bits of the head row, head build row + 1 (0 = empty)}:
  // a probe step is one 128-bit load = one DRAM access (two separate arrays cost two, and the
  // probe kernel is bound by random DRAM accesses: profiles/r1b_summary.md)
  unsigned long long* slots;
  // multiset
  unsigned long long* run_start;  // [capacity] offset into run_rows
  unsigned int* run_count;        // [capacity]
  long long* run_rows;            // build rows grouped by slot, insertion order
  // Dense integer keys (UNIQUE, one integer key column whose values span at most 4 x rows, e.g. a surrogate primary
  // key): no hashing and no slots, dense_rows[key - dense_min] = smallest build row with that key (kDenseEmpty = none).
  // 4 bytes per key value instead of a 16-byte slot at load 0.25-0.5: the C4 table is 50 MB (L2 resident) instead of
  // 537 MB, and a probe is one 4-byte load.
  unsigned int* dense_rows;
  unsigned int* dense_present;   // bit (key - dense_min): the key has a build row (range / 8 bytes: stays in L2)
  long long dense_min;
  unsigned long long dense_range;
};
enum : unsigned int { kDenseEmpty = 0xffffffffu };

// The replicated form of the sharded join: the build side is hash-partitioned over the
// ranks, every rank builds the table of its part, the tables are all-gathered, and a probe looks a
// key up in the table of the key's part. rhs rows are reported as row_offset[part] + row.
enum { kJoinMaxParts = 16 };
struct JoinParts {
  int32_t n_parts;                                  // 0 = one local table (JoinTable)
  const unsigned long long* slots[kJoinMaxParts];
  unsigned long long capacity[kJoinMaxParts];
  long long row_offset[kJoinMaxParts];
};

// Home slot of a key hash in a table of `capacity` slots (any capacity below 2^32, not only powers of two:
// the replicated tables of the sharded join travel over NVLink and are sized rows / 0.6): the low 32 bits
// of the hash scaled to [0, capacity). The high 32 bits choose the hash part (part_id_kernel).
__device__ __forceinline__ unsigned long long home_slot(unsigned long long h, unsigned long long capacity) {
  return ((h & 0xffffffffull) * capacity) >> 32;
}
__device__ __forceinline__ unsigned long long next_slot(unsigned long long s, unsigned long long capacity) {
  return s + 1 == capacity ? 0ull : s + 1;
}

__device__ __forceinline__ unsigned long long jmix64(unsigned long long x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}
__device__ __forceinline__ unsigned long long jload(const void* base, int phys, long long i) {
  switch (phys_width(phys)) {
    case 8: return static_cast<const unsigned long long*>(base)[i];
    // INT32 is sign-extended so that integer keys of different widths compare by value
    case 4: return phys == T_I32 ? static_cast<unsigned long long>(static_cast<long long>(static_cast<const int32_t*>(base)[i]))
                                 : static_cast<unsigned long long>(static_cast<const uint32_t*>(base)[i]);
    default: return static_cast<const uint8_t*>(base)[i];
  }
}
// Rows that can never match: a NULL in any key column (hash_join.cc:67-76,616-617,755-756), and a NaN in a
// floating-point key column -- the reference's hash set confirms a hit with operator== (row_hash_set.cc:
// 487-498), which no NaN satisfies, so a NaN key finds nothing and is found by nothing.
__device__ __forceinline__ bool key_has_null(const JoinKeys& k, long long row) {
  for (int c = 0; c < k.n_keys; ++c) {
    if (k.nulls[c] != nullptr && ((k.nulls[c][row >> 5] >> (row & 31)) & 1u)) return true;
    if (k.phys[c] == T_F64) { const double v = static_cast<const double*>(k.data[c])[row]; if (v != v) return true; }
    if (k.phys[c] == T_F32) { const float v = static_cast<const float*>(k.data[c])[row]; if (v != v) return true; }
  }
  return false;
}
__device__ __forceinline__ unsigned long long key_hash(const JoinKeys& k, long long row, unsigned long long* first) {
  unsigned long long h = 0x9E3779B97F4A7C15ull;
  for (int c = 0; c < k.n_keys; ++c) {
    const unsigned long long v = jload(k.data[c], k.phys[c], row);
    if (c == 0) *first = v;
    h = jmix64(h ^ v) + c;
  }
  return h;
}
__device__ __forceinline__ bool keys_equal(const JoinKeys& a, long long ra, const JoinKeys& b, long long rb) {
  for (int c = 1; c < a.n_keys; ++c) {   // column 0 was compared through slot_key
    if (jload(a.data[c], a.phys[c], ra) != jload(b.data[c], b.phys[c], rb)) return false;
  }
  return true;
}

// Finds the slot holding the key of (keys,row); -1 when absent. `build` are the build keys.
__device__ __forceinline__ long long lookup(const JoinTable& t, const JoinKeys& build, const JoinKeys& keys, long long row,
                                            long long* head_row) {
  unsigned long long first = 0;
  unsigned long long s = home_slot(key_hash(keys, row, &first), t.capacity);
  for (;;) {
    const ulonglong2 e = __ldg(reinterpret_cast<const ulonglong2*>(t.slots) + s);   // read-only during a probe
    const long long head = static_cast<long long>(e.y);
    if (head == 0) return -1;
    if (e.x == first && keys_equal(keys, row, build, head - 1)) { *head_row = head - 1; return static_cast<long long>(s); }
    s = next_slot(s, t.capacity);
  }
}

// Build: every non-NULL-key row finds or claims the slot of its key. slot_of[row] = slot, or
// -1 for NULL keys. The head of a slot ends as the smallest row with that key.
__global__ void __launch_bounds__(256) join_build_kernel(JoinTable t, JoinKeys build, long long rows,
                                                          long long* __restrict__ slot_of) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long row = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; row < rows; row += stride) {
    if (key_has_null(build, row)) { if (slot_of) slot_of[row] = -1; continue; }
    unsigned long long first = 0;
    unsigned long long s = home_slot(key_hash(build, row, &first), t.capacity);
    for (;;) {
      long long* slot_row = reinterpret_cast<long long*>(&t.slots[2 * s + 1]);
      unsigned long long* slot_key = &t.slots[2 * s];
      long long head = *reinterpret_cast<volatile long long*>(slot_row);
      if (head == 0) {
        // claim with a negative marker, publish the key word, then publish the row
        const long long old = static_cast<long long>(atomicCAS(reinterpret_cast<unsigned long long*>(slot_row), 0ull,
                                                               static_cast<unsigned long long>(-(row + 1))));
        if (old == 0) {
          *slot_key = first;
          __threadfence();
          atomicExch(reinterpret_cast<unsigned long long*>(slot_row), static_cast<unsigned long long>(row + 1));
          if (slot_of) slot_of[row] = static_cast<long long>(s);
          break;
        }
        head = old;
      }
      while (head < 0) head = *reinterpret_cast<volatile long long*>(slot_row);   // being published
      __threadfence();
      if (*reinterpret_cast<volatile unsigned long long*>(slot_key) == first && keys_equal(build, row, build, head - 1)) {
        // same key: keep the smallest row as head (insertion order of the reference)
        atomicMin(slot_row, row + 1);
        if (slot_of) slot_of[row] = static_cast<long long>(s);
        break;
      }
      s = next_slot(s, t.capacity);
    }
  }
}

// Smallest and largest key of the build side (NULL keys skipped), as signed 64-bit values (the sign- / zero-extended
// containers of jload): mm[0] = min, mm[1] = max, mm[2] = number of non-NULL keys.
__global__ void __launch_bounds__(256) join_minmax_kernel(JoinKeys build, long long rows, long long* __restrict__ mm) {
  long long lo = INT64_MAX, hi = INT64_MIN, cnt = 0;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long row = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; row < rows; row += stride) {
    if (key_has_null(build, row)) continue;
    const long long k = static_cast<long long>(jload(build.data[0], build.phys[0], row));
    lo = k < lo ? k : lo;
    hi = k > hi ? k : hi;
    ++cnt;
  }
  for (int d = 16; d > 0; d >>= 1) {
    const long long ol = __shfl_xor_sync(0xffffffffu, lo, d), oh = __shfl_xor_sync(0xffffffffu, hi, d);
    lo = ol < lo ? ol : lo;
    hi = oh > hi ? oh : hi;
    cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
  }
  if ((threadIdx.x & 31) == 0 && cnt > 0) {
    atomicMin(&mm[0], lo);
    atomicMax(&mm[1], hi);
    atomicAdd(reinterpret_cast<unsigned long long*>(&mm[2]), static_cast<unsigned long long>(cnt));
  }
}
__global__ void __launch_bounds__(256) join_build_dense_kernel(JoinTable t, JoinKeys build, long long rows) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long row = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; row < rows; row += stride) {
    if (key_has_null(build, row)) continue;
    const long long k = static_cast<long long>(jload(build.data[0], build.phys[0], row));
    // duplicates of a key keep the smallest row as the head, like the slot table
    const unsigned long long idx = static_cast<unsigned long long>(k - t.dense_min);
    atomicMin(&t.dense_rows[idx], static_cast<unsigned int>(row));
    atomicOr(&t.dense_present[idx >> 5], 1u << (idx & 31));
  }
}
// The materialising probe over a dense index reads its rhs result columns BY KEY: dst[key - min] = src[head row of the
// key] (zero where the key has no row), so a probe row costs one random access per rhs column and a presence bit,
// not an index access followed by a gather.
__global__ void __launch_bounds__(256) join_dense_spread_kernel(JoinTable t, const void* __restrict__ src, int w, void* __restrict__ dst) {
  const unsigned long long stride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
  for (unsigned long long idx = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < t.dense_range; idx += stride) {
    const unsigned int r = t.dense_rows[idx];
    const long long i = r == kDenseEmpty ? -1 : static_cast<long long>(r);
    if (w == 8) static_cast<unsigned long long*>(dst)[idx] = i >= 0 ? static_cast<const unsigned long long*>(src)[i] : 0ull;
    else if (w == 4) static_cast<uint32_t*>(dst)[idx] = i >= 0 ? static_cast<const uint32_t*>(src)[i] : 0u;
    else static_cast<unsigned char*>(dst)[idx] = i >= 0 ? static_cast<const unsigned char*>(src)[i] : static_cast<unsigned char>(0);
  }
}

// NOTE on atomicMin above: heads are positive once published, so min keeps the smallest row;
// a concurrent reader may compare keys against either row of the same key - both are equal.

__global__ void run_bounds_kernel(const unsigned long long* __restrict__ sorted_slots, long long n_valid,
                                  unsigned long long capacity, unsigned long long* __restrict__ run_start,
                                  unsigned int* __restrict__ run_count) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_valid; i += stride) {
    const unsigned long long s = sorted_slots[i];
    if (s >= capacity) continue;   // NULL-key marker (== capacity): owns no run, and the arrays hold capacity entries
    if (i == 0 || sorted_slots[i - 1] != s) run_start[s] = static_cast<unsigned long long>(i);
    atomicAdd(&run_count[s], 1u);
  }
}

__global__ void slot_keys_kernel(const long long* __restrict__ slot_of, long long rows, unsigned long long capacity,
                                 unsigned long long* __restrict__ keys, long long* __restrict__ vals) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < rows; i += stride) {
    keys[i] = slot_of[i] < 0 ? capacity : static_cast<unsigned long long>(slot_of[i]);   // NULL keys sort last
    vals[i] = i;
  }
}

// Probe pass 1: match slot and output count per lhs row.
__global__ void __launch_bounds__(256) join_count_kernel(JoinTable t, JoinKeys build, JoinKeys probe, long long rows,
                                                          int unique, int left_outer, long long* __restrict__ slot_of,
                                                          unsigned long long* __restrict__ counts) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long row = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; row < rows; row += stride) {
    long long s = -1, head = -1;
    if (!key_has_null(probe, row)) s = lookup(t, build, probe, row, &head);
    // UNIQUE keys: remember the matching build row itself, the emit pass then never touches the table
    slot_of[row] = unique ? (s >= 0 ? head : -1) : s;
    unsigned long long c = 0;
    if (s >= 0) c = unique ? 1ull : t.run_count[s];
    else if (left_outer) c = 1ull;
    counts[row] = c;
  }
}

// Probe pass 2: write the pairs at the scanned offsets.
__global__ void __launch_bounds__(256) join_emit_kernel(JoinTable t, long long rows, int unique,
                                                         const long long* __restrict__ slot_of,
                                                         const unsigned long long* __restrict__ offsets,
                                                         long long* __restrict__ lhs_out, long long* __restrict__ rhs_out) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long row = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; row < rows; row += stride) {
    const long long s = slot_of[row];
    const unsigned long long off = offsets[row];
    if (s < 0) {
      // counts were 0 (INNER) or 1 (LEFT_OUTER); offsets of the next row tell which
      const unsigned long long next = offsets[row + 1];
      if (next != off) { lhs_out[off] = row; rhs_out[off] = -1; }
    } else if (unique) {
      lhs_out[off] = row;
      rhs_out[off] = s;   // the count pass stored the build row
    } else {
      const unsigned long long start = t.run_start[s];
      const unsigned int cnt = t.run_count[s];
      for (unsigned int j = 0; j < cnt; ++j) { lhs_out[off + j] = row; rhs_out[off + j] = t.run_rows[start + j]; }
    }
  }
}


// UNIQUE keys: a lhs row yields at most one pair, so the probe is a stream compaction and runs
// as ONE pass: a CTA takes the next 1024-row tile (atomic ticket), every thread looks up four
// rows (the four table probes are in flight together), the matches are ranked in row order with
// ballots, the tile's output offset comes from the decoupled look-back over the tiles' match
// counts (device_utils.h), and the pairs are written at their final positions. LEFT_OUTER emits
// every lhs row, so positions are the row numbers and no prefix is needed.
// aux: [0] ticket, [1] total pairs, [2 ..] one status word per tile.
enum { kProbeThreads = 256, kProbeRows = 4, kProbeTile = kProbeThreads * kProbeRows };   // 8 rows per thread measured slower (114 registers: 9.7 ms vs 7.6 ms per 200M probes)

// EMIT: the probe writes the result COLUMNS itself (lhs columns gathered by the probe row, rhs columns by the matched
// build row) at the pairs' final positions, instead of the two row-id lists a gather per column would read back:
// per probe row 16 bytes less written and 16 + 8 per lhs column less read.
enum { kEmitMax = 8 };
struct JoinEmit {
  int32_t n_l, n_r;
  const void* l_src[kEmitMax]; void* l_dst[kEmitMax]; int32_t l_w[kEmitMax];
  const void* r_src[kEmitMax]; void* r_dst[kEmitMax]; int32_t r_w[kEmitMax];
  unsigned char* matched;   // LEFT_OUTER: 1 = the row found a build row (else the rhs cells are zero), or nullptr
  int32_t by_key;           // dense index: r_src are indexed by key - dense_min (join_dense_spread_kernel), not by build row
};
__device__ __forceinline__ void emit_cell(void* dst, unsigned long long o, const void* src, long long i, int w) {
  if (w == 8) static_cast<unsigned long long*>(dst)[o] = i >= 0 ? static_cast<const unsigned long long*>(src)[i] : 0ull;
  else if (w == 4) static_cast<uint32_t*>(dst)[o] = i >= 0 ? static_cast<const uint32_t*>(src)[i] : 0u;
  else static_cast<unsigned char*>(dst)[o] = i >= 0 ? static_cast<const unsigned char*>(src)[i] : static_cast<unsigned char>(0);
}

// A result cell read early (its load is in flight while the tile's ranks and prefix are computed) and written late.
enum { kEmitEarly = 2 };   // columns per side whose cells are read early (registers: 2 sides x 2 columns x 4 rows x 8 bytes)
__device__ __forceinline__ unsigned long long load_cell(const void* src, long long i, int w) {
  if (i < 0) return 0ull;
  if (w == 8) return static_cast<const unsigned long long*>(src)[i];
  if (w == 4) return static_cast<const uint32_t*>(src)[i];
  return static_cast<const unsigned char*>(src)[i];
}
__device__ __forceinline__ void store_cell(void* dst, unsigned long long o, unsigned long long v, int w) {
  if (w == 8) static_cast<unsigned long long*>(dst)[o] = v;
  else if (w == 4) static_cast<uint32_t*>(dst)[o] = static_cast<uint32_t>(v);
  else static_cast<unsigned char*>(dst)[o] = static_cast<unsigned char>(v);
}

template <bool PARTS, bool EMIT, bool DENSE>
__global__ void __launch_bounds__(kProbeThreads) join_probe_unique_kernel(JoinTable t, JoinKeys build, JoinKeys probe,
                                                                           long long rows, int left_outer,
                                                                           long long* __restrict__ lhs_out,
                                                                           long long* __restrict__ rhs_out,
                                                                           unsigned long long* __restrict__ aux,
                                                                           const __grid_constant__ JoinParts parts,
                                                                           const __grid_constant__ JoinEmit emit) {
  __shared__ unsigned int s_tile;
  __shared__ unsigned int wcnt[kProbeRows][kProbeThreads / 32];
  __shared__ unsigned long long s_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long tiles = (rows + kProbeTile - 1) / kProbeTile;
  for (;;) {
    __syncthreads();   // s_tile / wcnt / s_base of the previous tile are no longer read
    if (tid == 0) s_tile = atomicAdd(reinterpret_cast<unsigned int*>(aux), 1u);
    __syncthreads();
    const long long tile = s_tile;
    if (tile >= tiles) return;
    const long long r0 = tile * kProbeTile + tid;
    long long head[kProbeRows];
    unsigned int pos[kProbeRows];
    unsigned long long first[kProbeRows], slot[kProbeRows];
    ulonglong2 ent[kProbeRows];
    bool live[kProbeRows];
    // the table of the row's key: the one local table, or the table of the key's hash part
    const unsigned long long* tab[kProbeRows];
    unsigned long long cap[kProbeRows];
    long long offset[kProbeRows];
    if (DENSE) {
      // dense integer keys: the key is the index; the four loads are in flight together
      unsigned int dr[kProbeRows];
#pragma unroll
      for (int j = 0; j < kProbeRows; ++j) {
        const long long row = r0 + j * kProbeThreads;
        dr[j] = kDenseEmpty;
        if (row < rows && !key_has_null(probe, row)) {
          const unsigned long long idx = static_cast<unsigned long long>(static_cast<long long>(jload(probe.data[0], probe.phys[0], row)) - t.dense_min);
          if (idx < t.dense_range) {
            if (EMIT && emit.by_key) dr[j] = ((__ldg(&t.dense_present[idx >> 5]) >> (idx & 31)) & 1u) ? static_cast<unsigned int>(idx) : kDenseEmpty;
            else dr[j] = __ldg(&t.dense_rows[idx]);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < kProbeRows; ++j) head[j] = dr[j] == kDenseEmpty ? -1 : static_cast<long long>(dr[j]);
    } else {
    // hash and first table probe of all four rows before any of them is looked at
#pragma unroll
    for (int j = 0; j < kProbeRows; ++j) {
      const long long row = r0 + j * kProbeThreads;
      live[j] = row < rows && !key_has_null(probe, row);
      first[j] = 0;
      const unsigned long long h = live[j] ? key_hash(probe, row, &first[j]) : 0ull;
      if (PARTS) {
        const unsigned int part = static_cast<unsigned int>(((h >> 32) * static_cast<unsigned long long>(parts.n_parts)) >> 32);   // = part_id_kernel
        tab[j] = parts.slots[part];
        cap[j] = parts.capacity[part];
        offset[j] = parts.row_offset[part];
      } else {
        tab[j] = t.slots;
        cap[j] = t.capacity;
        offset[j] = 0;
      }
      slot[j] = home_slot(h, cap[j]);
    }
#pragma unroll
    for (int j = 0; j < kProbeRows; ++j) ent[j] = __ldg(reinterpret_cast<const ulonglong2*>(tab[j]) + slot[j]);
#pragma unroll
    for (int j = 0; j < kProbeRows; ++j) {
      const long long row = r0 + j * kProbeThreads;
      head[j] = -1;
      if (!live[j]) continue;
      ulonglong2 e = ent[j];
      unsigned long long sl = slot[j];
      for (;;) {
        const long long h = static_cast<long long>(e.y);
        if (h == 0) break;
        // PARTS: single-column keys only (host-checked), the slot word is the whole key
        if (e.x == first[j] && (PARTS || keys_equal(probe, row, build, h - 1))) { head[j] = offset[j] + h - 1; break; }
        sl = next_slot(sl, cap[j]);
        e = __ldg(reinterpret_cast<const ulonglong2*>(tab[j]) + sl);
      }
    }
    }   // !DENSE
    // EMIT: the first cells of every pair are requested now -- lhs cells depend on the row only, rhs cells on the
    // match -- so their (random, for the rhs) loads overlap the ranking and the look-back below instead of following them
    unsigned long long lv_early[kEmitEarly][kProbeRows], rv_early[kEmitEarly][kProbeRows];
    if (EMIT) {
#pragma unroll
      for (int c = 0; c < kEmitEarly; ++c) {
#pragma unroll
        for (int j = 0; j < kProbeRows; ++j) {
          const long long row = r0 + j * kProbeThreads;
          const bool wanted = row < rows && (left_outer || head[j] >= 0);
          lv_early[c][j] = (c < emit.n_l && wanted) ? load_cell(emit.l_src[c], row, emit.l_w[c]) : 0ull;
          rv_early[c][j] = (c < emit.n_r && wanted) ? load_cell(emit.r_src[c], head[j], emit.r_w[c]) : 0ull;
        }
      }
    }
    if (left_outer) {
#pragma unroll
      for (int j = 0; j < kProbeRows; ++j) {
        const long long row = r0 + j * kProbeThreads;
        if (row >= rows) continue;
        if (EMIT) {
#pragma unroll
          for (int c = 0; c < kEmitEarly; ++c) {
            if (c < emit.n_l) store_cell(emit.l_dst[c], static_cast<unsigned long long>(row), lv_early[c][j], emit.l_w[c]);
            if (c < emit.n_r) store_cell(emit.r_dst[c], static_cast<unsigned long long>(row), rv_early[c][j], emit.r_w[c]);
          }
          for (int c = kEmitEarly; c < emit.n_l; ++c) emit_cell(emit.l_dst[c], static_cast<unsigned long long>(row), emit.l_src[c], row, emit.l_w[c]);
          for (int c = kEmitEarly; c < emit.n_r; ++c) emit_cell(emit.r_dst[c], static_cast<unsigned long long>(row), emit.r_src[c], head[j], emit.r_w[c]);
          if (emit.matched != nullptr) emit.matched[row] = head[j] >= 0 ? 1 : 0;
        } else {
          lhs_out[row] = row;
          rhs_out[row] = head[j];
        }
      }
      if (tile == tiles - 1 && tid == 0) aux[1] = static_cast<unsigned long long>(rows);
      continue;
    }
    // rank of every match inside the tile, in row order: round j before round j + 1, then thread order
#pragma unroll
    for (int j = 0; j < kProbeRows; ++j) {
      const unsigned m = __ballot_sync(0xffffffffu, head[j] >= 0);
      pos[j] = __popc(m & ((1u << lane) - 1u));
      if (lane == 0) wcnt[j][warp] = __popc(m);
    }
    __syncthreads();
    unsigned int before[kProbeRows], total = 0;
#pragma unroll
    for (int j = 0; j < kProbeRows; ++j) {
      before[j] = total;
#pragma unroll
      for (int w = 0; w < kProbeThreads / 32; ++w) {
        const unsigned int c = wcnt[j][w];
        if (w < warp) before[j] += c;
        total += c;
      }
    }
    if (warp == 0) {
      const unsigned long long excl = tile_prefix_warp(aux + 2, static_cast<unsigned long long>(tile), total, lane);
      if (lane == 0) {
        s_base = excl;
        if (tile == tiles - 1) aux[1] = excl + total;
      }
    }
    __syncthreads();
    const unsigned long long base = s_base;
#pragma unroll
    for (int j = 0; j < kProbeRows; ++j) {
      if (head[j] >= 0) {
        const unsigned long long o = base + before[j] + pos[j];
        if (EMIT) {
          const long long row = r0 + j * kProbeThreads;
#pragma unroll
          for (int c = 0; c < kEmitEarly; ++c) {
            if (c < emit.n_l) store_cell(emit.l_dst[c], o, lv_early[c][j], emit.l_w[c]);
            if (c < emit.n_r) store_cell(emit.r_dst[c], o, rv_early[c][j], emit.r_w[c]);
          }
          for (int c = kEmitEarly; c < emit.n_l; ++c) emit_cell(emit.l_dst[c], o, emit.l_src[c], row, emit.l_w[c]);
          for (int c = kEmitEarly; c < emit.n_r; ++c) emit_cell(emit.r_dst[c], o, emit.r_src[c], head[j], emit.r_w[c]);
        } else {
          lhs_out[o] = r0 + j * kProbeThreads;
          rhs_out[o] = head[j];
        }
      }
    }
  }
}

// Part id per row for the multi-GPU redistribution: the high bits of the key hash scaled to
// [0, n_parts) (the table slot uses the low bits, so parts and slots stay independent); rows
// with a NULL key column never match and stay on `null_part`. Also counts the rows per part.
__global__ void __launch_bounds__(256) part_id_kernel(JoinKeys keys, long long rows, unsigned int n_parts,
                                                       unsigned int null_part, unsigned long long* __restrict__ part_of,
                                                       long long* __restrict__ row_of,
                                                       unsigned long long* __restrict__ counts) {
  __shared__ unsigned int cnt[256];
  cnt[threadIdx.x] = 0;
  __syncthreads();
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long row = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; row < rows; row += stride) {
    unsigned int part = null_part;
    if (!key_has_null(keys, row)) {
      unsigned long long first = 0;
      const unsigned long long h = key_hash(keys, row, &first);
      part = static_cast<unsigned int>(((h >> 32) * n_parts) >> 32);
    }
    part_of[row] = part;
    row_of[row] = row;
    atomicAdd(&cnt[part], 1u);
  }
  __syncthreads();
  if (threadIdx.x <= n_parts && threadIdx.x < 256 && cnt[threadIdx.x]) {   // slot n_parts: NULL keys set aside (null_part == n_parts)
    atomicAdd(&counts[threadIdx.x], static_cast<unsigned long long>(cnt[threadIdx.x]));
  }
}

static bool dense_enabled() {
  static const bool on = getenv("SSB200_JOIN_DENSE") == nullptr || atoi(getenv("SSB200_JOIN_DENSE")) != 0;
  return on;
}

bool join_dense_fits(long long lo, long long hi, long long count, long long rows) {
  if (!dense_enabled() || count <= 0 || hi < lo || rows >= (1LL << 32) - 1) return false;
  const unsigned long long span = static_cast<unsigned long long>(hi) - static_cast<unsigned long long>(lo);   // max - min, no overflow
  return span < static_cast<unsigned long long>(rows) * 4 && span < (1ull << 32) - 2;
}

int join_key_range(ssb_ctx* ctx, const ssb_column* key, int64_t rows, long long out[3], int* eligible) {
  out[0] = INT64_MAX; out[1] = INT64_MIN; out[2] = 0;
  JoinKeys k;
  memset(&k, 0, sizeof(k));
  k.n_keys = 1;
  k.phys[0] = phys_of(key->dtype);
  k.data[0] = key->data;
  k.nulls[0] = key->nulls;
  *eligible = (k.phys[0] == T_I32 || k.phys[0] == T_I64 || k.phys[0] == T_U32) && dense_enabled() ? 1 : 0;
  if (!*eligible || rows <= 0) return 0;
  long long* mm = nullptr;
  cudaError_t e = tmp_malloc(ctx, &mm, 32);
  if (e != cudaSuccess) return cuda_fail(ctx, e, "join key range");
  cudaMemcpyAsync(mm, out, 24, cudaMemcpyHostToDevice, ctx->stream);
  join_minmax_kernel<<<grid_1d(ctx, rows, 256), 256, 0, ctx->stream>>>(k, rows, mm);
  ++ctx->launches;
  e = cudaMemcpyAsync(out, mm, 24, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  tmp_free(ctx, mm);
  if (e != cudaSuccess) return cuda_fail(ctx, e, "join key range");
  return 0;
}

}  // namespace ssb

using namespace ssb;

struct ssb_join {
  ssb_ctx* ctx;
  JoinKeys build_keys;
  long long build_rows;
  int uniqueness;
  JoinTable table;
  JoinParts parts;        // n_parts > 0: an index over tables built elsewhere (ssb_join_attach_parts); owns none of them
  long long* lhs_out;
  long long* rhs_out;
};

extern "C" {

void ssb_join_destroy(ssb_join* j) {
  if (!j) return;
  ssb_ctx* ctx = j->ctx;
  cudaStreamSynchronize(ctx->stream);
  if (j->parts.n_parts == 0) tmp_free(ctx, j->table.slots);
  tmp_free(ctx, j->table.run_start);
  tmp_free(ctx, j->table.run_count);
  tmp_free(ctx, j->table.run_rows);
  tmp_free(ctx, j->table.dense_rows);
  tmp_free(ctx, j->table.dense_present);
  tmp_free(ctx, j->lhs_out);
  tmp_free(ctx, j->rhs_out);
  delete j;
}

static int fill_keys(ssb_ctx* ctx, int32_t n_keys, const ssb_column* keys, JoinKeys* out) {
  if (n_keys < 1 || n_keys > kJoinMaxKeys) return fail(ctx, SSB_ERROR_NOT_IMPLEMENTED, "hash join needs 1..8 key columns");
  memset(out, 0, sizeof(*out));
  out->n_keys = n_keys;
  for (int c = 0; c < n_keys; ++c) {
    out->phys[c] = phys_of(keys[c].dtype);
    if (out->phys[c] < 0) return fail(ctx, SSB_ERROR_INVALID_ARGUMENT_TYPE, "unsupported join key type");
    out->data[c] = keys[c].data;
    out->nulls[c] = keys[c].nulls;
  }
  return 0;
}

int ssb_join_build(ssb_ctx* ctx, int32_t n_keys, const ssb_column* keys, int64_t rows, int32_t uniqueness,
                   ssb_join** out) {
  *out = nullptr;
  if (rows < 0) return fail(ctx, SSB_ERROR_INVALID_ARGUMENT_VALUE, "negative row count");
  ssb_join* j = new ssb_join();
  memset(j, 0, sizeof(*j));
  j->ctx = ctx;
  if (int rc = fill_keys(ctx, n_keys, keys, &j->build_keys)) { delete j; return rc; }
  j->build_rows = rows;
  j->uniqueness = uniqueness;
  // a power of two >= 2 x rows (load 0.25 .. 0.5); SSB_KEYS_COMPACT_TABLE: rows / 0.6, for tables that travel
  const bool compact = (uniqueness & SSB_KEYS_COMPACT_TABLE) != 0;
  uniqueness &= ~SSB_KEYS_COMPACT_TABLE;
  j->uniqueness = uniqueness;
  unsigned long long cap = 1024;
  if (compact) {
    cap = static_cast<unsigned long long>(rows) * 5 / 3 + 1024;
  } else {
    while (cap < static_cast<unsigned long long>(rows) * 2) cap *= 2;
  }
  if (cap >= (1ull << 32)) { delete j; return fail(ctx, SSB_ERROR_NOT_IMPLEMENTED, "hash join build side beyond 2^31 rows"); }
  j->table.capacity = cap;
  TimedRegion timed(ctx);
  // dense integer keys (see JoinTable): decided from the key range; SSB200_JOIN_DENSE=0 keeps the slot table
  if (uniqueness == SSB_KEYS_UNIQUE && !compact && n_keys == 1 && rows >= 4096) {
    long long range[3] = {0, 0, 0};
    int eligible = 0;
    if (int rc0 = join_key_range(ctx, &keys[0], rows, range, &eligible)) { ssb_join_destroy(j); return rc0; }
    if (eligible && join_dense_fits(range[0], range[1], range[2], rows)) {
      const unsigned long long span = static_cast<unsigned long long>(range[1]) - static_cast<unsigned long long>(range[0]);
      j->table.dense_min = range[0];
      j->table.dense_range = span + 1;
      cudaError_t e0 = tmp_malloc(ctx, &j->table.dense_rows, static_cast<size_t>(span + 1) * 4);
      if (e0 == cudaSuccess) e0 = tmp_malloc(ctx, &j->table.dense_present, static_cast<size_t>(span / 32 + 2) * 4);
      if (e0 != cudaSuccess) { ssb_join_destroy(j); return cuda_fail(ctx, e0, "dense join index"); }
      cudaMemsetAsync(j->table.dense_rows, 0xff, static_cast<size_t>(span + 1) * 4, ctx->stream);
      cudaMemsetAsync(j->table.dense_present, 0, static_cast<size_t>(span / 32 + 2) * 4, ctx->stream);
      join_build_dense_kernel<<<grid_1d(ctx, rows, 256), 256, 0, ctx->stream>>>(j->table, j->build_keys, rows);
      ++ctx->launches;
      e0 = cudaGetLastError();
      if (e0 == cudaSuccess) e0 = cudaStreamSynchronize(ctx->stream);
      if (e0 != cudaSuccess) { ssb_join_destroy(j); return cuda_fail(ctx, e0, "dense join build"); }
      *out = j;
      return 0;
    }
  }
  cudaError_t e = tmp_malloc(ctx, &j->table.slots, cap * 16);
  if (e != cudaSuccess) { ssb_join_destroy(j); return cuda_fail(ctx, e, "join table"); }
  cudaMemsetAsync(j->table.slots, 0, cap * 16, ctx->stream);
  long long* slot_of = nullptr;
  const bool multi = uniqueness != SSB_KEYS_UNIQUE;
  if (multi && rows > 0) {
    e = tmp_malloc(ctx, &slot_of, static_cast<size_t>(rows) * 8);
    if (e != cudaSuccess) { ssb_join_destroy(j); return cuda_fail(ctx, e, "join build scratch"); }
  }
  if (rows > 0) {
    join_build_kernel<<<grid_1d(ctx, rows, 256), 256, 0, ctx->stream>>>(j->table, j->build_keys, rows, slot_of);
    ++ctx->launches;
  }
  int rc = 0;
  if (multi) {
    e = tmp_malloc(ctx, &j->table.run_start, cap * 8);
    if (e == cudaSuccess) e = tmp_malloc(ctx, &j->table.run_count, cap * 4);
    if (e != cudaSuccess) { tmp_free(ctx, slot_of); ssb_join_destroy(j); return cuda_fail(ctx, e, "join runs"); }
    cudaMemsetAsync(j->table.run_count, 0, cap * 4, ctx->stream);
    if (rows > 0) {
      unsigned long long *k0 = nullptr, *k1 = nullptr;
      long long *v0 = nullptr, *v1 = nullptr;
      e = tmp_malloc(ctx, &k0, static_cast<size_t>(rows) * 8);
      if (e == cudaSuccess) e = tmp_malloc(ctx, &k1, static_cast<size_t>(rows) * 8);
      if (e == cudaSuccess) e = tmp_malloc(ctx, &v0, static_cast<size_t>(rows) * 8);
      if (e == cudaSuccess) e = tmp_malloc(ctx, &v1, static_cast<size_t>(rows) * 8);
      if (e != cudaSuccess) {
        tmp_free(ctx, k0); tmp_free(ctx, k1); tmp_free(ctx, v0); tmp_free(ctx, v1); tmp_free(ctx, slot_of);
        ssb_join_destroy(j);
        return cuda_fail(ctx, e, "join sort scratch");
      }
      slot_keys_kernel<<<grid_1d(ctx, rows, 256), 256, 0, ctx->stream>>>(slot_of, rows, cap, k0, v0);
      ++ctx->launches;
      int bits = 1;
      while ((1ull << bits) <= cap) ++bits;   // slots < cap, NULL marker == cap
      bits = (bits + 7) / 8 * 8;
      rc = radix_sort_pairs(ctx, &k0, &v0, &k1, &v1, static_cast<unsigned long long>(rows), 0, bits);
      if (rc == 0) {
        // rows with NULL keys (marker cap) sort last and own no run
        run_bounds_kernel<<<grid_1d(ctx, rows, 256), 256, 0, ctx->stream>>>(k0, rows, cap, j->table.run_start, j->table.run_count);
        ++ctx->launches;
      }
      j->table.run_rows = v0;
      cudaStreamSynchronize(ctx->stream);
      tmp_free(ctx, k0); tmp_free(ctx, k1); tmp_free(ctx, v1);
    }
  }
  e = cudaGetLastError();
  cudaStreamSynchronize(ctx->stream);
  tmp_free(ctx, slot_of);
  if (rc == 0 && e != cudaSuccess) rc = cuda_fail(ctx, e, "join build");
  if (rc) { ssb_join_destroy(j); return rc; }
  *out = j;
  return 0;
}

int ssb_join_probe(ssb_join* j, const ssb_column* keys, int64_t rows, int32_t join_type, int64_t* n_pairs,
                   const int64_t** d_lhs_rows, const int64_t** d_rhs_rows) {
  ssb_ctx* ctx = j->ctx;
  *n_pairs = 0;
  *d_lhs_rows = nullptr;
  *d_rhs_rows = nullptr;
  if (rows < 0) return fail(ctx, SSB_ERROR_INVALID_ARGUMENT_VALUE, "negative row count");
  if (join_type != SSB_JOIN_INNER && join_type != SSB_JOIN_LEFT_OUTER) return fail(ctx, SSB_ERROR_NOT_IMPLEMENTED, "join type");
  if (rows == 0) return 0;
  JoinKeys probe;
  if (int rc = fill_keys(ctx, j->build_keys.n_keys, keys, &probe)) return rc;
  for (int c = 0; c < probe.n_keys; ++c) {
    const int a = probe.phys[c], b = j->build_keys.phys[c];
    const bool ints = (a == T_I32 || a == T_I64 || a == T_U32 || a == T_U64) && (b == T_I32 || b == T_I64 || b == T_U32 || b == T_U64);
    if (a != b && !ints) return fail(ctx, SSB_ERROR_INVALID_ARGUMENT_TYPE, "probe key types differ from the build keys");
  }
  TimedRegion timed(ctx);
  if (j->uniqueness == SSB_KEYS_UNIQUE) {
    // single pass: pairs <= rows, so the result buffers are sized by the lhs
    const long long tiles = div_up(rows, kProbeTile);
    unsigned long long* aux = nullptr;
    tmp_free(ctx, j->lhs_out); tmp_free(ctx, j->rhs_out);
    j->lhs_out = j->rhs_out = nullptr;
    cudaError_t e = tmp_malloc(ctx, &aux, static_cast<size_t>(tiles + 2) * 8);
    if (e == cudaSuccess) e = tmp_malloc(ctx, &j->lhs_out, static_cast<size_t>(rows + 1) * 8);
    if (e == cudaSuccess) e = tmp_malloc(ctx, &j->rhs_out, static_cast<size_t>(rows + 1) * 8);
    if (e != cudaSuccess) { tmp_free(ctx, aux); return cuda_fail(ctx, e, "join probe buffers"); }
    cudaMemsetAsync(aux, 0, static_cast<size_t>(tiles + 2) * 8, ctx->stream);
    long long grid = static_cast<long long>(ctx->num_sms) * 4;
    if (grid > tiles) grid = tiles;
    JoinEmit no_emit;
    memset(&no_emit, 0, sizeof(no_emit));
    if (j->parts.n_parts > 0) {
      join_probe_unique_kernel<true, false, false><<<static_cast<unsigned>(grid), kProbeThreads, 0, ctx->stream>>>(
          j->table, j->build_keys, probe, rows, join_type == SSB_JOIN_LEFT_OUTER ? 1 : 0, j->lhs_out, j->rhs_out, aux, j->parts, no_emit);
    } else if (j->table.dense_rows != nullptr) {
      join_probe_unique_kernel<false, false, true><<<static_cast<unsigned>(grid), kProbeThreads, 0, ctx->stream>>>(
          j->table, j->build_keys, probe, rows, join_type == SSB_JOIN_LEFT_OUTER ? 1 : 0, j->lhs_out, j->rhs_out, aux, j->parts, no_emit);
    } else {
      join_probe_unique_kernel<false, false, false><<<static_cast<unsigned>(grid), kProbeThreads, 0, ctx->stream>>>(
          j->table, j->build_keys, probe, rows, join_type == SSB_JOIN_LEFT_OUTER ? 1 : 0, j->lhs_out, j->rhs_out, aux, j->parts, no_emit);
    }
    ++ctx->launches;
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->h_count, aux + 1, 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    tmp_free(ctx, aux);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "join probe");
    *n_pairs = *ctx->h_count;
    *d_lhs_rows = reinterpret_cast<const int64_t*>(j->lhs_out);
    *d_rhs_rows = reinterpret_cast<const int64_t*>(j->rhs_out);
    return 0;
  }
  long long* slot_of = nullptr;
  unsigned long long* counts = nullptr;
  unsigned long long* d_total = nullptr;
  cudaError_t e = tmp_malloc(ctx, &slot_of, static_cast<size_t>(rows) * 8);
  if (e == cudaSuccess) e = tmp_malloc(ctx, &counts, static_cast<size_t>(rows + 1) * 8);
  if (e == cudaSuccess) e = tmp_malloc(ctx, &d_total, 8);
  if (e != cudaSuccess) { tmp_free(ctx, slot_of); tmp_free(ctx, counts); tmp_free(ctx, d_total); return cuda_fail(ctx, e, "join probe scratch"); }
  const int unique = j->uniqueness == SSB_KEYS_UNIQUE ? 1 : 0;
  join_count_kernel<<<grid_1d(ctx, rows, 256), 256, 0, ctx->stream>>>(j->table, j->build_keys, probe, rows, unique,
                                                                       join_type == SSB_JOIN_LEFT_OUTER ? 1 : 0, slot_of, counts);
  ++ctx->launches;
  cudaMemsetAsync(counts + rows, 0, 8, ctx->stream);
  int rc = exclusive_scan_u64(ctx, counts, static_cast<unsigned long long>(rows) + 1, d_total);
  unsigned long long total = 0;
  if (rc == 0) {
    cudaMemcpyAsync(ctx->h_count, d_total, 8, cudaMemcpyDeviceToHost, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    total = static_cast<unsigned long long>(*ctx->h_count);
    tmp_free(ctx, j->lhs_out); tmp_free(ctx, j->rhs_out);
    j->lhs_out = j->rhs_out = nullptr;
    e = tmp_malloc(ctx, &j->lhs_out, (total + 1) * 8);
    if (e == cudaSuccess) e = tmp_malloc(ctx, &j->rhs_out, (total + 1) * 8);
    if (e != cudaSuccess) rc = cuda_fail(ctx, e, "join result");
  }
  if (rc == 0 && total > 0) {
    join_emit_kernel<<<grid_1d(ctx, rows, 256), 256, 0, ctx->stream>>>(j->table, rows, unique, slot_of, counts, j->lhs_out, j->rhs_out);
    ++ctx->launches;
    e = cudaGetLastError();
    if (e != cudaSuccess) rc = cuda_fail(ctx, e, "join probe");
  }
  cudaStreamSynchronize(ctx->stream);
  tmp_free(ctx, slot_of); tmp_free(ctx, counts); tmp_free(ctx, d_total);
  if (rc) return rc;
  *n_pairs = static_cast<int64_t>(total);
  *d_lhs_rows = reinterpret_cast<const int64_t*>(j->lhs_out);
  *d_rhs_rows = reinterpret_cast<const int64_t*>(j->rhs_out);
  return 0;
}

int ssb_join_probe_materialize(ssb_join* j, const ssb_column* keys, int64_t rows, int32_t join_type, int32_t n_lhs,
                               const ssb_column* lhs_cols, int32_t n_rhs, const ssb_column* rhs_cols, const ssb_column* out_cols,
                               uint8_t* d_matched, int64_t* n_rows) {
  ssb_ctx* ctx = j->ctx;
  *n_rows = 0;
  if (rows < 0) return fail(ctx, SSB_ERROR_INVALID_ARGUMENT_VALUE, "negative row count");
  if (join_type != SSB_JOIN_INNER && join_type != SSB_JOIN_LEFT_OUTER) return fail(ctx, SSB_ERROR_NOT_IMPLEMENTED, "join type");
  if (j->uniqueness != SSB_KEYS_UNIQUE) return fail(ctx, SSB_ERROR_NOT_IMPLEMENTED, "the materialising probe serves UNIQUE keys; use ssb_join_probe + ssb_gather");
  if (n_lhs < 0 || n_rhs < 0 || n_lhs > kEmitMax || n_rhs > kEmitMax) return fail(ctx, SSB_ERROR_NOT_IMPLEMENTED, "at most 8 result columns per side");
  JoinEmit emit;
  memset(&emit, 0, sizeof(emit));
  emit.n_l = n_lhs;
  emit.n_r = n_rhs;
  for (int c = 0; c < n_lhs + n_rhs; ++c) {
    const ssb_column& src = c < n_lhs ? lhs_cols[c] : rhs_cols[c - n_lhs];
    const int w = width_of(src.dtype);
    if (w == 0 || width_of(out_cols[c].dtype) != w) return fail(ctx, SSB_ERROR_INVALID_ARGUMENT_TYPE, "result column type");
    if (src.nulls != nullptr) return fail(ctx, SSB_ERROR_NOT_IMPLEMENTED, "nullable result columns: use ssb_join_probe + ssb_gather");
    if (c < n_lhs) { emit.l_src[c] = src.data; emit.l_dst[c] = out_cols[c].data; emit.l_w[c] = w; }
    else { emit.r_src[c - n_lhs] = src.data; emit.r_dst[c - n_lhs] = out_cols[c].data; emit.r_w[c - n_lhs] = w; }
  }
  emit.matched = join_type == SSB_JOIN_LEFT_OUTER ? d_matched : nullptr;
  if (rows == 0) return 0;
  JoinKeys probe;
  if (int rc = fill_keys(ctx, j->build_keys.n_keys, keys, &probe)) return rc;
  for (int c = 0; c < probe.n_keys; ++c) {
    const int a = probe.phys[c], b = j->build_keys.phys[c];
    const bool ints = (a == T_I32 || a == T_I64 || a == T_U32 || a == T_U64) && (b == T_I32 || b == T_I64 || b == T_U32 || b == T_U64);
    if (a != b && !ints) return fail(ctx, SSB_ERROR_INVALID_ARGUMENT_TYPE, "probe key types differ from the build keys");
  }
  TimedRegion timed(ctx);
  const long long tiles = div_up(rows, kProbeTile);
  unsigned long long* aux = nullptr;
  cudaError_t e = tmp_malloc(ctx, &aux, static_cast<size_t>(tiles + 2) * 8);
  if (e != cudaSuccess) return cuda_fail(ctx, e, "join probe buffers");
  cudaMemsetAsync(aux, 0, static_cast<size_t>(tiles + 2) * 8, ctx->stream);
  long long grid = static_cast<long long>(ctx->num_sms) * 4;
  if (grid > tiles) grid = tiles;
  if (j->parts.n_parts > 0) {
    join_probe_unique_kernel<true, true, false><<<static_cast<unsigned>(grid), kProbeThreads, 0, ctx->stream>>>(
        j->table, j->build_keys, probe, rows, join_type == SSB_JOIN_LEFT_OUTER ? 1 : 0, nullptr, nullptr, aux, j->parts, emit);
  } else if (j->table.dense_rows != nullptr) {
    // enough probe rows to pay for it: the rhs result columns are laid out by key first (one pass over the key range)
    void* spread[kEmitMax];
    for (int c = 0; c < kEmitMax; ++c) spread[c] = nullptr;
    if (n_rhs > 0 && static_cast<unsigned long long>(rows) >= j->table.dense_range) {
      for (int c = 0; c < n_rhs && e == cudaSuccess; ++c) e = tmp_malloc_bytes(ctx, &spread[c], static_cast<size_t>(j->table.dense_range) * emit.r_w[c] + 64);
      if (e != cudaSuccess) {
        for (int c = 0; c < n_rhs; ++c) tmp_free(ctx, spread[c]);
        tmp_free(ctx, aux);
        return cuda_fail(ctx, e, "join result columns by key");
      }
      for (int c = 0; c < n_rhs; ++c) {
        join_dense_spread_kernel<<<grid_1d(ctx, static_cast<long long>(j->table.dense_range), 256), 256, 0, ctx->stream>>>(j->table, emit.r_src[c], emit.r_w[c], spread[c]);
        ++ctx->launches;
        emit.r_src[c] = spread[c];
      }
      emit.by_key = 1;
    }
    join_probe_unique_kernel<false, true, true><<<static_cast<unsigned>(grid), kProbeThreads, 0, ctx->stream>>>(
        j->table, j->build_keys, probe, rows, join_type == SSB_JOIN_LEFT_OUTER ? 1 : 0, nullptr, nullptr, aux, j->parts, emit);
    ++ctx->launches;
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->h_count, aux + 1, 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    for (int c = 0; c < n_rhs; ++c) tmp_free(ctx, spread[c]);
    tmp_free(ctx, aux);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "join probe");
    *n_rows = *ctx->h_count;
    return 0;
  } else {
    join_probe_unique_kernel<false, true, false><<<static_cast<unsigned>(grid), kProbeThreads, 0, ctx->stream>>>(
        j->table, j->build_keys, probe, rows, join_type == SSB_JOIN_LEFT_OUTER ? 1 : 0, nullptr, nullptr, aux, j->parts, emit);
  }
  ++ctx->launches;
  e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->h_count, aux + 1, 8, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  tmp_free(ctx, aux);
  if (e != cudaSuccess) return cuda_fail(ctx, e, "join probe");
  *n_rows = *ctx->h_count;
  return 0;
}

int ssb_join_table(const ssb_join* j, const void** d_slots, int64_t* capacity) {
  if (j->parts.n_parts > 0) return fail(j->ctx, SSB_ERROR_INVALID_ARGUMENT_VALUE, "an attached index owns no table");
  if (j->table.dense_rows != nullptr) {
    *d_slots = nullptr; *capacity = 0;
    return fail(j->ctx, SSB_ERROR_INVALID_ARGUMENT_VALUE, "this index has no slot table (dense integer keys); build with SSB_KEYS_COMPACT_TABLE to export one");
  }
  *d_slots = j->table.slots;
  *capacity = static_cast<int64_t>(j->table.capacity);
  return 0;
}

int ssb_join_attach_parts(ssb_ctx* ctx, int32_t key_type, int32_t n_parts, const void* const* d_slots,
                          const int64_t* capacities, const int64_t* row_offsets, ssb_join** out) {
  *out = nullptr;
  if (n_parts < 1 || n_parts > kJoinMaxParts) return fail(ctx, SSB_ERROR_INVALID_ARGUMENT_VALUE, "1..16 table parts");
  if (phys_of(key_type) < 0) return fail(ctx, SSB_ERROR_INVALID_ARGUMENT_TYPE, "unsupported join key type");
  ssb_join* j = new ssb_join();
  memset(j, 0, sizeof(*j));
  j->ctx = ctx;
  j->build_keys.n_keys = 1;          // the slot word holds the whole key: no build column is read by a probe
  j->build_keys.phys[0] = phys_of(key_type);
  j->uniqueness = SSB_KEYS_UNIQUE;
  j->parts.n_parts = n_parts;
  for (int p = 0; p < n_parts; ++p) {
    const int64_t cap = capacities[p];
    if (cap < 1 || cap >= (1LL << 32) || d_slots[p] == nullptr) { delete j; return fail(ctx, SSB_ERROR_INVALID_ARGUMENT_VALUE, "table part: bad capacity"); }
    j->parts.slots[p] = static_cast<const unsigned long long*>(d_slots[p]);
    j->parts.capacity[p] = static_cast<unsigned long long>(cap);
    j->parts.row_offset[p] = row_offsets[p];
    j->build_rows += 0;
  }
  *out = j;
  return 0;
}

int ssb_partition_rows(ssb_ctx* ctx, int32_t n_keys, const ssb_column* keys, int64_t rows, int32_t n_parts,
                       int32_t null_part, int64_t* d_perm, int64_t* h_counts) {
  if (rows < 0) return fail(ctx, SSB_ERROR_INVALID_ARGUMENT_VALUE, "negative row count");
  const bool aside = null_part == n_parts;   // NULL keys form an extra part behind the hash parts
  if (n_parts < 1 || n_parts > (aside ? 255 : 256) || null_part < 0 || null_part > n_parts) {
    return fail(ctx, SSB_ERROR_INVALID_ARGUMENT_VALUE, "partition: 1..256 parts, null_part inside (or == n_parts with at most 255 parts)");
  }
  const int n_counts = n_parts + (aside ? 1 : 0);
  for (int p = 0; p < n_counts; ++p) h_counts[p] = 0;
  if (rows == 0) return 0;
  JoinKeys jk;
  if (int rc = fill_keys(ctx, n_keys, keys, &jk)) return rc;
  TimedRegion timed(ctx);
  unsigned long long *k0 = nullptr, *k1 = nullptr, *d_counts = nullptr;
  long long* v0 = nullptr;
  cudaError_t e = tmp_malloc(ctx, &k0, static_cast<size_t>(rows) * 8);
  if (e == cudaSuccess) e = tmp_malloc(ctx, &k1, static_cast<size_t>(rows) * 8);
  if (e == cudaSuccess) e = tmp_malloc(ctx, &v0, static_cast<size_t>(rows) * 8);
  if (e == cudaSuccess) e = tmp_malloc(ctx, &d_counts, 256 * 8);
  if (e != cudaSuccess) { tmp_free(ctx, k0); tmp_free(ctx, k1); tmp_free(ctx, v0); tmp_free(ctx, d_counts); return cuda_fail(ctx, e, "partition scratch"); }
  cudaMemsetAsync(d_counts, 0, 256 * 8, ctx->stream);
  part_id_kernel<<<grid_1d(ctx, rows, 256), 256, 0, ctx->stream>>>(jk, rows, static_cast<unsigned>(n_parts),
                                                                    static_cast<unsigned>(null_part), k0, v0, d_counts);
  ++ctx->launches;
  // one stable 8-bit pass groups the row ids by part; the sorted ids land in the buffer that
  // holds them after the last pass (d_perm or v0)
  long long* va = v0;
  long long* vb = reinterpret_cast<long long*>(d_perm);
  int rc = radix_sort_pairs(ctx, &k0, &va, &k1, &vb, static_cast<unsigned long long>(rows), 0, 8);
  if (rc == 0 && va != reinterpret_cast<long long*>(d_perm)) {
    cudaMemcpyAsync(d_perm, va, static_cast<size_t>(rows) * 8, cudaMemcpyDeviceToDevice, ctx->stream);
  }
  unsigned long long h[256];
  if (rc == 0) {
    e = cudaMemcpyAsync(h, d_counts, static_cast<size_t>(n_counts) * 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) rc = cuda_fail(ctx, e, "partition");
  }
  cudaStreamSynchronize(ctx->stream);
  tmp_free(ctx, k0); tmp_free(ctx, k1); tmp_free(ctx, v0); tmp_free(ctx, d_counts);
  if (rc) return rc;
  for (int p = 0; p < n_counts; ++p) h_counts[p] = static_cast<int64_t>(h[p]);
  return 0;
}

}  // extern "C"
