// expr_kernel.cu -- the fused Compute / Project / Filter kernel for sm_100a.
//
// Replaces, in one launch over a whole shard, the reference's per-1024-row-block chain
//   ViewCursor::Next -> BoundExpressionTree::Evaluate -> VectorBinaryPrimitive loops
//   -> FilterCursor::PrepareInputRowIds -> SelectiveViewCopier gather
// (cursor/core/compute.cc:49-56, cursor/core/filter.cc:96-230,
//  expression/vector/vector_primitives.h:70-353, base/infrastructure/copy_column.cc:200-286).
//
// Shape of the kernel (HBM-bound; no tensor cores because nothing here is a contraction):
//  * persistent CTAs, tile = 1024 rows; tile t belongs to CTA t % gridDim.x
//  * input column tiles are staged into shared memory by TMA bulk copies
//    (cp.async.bulk, one elected thread, mbarrier complete_tx), `stages` tiles in flight
//    per CTA, so HBM latency is hidden by bytes in flight rather than by occupancy
//  * the expression program (bytecode in the constant bank) is interpreted warp-uniformly;
//    each thread owns 4 rows whose values live in registers (the accumulator); shared
//    memory slots hold operands only
//  * Filter: warp ballots + one 32-entry scan give in-tile offsets, decoupled look-back over
//    per-tile status words gives the global offset, so output order = input order with a
//    single pass over the data
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.h"
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
// streaming (evict-first) global accesses: every byte is touched once
__device__ __forceinline__ void st_cs_u64(void* p, uint64_t v) {
  asm volatile("st.global.cs.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_cs_u32(void* p, uint32_t v) {
  asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

static constexpr unsigned long long kFlagAggregate = 1ull << 62;
static constexpr unsigned long long kFlagPrefix = 2ull << 62;
static constexpr unsigned long long kValueMask = (1ull << 62) - 1;

template <int NT, int R>
struct Machine {
  static constexpr int TILE = NT * R;
  static constexpr int NW = NT / 32;
  static constexpr int NSEG = NW * R;
  static_assert(NSEG <= 32, "one warp scans the segment counts");

  const ExprParams& p;
  unsigned char* smem;
  int tid, lane, warp;

  __device__ __forceinline__ unsigned char* slot_data(int slot, int stage) const {
    return slot < p.n_in ? smem + p.off_data + stage * p.stage_bytes + p.in_off[slot]
                         : smem + p.off_tmp + (slot - p.n_in) * (TILE * 8);
  }
  // null words of a slot for the current stage; inputs without a null-word row return NULL
  __device__ __forceinline__ uint32_t* slot_nullw(int slot, int stage) const {
    uint32_t* base = reinterpret_cast<uint32_t*>(smem + p.off_nullw);
    if (slot < p.n_in) {
      const int row = p.in_nullw[slot];
      return row < 0 ? nullptr : base + (stage * p.stage_nullw + row) * (TILE / 32);
    }
    return base + (p.stages * p.stage_nullw + (slot - p.n_in)) * (TILE / 32);
  }

  __device__ __forceinline__ void load_vals(const unsigned char* base, int width, u64 (&v)[R]) const {
    if (width == 8) {
      const u64* s = reinterpret_cast<const u64*>(base);
#pragma unroll
      for (int k = 0; k < R; ++k) v[k] = s[k * NT + tid];
    } else if (width == 4) {
      const uint32_t* s = reinterpret_cast<const uint32_t*>(base);
#pragma unroll
      for (int k = 0; k < R; ++k) v[k] = s[k * NT + tid];
    } else {
#pragma unroll
      for (int k = 0; k < R; ++k) v[k] = base[k * NT + tid];
    }
  }
  __device__ __forceinline__ uint32_t load_nulls(const uint32_t* words) const {
    uint32_t n = 0;
#pragma unroll
    for (int k = 0; k < R; ++k) n |= ((words[k * NW + warp] >> lane) & 1u) << k;
    return n;
  }
  __device__ __forceinline__ void store_vals(unsigned char* base, int width, const u64 (&v)[R]) const {
    if (width == 8) {
      u64* s = reinterpret_cast<u64*>(base);
#pragma unroll
      for (int k = 0; k < R; ++k) s[k * NT + tid] = v[k];
    } else if (width == 4) {
      uint32_t* s = reinterpret_cast<uint32_t*>(base);
#pragma unroll
      for (int k = 0; k < R; ++k) s[k * NT + tid] = static_cast<uint32_t>(v[k]);
    } else {
#pragma unroll
      for (int k = 0; k < R; ++k) base[k * NT + tid] = static_cast<unsigned char>(v[k]);
    }
  }
  __device__ __forceinline__ void store_nulls(uint32_t* words, uint32_t n) const {
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const uint32_t w = __ballot_sync(0xffffffffu, (n >> k) & 1u);
      if (lane == 0) words[k * NW + warp] = w;
    }
    __syncwarp();   // the words are read back by the lanes of this warp only
  }

  // operand fetch for K_LOAD / K_ALU2 / K_ALU3 (first operand)
  __device__ __forceinline__ void fetch(const Insn& in, int slot_or_imm, bool is_imm, bool null_const,
                                        bool nullable, int stage, u64 (&v)[R], uint32_t& n) const {
    const uint32_t all = (1u << R) - 1u;
    if (is_imm) {
      const u64 c = p.imm[slot_or_imm];
#pragma unroll
      for (int k = 0; k < R; ++k) v[k] = c;
      n = null_const ? all : 0u;
    } else {
      load_vals(slot_data(slot_or_imm, stage), in.rw, v);
      n = 0;
      if (nullable) {
        const uint32_t* w = slot_nullw(slot_or_imm, stage);
        if (w != nullptr) n = load_nulls(w);
      }
    }
  }

  // Runs the program for the tile staged in `stage`. Returns the pass bits (predicate).
  __device__ __forceinline__ uint32_t run(int stage, uint32_t live, uint32_t& fail) const {
    u64 acc[R];
    u64 rhs[R];
    u64 rhs2[R];
    uint32_t accn = 0, pass = live;
#pragma unroll
    for (int k = 0; k < R; ++k) { acc[k] = 0; rhs[k] = 0; rhs2[k] = 0; }
    for (int pc = 0; pc < p.n_insn; ++pc) {
      const Insn in = p.insn[pc];
      switch (in.kind) {
        case K_LOAD:
          fetch(in, in.a, in.flags & F_RHS_IMM, in.flags & F_RHS_NULLK, in.rhs_nullable & 1, stage, acc, accn);
          break;
        case K_STORE:
          store_vals(slot_data(in.a, stage), in.rw, acc);
          if (in.rhs_nullable & 1) store_nulls(slot_nullw(in.a, stage), accn);
          break;
        case K_ALU1:
          alu<R>(in, acc, accn, rhs, 0u, rhs2, 0u, live, fail);
          break;
        case K_ALU2: {
          uint32_t rn;
          fetch(in, in.a, in.flags & F_RHS_IMM, in.flags & F_RHS_NULLK, in.rhs_nullable & 1, stage, rhs, rn);
          alu<R>(in, acc, accn, rhs, rn, rhs2, 0u, live, fail);
        } break;
        case K_ALU3: {
          uint32_t rn, rn2;
          fetch(in, in.a, in.flags & F_RHS_IMM, in.flags & F_RHS_NULLK, in.rhs_nullable & 1, stage, rhs, rn);
          fetch(in, in.b, in.flags & F_RHS2_IMM, in.rhs_nullable & 4, in.rhs_nullable & 2, stage, rhs2, rn2);
          alu<R>(in, acc, accn, rhs, rn, rhs2, rn2, live, fail);
        } break;
        case K_PRED: {
          uint32_t t = 0;
#pragma unroll
          for (int k = 0; k < R; ++k) t |= (Codec<bool>::dec(acc[k]) ? 1u : 0u) << k;
          pass = t & ~accn & live;
        } break;
        default: break;
      }
    }
    return pass;
  }
};

template <int NT, int R>
__global__ void __launch_bounds__(NT) expr_kernel(const __grid_constant__ ExprParams p) {
  constexpr int TILE = NT * R;
  constexpr int NW = NT / 32;
  constexpr int NSEG = NW * R;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~static_cast<uintptr_t>(127));

  Machine<NT, R> m{p, smem, static_cast<int>(threadIdx.x), static_cast<int>(threadIdx.x & 31),
                   static_cast<int>(threadIdx.x >> 5)};
  const int tid = m.tid, lane = m.lane, warp = m.warp;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.off_bar);
  uint32_t* seg_cnt = reinterpret_cast<uint32_t*>(smem + p.off_scan);        // [NSEG]
  uint32_t* seg_off = seg_cnt + 32;                                          // [NSEG]
  long long* s_base = reinterpret_cast<long long*>(seg_cnt + 64);            // 8-byte slot

  const long long G = gridDim.x;
  const long long bid = blockIdx.x;
  const long long n_my = (p.num_tiles - bid + G - 1) / G;
  const int S = p.stages;

  if (tid == 0 && p.use_tma) {
    for (int s = 0; s < S; ++s) mbar_init(&bars[s], 1);
    fence_barrier_init();
  }
  __syncthreads();

  // Issues the TMA fill of `stage` with tile `tile` (full tiles only).
  auto issue = [&](long long tile, int stage) {
    const long long row0 = tile * TILE;
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
        tma_load_1d(m.slot_nullw(i, stage), p.in_nulls[i] + row0 / 32, TILE / 8, &bars[stage]);
      }
    }
  };

  if (tid == 0) {
    for (int s = 0; s < S && s < n_my; ++s) issue(bid + s * G, s);
  }

  uint32_t fail = 0;
  for (long long it = 0; it < n_my; ++it) {
    const long long tile = bid + it * G;
    const int stage = static_cast<int>(it % S);
    const uint32_t parity = static_cast<uint32_t>((it / S) & 1);
    const long long row0 = tile * TILE;
    const int n = static_cast<int>(p.rows - row0 < TILE ? p.rows - row0 : TILE);
    const bool via_tma = p.use_tma && n == TILE;

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
      __syncthreads();
    }
    // Null words of inputs declared nullable: TMA delivered them when the column has a
    // bitmap; otherwise (no bitmap in this run, or the plain-load path) fill them here.
    bool filled = false;
    for (int i = 0; i < p.n_in; ++i) {
      if (p.in_nullw[i] < 0) continue;
      const bool have = p.in_nulls[i] != nullptr;
      if (via_tma && have) continue;
      uint32_t* wdst = m.slot_nullw(i, stage);
      if (tid < TILE / 32) {
        uint32_t wv = 0;
        if (have && tid * 32 < n) wv = p.in_nulls[i][row0 / 32 + tid];
        wdst[tid] = wv;
      }
      filled = true;
    }
    if (filled) __syncthreads();

    uint32_t live = 0;
#pragma unroll
    for (int k = 0; k < R; ++k) live |= (k * NT + tid < n ? 1u : 0u) << k;

    const uint32_t pass = m.run(stage, live, fail);
    __syncthreads();   // temporaries written by other warps' ballots are complete

    if (!p.has_pred) {
      // Compute / Project: row i of the tile goes to row row0 + i
      for (int j = 0; j < p.n_out; ++j) {
        const unsigned char* src = m.slot_data(p.out_slot[j], stage);
        unsigned char* dst = static_cast<unsigned char*>(p.out_data[j]);
        const int w = p.out_width[j];
#pragma unroll
        for (int k = 0; k < R; ++k) {
          const int r = k * NT + tid;
          if (r < n) {
            if (w == 8) st_cs_u64(dst + (row0 + r) * 8, reinterpret_cast<const u64*>(src)[r]);
            else if (w == 4) st_cs_u32(dst + (row0 + r) * 4, reinterpret_cast<const uint32_t*>(src)[r]);
            else dst[row0 + r] = src[r];
          }
        }
        if (p.out_nullable[j] && p.out_nulls[j] != nullptr) {
          const uint32_t* wsrc = m.slot_nullw(p.out_slot[j], stage);
          if (tid < TILE / 32 && tid * 32 < n) {
            uint32_t wv = wsrc != nullptr ? wsrc[tid] : 0u;
            const int rem = n - tid * 32;
            if (rem < 32) wv &= (1u << rem) - 1u;
            p.out_nulls[j][row0 / 32 + tid] = wv;
          }
        }
      }
    } else {
      // Filter: order-preserving stream compaction
      uint32_t mask[R];
#pragma unroll
      for (int k = 0; k < R; ++k) {
        mask[k] = __ballot_sync(0xffffffffu, (pass >> k) & 1u);
        if (lane == 0) seg_cnt[k * NW + warp] = __popc(mask[k]);
      }
      __syncthreads();
      if (warp == 0) {
        const uint32_t c = lane < NSEG ? seg_cnt[lane] : 0u;
        uint32_t incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
          if (lane >= d) incl += y;
        }
        seg_off[lane] = incl - c;
        const unsigned long long total = __shfl_sync(0xffffffffu, incl, 31);
        // decoupled look-back over the tile status words
        unsigned long long excl = 0;
        if (tile == 0) {
          if (lane == 0) st_relaxed(&p.tile_status[0], kFlagPrefix | total);
        } else {
          if (lane == 0) st_relaxed(&p.tile_status[tile], kFlagAggregate | total);
          long long idx = tile - 1;
          for (;;) {
            const long long j = idx - lane;
            unsigned long long w = kFlagPrefix;   // virtual prefix 0 in front of tile 0
            if (j >= 0) {
              do { w = ld_relaxed(&p.tile_status[j]); } while ((w >> 62) == 0);
            }
            const unsigned pm = __ballot_sync(0xffffffffu, (w >> 62) == 2);
            unsigned long long v = w & kValueMask;
            if (pm) {
              const int first = __ffs(pm) - 1;
              if (lane > first) v = 0;
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
            excl += v;
            if (pm) break;
            idx -= 32;
          }
          if (lane == 0) st_relaxed(&p.tile_status[tile], kFlagPrefix | (excl + total));
        }
        if (lane == 0) {
          *s_base = static_cast<long long>(excl);
          if (tile == p.num_tiles - 1 && p.d_out_rows != nullptr) *p.d_out_rows = static_cast<long long>(excl + total);
        }
      }
      __syncthreads();
      const long long base = *s_base;
      const uint32_t lt = (1u << lane) - 1u;
      for (int j = 0; j < p.n_out; ++j) {
        const unsigned char* src = m.slot_data(p.out_slot[j], stage);
        unsigned char* dst = static_cast<unsigned char*>(p.out_data[j]);
        const int w = p.out_width[j];
        const uint32_t* wsrc = (p.out_nullable[j] && p.out_nulls[j] != nullptr)
                                   ? m.slot_nullw(p.out_slot[j], stage) : nullptr;
#pragma unroll
        for (int k = 0; k < R; ++k) {
          if ((pass >> k) & 1u) {
            const int r = k * NT + tid;
            const long long pos = base + seg_off[k * NW + warp] + __popc(mask[k] & lt);
            if (w == 8) st_cs_u64(dst + pos * 8, reinterpret_cast<const u64*>(src)[r]);
            else if (w == 4) st_cs_u32(dst + pos * 4, reinterpret_cast<const uint32_t*>(src)[r]);
            else dst[pos] = src[r];
            if (wsrc != nullptr && ((wsrc[k * NW + warp] >> lane) & 1u)) {
              atomicOr(&p.out_nulls[j][pos >> 5], 1u << (pos & 31));
            }
          }
        }
      }
    }
    __syncthreads();   // every read of this stage and of the temporaries is done
    if (tid == 0 && it + S < n_my) issue(tile + S * G, stage);
  }
  if (!p.has_pred && p.d_out_rows != nullptr && blockIdx.x == 0 && tid == 0) *p.d_out_rows = p.rows;
  if (fail && p.d_fail != nullptr) atomicOr(p.d_fail, 1);
}

// ------------------------------------------------------------------ host side
static int launch_program(ssb_program* sp, const ssb_column* inputs, int64_t rows,
                          const ssb_column* outputs, int64_t* d_out_rows) {
  ssb_ctx* ctx = sp->ctx;
  Program& prog = sp->prog;
  ExprParams p = prog.params;
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
  }
  p.rows = rows;
  p.num_tiles = div_up(rows, kTile);
  p.use_tma = aligned ? 1 : 0;
  p.d_out_rows = d_out_rows;
  p.d_fail = prog.has_signaling ? ctx->d_fail : nullptr;
  if (p.has_pred) {
    void* st = nullptr;
    if (int rc = scratch(ctx, static_cast<size_t>(p.num_tiles) * 8, &st)) return rc;
    p.tile_status = static_cast<unsigned long long*>(st);
    SSB_CUDA(ctx, cudaMemsetAsync(st, 0, static_cast<size_t>(p.num_tiles) * 8, ctx->stream));
    for (int j = 0; j < p.n_out; ++j) {
      if (p.out_nulls[j]) {
        SSB_CUDA(ctx, cudaMemsetAsync(p.out_nulls[j], 0, static_cast<size_t>(div_up(rows, 32) + 1) * 4, ctx->stream));
      }
    }
  }
  long long grid = static_cast<long long>(ctx->num_sms) * sp->max_ctas_per_sm;
  if (grid > p.num_tiles) grid = p.num_tiles;
  TimedRegion timed(ctx);
  expr_kernel<kThreads, kRowsPerThread><<<static_cast<unsigned>(grid), kThreads, prog.smem_bytes, ctx->stream>>>(p);
  ++ctx->launches;
  SSB_CUDA(ctx, cudaGetLastError());
  return 0;
}

}  // namespace ssb

using namespace ssb;

extern "C" {

int ssb_program_create(ssb_ctx* ctx, const ssb_expr_node* nodes, int32_t n_nodes,
                       int32_t n_inputs, const int32_t* input_types,
                       const int32_t* input_nullable, const int32_t* outputs,
                       int32_t n_outputs, int32_t predicate, ssb_program** out) {
  *out = nullptr;
  ssb_program* sp = new ssb_program;
  sp->ctx = ctx;
  std::string err;
  // Aim for two resident CTAs per SM (look-back latency of one hides behind the other).
  const uint32_t budget = static_cast<uint32_t>(ctx->smem_optin / 2 > 2048 ? ctx->smem_optin / 2 - 1024 : ctx->smem_optin);
  int rc = compile_program(nodes, n_nodes, n_inputs, input_types, input_nullable, outputs, n_outputs,
                           predicate, budget, static_cast<uint32_t>(ctx->smem_optin), &sp->prog, &err);
  if (rc) { delete sp; return fail(ctx, rc, err); }
  cudaError_t e = cudaFuncSetAttribute(expr_kernel<kThreads, kRowsPerThread>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(ctx->smem_optin));
  if (e != cudaSuccess) { delete sp; return cuda_fail(ctx, e, "cudaFuncSetAttribute(expr_kernel)"); }
  int occ = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, expr_kernel<kThreads, kRowsPerThread>, kThreads,
                                                    sp->prog.smem_bytes);
  if (e != cudaSuccess || occ < 1) { delete sp; return cuda_fail(ctx, e, "occupancy(expr_kernel)"); }
  sp->max_ctas_per_sm = occ;
  *out = sp;
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
