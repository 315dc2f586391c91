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

static constexpr unsigned long long kValid = 1ull << 63;

// Straight-line binary step for the hot (op, type) pairs: acc = acc (op) rhs.
template <int R, typename T, typename F>
__device__ __forceinline__ void fast_arith(u64 (&acc)[R], const u64 (&rhs)[R], bool rev, F f) {
#pragma unroll
  for (int k = 0; k < R; ++k) {
    const T x = Codec<T>::dec(acc[k]), y = Codec<T>::dec(rhs[k]);
    acc[k] = Codec<T>::enc(rev ? f(y, x) : f(x, y));
  }
}
template <int R, typename T, typename F>
__device__ __forceinline__ void fast_cmp(u64 (&acc)[R], const u64 (&rhs)[R], bool rev, bool neg, F f) {
#pragma unroll
  for (int k = 0; k < R; ++k) {
    const T x = Codec<T>::dec(acc[k]), y = Codec<T>::dec(rhs[k]);
    acc[k] = ((rev ? f(y, x) : f(x, y)) != neg) ? 1u : 0u;
  }
}

template <int NT, int R>
__global__ void __launch_bounds__(NT, 2) expr_kernel(const __grid_constant__ ExprParams p) {
  constexpr int TILE = NT * R;
  constexpr int NW = NT / 32;
  constexpr int NSEG = NW * R;
  static_assert(NSEG <= 32, "one warp scans the segment counts");
  extern __shared__ __align__(1024) unsigned char smem[];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.off_bar);
  uint32_t* seg_cnt = reinterpret_cast<uint32_t*>(smem + p.off_scan);          // [32]
  uint32_t* seg_off = seg_cnt + 32;                                            // [32]
  unsigned long long* red = reinterpret_cast<unsigned long long*>(seg_cnt + 64);   // [2 * NW] reduce scratch
  unsigned long long* s_meta = red + 2 * NW;   // [0..1] tile totals of the two output buffers, [2] base, [3] wave base
  uint32_t* nullw_base = reinterpret_cast<uint32_t*>(smem + p.off_nullw);

  const long long G = gridDim.x;
  const long long bid = blockIdx.x;
  const long long n_my = (p.num_tiles - bid + G - 1) / G;
  const int S = p.stages;

  if (tid == 0) {
    if (p.use_tma) {
      for (int s = 0; s < S; ++s) mbar_init(&bars[s], 1);
      fence_barrier_init();
    }
    s_meta[3] = 0;
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
        tma_load_1d(slot_nullw(i, stage), p.in_nulls[i] + row0 / 32, TILE / 8, &bars[stage]);
      }
    }
  };

  if (tid == 0) {
    for (int s = 0; s < S && s < n_my; ++s) issue(bid + s * G, s);
  }

  const uint32_t all = (1u << R) - 1u;
  const uint32_t lt = (1u << lane) - 1u;
  uint32_t fail = 0;

  // One extra iteration drains the last deferred tile.
  for (long long it = 0; it <= n_my; ++it) {
    // ======================================================== evaluate tile `it`
    if (it < n_my) {
      const long long tile = bid + it * G;
      const int stage = static_cast<int>(it % S);
      const uint32_t parity = static_cast<uint32_t>((it / S) & 1);
      const long long row0 = tile * TILE;
      const int n = static_cast<int>(p.rows - row0 < TILE ? p.rows - row0 : TILE);
      const bool via_tma = p.use_tma && n == TILE;
      unsigned char* obuf = smem + p.off_out + static_cast<int>(it % kOutBuffers) * p.out_bytes;

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
        uint32_t* wdst = slot_nullw(i, stage);
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

      // ---- the accumulator machine
      u64 acc[R];
      uint32_t accn = 0, pass = live;
      int pos[R];
#pragma unroll
      for (int k = 0; k < R; ++k) { acc[k] = 0; pos[k] = k * NT + tid; }
      // Without a predicate no barrier separates this tile's staging writes from the copy-out
      // of the tile evaluated two iterations ago (same buffer): add one.
      if (!p.has_pred) __syncthreads();

      for (int pc = 0; pc < p.n_insn; ++pc) {
        const Insn in = p.insn[pc];
        if (in.kind == K_ALU2 || in.kind == K_LOAD) {
          // operand fetch
          u64 rhs[R];
          uint32_t rn = 0;
          if (in.flags & F_RHS_IMM) {
            const u64 c = p.imm[in.a];
#pragma unroll
            for (int k = 0; k < R; ++k) rhs[k] = c;
            rn = (in.flags & F_RHS_NULLK) ? all : 0u;
          } else {
            const unsigned char* base = slot_data(in.a, stage);
            if (in.rw == 8) {
#pragma unroll
              for (int k = 0; k < R; ++k) rhs[k] = reinterpret_cast<const u64*>(base)[k * NT + tid];
            } else if (in.rw == 4) {
#pragma unroll
              for (int k = 0; k < R; ++k) rhs[k] = reinterpret_cast<const uint32_t*>(base)[k * NT + tid];
            } else {
#pragma unroll
              for (int k = 0; k < R; ++k) rhs[k] = base[k * NT + tid];
            }
            if (in.rhs_nullable & 1) {
              const uint32_t* w = slot_nullw(in.a, stage);
              if (w != nullptr) {
#pragma unroll
                for (int k = 0; k < R; ++k) rn |= ((w[k * NW + warp] >> lane) & 1u) << k;
              }
            }
          }
          if (in.kind == K_LOAD) {
#pragma unroll
            for (int k = 0; k < R; ++k) acc[k] = rhs[k];
            accn = rn;
            continue;
          }
          const bool rev = (in.flags & F_REV) != 0;
          const bool neg = (in.flags & F_NEGATE) != 0;
          switch (in.code) {
            case C_ADD_I64: fast_arith<R, int64_t>(acc, rhs, false, [](int64_t x, int64_t y) { return Arith<int64_t>::add(x, y); }); accn |= rn; break;
            case C_SUB_I64: fast_arith<R, int64_t>(acc, rhs, rev, [](int64_t x, int64_t y) { return Arith<int64_t>::sub(x, y); }); accn |= rn; break;
            case C_MUL_I64: fast_arith<R, int64_t>(acc, rhs, false, [](int64_t x, int64_t y) { return Arith<int64_t>::mul(x, y); }); accn |= rn; break;
            case C_LT_I64: fast_cmp<R, int64_t>(acc, rhs, rev, neg, [](int64_t x, int64_t y) { return x < y; }); accn |= rn; break;
            case C_EQ_I64: fast_cmp<R, int64_t>(acc, rhs, false, neg, [](int64_t x, int64_t y) { return x == y; }); accn |= rn; break;
            case C_ADD_F64: fast_arith<R, double>(acc, rhs, false, [](double x, double y) { return x + y; }); accn |= rn; break;
            case C_SUB_F64: fast_arith<R, double>(acc, rhs, rev, [](double x, double y) { return x - y; }); accn |= rn; break;
            case C_MUL_F64: fast_arith<R, double>(acc, rhs, false, [](double x, double y) { return x * y; }); accn |= rn; break;
            case C_LT_F64: fast_cmp<R, double>(acc, rhs, rev, neg, [](double x, double y) { return x < y; }); accn |= rn; break;
            case C_EQ_F64: fast_cmp<R, double>(acc, rhs, false, neg, [](double x, double y) { return x == y; }); accn |= rn; break;
            case C_ADD_I32: fast_arith<R, int32_t>(acc, rhs, false, [](int32_t x, int32_t y) { return Arith<int32_t>::add(x, y); }); accn |= rn; break;
            case C_SUB_I32: fast_arith<R, int32_t>(acc, rhs, rev, [](int32_t x, int32_t y) { return Arith<int32_t>::sub(x, y); }); accn |= rn; break;
            case C_MUL_I32: fast_arith<R, int32_t>(acc, rhs, false, [](int32_t x, int32_t y) { return Arith<int32_t>::mul(x, y); }); accn |= rn; break;
            case C_LT_I32: fast_cmp<R, int32_t>(acc, rhs, rev, neg, [](int32_t x, int32_t y) { return x < y; }); accn |= rn; break;
            case C_EQ_I32: fast_cmp<R, int32_t>(acc, rhs, false, neg, [](int32_t x, int32_t y) { return x == y; }); accn |= rn; break;
            case C_AND3:
            case C_OR3: {
              uint32_t av = 0, bv = 0;
#pragma unroll
              for (int k = 0; k < R; ++k) { av |= (acc[k] & 1u) << k; bv |= (rhs[k] & 1u) << k; }
              const uint32_t at = av & ~accn, af = ~av & ~accn & all, bt = bv & ~rn, bf = ~bv & ~rn & all;
              uint32_t val, nul;
              if (in.code == C_OR3) { val = at | bt; nul = (accn | rn) & ~(at | bt); }
              else { val = at & bt; nul = (accn | rn) & ~(af | bf); }
#pragma unroll
              for (int k = 0; k < R; ++k) acc[k] = (val >> k) & 1u;
              accn = nul & all;
            } break;
            default: {
              u64 rhs2[R];
#pragma unroll
              for (int k = 0; k < R; ++k) rhs2[k] = 0;
              alu<R>(in, acc, accn, rhs, rn, rhs2, 0u, live, fail);
            } break;
          }
          continue;
        }
        switch (in.kind) {
          case K_STORE: {
            unsigned char* base = slot_data(in.a, stage);
            if (in.rw == 8) {
#pragma unroll
              for (int k = 0; k < R; ++k) reinterpret_cast<u64*>(base)[k * NT + tid] = acc[k];
            } else if (in.rw == 4) {
#pragma unroll
              for (int k = 0; k < R; ++k) reinterpret_cast<uint32_t*>(base)[k * NT + tid] = static_cast<uint32_t>(acc[k]);
            } else {
#pragma unroll
              for (int k = 0; k < R; ++k) base[k * NT + tid] = static_cast<unsigned char>(acc[k]);
            }
            if (in.rhs_nullable & 1) {
              uint32_t* w = slot_nullw(in.a, stage);
#pragma unroll
              for (int k = 0; k < R; ++k) {
                const uint32_t b = __ballot_sync(0xffffffffu, (accn >> k) & 1u);
                if (lane == 0) w[k * NW + warp] = b;
              }
              __syncwarp();   // read back by the lanes of this warp only
            }
          } break;
          case K_ALU1: {
            u64 z[R];
#pragma unroll
            for (int k = 0; k < R; ++k) z[k] = 0;
            alu<R>(in, acc, accn, z, 0u, z, 0u, live, fail);
          } break;
          case K_ALU3: {
            u64 r1[R], r2[R];
            uint32_t n1 = 0, n2 = 0;
            // both operands have the element width rw
            auto fetch = [&](int idx, bool is_imm, bool null_const, bool nullable, u64 (&v)[R], uint32_t& nn) {
              if (is_imm) {
#pragma unroll
                for (int k = 0; k < R; ++k) v[k] = p.imm[idx];
                nn = null_const ? all : 0u;
              } else {
                const unsigned char* base = slot_data(idx, stage);
#pragma unroll
                for (int k = 0; k < R; ++k) {
                  const int r = k * NT + tid;
                  v[k] = in.rw == 8 ? reinterpret_cast<const u64*>(base)[r]
                                    : (in.rw == 4 ? static_cast<u64>(reinterpret_cast<const uint32_t*>(base)[r]) : static_cast<u64>(base[r]));
                }
                nn = 0;
                if (nullable) {
                  const uint32_t* w = slot_nullw(idx, stage);
                  if (w != nullptr) {
#pragma unroll
                    for (int k = 0; k < R; ++k) nn |= ((w[k * NW + warp] >> lane) & 1u) << k;
                  }
                }
              }
            };
            fetch(in.a, in.flags & F_RHS_IMM, in.flags & F_RHS_NULLK, in.rhs_nullable & 1, r1, n1);
            fetch(in.b, in.flags & F_RHS2_IMM, in.rhs_nullable & 4, in.rhs_nullable & 2, r2, n2);
            alu<R>(in, acc, accn, r1, n1, r2, n2, live, fail);
          } break;
          case K_PRED: {
            uint32_t t = 0;
#pragma unroll
            for (int k = 0; k < R; ++k) t |= static_cast<uint32_t>(acc[k] & 1u) << k;
            pass = t & ~accn & live;
            // in-tile compaction: ballots, one 32-entry scan, positions
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
              const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
              if (lane == 0) {
                s_meta[it % kOutBuffers] = total;
                st_relaxed(&p.tile_status[tile], kValid | total);   // this tile's share of its wave
              }
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < R; ++k) pos[k] = static_cast<int>(seg_off[k * NW + warp]) + __popc(mask[k] & lt);
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
      __syncthreads();   // the stage and the temporaries are free; the staging buffer is complete
      if (tid == 0 && it + S < n_my) issue(tile + S * G, stage);
    }

    // ======================================================== write out tile `it - 1`
    if (it >= 1) {
      const long long wave = it - 1;
      const long long tile = bid + wave * G;
      const unsigned char* obuf = smem + p.off_out + static_cast<int>(wave % kOutBuffers) * p.out_bytes;
      long long total;
      long long base;
      if (p.has_pred) {
        // Wave-synchronous prefix: every CTA of this wave published its count when it evaluated
        // the tile (one iteration ago), so this read normally does not spin.
        const long long w0 = wave * G;
        unsigned long long before = 0, sum = 0;
        for (long long j = w0 + tid; j < w0 + G && j < p.num_tiles; j += NT) {
          unsigned long long v;
          do { v = ld_relaxed(&p.tile_status[j]); } while (!(v & kValid));
          v &= ~kValid;
          sum += v;
          if (j < tile) before += v;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
          before += __shfl_xor_sync(0xffffffffu, before, d);
          sum += __shfl_xor_sync(0xffffffffu, sum, d);
        }
        if (lane == 0) { red[warp] = before; red[NW + warp] = sum; }
        __syncthreads();
        if (tid == 0) {
          unsigned long long b = 0, s = 0;
          for (int w = 0; w < NW; ++w) { b += red[w]; s += red[NW + w]; }
          s_meta[2] = s_meta[3] + b;
          s_meta[3] += s;
        }
        __syncthreads();
        base = static_cast<long long>(s_meta[2]);
        total = static_cast<long long>(s_meta[wave % kOutBuffers]);
        if (tile == p.num_tiles - 1 && tid == 0 && p.d_out_rows != nullptr) *p.d_out_rows = base + total;
      } else {
        base = tile * static_cast<long long>(TILE);
        total = p.rows - base < TILE ? p.rows - base : TILE;
      }
      for (int j = 0; j < p.n_out; ++j) {
        const unsigned char* src = obuf + p.out_off[j];
        unsigned char* dst = static_cast<unsigned char*>(p.out_data[j]);
        const int w = p.out_width[j];
        if (w == 8) {
          for (int i = tid; i < total; i += NT) st_cs_u64(dst + (base + i) * 8, reinterpret_cast<const u64*>(src)[i]);
        } else if (w == 4) {
          for (int i = tid; i < total; i += NT) st_cs_u32(dst + (base + i) * 4, reinterpret_cast<const uint32_t*>(src)[i]);
        } else {
          for (int i = tid; i < total; i += NT) dst[base + i] = src[i];
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
              const long long w0 = g;   // g is a multiple of 32 for lane 0
              if (w0 >= base && w0 + 32 <= end) p.out_nulls[j][w0 >> 5] = word;
              else atomicOr(&p.out_nulls[j][w0 >> 5], word);
            }
          }
        }
      }
      // No barrier needed here: the next write into this staging buffer happens two
      // evaluations later, behind at least one __syncthreads of the next iteration.
    }
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
    // every nullable output bitmap is zeroed: only non-zero words are written by the kernel
    if (p.out_nulls[j]) {
      SSB_CUDA(ctx, cudaMemsetAsync(p.out_nulls[j], 0, static_cast<size_t>(div_up(rows, 32) + 1) * 4, ctx->stream));
    }
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
