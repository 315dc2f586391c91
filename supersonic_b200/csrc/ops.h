// ops.h -- scalar semantics of the expression micro-ops, shared by the sm_100a kernel
// (expr_kernel.cu) and by the bytecode compiler's host-side unit tests (tests/ only).
//
// Every micro-op restates one functor of the reference's
// supersonic/base/infrastructure/operators.h (file:line cited at each op). Values travel
// in 64-bit containers (`u64`); a per-row NULL bit travels beside them. A thread owns R
// rows, so all loops here run over k < R on registers.
#ifndef SSB_CSRC_OPS_H_
#define SSB_CSRC_OPS_H_

#if defined(__CUDACC_RTC__)
#include "jit_rt.h"   // run-time compilation (csrc/jit.cu): no libc headers
#else
#include <stdint.h>
#include <string.h>
#endif

#if defined(__CUDACC__)
#define SSB_HD __host__ __device__ __forceinline__
#else
#define SSB_HD inline
#endif

namespace ssb {

typedef uint64_t u64;

// Physical value types (what the bits in a container mean).
enum PhysType { T_I32 = 0, T_I64 = 1, T_U32 = 2, T_U64 = 3, T_F32 = 4, T_F64 = 5, T_B8 = 6,
                T_COUNT = 7 };

SSB_HD int phys_width(int t) { return (t == T_B8) ? 1 : ((t == T_I32 || t == T_U32 || t == T_F32) ? 4 : 8); }

// Micro-ops executed by the accumulator machine.
enum Mop {
  M_NOP = 0,
  // acc = acc (op) rhs, same physical type `t` on both sides
  M_ADD, M_SUB, M_MUL, M_DIV, M_MOD,
  M_LT, M_EQ,            // -> B8; `t` = lhs type, `t2` = rhs type (mixed integer pairs allowed)
  M_AND3, M_OR3, M_XOR3, M_ANDNOT3,   // three-valued logic on B8
  M_BAND, M_BOR, M_BXOR, M_BANDNOT, M_SHL, M_SHR,
  M_IFNULL,              // acc = acc is NULL ? rhs : acc
  M_SEL,                 // acc = (acc true and not NULL) ? rhs : rhs2
  // unary
  M_NEG, M_NOT, M_BNOT, M_ISODD, M_ISNULL,
  M_CAST,                // `t` -> `t2`
  M_D2DT,                // DATE (I32) -> DATETIME (I64)
  M_COUNT_
};

// Instruction flags.
enum {
  F_REV = 1,          // operands swapped: acc = rhs (op) acc
  F_NEGATE = 2,       // M_LT / M_EQ / M_ISODD: logical negation of the result
  F_ZERO_NULLS = 4,   // M_DIV / M_MOD: zero divisor -> NULL
  F_ZERO_FAILS = 8,   // M_DIV / M_MOD: zero divisor on a non-NULL row -> failure flag
  F_NULLING = 16,     // M_SEL: NULL condition -> NULL result (NullingIf)
  F_RHS_IMM = 32,     // rhs is imm[a] (with F_RHS_NULLK: a NULL constant)
  F_RHS_NULLK = 64,
  F_RHS2_IMM = 128,   // M_SEL: rhs2 is imm[b]
  F_THEN_PRED = 256,  // fast compare directly followed by K_PRED: the row mask feeds the compaction
  F_GUARDED = 512,    // M_DIV / M_MOD with F_ZERO_FAILS as K_ALU3: rhs2 (BOOL) = rows on which the failure counts
};

// Instruction kinds.
enum Kind {
  K_END = 0,
  K_LOAD = 1,      // acc = slot[a] | imm[a]
  K_STORE = 2,     // slot[a] = acc (width rw; null words when rhs_nullable bit0)
  K_ALU1 = 3,      // acc = mop(acc)
  K_ALU2 = 4,      // acc = mop(acc, rhs)             rhs = slot[a] | imm[a]
  K_ALU3 = 5,      // acc = mop(acc, rhs, rhs2)       rhs2 = slot[b] | imm[b]
  K_PRED = 6,      // pass = acc && !null; in-tile compaction offsets are computed here
  K_OUT = 7,       // output column a = acc, written to the tile's output staging (compacted)
};

// Fast-path codes. Instructions whose operands cannot be NULL and whose type is one of the
// hot ones get a pre-decoded, straight-line case in the kernel (operand address = one add);
// everything else runs through the generic path (C_GENERIC) and alu() below.
enum Code {
  C_GENERIC = 0,
  C_LOAD8, C_LOAD4, C_LOADK,           // acc = slot / immediate, no NULLs
  C_OUT8, C_OUT4,                      // staged output, column not nullable
  C_PRED,
  C_AND3_S, C_OR3_S,
  // Binary ops: 4 consecutive codes per (type, op):
  //   +0 acc (op) slot   +1 acc (op) imm   +2 slot (op) slot   +3 slot (op) imm
  // (the last two are a LOAD fused into the operation by the compiler's peephole)
  C_BIN_BASE,
  C_BIN_I64 = C_BIN_BASE,              // ADD SUB SUBR MUL LT GT EQ
  C_BIN_F64 = C_BIN_I64 + 28,
  C_BIN_I32 = C_BIN_F64 + 28,
  C_BIN_END = C_BIN_I32 + 28,
  // Multiply-add (a * b + c, each step rounded / wrapped like the separate ops):
  //   +0 acc * slot a + slot b     +1 slot c * slot a + slot b
  C_MAD_I64 = C_BIN_END, C_MAD_F64 = C_MAD_I64 + 2, C_MAD_I32 = C_MAD_F64 + 2,
  C_MAD_END = C_MAD_I32 + 2,
  C_STORE8 = C_MAD_END, C_STORE4,      // slot = acc, slot not nullable
};
// SUBR / GT are SUB / LT with the operands swapped (F_REV resolved at compile time).
enum { B_ADD = 0, B_SUB = 1, B_SUBR = 2, B_MUL = 3, B_LT = 4, B_GT = 5, B_EQ = 6 };


struct Insn {
  uint8_t kind;
  uint8_t mop;
  uint8_t t;       // operand physical type (K_LOAD/K_STORE: type of the slot element)
  uint8_t t2;      // second type (casts, mixed compares); K_ALU3: type of rhs/rhs2
  uint16_t flags;
  uint8_t rhs_nullable;   // bit0: slot a carries null words, bit1: slot b does
  uint8_t rw;      // byte width of the elements of slot a (and b): 1, 4 or 8
  int16_t a;       // slot or immediate index
  int16_t b;
  uint16_t code;   // Code: fast path selector
  uint16_t pad2;
  uint32_t off_a;  // byte offset of slot a from the shared-memory base (stage 0 for inputs)
  uint32_t off_b;  // same for slot b; bit 31 of either: add the current stage's offset
};

// ---- container encode / decode ------------------------------------------------------------
template <typename T> struct Codec;
template <> struct Codec<int32_t> { static SSB_HD int32_t dec(u64 x) { return (int32_t)(uint32_t)x; }
                                    static SSB_HD u64 enc(int32_t v) { return (u64)(uint32_t)v; } };
template <> struct Codec<uint32_t> { static SSB_HD uint32_t dec(u64 x) { return (uint32_t)x; }
                                     static SSB_HD u64 enc(uint32_t v) { return (u64)v; } };
template <> struct Codec<int64_t> { static SSB_HD int64_t dec(u64 x) { return (int64_t)x; }
                                    static SSB_HD u64 enc(int64_t v) { return (u64)v; } };
template <> struct Codec<uint64_t> { static SSB_HD uint64_t dec(u64 x) { return x; }
                                     static SSB_HD u64 enc(uint64_t v) { return v; } };
template <> struct Codec<float> {
  static SSB_HD float dec(u64 x) { uint32_t b = (uint32_t)x; float f; memcpy(&f, &b, 4); return f; }
  static SSB_HD u64 enc(float f) { uint32_t b; memcpy(&b, &f, 4); return (u64)b; } };
template <> struct Codec<double> {
  static SSB_HD double dec(u64 x) { double d; memcpy(&d, &x, 8); return d; }
  static SSB_HD u64 enc(double d) { u64 b; memcpy(&b, &d, 8); return b; } };
template <> struct Codec<bool> { static SSB_HD bool dec(u64 x) { return (x & 0xff) != 0; }
                                 static SSB_HD u64 enc(bool v) { return v ? 1u : 0u; } };

// ---- wrapping integer arithmetic (the reference's signed overflow is UB; two's complement
// wrap is what its x86 build does in practice) ------------------------------------------------
template <typename T> struct Arith {
  static SSB_HD T add(T a, T b) { return a + b; }
  static SSB_HD T sub(T a, T b) { return a - b; }
  static SSB_HD T mul(T a, T b) { return a * b; }
  static SSB_HD T neg(T a) { return -a; }
};
template <> struct Arith<int32_t> {
  static SSB_HD int32_t add(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
  static SSB_HD int32_t sub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }
  static SSB_HD int32_t mul(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }
  static SSB_HD int32_t neg(int32_t a) { return (int32_t)(0u - (uint32_t)a); }
};
template <> struct Arith<int64_t> {
  static SSB_HD int64_t add(int64_t a, int64_t b) { return (int64_t)((u64)a + (u64)b); }
  static SSB_HD int64_t sub(int64_t a, int64_t b) { return (int64_t)((u64)a - (u64)b); }
  static SSB_HD int64_t mul(int64_t a, int64_t b) { return (int64_t)((u64)a * (u64)b); }
  static SSB_HD int64_t neg(int64_t a) { return (int64_t)(0ull - (u64)a); }
};

template <typename T> struct IsZero { static SSB_HD bool f(T v) { return v == (T)0; } };

// operators.h:88-91 Divide (C++ '/'). Integer division by zero is undefined in the
// reference (it traps on x86); here it yields 0. INT_MIN / -1 yields INT_MIN.
template <typename T> struct Div {
  static SSB_HD T f(T a, T b) { return b == (T)0 ? (T)0 : (T)(a / b); }
};
template <> struct Div<int32_t> {
  static SSB_HD int32_t f(int32_t a, int32_t b) {
    if (b == 0) return 0;
    if (b == -1) return Arith<int32_t>::neg(a);
    return a / b;
  }
};
template <> struct Div<int64_t> {
  static SSB_HD int64_t f(int64_t a, int64_t b) {
    if (b == 0) return 0;
    if (b == -1) return Arith<int64_t>::neg(a);
    return a / b;
  }
};
template <> struct Div<float> { static SSB_HD float f(float a, float b) { return a / b; } };
template <> struct Div<double> { static SSB_HD double f(double a, double b) { return a / b; } };

// double -> int64 as the reference's x86 build converts (cvttsd2si): out-of-range and NaN
// give INT64_MIN ("integer indefinite"), unlike CUDA's saturating conversion.
SSB_HD int64_t f64_to_i64_x86(double d) {
  if (!(d >= -9223372036854775808.0 && d < 9223372036854775808.0)) return INT64_MIN;
  return (int64_t)d;
}
SSB_HD int32_t f64_to_i32_x86(double d) {
  if (!(d > -2147483649.0 && d < 2147483648.0)) return INT32_MIN;
  return (int32_t)d;
}

// operators.h:93-106 Modulus: '%' for integers; FLOAT/DOUBLE: int64(a) % int64(b) -> INT64.
template <typename T> struct Mod {
  static SSB_HD u64 f(T a, T b) { return b == (T)0 ? 0 : Codec<T>::enc((T)(a % b)); }
};
template <> struct Mod<int32_t> {
  static SSB_HD u64 f(int32_t a, int32_t b) { return (b == 0 || b == -1) ? 0 : Codec<int32_t>::enc(a % b); }
};
template <> struct Mod<int64_t> {
  static SSB_HD u64 f(int64_t a, int64_t b) { return (b == 0 || b == -1) ? 0 : Codec<int64_t>::enc(a % b); }
};
template <> struct Mod<double> {
  static SSB_HD u64 f(double a, double b) {
    return Mod<int64_t>::f(f64_to_i64_x86(a), f64_to_i64_x86(b));
  }
};
template <> struct Mod<float> {
  static SSB_HD u64 f(float a, float b) {
    return Mod<int64_t>::f(f64_to_i64_x86((double)a), f64_to_i64_x86((double)b));
  }
};
// zero test of the divisor as the reference's failers/nullers see it
// (expression/vector/column_validity_checkers.h: the right operand == 0).
template <typename T> struct ModZero { static SSB_HD bool f(T b) { return b == (T)0; } };

// operators.h:241-280 Less, including every signed/unsigned overload. Note the reference's
// Less(uint32, int64) narrows b to uint32 (operators.h:261-263); reproduced bit for bit.
template <typename A, typename B> struct Less { static SSB_HD bool f(A a, B b) { return a < b; } };
template <> struct Less<int32_t, uint32_t> { static SSB_HD bool f(int32_t a, uint32_t b) { return a < 0 || (uint32_t)a < b; } };
template <> struct Less<int32_t, uint64_t> { static SSB_HD bool f(int32_t a, uint64_t b) { return a < 0 || (uint32_t)a < b; } };
template <> struct Less<int64_t, uint32_t> { static SSB_HD bool f(int64_t a, uint32_t b) { return a < 0 || (uint64_t)a < b; } };
template <> struct Less<int64_t, uint64_t> { static SSB_HD bool f(int64_t a, uint64_t b) { return a < 0 || (uint64_t)a < b; } };
template <> struct Less<uint32_t, int32_t> { static SSB_HD bool f(uint32_t a, int32_t b) { return b >= 0 && a < (uint32_t)b; } };
template <> struct Less<uint32_t, int64_t> { static SSB_HD bool f(uint32_t a, int64_t b) { return b >= 0 && a < (uint32_t)b; } };
template <> struct Less<uint64_t, int32_t> { static SSB_HD bool f(uint64_t a, int32_t b) { return b >= 0 && a < (uint32_t)b; } };
template <> struct Less<uint64_t, int64_t> { static SSB_HD bool f(uint64_t a, int64_t b) { return b >= 0 && a < (uint64_t)b; } };

// operators.h:185-217 Equal, with its signed/unsigned overloads.
template <typename A, typename B> struct Equal { static SSB_HD bool f(A a, B b) { return a == b; } };
template <> struct Equal<int32_t, uint32_t> { static SSB_HD bool f(int32_t a, uint32_t b) { return a >= 0 && (uint32_t)a == b; } };
template <> struct Equal<int32_t, uint64_t> { static SSB_HD bool f(int32_t a, uint64_t b) { return a >= 0 && (uint32_t)a == b; } };
template <> struct Equal<int64_t, uint32_t> { static SSB_HD bool f(int64_t a, uint32_t b) { return a >= 0 && (uint64_t)a == b; } };
template <> struct Equal<int64_t, uint64_t> { static SSB_HD bool f(int64_t a, uint64_t b) { return a >= 0 && (uint64_t)a == b; } };
template <> struct Equal<uint32_t, int32_t> { static SSB_HD bool f(uint32_t a, int32_t b) { return b >= 0 && a == (uint32_t)b; } };
template <> struct Equal<uint32_t, int64_t> { static SSB_HD bool f(uint32_t a, int64_t b) { return b >= 0 && a == (uint64_t)b; } };
template <> struct Equal<uint64_t, int32_t> { static SSB_HD bool f(uint64_t a, int32_t b) { return b >= 0 && a == (uint32_t)b; } };
template <> struct Equal<uint64_t, int64_t> { static SSB_HD bool f(uint64_t a, int64_t b) { return b >= 0 && a == (uint64_t)b; } };

// operators.h:50-57 Cast = C++ conversion. Floating -> integer conversions follow x86.
template <typename F, typename T> struct CastOp { static SSB_HD T f(F v) { return (T)v; } };
template <> struct CastOp<double, int64_t> { static SSB_HD int64_t f(double v) { return f64_to_i64_x86(v); } };
template <> struct CastOp<float, int64_t> { static SSB_HD int64_t f(float v) { return f64_to_i64_x86((double)v); } };
template <> struct CastOp<double, int32_t> { static SSB_HD int32_t f(double v) { return f64_to_i32_x86(v); } };
template <> struct CastOp<float, int32_t> { static SSB_HD int32_t f(float v) { return f64_to_i32_x86((double)v); } };
template <> struct CastOp<double, uint32_t> { static SSB_HD uint32_t f(double v) { return (uint32_t)f64_to_i64_x86(v); } };
template <> struct CastOp<float, uint32_t> { static SSB_HD uint32_t f(float v) { return (uint32_t)f64_to_i64_x86((double)v); } };
template <> struct CastOp<double, uint64_t> {
  // gcc's x86-64 sequence: values >= 2^63 are converted after subtracting 2^63.
  static SSB_HD uint64_t f(double v) {
    if (v >= 9223372036854775808.0) return (uint64_t)f64_to_i64_x86(v - 9223372036854775808.0) ^ 0x8000000000000000ull;
    return (uint64_t)f64_to_i64_x86(v);
  }
};
template <> struct CastOp<float, uint64_t> { static SSB_HD uint64_t f(float v) { return CastOp<double, uint64_t>::f((double)v); } };
template <typename F> struct CastOp<F, bool> { static SSB_HD bool f(F v) { return v != (F)0; } };

#define SSB_FOR_K for (int k = 0; k < R; ++k)
#if defined(__CUDA_ARCH__)
#define SSB_UNROLL _Pragma("unroll")
#else
#define SSB_UNROLL
#endif

template <typename T> struct UnsignedOf { typedef T type; };
template <> struct UnsignedOf<int32_t> { typedef uint32_t type; };
template <> struct UnsignedOf<int64_t> { typedef uint64_t type; };

#define SSB_NUM_TYPES(t, X)                                        \
  switch (t) {                                                     \
    case T_I32: { typedef int32_t T; X } break;                    \
    case T_I64: { typedef int64_t T; X } break;                    \
    case T_U32: { typedef uint32_t T; X } break;                   \
    case T_U64: { typedef uint64_t T; X } break;                   \
    case T_F32: { typedef float T; X } break;                      \
    case T_F64: { typedef double T; X } break;                     \
    default: break;                                                \
  }
#define SSB_INT_TYPES(t, X)                                        \
  switch (t) {                                                     \
    case T_I32: { typedef int32_t T; X } break;                    \
    case T_I64: { typedef int64_t T; X } break;                    \
    case T_U32: { typedef uint32_t T; X } break;                   \
    case T_U64: { typedef uint64_t T; X } break;                   \
    default: break;                                                \
  }
#define SSB_INT_TYPES2(t, X)                                       \
  switch (t) {                                                     \
    case T_I32: { typedef int32_t T2; X } break;                   \
    case T_I64: { typedef int64_t T2; X } break;                   \
    case T_U32: { typedef uint32_t T2; X } break;                  \
    case T_U64: { typedef uint64_t T2; X } break;                  \
    default: break;                                                \
  }
#define SSB_ALL_TYPES2(t, X)                                       \
  switch (t) {                                                     \
    case T_I32: { typedef int32_t T2; X } break;                   \
    case T_I64: { typedef int64_t T2; X } break;                   \
    case T_U32: { typedef uint32_t T2; X } break;                  \
    case T_U64: { typedef uint64_t T2; X } break;                  \
    case T_F32: { typedef float T2; X } break;                     \
    case T_F64: { typedef double T2; X } break;                    \
    case T_B8: { typedef bool T2; X } break;                       \
    default: break;                                                \
  }

#define SSB_MAP2(EXPR)                                             \
  SSB_UNROLL SSB_FOR_K {                                    \
    const T a = Codec<T>::dec(acc[k]);                             \
    const T b = Codec<T>::dec(rhs[k]);                             \
    acc[k] = Codec<T>::enc((T)(EXPR));                             \
  }

// One ALU step over the R rows a thread owns.
//   acc/accn   accumulator values and NULL bits (bit k = row k)
//   rhs/rhsn   right operand (K_ALU2/3), rhs2/rhs2n third operand (K_ALU3)
//   fail       set non-zero when an F_ZERO_FAILS op meets a zero divisor on a live row
//   live       bit k set = row k exists (tail tiles)
// Rows on which a guarded signaling op may fail: the guard is TRUE and not NULL (all rows without a guard).
template <int R>
SSB_HD uint32_t guard_rows(const Insn& in, const u64 (&g)[R], uint32_t gn) {
  if (!(in.flags & F_GUARDED)) return 0xffffffffu;
  uint32_t m = 0;
  SSB_UNROLL SSB_FOR_K { m |= ((g[k] & 1u) ? 1u : 0u) << k; }
  return m & ~gn;
}

template <int R>
SSB_HD void alu(const Insn& in, u64 (&acc)[R], uint32_t& accn, u64 (&rhs)[R], uint32_t rhsn,
                const u64 (&rhs2)[R], uint32_t rhs2n, uint32_t live, uint32_t& fail) {
  const uint32_t all = (R >= 32) ? 0xffffffffu : ((1u << R) - 1u);
  if (in.flags & F_REV) {
    SSB_UNROLL SSB_FOR_K { const u64 x = acc[k]; acc[k] = rhs[k]; rhs[k] = x; }
    const uint32_t n = accn; accn = rhsn; rhsn = n;
  }
  switch (in.mop) {
    case M_ADD: SSB_NUM_TYPES(in.t, SSB_MAP2(Arith<T>::add(a, b))) accn |= rhsn; break;
    case M_SUB: SSB_NUM_TYPES(in.t, SSB_MAP2(Arith<T>::sub(a, b))) accn |= rhsn; break;
    case M_MUL: SSB_NUM_TYPES(in.t, SSB_MAP2(Arith<T>::mul(a, b))) accn |= rhsn; break;
    case M_DIV: {
      accn |= rhsn;
      uint32_t zero = 0;
      SSB_NUM_TYPES(in.t,
        SSB_UNROLL SSB_FOR_K {
          const T a = Codec<T>::dec(acc[k]);
          const T b = Codec<T>::dec(rhs[k]);
          if (IsZero<T>::f(b)) zero |= 1u << k;
          acc[k] = Codec<T>::enc(Div<T>::f(a, b));
        })
      if (in.flags & F_ZERO_FAILS) fail |= (zero & ~accn & live & guard_rows<R>(in, rhs2, rhs2n));
      if (in.flags & F_ZERO_NULLS) accn |= zero;
    } break;
    case M_MOD: {
      accn |= rhsn;
      uint32_t zero = 0;
      SSB_NUM_TYPES(in.t,
        SSB_UNROLL SSB_FOR_K {
          const T a = Codec<T>::dec(acc[k]);
          const T b = Codec<T>::dec(rhs[k]);
          if (ModZero<T>::f(b)) zero |= 1u << k;
          acc[k] = Mod<T>::f(a, b);
        })
      if (in.flags & F_ZERO_FAILS) fail |= (zero & ~accn & live & guard_rows<R>(in, rhs2, rhs2n));
      if (in.flags & F_ZERO_NULLS) accn |= zero;
    } break;
    case M_LT:
    case M_EQ: {
      const bool is_lt = in.mop == M_LT;
      const bool neg = (in.flags & F_NEGATE) != 0;
      if (in.t == in.t2) {
        switch (in.t) {
          case T_B8: {
            SSB_UNROLL SSB_FOR_K {
              const bool a = Codec<bool>::dec(acc[k]), b = Codec<bool>::dec(rhs[k]);
              acc[k] = ((is_lt ? (a < b) : (a == b)) != neg) ? 1u : 0u;
            }
          } break;
          default:
            SSB_NUM_TYPES(in.t,
              SSB_UNROLL SSB_FOR_K {
                const T a = Codec<T>::dec(acc[k]);
                const T b = Codec<T>::dec(rhs[k]);
                acc[k] = ((is_lt ? Less<T, T>::f(a, b) : Equal<T, T>::f(a, b)) != neg) ? 1u : 0u;
              })
        }
      } else {
        SSB_INT_TYPES(in.t, SSB_INT_TYPES2(in.t2,
          SSB_UNROLL SSB_FOR_K {
            const T a = Codec<T>::dec(acc[k]);
            const T2 b = Codec<T2>::dec(rhs[k]);
            acc[k] = ((is_lt ? Less<T, T2>::f(a, b) : Equal<T, T2>::f(a, b)) != neg) ? 1u : 0u;
          }))
      }
      accn |= rhsn;
    } break;
    // SQL three-valued logic (elementary_bound_expressions.cc:270-506): FALSE AND x = FALSE,
    // TRUE OR x = TRUE even when x is NULL.
    case M_AND3:
    case M_ANDNOT3:
    case M_OR3: {
      uint32_t av = 0, bv = 0;
      SSB_UNROLL SSB_FOR_K {
        av |= (Codec<bool>::dec(acc[k]) ? 1u : 0u) << k;
        bv |= (Codec<bool>::dec(rhs[k]) ? 1u : 0u) << k;
      }
      if (in.mop == M_ANDNOT3) av = ~av & all;   // (!a) && b, operators.h:136-139
      const uint32_t at = av & ~accn, af = ~av & ~accn & all;  // known true / known false
      const uint32_t bt = bv & ~rhsn, bf = ~bv & ~rhsn & all;
      uint32_t val, nul;
      if (in.mop == M_OR3) { val = at | bt; nul = (accn | rhsn) & ~(at | bt); }
      else { val = at & bt; nul = (accn | rhsn) & ~(af | bf); }
      SSB_UNROLL SSB_FOR_K { acc[k] = (val >> k) & 1u; }
      accn = nul & all;
    } break;
    case M_XOR3: {
      SSB_UNROLL SSB_FOR_K {
        acc[k] = (Codec<bool>::dec(acc[k]) != Codec<bool>::dec(rhs[k])) ? 1u : 0u;
      }
      accn |= rhsn;
    } break;
    case M_BAND: SSB_INT_TYPES(in.t, SSB_MAP2(a & b)) accn |= rhsn; break;
    case M_BOR: SSB_INT_TYPES(in.t, SSB_MAP2(a | b)) accn |= rhsn; break;
    case M_BXOR: SSB_INT_TYPES(in.t, SSB_MAP2(a ^ b)) accn |= rhsn; break;
    case M_BANDNOT: SSB_INT_TYPES(in.t, SSB_MAP2((~a) & b)) accn |= rhsn; break;
    case M_SHL:   // operators.h:172-176; the shift count has type t2. Counts >= width are UB
    case M_SHR: { // in the reference; x86 masks the count, and so does this.
      const bool left = in.mop == M_SHL;
      SSB_INT_TYPES(in.t, SSB_INT_TYPES2(in.t2,
        SSB_UNROLL SSB_FOR_K {
          const T a = Codec<T>::dec(acc[k]);
          const unsigned s = (unsigned)Codec<T2>::dec(rhs[k]) & (sizeof(T) * 8 - 1);
          typedef typename UnsignedOf<T>::type U;
          acc[k] = Codec<T>::enc(left ? (T)((U)a << s) : (T)(a >> s));
        }))
      accn |= rhsn;
    } break;
    case M_IFNULL: {
      SSB_UNROLL SSB_FOR_K { if ((accn >> k) & 1u) acc[k] = rhs[k]; }
      accn &= rhsn;
    } break;
    case M_SEL: {
      uint32_t n = 0;
      SSB_UNROLL SSB_FOR_K {
        const bool cn = (accn >> k) & 1u;
        const bool c = Codec<bool>::dec(acc[k]) && !cn;
        acc[k] = c ? rhs[k] : rhs2[k];
        uint32_t bit = c ? ((rhsn >> k) & 1u) : ((rhs2n >> k) & 1u);
        if ((in.flags & F_NULLING) && cn) bit = 1u;
        n |= bit << k;
      }
      accn = n;
    } break;
    case M_NEG: {
      // operators.h:63-71: UINT32/UINT64 negate to INT64 (the compiler inserts the cast).
      SSB_NUM_TYPES(in.t,
        SSB_UNROLL SSB_FOR_K { acc[k] = Codec<T>::enc(Arith<T>::neg(Codec<T>::dec(acc[k]))); })
    } break;
    case M_NOT: {
      SSB_UNROLL SSB_FOR_K { acc[k] = Codec<bool>::dec(acc[k]) ? 0u : 1u; }
    } break;
    case M_BNOT: {
      SSB_INT_TYPES(in.t,
        SSB_UNROLL SSB_FOR_K { acc[k] = Codec<T>::enc((T)~Codec<T>::dec(acc[k])); })
    } break;
    case M_ISODD: {
      // operators.h:108-128: arg % 2 (FLOAT/DOUBLE through int64).
      const bool neg = (in.flags & F_NEGATE) != 0;
      switch (in.t) {
        case T_F32:
          SSB_UNROLL SSB_FOR_K {
            acc[k] = (((f64_to_i64_x86((double)Codec<float>::dec(acc[k])) % 2) != 0) != neg) ? 1u : 0u; }
          break;
        case T_F64:
          SSB_UNROLL SSB_FOR_K {
            acc[k] = (((f64_to_i64_x86(Codec<double>::dec(acc[k])) % 2) != 0) != neg) ? 1u : 0u; }
          break;
        default:
          SSB_INT_TYPES(in.t,
            SSB_UNROLL SSB_FOR_K { acc[k] = (((Codec<T>::dec(acc[k]) % 2) != 0) != neg) ? 1u : 0u; })
      }
    } break;
    case M_ISNULL: {
      SSB_UNROLL SSB_FOR_K { acc[k] = (accn >> k) & 1u; }
      accn = 0;
    } break;
    case M_CAST: {
      switch (in.t) {
        case T_B8:
          SSB_ALL_TYPES2(in.t2,
            SSB_UNROLL SSB_FOR_K { acc[k] = Codec<T2>::enc((T2)(Codec<bool>::dec(acc[k]) ? 1 : 0)); })
          break;
        default:
          SSB_NUM_TYPES(in.t, SSB_ALL_TYPES2(in.t2,
            SSB_UNROLL SSB_FOR_K { acc[k] = Codec<T2>::enc(CastOp<T, T2>::f(Codec<T>::dec(acc[k]))); }))
      }
    } break;
    case M_D2DT: {
      // operators.h:59-61
      SSB_UNROLL SSB_FOR_K {
        acc[k] = Codec<int64_t>::enc(Arith<int64_t>::mul((int64_t)Codec<int32_t>::dec(acc[k]), 24LL * 3600000000LL));
      }
    } break;
    default: break;
  }
}

}  // namespace ssb
#endif  // SSB_CSRC_OPS_H_
