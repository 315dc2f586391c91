// compile.cc -- bound expression DAG -> accumulator-machine bytecode + shared-memory plan.
//
// The reference evaluates a bound tree node by node, each node writing a 1024-row scratch
// Block (expression/templated/abstract_bound_expressions.h:129-147,
// expression/infrastructure/basic_bound_expression.h:49-82). Here the whole DAG becomes one
// straight-line program: a value lives in per-thread registers (the accumulator) and touches
// shared memory only when it is an operand of a later instruction (Sethi-Ullman style:
// the deeper operand is evaluated first and parked in a shared-memory slot).
#include "program.h"

#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "common.h"

namespace ssb {
namespace {

struct Ref {           // an operand the ALU can read directly
  bool imm;
  bool null_const;
  int idx;             // slot or immediate index
  int phys;
  bool nullable;
};

class Compiler {
 public:
  Compiler(const ssb_expr_node* nodes, int n, int n_in, const int32_t* in_types,
           const int32_t* in_nullable, int tile, Program* prog, std::string* err)
      : nodes_(nodes), n_(n), n_in_(n_in), in_types_(in_types), in_nullable_(in_nullable),
        tile_(tile), prog_(prog), err_(err), code_(0) {
    memset(&prog->params, 0, sizeof(prog->params));
    info_.resize(n);
    tmp_used_.assign(kMaxTmp, false);
    tmp_nullable_.assign(kMaxTmp, false);
    n_tmp_high_ = 0;
  }

  int Fail(int code, const std::string& msg) {
    if (code_ == 0) { code_ = code; *err_ = msg; }
    return code;
  }
  int code() const { return code_; }

  // ---- pass 1: validate and annotate
  int Analyze() {
    for (int i = 0; i < n_; ++i) {
      const ssb_expr_node& nd = nodes_[i];
      Info& in = info_[i];
      in.phys = phys_of(nd.out_type);
      if (in.phys < 0) return Fail(SSB_ERROR_INVALID_ARGUMENT_TYPE, "unsupported node type");
      for (int a = 0; a < 3; ++a) {
        if (nd.arg[a] >= i && nd.op != SSB_OP_INPUT) return Fail(SSB_ERROR_INVALID_ARGUMENT_VALUE, "node argument is not an earlier node");
      }
      const int arity = Arity(nd.op);
      if (arity < 0) return Fail(SSB_ERROR_NOT_IMPLEMENTED, "unknown expression op");
      for (int a = 0; a < arity; ++a) {
        if (nd.arg[a] < 0) return Fail(SSB_ERROR_INVALID_ARGUMENT_VALUE, "missing node argument");
      }
      if (nd.flags & SSB_NODE_GUARDED) {
        if (!Guarded(nd)) return Fail(SSB_ERROR_INVALID_ARGUMENT_VALUE, "SSB_NODE_GUARDED needs a signaling DIV / MOD");
        if (nd.arg[2] < 0 || nd.arg[2] >= i) return Fail(SSB_ERROR_INVALID_ARGUMENT_VALUE, "guard is not an earlier node");
        if (info_[nd.arg[2]].phys != T_B8) return Fail(SSB_ERROR_INVALID_ARGUMENT_TYPE, "guard must be BOOL");
      }
      if (int rc = CheckTypes(i)) return rc;
      in.nullable = Nullable(i);
    }
    return 0;
  }

  void CountRef(int i) {
    if (info_[i].refs++ > 0) return;
    const ssb_expr_node& nd = nodes_[i];
    const int arity = Arity(nd.op);
    for (int a = 0; a < arity; ++a) CountRef(nd.arg[a]);
    if (Guarded(nd)) CountRef(nd.arg[2]);
  }

  // ---- pass 2: code generation
  void Emit(const Insn& in) {
    if (code_) return;
    if (prog_->params.n_insn >= kMaxInsn - 1) { Fail(SSB_ERROR_NOT_IMPLEMENTED, "expression too large (instruction budget)"); return; }
    prog_->params.insn[prog_->params.n_insn++] = in;
  }

  int AllocTmp(bool nullable) {
    for (int t = 0; t < kMaxTmp; ++t) {
      if (!tmp_used_[t]) {
        tmp_used_[t] = true;
        tmp_nullable_[t] = tmp_nullable_[t] || nullable;
        n_tmp_high_ = std::max(n_tmp_high_, t + 1);
        return t;
      }
    }
    Fail(SSB_ERROR_NOT_IMPLEMENTED, "expression too large (temporary budget)");
    return 0;
  }
  void FreeTmp(int slot) { if (slot >= n_in_) tmp_used_[slot - n_in_] = false; }

  int AddImm(const ssb_expr_node& nd) {
    uint64_t bits = 0;
    switch (phys_of(nd.out_type)) {
      case T_I32: bits = Codec<int32_t>::enc(nd.imm.i32); break;
      case T_U32: bits = Codec<uint32_t>::enc(nd.imm.u32); break;
      case T_I64: bits = Codec<int64_t>::enc(nd.imm.i64); break;
      case T_U64: bits = nd.imm.u64; break;
      case T_F32: bits = Codec<float>::enc(nd.imm.f32); break;
      case T_F64: bits = Codec<double>::enc(nd.imm.f64); break;
      case T_B8: bits = nd.imm.b ? 1 : 0; break;
    }
    ExprParams& p = prog_->params;
    int n_imm = static_cast<int>(imm_count_);
    for (int i = 0; i < n_imm; ++i) if (p.imm[i] == bits) return i;
    if (n_imm >= kMaxImm) { Fail(SSB_ERROR_NOT_IMPLEMENTED, "too many constants"); return 0; }
    p.imm[imm_count_++] = bits;
    return n_imm;
  }

  bool Simple(int i) const {
    const ssb_expr_node& nd = nodes_[i];
    return nd.op == SSB_OP_INPUT || nd.op == SSB_OP_CONST || info_[i].slot >= 0;
  }

  Ref RefOf(int i) {
    const ssb_expr_node& nd = nodes_[i];
    Ref r;
    r.phys = info_[i].phys;
    r.nullable = info_[i].nullable;
    r.null_const = false;
    if (nd.op == SSB_OP_CONST) {
      r.imm = true;
      r.null_const = (nd.flags & SSB_NODE_NULL) != 0;
      r.idx = AddImm(nd);
    } else {
      r.imm = false;
      r.idx = (nd.op == SSB_OP_INPUT) ? nd.arg[0] : info_[i].slot;
    }
    return r;
  }

  void SetRhs(Insn* in, const Ref& r) {
    in->a = static_cast<int16_t>(r.idx);
    in->rw = static_cast<uint8_t>(phys_width(r.phys));
    if (r.imm) {
      in->flags |= F_RHS_IMM;
      if (r.null_const) in->flags |= F_RHS_NULLK;
    } else if (r.nullable) {
      in->rhs_nullable |= 1;
    }
  }

  void EmitLoad(const Ref& r) {
    Insn in;
    memset(&in, 0, sizeof(in));
    in.kind = K_LOAD;
    in.t = static_cast<uint8_t>(r.phys);
    SetRhs(&in, r);
    Emit(in);
  }

  // Parks the accumulator (value of node i) in a fresh temporary slot.
  int EmitStore(int i) {
    const int t = AllocTmp(info_[i].nullable);
    Insn in;
    memset(&in, 0, sizeof(in));
    in.kind = K_STORE;
    in.t = static_cast<uint8_t>(info_[i].phys);
    in.rw = static_cast<uint8_t>(phys_width(info_[i].phys));
    in.a = static_cast<int16_t>(n_in_ + t);
    if (info_[i].nullable) in.rhs_nullable = 1;
    Emit(in);
    return n_in_ + t;
  }

  // Makes node i directly readable (a slot or an immediate), evaluating it if needed.
  // *owned = slot to free afterwards, or -1.
  Ref Materialize(int i, int* owned) {
    *owned = -1;
    if (!Simple(i)) {
      Gen(i);
      if (info_[i].slot < 0) {         // not parked by Gen (single use): park it now
        const int s = EmitStore(i);
        *owned = s;
        Ref r;
        r.imm = false; r.null_const = false; r.idx = s; r.phys = info_[i].phys;
        r.nullable = info_[i].nullable;
        return r;
      }
    }
    return RefOf(i);
  }

  // Leaves the value of node i in the accumulator.
  void Gen(int i) {
    if (code_) return;
    const ssb_expr_node& nd = nodes_[i];
    if (Simple(i)) { EmitLoad(RefOf(i)); return; }
    Insn in;
    memset(&in, 0, sizeof(in));
    const int p0 = nd.arg[0] >= 0 ? info_[nd.arg[0]].phys : 0;
    switch (nd.op) {
      case SSB_OP_CAST:
        Gen(nd.arg[0]);
        if (p0 != info_[i].phys) {
          in.kind = K_ALU1; in.mop = M_CAST; in.t = p0; in.t2 = info_[i].phys;
          Emit(in);
        }
        break;
      case SSB_OP_DATE_TO_DATETIME:
        Gen(nd.arg[0]); in.kind = K_ALU1; in.mop = M_D2DT; in.t = T_I32; in.t2 = T_I64; Emit(in);
        break;
      case SSB_OP_NEGATE:
        Gen(nd.arg[0]);
        if (p0 == T_U32 || p0 == T_U64) {   // operators.h:69-70: -static_cast<int64>(arg)
          Insn c; memset(&c, 0, sizeof(c));
          c.kind = K_ALU1; c.mop = M_CAST; c.t = p0; c.t2 = T_I64; Emit(c);
          in.t = T_I64;
        } else {
          in.t = p0;
        }
        in.kind = K_ALU1; in.mop = M_NEG; Emit(in);
        break;
      case SSB_OP_NOT: Gen(nd.arg[0]); in.kind = K_ALU1; in.mop = M_NOT; in.t = T_B8; Emit(in); break;
      case SSB_OP_BIT_NOT: Gen(nd.arg[0]); in.kind = K_ALU1; in.mop = M_BNOT; in.t = p0; Emit(in); break;
      case SSB_OP_IS_ODD:
      case SSB_OP_IS_EVEN:
        Gen(nd.arg[0]); in.kind = K_ALU1; in.mop = M_ISODD; in.t = p0;
        if (nd.op == SSB_OP_IS_EVEN) in.flags |= F_NEGATE;
        Emit(in);
        break;
      case SSB_OP_IS_NULL: Gen(nd.arg[0]); in.kind = K_ALU1; in.mop = M_ISNULL; in.t = p0; Emit(in); break;
      case SSB_OP_IF:
      case SSB_OP_NULLING_IF: {
        int own1, own2;
        Ref r1 = Materialize(nd.arg[1], &own1);
        Ref r2 = Materialize(nd.arg[2], &own2);
        Gen(nd.arg[0]);
        in.kind = K_ALU3; in.mop = M_SEL; in.t = T_B8; in.t2 = info_[i].phys;
        if (nd.op == SSB_OP_NULLING_IF) in.flags |= F_NULLING;
        SetRhs(&in, r1);
        in.b = static_cast<int16_t>(r2.idx);
        if (r2.imm) {
          in.flags |= F_RHS2_IMM;
          if (r2.null_const) in.rhs_nullable |= 4;   // bit2: rhs2 is a NULL constant
        } else if (r2.nullable) {
          in.rhs_nullable |= 2;
        }
        Emit(in);
        if (own1 >= 0) FreeTmp(own1);
        if (own2 >= 0) FreeTmp(own2);
      } break;
      default: GenBinary(i); break;
    }
    if (info_[i].refs > 1 && info_[i].slot < 0) {
      info_[i].slot = EmitStore(i);   // shared sub-expression: keep for the rest of the tile
    }
  }

  void GenBinary(int i) {
    const ssb_expr_node& nd = nodes_[i];
    int l = nd.arg[0], r = nd.arg[1];
    Insn in;
    memset(&in, 0, sizeof(in));
    in.kind = K_ALU2;
    switch (nd.op) {
      case SSB_OP_ADD: in.mop = M_ADD; break;
      case SSB_OP_SUB: in.mop = M_SUB; break;
      case SSB_OP_MUL: in.mop = M_MUL; break;
      case SSB_OP_DIV: in.mop = M_DIV; break;
      case SSB_OP_MOD: in.mop = M_MOD; break;
      case SSB_OP_EQ: in.mop = M_EQ; break;
      case SSB_OP_NE: in.mop = M_EQ; in.flags |= F_NEGATE; break;
      case SSB_OP_LT: in.mop = M_LT; break;
      // operators.h:282-306: Greater(a,b) = less(b,a); LessOrEqual(a,b) = !less(b,a);
      // GreaterOrEqual(a,b) = !less(a,b)
      case SSB_OP_GT: in.mop = M_LT; std::swap(l, r); break;
      case SSB_OP_LE: in.mop = M_LT; std::swap(l, r); in.flags |= F_NEGATE; break;
      case SSB_OP_GE: in.mop = M_LT; in.flags |= F_NEGATE; break;
      case SSB_OP_AND: in.mop = M_AND3; break;
      case SSB_OP_OR: in.mop = M_OR3; break;
      case SSB_OP_XOR: in.mop = M_XOR3; break;
      case SSB_OP_AND_NOT: in.mop = M_ANDNOT3; break;
      case SSB_OP_BIT_AND: in.mop = M_BAND; break;
      case SSB_OP_BIT_OR: in.mop = M_BOR; break;
      case SSB_OP_BIT_XOR: in.mop = M_BXOR; break;
      case SSB_OP_BIT_AND_NOT: in.mop = M_BANDNOT; break;
      case SSB_OP_SHL: in.mop = M_SHL; break;
      case SSB_OP_SHR: in.mop = M_SHR; break;
      case SSB_OP_IF_NULL: in.mop = M_IFNULL; break;
      default: Fail(SSB_ERROR_NOT_IMPLEMENTED, "unknown binary op"); return;
    }
    if (nd.flags & SSB_NODE_ZERO_NULLS) in.flags |= F_ZERO_NULLS;
    if (nd.flags & SSB_NODE_ZERO_FAILS) { in.flags |= F_ZERO_FAILS; prog_->has_signaling = true; }
    in.t = static_cast<uint8_t>(info_[l].phys);    // true left operand
    in.t2 = static_cast<uint8_t>(info_[r].phys);   // true right operand
    if (Guarded(nd)) {
      // the guard travels as the third operand (rhs2) of a K_ALU3 form
      int own_g, own_r;
      Ref rg = Materialize(nd.arg[2], &own_g);
      Ref rr = Materialize(r, &own_r);
      Gen(l);
      in.kind = K_ALU3;
      in.flags |= F_GUARDED;
      SetRhs(&in, rr);
      in.b = static_cast<int16_t>(rg.idx);
      if (rg.imm) {
        in.flags |= F_RHS2_IMM;
        if (rg.null_const) in.rhs_nullable |= 4;
      } else if (rg.nullable) {
        in.rhs_nullable |= 2;
      }
      Emit(in);
      if (own_r >= 0) FreeTmp(own_r);
      if (own_g >= 0) FreeTmp(own_g);
      return;
    }
    if (Simple(r)) {
      Gen(l);
      SetRhs(&in, RefOf(r));
      Emit(in);
    } else if (Simple(l)) {
      Gen(r);
      in.flags |= F_REV;
      SetRhs(&in, RefOf(l));
      Emit(in);
    } else {
      int own;
      Ref rr = Materialize(r, &own);
      Gen(l);
      SetRhs(&in, rr);
      Emit(in);
      if (own >= 0) FreeTmp(own);
    }
  }

  uint32_t sink_bytes_ = 0;   // > 0: aggregation sink instead of output staging (public: set by compile_program)

  int Finish(const int32_t* outputs, int n_out, int predicate, uint32_t smem_budget,
             uint32_t smem_max) {
    ExprParams& p = prog_->params;
    if (n_out > kMaxOut) return Fail(SSB_ERROR_NOT_IMPLEMENTED, "too many output columns");
    if (n_in_ > kMaxIn) return Fail(SSB_ERROR_NOT_IMPLEMENTED, "too many input columns");
    for (int j = 0; j < n_out; ++j) {
      if (outputs[j] < 0 || outputs[j] >= n_) return Fail(SSB_ERROR_INVALID_ARGUMENT_VALUE, "bad output node");
      CountRef(outputs[j]);
    }
    if (predicate >= 0) {
      if (predicate >= n_) return Fail(SSB_ERROR_INVALID_ARGUMENT_VALUE, "bad predicate node");
      if (info_[predicate].phys != T_B8) return Fail(SSB_ERROR_INVALID_ARGUMENT_TYPE, "predicate must be BOOL");
      CountRef(predicate);
    }
    // The predicate runs first: K_PRED turns the pass bits into in-tile positions, and every
    // output is then stored already compacted into the tile's output staging buffer.
    p.n_in = n_in_;
    p.n_out = n_out;
    if (predicate >= 0) {
      Gen(predicate);
      Insn in;
      memset(&in, 0, sizeof(in));
      in.kind = K_PRED;
      Emit(in);
      p.has_pred = 1;
    }
    for (int j = 0; j < n_out; ++j) {
      const int node = outputs[j];
      const ssb_expr_node& nd = nodes_[node];
      if (sink_bytes_) {
        // aggregation sink: nothing is staged. The sink reads the value from the slot that holds it
        // -- the input column itself for a pass-through output, else a temporary kept for the tile.
        int owned;
        Ref r = Materialize(node, &owned);
        if (r.imm) {   // a constant output: give it a slot
          EmitLoad(r);
          r.idx = EmitStore(node);
          r.imm = false;
        }
        sink_slot_[j] = r.idx;
        p.sink_src_slot[j] = static_cast<int16_t>(r.idx);
        p.sink_src_w[j] = static_cast<uint8_t>(phys_width(r.phys));
        p.sink_src_nullable[j] = r.nullable ? 1 : 0;
        p.out_width[j] = static_cast<uint8_t>(phys_width(info_[node].phys));
        p.out_nullable[j] = info_[node].nullable ? 1 : 0;
        prog_->out_types.push_back(nd.out_type);
        prog_->out_nullable.push_back(info_[node].nullable ? 1 : 0);
        continue;
      }
      Gen(node);
      Insn in;
      memset(&in, 0, sizeof(in));
      in.kind = K_OUT;
      in.t = static_cast<uint8_t>(info_[node].phys);
      in.rw = static_cast<uint8_t>(phys_width(info_[node].phys));
      in.a = static_cast<int16_t>(j);
      if (info_[node].nullable) in.rhs_nullable = 1;
      Emit(in);
      p.out_width[j] = static_cast<uint8_t>(phys_width(info_[node].phys));
      p.out_nullable[j] = info_[node].nullable ? 1 : 0;
      prog_->out_types.push_back(nd.out_type);
      prog_->out_nullable.push_back(info_[node].nullable ? 1 : 0);
    }
    Insn end;
    memset(&end, 0, sizeof(end));
    end.kind = K_END;
    p.insn[p.n_insn] = end;
    if (code_) return code_;

    // ---- shared-memory plan
    p.n_tmp = n_tmp_high_;
    uint32_t off = 0;
    int nullw = 0;
    prog_->bytes_in_row = 0;
    uint32_t tx = 0;
    for (int i = 0; i < n_in_; ++i) {
      const int w = width_of(in_types_[i]);
      p.in_width[i] = static_cast<uint8_t>(w);
      p.in_off[i] = off;
      off += tile_ * w;                 // tile_ * w is a multiple of 128: 16-byte aligned
      tx += tile_ * w;
      p.in_nullable[i] = in_nullable_[i] ? 1 : 0;
      p.in_nullw[i] = in_nullable_[i] ? nullw++ : -1;
      prog_->bytes_in_row += w;
    }
    p.stage_bytes = off;
    p.stage_nullw = nullw;
    p.stage_tx_bytes = tx;     // null words are added per run (a column may come without bitmap)
    prog_->bytes_out_row = 0;
    for (int j = 0; j < n_out; ++j) prog_->bytes_out_row += p.out_width[j];

    // output staging: per column tile_ elements (+ tile_ null bytes when nullable), 16-byte aligned
    uint32_t ooff = 0;
    for (int j = 0; j < n_out; ++j) {
      p.out_off[j] = ooff;
      ooff += tile_ * p.out_width[j];
      ooff = (ooff + 15) & ~15u;
      if (p.out_nullable[j]) { p.out_null_off[j] = ooff; ooff += tile_; } else { p.out_null_off[j] = 0xffffffffu; }
    }
    p.out_bytes = (ooff + 127) & ~127u;
    const uint32_t tmp_bytes = static_cast<uint32_t>(p.n_tmp) * tile_ * 8;
    const uint32_t tmp_nullw = static_cast<uint32_t>(p.n_tmp);
    // header: barriers (64 B) + scan / reduce scratch (512 B)
    p.off_bar = 0;
    p.off_scan = 64;
    p.off_nullw = 64 + 512;
    // Filter defers the copy-out by two tiles (the wave's counts are then always published);
    // without a predicate the output position is known at once and one tile of slack suffices.
    uint32_t defer = sink_bytes_ ? 0 : (p.has_pred ? kMaxDefer : 1);
    if (const char* env = getenv("SSB200_EXPR_DEFER")) {   // experiment: copy-out lag of Filter plans (1 or 2 tiles)
      const int d = atoi(env);
      if (!sink_bytes_ && p.has_pred && d >= 1 && d <= kMaxDefer) defer = static_cast<uint32_t>(d);
    }
    p.defer = static_cast<int32_t>(defer);
    if (sink_bytes_) p.out_bytes = 0;
    int stages = kMaxStages;
    for (;; --stages) {
      const uint32_t nullw_bytes = (stages * p.stage_nullw + tmp_nullw) * (tile_ / 32) * 4;
      // per-stage table of pre-resolved instructions (16 bytes each, + terminator)
      const uint32_t itab_off = (p.off_nullw + nullw_bytes + 15) & ~15u;
      const uint32_t itab_bytes = static_cast<uint32_t>(stages) * (p.n_insn + 1) * 16;
      const uint32_t data_off = (itab_off + itab_bytes + 1023) & ~1023u;
      const uint32_t sink_off = (data_off + stages * p.stage_bytes + tmp_bytes + 15) & ~15u;
      const uint32_t total = sink_bytes_ ? sink_off + sink_bytes_
                                         : data_off + stages * p.stage_bytes + tmp_bytes + (defer + 1) * p.out_bytes;
      if ((total <= smem_budget && stages >= 2) || stages == 1 || (stages == 2 && total <= smem_max)) {
        if (total > smem_max) return Fail(SSB_ERROR_NOT_IMPLEMENTED, "expression needs more shared memory than one SM has");
        p.stages = stages;
        p.tile = tile_;
        p.off_itab = itab_off;
        p.off_data = data_off;
        p.off_tmp = data_off + stages * p.stage_bytes;
        p.off_out = p.off_tmp + tmp_bytes;
        p.sink_off = sink_off;
        prog_->smem_bytes = total;
        break;
      }
    }
    AssignFastCodes();
    if (sink_bytes_) {
      for (int j = 0; j < n_out; ++j) p.sink_src_off[j] = SlotOffset(sink_slot_[j], true);
    }
    return 0;
  }

  // Pre-decodes operand addresses and marks the instructions that have a straight-line case
  // in the kernel. Runs after the shared-memory plan is final.
  void AssignFastCodes() {
    ExprParams& p = prog_->params;
    for (int i = 0; i < p.n_insn; ++i) {
      Insn& in = p.insn[i];
      in.code = C_GENERIC;
      in.off_a = SlotOffset(in.a, in.kind != K_OUT && !(in.flags & F_RHS_IMM));
      in.off_b = SlotOffset(in.b, in.kind == K_ALU3 && !(in.flags & F_RHS2_IMM));
      const bool imm = (in.flags & F_RHS_IMM) != 0;
      const bool rhs_clean = imm ? !(in.flags & F_RHS_NULLK) : !(in.rhs_nullable & 1);
      switch (in.kind) {
        case K_LOAD:
          if (!rhs_clean) break;
          if (imm) in.code = C_LOADK;
          else if (in.rw == 8) in.code = C_LOAD8;
          else if (in.rw == 4) in.code = C_LOAD4;
          break;
        case K_OUT:
          if (in.rhs_nullable & 1) break;
          if (in.rw == 8) in.code = C_OUT8; else if (in.rw == 4) in.code = C_OUT4;
          break;
        case K_PRED: in.code = C_PRED; break;
        case K_STORE:
          if (in.rhs_nullable & 1) break;
          if (in.rw == 8) in.code = C_STORE8; else if (in.rw == 4) in.code = C_STORE4;
          break;
        case K_ALU2: {
          if (!rhs_clean) break;
          if (!imm && in.rw != phys_width(in.t)) break;
          if ((in.mop == M_LT || in.mop == M_EQ) && in.t != in.t2) break;
          if (in.mop == M_AND3 && !imm && !(in.flags & F_REV)) { in.code = C_AND3_S; break; }
          if (in.mop == M_OR3 && !imm) { in.code = C_OR3_S; break; }
          int base = -1;
          if (in.t == T_I64) base = C_BIN_I64; else if (in.t == T_F64) base = C_BIN_F64; else if (in.t == T_I32) base = C_BIN_I32;
          if (base < 0) break;
          int op = -1;
          const bool rev = (in.flags & F_REV) != 0;
          switch (in.mop) {
            case M_ADD: op = B_ADD; break;
            case M_SUB: op = rev ? B_SUBR : B_SUB; break;
            case M_MUL: op = B_MUL; break;
            case M_LT: op = rev ? B_GT : B_LT; break;
            case M_EQ: op = B_EQ; break;
            default: break;
          }
          if (op >= 0) in.code = static_cast<uint16_t>(base + 4 * op + (imm ? 1 : 0));
        } break;
        default: break;
      }
    }
    // the unfused list (fast codes of the +0 / +1 forms only) is what group.cu's row evaluator runs
    prog_->generic.assign(p.insn, p.insn + p.n_insn);
    // Peephole: a clean LOAD followed by a fast binary op becomes one two-operand instruction
    // (the left operand is read from its slot instead of the accumulator).
    int w = 0;
    for (int i = 0; i < p.n_insn; ++i) {
      Insn cur = p.insn[i];
      if (i + 1 < p.n_insn && (cur.code == C_LOAD8 || cur.code == C_LOAD4)) {
        Insn& next = p.insn[i + 1];
        const bool bin = next.code >= C_BIN_BASE && next.code < C_BIN_END && ((next.code - C_BIN_BASE) % 4) < 2;
        // operand order (F_REV) is part of the fast code, so the fused form needs no care
        if (bin && phys_width(next.t) == cur.rw) {
          next.off_b = cur.off_a;
          next.code = static_cast<uint16_t>(next.code + 2);
          continue;   // drop the LOAD
        }
      }
      p.insn[w++] = cur;
    }
    p.n_insn = w;
    // Peephole 2: MUL then ADD of a clean slot -> one multiply-add; a fast compare directly
    // followed by K_PRED -> the compare feeds the compaction itself (no 0/1 materialisation).
    w = 0;
    for (int i = 0; i < p.n_insn; ++i) {
      Insn cur = p.insn[i];
      if (i + 1 < p.n_insn && cur.code >= C_BIN_BASE && cur.code < C_BIN_END) {
        const Insn& next = p.insn[i + 1];
        const int rel = cur.code - C_BIN_BASE;
        const int type_base = rel / 28, op = (rel % 28) / 4, form = rel % 4;
        const bool next_add_slot = next.code >= C_BIN_BASE && next.code < C_BIN_END &&
                                   (next.code - C_BIN_BASE) / 28 == type_base &&
                                   ((next.code - C_BIN_BASE) % 28) == 4 * B_ADD + 0;
        if (op == B_MUL && (form == 0 || form == 2) && next_add_slot) {
          // x = slot a (multiplier), y = slot b of the ADD, z = left slot of the fused MUL
          const uint32_t left = cur.off_b;
          cur.off_b = next.off_a;
          cur.code = static_cast<uint16_t>(C_MAD_I64 + 2 * type_base + (form == 2 ? 1 : 0));
          cur.pad2 = 0;
          mad_left_[w] = left;
          p.insn[w++] = cur;
          ++i;
          continue;
        }
        if ((op == B_LT || op == B_GT || op == B_EQ) && next.code == C_PRED) {
          cur.flags |= F_THEN_PRED;
          p.insn[w++] = cur;
          ++i;
          continue;
        }
      }
      mad_left_[w] = 0;
      p.insn[w++] = cur;
    }
    p.n_insn = w;
    for (int i = 0; i < p.n_insn; ++i) p.insn_c[i] = mad_left_[i];
    Insn end;
    memset(&end, 0, sizeof(end));
    end.kind = K_END;
    p.insn[p.n_insn] = end;
  }

  // Byte offset of a slot's data from the shared-memory base; inputs are stage relative
  // (bit 31 set: the kernel adds stage * stage_bytes).
  uint32_t SlotOffset(int slot, bool used) const {
    const ExprParams& p = prog_->params;
    if (!used || slot < 0) return 0;
    if (slot < p.n_in) return (p.off_data + p.in_off[slot]) | 0x80000000u;
    return p.off_tmp + static_cast<uint32_t>(slot - p.n_in) * tile_ * 8;
  }

 private:
  uint32_t mad_left_[kMaxInsn + 1];
  int sink_slot_[kMaxOut] = {0};
  struct Info {
    int phys;
    bool nullable;
    int refs;
    int slot;     // slot holding the value for the rest of the tile, or -1
    Info() : phys(0), nullable(false), refs(0), slot(-1) {}
  };

  static int Arity(int op) {
    switch (op) {
      case SSB_OP_INPUT: case SSB_OP_CONST: return 0;
      case SSB_OP_CAST: case SSB_OP_DATE_TO_DATETIME: case SSB_OP_NEGATE: case SSB_OP_NOT:
      case SSB_OP_BIT_NOT: case SSB_OP_IS_ODD: case SSB_OP_IS_EVEN: case SSB_OP_IS_NULL:
        return 1;
      case SSB_OP_ADD: case SSB_OP_SUB: case SSB_OP_MUL: case SSB_OP_DIV: case SSB_OP_MOD:
      case SSB_OP_EQ: case SSB_OP_NE: case SSB_OP_LT: case SSB_OP_LE: case SSB_OP_GT:
      case SSB_OP_GE: case SSB_OP_AND: case SSB_OP_OR: case SSB_OP_XOR: case SSB_OP_AND_NOT:
      case SSB_OP_BIT_AND: case SSB_OP_BIT_OR: case SSB_OP_BIT_XOR: case SSB_OP_BIT_AND_NOT:
      case SSB_OP_SHL: case SSB_OP_SHR: case SSB_OP_IF_NULL:
        return 2;
      case SSB_OP_IF: case SSB_OP_NULLING_IF: return 3;
      default: return -1;
    }
  }

  static bool Guarded(const ssb_expr_node& nd) {
    return (nd.flags & SSB_NODE_GUARDED) && (nd.flags & SSB_NODE_ZERO_FAILS) && (nd.op == SSB_OP_DIV || nd.op == SSB_OP_MOD);
  }
  static bool IsInt(int p) { return p == T_I32 || p == T_I64 || p == T_U32 || p == T_U64; }
  static bool IsNum(int p) { return p != T_B8; }

  int CheckTypes(int i) {
    const ssb_expr_node& nd = nodes_[i];
    const int out = info_[i].phys;
    const int a = nd.arg[0] >= 0 && nd.op != SSB_OP_INPUT ? info_[nd.arg[0]].phys : -1;
    const int b = nd.arg[1] >= 0 && Arity(nd.op) >= 2 ? info_[nd.arg[1]].phys : -1;
    const int c = nd.arg[2] >= 0 && Arity(nd.op) >= 3 ? info_[nd.arg[2]].phys : -1;
    const char* bad = NULL;
    switch (nd.op) {
      case SSB_OP_INPUT:
        if (nd.arg[0] < 0 || nd.arg[0] >= n_in_) return Fail(SSB_ERROR_INVALID_ARGUMENT_VALUE, "input index out of range");
        if (phys_of(in_types_[nd.arg[0]]) != out) bad = "input node type differs from the column type";
        break;
      case SSB_OP_CONST: break;
      case SSB_OP_CAST: break;
      case SSB_OP_DATE_TO_DATETIME: if (a != T_I32 || out != T_I64) bad = "DATE_TO_DATETIME"; break;
      case SSB_OP_NEGATE:
        if (!IsNum(a)) bad = "NEGATE of a non-numeric value";
        else if ((a == T_U32 || a == T_U64) ? out != T_I64 : out != a) bad = "NEGATE result type";
        break;
      case SSB_OP_ADD: case SSB_OP_SUB: case SSB_OP_MUL: case SSB_OP_DIV:
        if (!IsNum(a) || a != b || out != a) bad = "arithmetic needs equal numeric operand and result types";
        break;
      case SSB_OP_MOD:
        if (!IsNum(a) || a != b) bad = "MOD needs equal numeric operand types";
        else if ((a == T_F32 || a == T_F64) ? out != T_I64 : out != a) bad = "MOD result type";
        break;
      case SSB_OP_IS_ODD: case SSB_OP_IS_EVEN:
        if (!IsNum(a) || out != T_B8) bad = "IS_ODD/IS_EVEN"; break;
      case SSB_OP_EQ: case SSB_OP_NE: case SSB_OP_LT: case SSB_OP_LE: case SSB_OP_GT: case SSB_OP_GE:
        if (out != T_B8) bad = "comparison result must be BOOL";
        else if (a != b && !(IsInt(a) && IsInt(b))) bad = "comparison of different non-integer types";
        break;
      case SSB_OP_AND: case SSB_OP_OR: case SSB_OP_XOR: case SSB_OP_AND_NOT:
        if (a != T_B8 || b != T_B8 || out != T_B8) bad = "logic needs BOOL operands"; break;
      case SSB_OP_NOT: if (a != T_B8 || out != T_B8) bad = "NOT needs BOOL"; break;
      case SSB_OP_BIT_AND: case SSB_OP_BIT_OR: case SSB_OP_BIT_XOR: case SSB_OP_BIT_AND_NOT:
        if (!IsInt(a) || a != b || out != a) bad = "bitwise ops need equal integer types"; break;
      case SSB_OP_BIT_NOT: if (!IsInt(a) || out != a) bad = "BIT_NOT needs an integer"; break;
      case SSB_OP_SHL: case SSB_OP_SHR:
        if (!IsInt(a) || !IsInt(b) || out != a) bad = "shift needs integer operands"; break;
      case SSB_OP_IS_NULL: if (out != T_B8) bad = "IS_NULL result must be BOOL"; break;
      case SSB_OP_IF_NULL: if (a != b || out != a) bad = "IF_NULL needs equal types"; break;
      case SSB_OP_IF: case SSB_OP_NULLING_IF:
        if (a != T_B8 || b != c || out != b) bad = "IF needs a BOOL condition and equal branch types"; break;
      default: return Fail(SSB_ERROR_NOT_IMPLEMENTED, "unknown expression op");
    }
    if (bad) return Fail(SSB_ERROR_INVALID_ARGUMENT_TYPE, bad);
    return 0;
  }

  bool Nullable(int i) const {
    const ssb_expr_node& nd = nodes_[i];
    const bool a = nd.arg[0] >= 0 && nd.op != SSB_OP_INPUT && info_[nd.arg[0]].nullable;
    const bool b = Arity(nd.op) >= 2 && info_[nd.arg[1]].nullable;
    const bool c = Arity(nd.op) >= 3 && info_[nd.arg[2]].nullable;
    switch (nd.op) {
      case SSB_OP_INPUT: return in_nullable_[nd.arg[0]] != 0;
      case SSB_OP_CONST: return (nd.flags & SSB_NODE_NULL) != 0;
      case SSB_OP_IS_NULL: return false;
      case SSB_OP_IF_NULL: return a && b;
      case SSB_OP_IF: return b || c;
      case SSB_OP_NULLING_IF: return a || b || c;
      case SSB_OP_DIV: case SSB_OP_MOD: return a || b || (nd.flags & SSB_NODE_ZERO_NULLS);
      default: return a || b;
    }
  }

  const ssb_expr_node* nodes_;
  int n_, n_in_;
  const int32_t* in_types_;
  const int32_t* in_nullable_;
  int tile_;
  Program* prog_;
  std::string* err_;
  int code_;
  std::vector<Info> info_;
  std::vector<bool> tmp_used_, tmp_nullable_;
  int n_tmp_high_;
  size_t imm_count_ = 0;
};

}  // namespace

int compile_program(const ssb_expr_node* nodes, int32_t n_nodes, int32_t n_inputs,
                    const int32_t* input_types, const int32_t* input_nullable,
                    const int32_t* outputs, int32_t n_outputs, int32_t predicate,
                    int32_t tile, uint32_t smem_budget, uint32_t smem_max, Program* prog, std::string* err,
                    uint32_t sink_bytes_per_thread, uint32_t sink_fixed_bytes, int32_t threads) {
  if (n_nodes <= 0 || n_inputs < 0 || n_outputs < 0 || (n_outputs == 0 && predicate < 0)) {
    *err = "empty program";
    return SSB_ERROR_INVALID_ARGUMENT_VALUE;
  }
  if (n_inputs > kMaxIn) { *err = "too many input columns"; return SSB_ERROR_NOT_IMPLEMENTED; }
  for (int i = 0; i < n_inputs; ++i) {
    if (phys_of(input_types[i]) < 0) { *err = "unsupported input column type"; return SSB_ERROR_INVALID_ARGUMENT_TYPE; }
  }
  prog->nodes.assign(nodes, nodes + n_nodes);
  prog->input_types.assign(input_types, input_types + n_inputs);
  prog->input_nullable.assign(input_nullable, input_nullable + n_inputs);
  prog->outputs.assign(outputs, outputs + n_outputs);
  prog->predicate = predicate;
  prog->out_types.clear();
  prog->out_nullable.clear();
  prog->has_signaling = false;
  Compiler c(nodes, n_nodes, n_inputs, input_types, input_nullable, tile, prog, err);
  if (int rc = c.Analyze()) return rc;
  if (sink_bytes_per_thread) c.sink_bytes_ = sink_fixed_bytes + sink_bytes_per_thread * static_cast<uint32_t>(threads);
  return c.Finish(outputs, n_outputs, predicate, smem_budget, smem_max);
}

}  // namespace ssb

extern "C" int ssb_program_plan(const ssb_expr_node* nodes, int32_t n_nodes, int32_t n_inputs, const int32_t* input_types,
                                const int32_t* input_nullable, const int32_t* outputs, int32_t n_outputs, int32_t predicate,
                                int32_t tile, uint32_t smem_budget, ssb_plan_info* info, char* err, int32_t err_len) {
  std::string message;
  int rc = 0;
  ssb::Program prog;
  if (nodes == NULL || info == NULL || (n_inputs > 0 && (input_types == NULL || input_nullable == NULL)) ||
      (n_outputs > 0 && outputs == NULL)) {
    rc = SSB_ERROR_INVALID_ARGUMENT_VALUE;
    message = "null argument";
  } else if (tile <= 0 || tile % 128 != 0) {
    rc = SSB_ERROR_INVALID_ARGUMENT_VALUE;
    message = "tile must be a positive multiple of 128 rows";
  } else {
    rc = ssb::compile_program(nodes, n_nodes, n_inputs, input_types, input_nullable, outputs, n_outputs, predicate, tile,
                              smem_budget, 232448u /* 227 KB: the most one CTA can opt into on sm_100a */, &prog, &message);
  }
  if (err != NULL && err_len > 0) {
    strncpy(err, message.c_str(), static_cast<size_t>(err_len) - 1);
    err[err_len - 1] = 0;
  }
  if (rc != 0) return rc;
  memset(info, 0, sizeof(*info));
  info->tile = prog.params.tile;
  info->stages = prog.params.stages;
  info->smem_bytes = static_cast<int32_t>(prog.smem_bytes);
  info->n_insn = prog.params.n_insn;
  info->n_tmp = prog.params.n_tmp;
  info->bytes_per_input_row = prog.bytes_in_row;
  info->bytes_per_output_row = prog.bytes_out_row;
  info->has_signaling = prog.has_signaling ? 1 : 0;
  info->n_outputs = static_cast<int32_t>(prog.out_types.size());
  for (size_t j = 0; j < prog.out_types.size() && j < 16; ++j) {
    info->out_types[j] = prog.out_types[j];
    info->out_nullable[j] = prog.out_nullable[j];
  }
  return 0;
}
