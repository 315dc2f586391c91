// program.h -- the compiled form of a bound expression DAG: bytecode for the accumulator
// machine of expr_kernel.cu plus the shared-memory plan of one CTA.
#ifndef SSB_CSRC_PROGRAM_H_
#define SSB_CSRC_PROGRAM_H_

#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/supersonic_b200.h"
#include "ops.h"

namespace ssb {

enum {
  kMaxTile = 4096,            // rows per tile of the largest kernel variant
  kMaxInsn = 64,
  kMaxImm = 16,
  kMaxIn = 12,
  kMaxOut = 16,
  kMaxTmp = 10,
  kMaxStages = 4,
};

// Kernel parameters (passed by value, lives in the constant bank).
struct ExprParams {
  Insn insn[kMaxInsn];
  uint64_t imm[kMaxImm];
  uint32_t insn_c[kMaxInsn];            // third operand offset of multiply-add instructions (encoded like Insn::off_a)
  int32_t n_insn;
  // ---- inputs: slot i < n_in is input column i of the current pipeline stage
  int32_t n_in;
  const void* in_data[kMaxIn];
  const uint32_t* in_nulls[kMaxIn];     // NULL = column has no bitmap in this run
  uint8_t in_width[kMaxIn];
  uint8_t in_nullable[kMaxIn];          // compiled with null words for this input
  uint32_t in_off[kMaxIn];              // byte offset of the column tile inside a stage
  int32_t in_nullw[kMaxIn];             // index of its null-word row inside a stage, -1
  // ---- temporaries: slot n_in + t (one copy; dead once the tile is evaluated)
  int32_t n_tmp;
  // ---- outputs: staged (compacted) in shared memory, copied out one tile later
  int32_t n_out;
  void* out_data[kMaxOut];
  uint32_t* out_nulls[kMaxOut];
  uint32_t out_off[kMaxOut];            // byte offset of the column inside an output buffer
  uint32_t out_null_off[kMaxOut];       // byte offset of its null bytes, or 0xffffffff
  uint8_t out_width[kMaxOut];
  uint8_t out_nullable[kMaxOut];
  // ---- shared memory plan (byte offsets from the 1024-byte aligned base)
  uint32_t off_bar, off_scan, off_nullw, off_itab, off_data, off_tmp, off_out;
  uint32_t stage_bytes;                 // data bytes of one input stage
  uint32_t stage_nullw;                 // null-word rows per stage (nullable inputs)
  uint32_t stage_tx_bytes;              // bytes one full-tile TMA fill delivers (data only)
  uint32_t out_bytes;                   // bytes of one output buffer
  int32_t stages;
  int32_t tile;                         // rows per tile (threads * rows per thread of the variant)
  int32_t defer;                        // copy-out lags evaluation by this many tiles (out buffers = defer + 1)
  // ---- run
  int64_t rows;
  int64_t num_tiles;
  int32_t has_pred;
  int32_t use_tma;
  int32_t debug_nowait;
  int32_t fill_nullw_tma, fill_nullw_plain;   // some null words are not delivered by TMA / any nullable input
  unsigned long long* tile_status;      // per-tile kept-row counts | valid bit (Filter)
  int64_t* d_out_rows;
  int32_t* d_fail;
  // ---- aggregation sink (expr_kernel<..., SINK = true>): the outputs are the group-by keys
  // followed by the aggregate inputs and go straight into the aggregation, nothing is staged
  const void* sink_gp;                  // GroupParams in device memory (group_device.h)
  int32_t sink_n_keys, sink_n_aggs, sink_groups;   // local group entries per CTA (<= kTinyGroups)
  uint32_t sink_off;                    // shared-memory offset of the sink area
  uint32_t sink_out_aggs[kMaxOut];      // bit a: aggregate a accumulates output column j
  uint32_t sink_count_star;             // bit a: aggregate a is COUNT(*)
  uint32_t sink_pad;                    // two bits per aggregate: TA_* accumulate code
  // output j of a sink program is not staged: the sink reads it from the shared-memory slot that holds it
  // (an input column of the current stage, or a temporary the program stored)
  uint32_t sink_src_off[kMaxOut];       // byte offset of the slot, encoded like Insn::off_a (bit 31: stage relative)
  int16_t sink_src_slot[kMaxOut];       // slot index (for its null words)
  uint8_t sink_src_w[kMaxOut];          // element width: 1, 4 or 8
  uint8_t sink_src_nullable[kMaxOut];
  uint32_t sink_seen_mask;              // bit a: aggregate a has a nullable input (its partials carry a seen byte)
  int32_t sink_keys_not_null;           // no key output can be NULL: the lookup skips the NULL words
};

enum { kMaxDefer = 2 };   // Filter: tile i is copied out while tile i + defer is evaluated

struct Program {
  // compile-time description
  std::vector<ssb_expr_node> nodes;
  std::vector<int32_t> input_types;
  std::vector<int32_t> input_nullable;
  std::vector<int32_t> outputs;
  int32_t predicate;
  // compiled
  ExprParams params;          // pointers / run fields are filled per run
  std::vector<int32_t> out_types;
  std::vector<int32_t> out_nullable;
  uint32_t smem_bytes;
  int32_t variant;            // kernel variant (threads x rows per thread)
  int32_t bytes_in_row, bytes_out_row;
  bool has_signaling;
  // The accumulator-machine program before the fast-path peepholes (K_LOAD / K_STORE / K_ALU* /
  // K_PRED / K_OUT only, operands by slot index): what group.cu's fused row evaluator runs.
  std::vector<Insn> generic;
};

// Compiles nodes into prog->params (bytecode + shared-memory plan for `smem_budget` bytes per
// CTA). Returns 0 or an SSB_ERROR_* with *err set. Pure host code (unit-tested without a GPU).
// sink_bytes_per_thread > 0 compiles for the aggregation sink: no output staging buffers, and
// sink_fixed_bytes + sink_bytes_per_thread * consumer threads of shared memory reserved at
// params.sink_off.
int compile_program(const ssb_expr_node* nodes, int32_t n_nodes, int32_t n_inputs,
                    const int32_t* input_types, const int32_t* input_nullable,
                    const int32_t* outputs, int32_t n_outputs, int32_t predicate,
                    int32_t tile, uint32_t smem_budget, uint32_t smem_max, Program* prog, std::string* err,
                    uint32_t sink_bytes_per_thread = 0, uint32_t sink_fixed_bytes = 0, int32_t threads = 0);

}  // namespace ssb

struct ssb_program;
namespace ssb {
// Aggregation sink of expr_kernel (expr_kernel.cu), driven by group.cu.
int sink_program_for(ssb_program* base, int n_keys, int n_aggs, int groups, ssb_program** out);
int launch_program_sink(ssb_program* sp, const ssb_column* inputs, int64_t rows, const void* d_gp, int n_keys,
                        int n_aggs, int groups, const uint32_t* out_aggs, uint32_t count_star, uint32_t pad_codes,
                        uint32_t seen_mask);
}  // namespace ssb

struct ssb_program {
  ssb_ctx* ctx;
  ssb::Program prog;
  int max_ctas_per_sm;
  // lazily compiled twin for the aggregation sink (keyed by the sink's shared-memory need)
  ssb_program* sink = nullptr;
  uint32_t sink_key = 0;
  // Wide plans run as column groups: each part evaluates the predicate and a contiguous range of
  // the outputs from only the input columns it needs (empty = one kernel does everything).
  struct Part {
    ssb_program* prog;
    int first_out, n_out;
    std::vector<int> inputs;   // indices into the whole program's input list
  };
  std::vector<Part> parts;
  ~ssb_program() {
    delete sink;
    for (size_t i = 0; i < parts.size(); ++i) delete parts[i].prog;
  }
};

#endif  // SSB_CSRC_PROGRAM_H_
