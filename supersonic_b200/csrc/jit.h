// jit.h -- plan-specialised kernels compiled at run time (NVRTC -> sm_100a cubin -> cudaLibrary).
#ifndef SSB_CSRC_JIT_H_
#define SSB_CSRC_JIT_H_

#include <string>
#include <vector>

#include "common.h"
#include "program.h"

namespace ssb {

struct JitKernel {
  void* kernel;        // cudaKernel_t (accepted by cudaLaunchKernel / cudaFuncSetAttribute as a function pointer)
  int regs;            // registers per thread
  double compile_ms;   // NVRTC + load time of the first use
};

// The aggregation side of a fused Filter -> Compute -> GroupAggregate plan (csrc/jit_rows.h).
struct JitRowsShape {
  int n_keys, n_aggs;
  int fn[16], in_phys[16], out_phys[16], out[16];   // out: program output feeding aggregate a, -1 = COUNT(*)
  int groups;          // CTA-local group entries (1 .. kTinyGroups); 0 = many groups: every row goes to the global table
  int threads, rows_per_thread, min_ctas;
  int prefetch;        // the next step's inputs are loaded before this step's are evaluated
};

// Fills threads / rows_per_thread / min_ctas (defaults or the SSB200_JIT_* environment).
void jit_rows_tune(JitRowsShape* shape);
// CUDA source of ssb_jit_rows for this program and shape ("" and *err when the plan does not fit).
std::string jit_rows_source(const Program& prog, const JitRowsShape& shape, std::string* err);
// Dynamic shared memory of one CTA of that kernel.
size_t jit_rows_smem(const JitRowsShape& shape);

// NVRTC only: needs no device. Returns 0 or an SSB_ERROR_* with the compiler log in *log.
int jit_compile(const std::string& source, std::vector<char>* cubin, std::string* log);
// Compiled and loaded once per distinct source and process; later calls return the cached kernel.
int jit_get_kernel(ssb_ctx* ctx, const std::string& source, const char* name, JitKernel* out);

// Counts a launch of a compiled kernel (ssb_jit_stats).
void jit_note_launch();

}  // namespace ssb
#endif  // SSB_CSRC_JIT_H_
