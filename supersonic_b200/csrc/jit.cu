// jit.cu -- run-time specialisation of the fused Filter -> Compute -> GroupAggregate kernel.
//
// The interpreting kernels (expr_kernel.cu, group.cu) dispatch one bytecode instruction per tile
// row group; on plans that are arithmetic-heavy per byte (the TPC-H Q1 shape: 7 columns in,
// 2 keys + 5 SUMs + COUNT out) they are issue-bound at a fifth of the HBM roofline
// (profiles/r2d_q1_sink_ncu_summary.txt: 12 warp instructions per row, dispatch + operand
// traffic through shared memory). For large inputs the plan is instead compiled: the program is
// written out as three X-macro lists, csrc/jit_rows.h + ops.h + group_device.h (embedded in this
// library at build time) are handed to NVRTC for sm_100a, the cubin is loaded with
// cudaLibraryLoadData and cached per distinct source. ops.h's alu() is the single definition of
// every operator's semantics for both forms. NVRTC is bound with dlopen, like NCCL in comm.cu:
// a box without it keeps the interpreting kernels.
#include "jit.h"

#include <dlfcn.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <mutex>

#include "group_device.h"
#include "jit_rows.h"
#include "jit_embed.inc"

namespace ssb {
namespace {

typedef struct _nvrtcProgram* nvrtcProgram;
typedef int nvrtcResult;

struct Nvrtc {
  void* lib = nullptr;
  nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  nvrtcResult (*DestroyProgram)(nvrtcProgram*) = nullptr;
  nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
  nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
  nvrtcResult (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
  nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
  nvrtcResult (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
  const char* (*GetErrorString)(nvrtcResult) = nullptr;
  std::string error;
};

Nvrtc* nvrtc() {
  static Nvrtc* n = [] {
    Nvrtc* r = new Nvrtc();
    const char* names[] = {getenv("SSB200_NVRTC"), "libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so",
                           "/usr/local/cuda/lib64/libnvrtc.so"};
    for (const char* name : names) {
      if (name == nullptr || *name == 0) continue;
      r->lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
      if (r->lib != nullptr) break;
    }
    if (r->lib == nullptr) { r->error = "libnvrtc.so.12 not found (set SSB200_NVRTC to its path)"; return r; }
#define SSB_NVRTC_SYM(field, sym)                                                             \
  r->field = reinterpret_cast<decltype(r->field)>(dlsym(r->lib, sym));                        \
  if (r->field == nullptr) { r->error = std::string("libnvrtc: missing ") + sym; r->lib = nullptr; return r; }
    SSB_NVRTC_SYM(CreateProgram, "nvrtcCreateProgram")
    SSB_NVRTC_SYM(DestroyProgram, "nvrtcDestroyProgram")
    SSB_NVRTC_SYM(CompileProgram, "nvrtcCompileProgram")
    SSB_NVRTC_SYM(GetCUBINSize, "nvrtcGetCUBINSize")
    SSB_NVRTC_SYM(GetCUBIN, "nvrtcGetCUBIN")
    SSB_NVRTC_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize")
    SSB_NVRTC_SYM(GetProgramLog, "nvrtcGetProgramLog")
    SSB_NVRTC_SYM(GetErrorString, "nvrtcGetErrorString")
#undef SSB_NVRTC_SYM
    return r;
  }();
  return n;
}

struct CacheEntry { JitKernel k; };
std::mutex g_mu;
long long g_compiled = 0, g_launches = 0;
double g_compile_ms = 0;
std::map<std::string, CacheEntry>* cache() {
  static std::map<std::string, CacheEntry>* c = new std::map<std::string, CacheEntry>();
  return c;
}

void appendf(std::string* s, const char* fmt, ...) __attribute__((format(printf, 2, 3)));
void appendf(std::string* s, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  *s += buf;
}

}  // namespace

size_t jit_rows_smem(const JitRowsShape& sh);

// Launch shape of ssb_jit_rows (call after n_aggs / groups are set). From the B200 sweep on the Q1 shape
// (profiles/r2h_q1_jit.txt): 192 threads x 1 row per step with the next step's loads in flight, as many CTAs per SM as
// the per-thread accumulators in shared memory and 84 registers per thread allow (3 for 6 groups x 6 aggregates).
// Two rows per thread with prefetch spill (117 registers wanted) and lose a factor of two.
// SSB200_JIT_THREADS / SSB200_JIT_ROWS / SSB200_JIT_PREFETCH / SSB200_JIT_MIN_CTAS pin them for A/B runs.
void jit_rows_tune(JitRowsShape* sh) {
  auto env = [](const char* name, int dflt) { const char* e = getenv(name); return (e && *e) ? atoi(e) : dflt; };
  sh->threads = env("SSB200_JIT_THREADS", 192);
  sh->rows_per_thread = env("SSB200_JIT_ROWS", 1);
  sh->prefetch = env("SSB200_JIT_PREFETCH", 1);
  long long fit = (220 * 1024) / static_cast<long long>(jit_rows_smem(*sh) + 2560);
  const long long by_regs = 65536 / (static_cast<long long>(sh->threads) * 84);
  if (fit > by_regs) fit = by_regs;
  if (fit < 1) fit = 1;
  if (fit > 8) fit = 8;
  sh->min_ctas = env("SSB200_JIT_MIN_CTAS", static_cast<int>(fit));
}

size_t jit_rows_smem(const JitRowsShape& sh) {
  return static_cast<size_t>(sh.groups) * sh.n_aggs * sh.threads * 8 + static_cast<size_t>(sh.groups) * sh.threads * 4;
}

std::string jit_rows_source(const Program& prog, const JitRowsShape& sh, std::string* err) {
  const int n_in = static_cast<int>(prog.input_types.size());
  const int n_out = static_cast<int>(prog.outputs.size());
  const int n_slot = n_in + prog.params.n_tmp;
  if (n_in > kJitMaxIn || sh.n_aggs < 1 || sh.n_aggs > kLocalMaxAggs || sh.n_keys > kMaxKeys || sh.groups < 0 || sh.groups > kTinyGroups ||
      sh.rows_per_thread < 1 || sh.rows_per_thread > 8 || sh.threads % 32 != 0 || sh.threads < 32 || sh.threads > 1024) {
    if (err) *err = "plan outside the limits of the specialised aggregation kernel";
    return std::string();
  }
  std::string s;
  s += "// generated by csrc/jit.cu (jit_rows_source)\n";
  appendf(&s, "#include \"jit_rt.h\"\nnamespace ssb {\nstruct Spec {\n  enum { T = %d, R = %d, G = %d, MIN_CTAS = %d, PREFETCH = %d, N_IN = %d, N_SLOT = %d, N_OUT = %d, NK = %d, A = %d };\n};\n}\n",
          sh.threads, sh.rows_per_thread, sh.groups, sh.min_ctas, sh.prefetch, n_in, n_slot, n_out, sh.n_keys, sh.n_aggs);
  s += "#define SSB_JIT_INPUTS(X)";
  for (int c = 0; c < n_in; ++c) {
    const int ph = phys_of(prog.input_types[c]);
    if (ph < 0) { if (err) *err = "input type without a physical form"; return std::string(); }
    appendf(&s, " \\\n  X(%d, %d, %d)", c, ph, prog.input_nullable[c] ? 1 : 0);
  }
  s += "\n#define SSB_JIT_PROGRAM(X)";
  for (size_t i = 0; i < prog.generic.size(); ++i) {
    const Insn& in = prog.generic[i];
    if (in.kind == K_END) break;
    const bool slot_a = (in.kind == K_STORE) || ((in.kind == K_LOAD || in.kind == K_ALU2 || in.kind == K_ALU3) && !(in.flags & F_RHS_IMM));
    const bool imm_a = (in.kind == K_LOAD || in.kind == K_ALU2 || in.kind == K_ALU3) && (in.flags & F_RHS_IMM);
    const bool slot_b = in.kind == K_ALU3 && !(in.flags & F_RHS2_IMM);
    const bool imm_b = in.kind == K_ALU3 && (in.flags & F_RHS2_IMM);
    if ((slot_a && (in.a < 0 || in.a >= n_slot)) || (slot_b && (in.b < 0 || in.b >= n_slot)) || (imm_a && (in.a < 0 || in.a >= kMaxImm)) ||
        (imm_b && (in.b < 0 || in.b >= kMaxImm)) || (in.kind == K_OUT && (in.a < 0 || in.a >= n_out))) {
      if (err) *err = "program operand out of range";
      return std::string();
    }
    appendf(&s, " \\\n  X(%d, %d, %d, %d, %d, %d, %d, %d, 0x%llxull, 0x%llxull)", in.kind, in.mop, in.t, in.t2, in.flags, in.rhs_nullable,
            (slot_a || in.kind == K_OUT) ? in.a : 0, slot_b ? in.b : 0,
            imm_a ? static_cast<unsigned long long>(prog.params.imm[in.a]) : 0ull, imm_b ? static_cast<unsigned long long>(prog.params.imm[in.b]) : 0ull);
  }
  s += "\n#define SSB_JIT_AGGS(X)";
  for (int a = 0; a < sh.n_aggs; ++a) {
    const int fn = sh.fn[a], op = sh.out_phys[a];
    const int pad = fn == SSB_AGG_COUNT ? TA_COUNT
                    : (fn == SSB_AGG_SUM && op == T_F64) ? TA_SUM_F64
                    : (fn == SSB_AGG_SUM && (op == T_I64 || op == T_U64)) ? TA_SUM_U64 : TA_OTHER;
    if (sh.out[a] >= n_out) { if (err) *err = "aggregate input out of range"; return std::string(); }
    appendf(&s, " \\\n  X(%d, %d, %d, %d, %d, %d)", a, fn, sh.in_phys[a], op, sh.out[a], pad);
  }
  s += "\n#include \"jit_rows.h\"\n";
  return s;
}

int jit_compile(const std::string& source, std::vector<char>* cubin, std::string* log) {
  Nvrtc* n = nvrtc();
  if (n->lib == nullptr) { if (log) *log = n->error; return SSB_ERROR_NOT_IMPLEMENTED; }
  nvrtcProgram prog = nullptr;
  nvrtcResult r = n->CreateProgram(&prog, source.c_str(), "ssb_jit.cu", kEmbeddedCount, kEmbeddedSources, kEmbeddedNames);
  if (r != 0) { if (log) *log = std::string("nvrtcCreateProgram: ") + n->GetErrorString(r); return SSB_ERROR_UNKNOWN; }
  const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "--device-as-default-execution-space"};
  r = n->CompileProgram(prog, static_cast<int>(sizeof(opts) / sizeof(opts[0])), opts);
  size_t log_size = 0;
  n->GetProgramLogSize(prog, &log_size);
  if (log != nullptr && log_size > 1) { log->resize(log_size); n->GetProgramLog(prog, &(*log)[0]); }
  if (r != 0) {
    if (log) *log = std::string("nvrtcCompileProgram: ") + n->GetErrorString(r) + "\n" + *log;
    n->DestroyProgram(&prog);
    return SSB_ERROR_UNKNOWN;
  }
  size_t size = 0;
  r = n->GetCUBINSize(prog, &size);
  if (r == 0 && size > 0) { cubin->resize(size); r = n->GetCUBIN(prog, cubin->data()); }
  n->DestroyProgram(&prog);
  if (r != 0 || size == 0) { if (log) *log = "nvrtcGetCUBIN failed"; return SSB_ERROR_UNKNOWN; }
  return 0;
}

int jit_get_kernel(ssb_ctx* ctx, const std::string& source, const char* name, JitKernel* out) {
  std::lock_guard<std::mutex> lock(g_mu);
  auto it = cache()->find(source);
  if (it != cache()->end()) { *out = it->second.k; return out->kernel != nullptr ? 0 : SSB_ERROR_NOT_IMPLEMENTED; }
  const auto t0 = std::chrono::steady_clock::now();
  CacheEntry e;
  memset(&e.k, 0, sizeof(e.k));
  std::vector<char> cubin;
  std::string log;
  int rc = jit_compile(source, &cubin, &log);
  if (rc == 0) {
    cudaLibrary_t lib = nullptr;
    cudaKernel_t kernel = nullptr;
    cudaError_t ce = cudaLibraryLoadData(&lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
    if (ce == cudaSuccess) ce = cudaLibraryGetKernel(&kernel, lib, name);
    cudaFuncAttributes attr;
    memset(&attr, 0, sizeof(attr));
    if (ce == cudaSuccess) ce = cudaFuncGetAttributes(&attr, reinterpret_cast<const void*>(kernel));
    if (ce != cudaSuccess) { log = std::string("loading the compiled kernel: ") + cudaGetErrorString(ce); cudaGetLastError(); rc = SSB_ERROR_UNKNOWN; }
    else { e.k.kernel = reinterpret_cast<void*>(kernel); e.k.regs = attr.numRegs; }
  }
  e.k.compile_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  if (getenv("SSB200_DEBUG_PLAN") != nullptr) {
    fprintf(stderr, "[ssb200] jit: %s in %.0f ms, %d registers per thread%s%s\n", rc == 0 ? "compiled" : "FAILED", e.k.compile_ms, e.k.regs,
            log.size() > 1 ? "\n" : "", log.size() > 1 ? log.c_str() : "");
  }
  if (rc == 0) { ++g_compiled; g_compile_ms += e.k.compile_ms; }
  (*cache())[source] = e;   // failures are cached too: the caller falls back to the interpreting kernels once, not per call
  *out = e.k;
  if (rc != 0) { ctx->last_error = "jit: " + log; return rc; }
  return 0;
}

void jit_note_launch() {
  std::lock_guard<std::mutex> lock(g_mu);
  ++g_launches;
}

}  // namespace ssb

extern "C" {

void ssb_jit_stats(int64_t* kernels_compiled, double* compile_ms, int64_t* launches) {
  std::lock_guard<std::mutex> lock(ssb::g_mu);
  if (kernels_compiled) *kernels_compiled = ssb::g_compiled;
  if (compile_ms) *compile_ms = ssb::g_compile_ms;
  if (launches) *launches = ssb::g_launches;
}

// Test / tooling entry (no device needed): compiles the specialised aggregation kernel of a plan to an sm_100a cubin.
int ssb_jit_rows_compile(const ssb_expr_node* nodes, int32_t n_nodes, int32_t n_inputs, const int32_t* input_types,
                         const int32_t* input_nullable, const int32_t* outputs, int32_t n_outputs, int32_t predicate,
                         int32_t n_keys, int32_t n_aggs, const ssb_agg_spec* aggs, int32_t groups, int32_t threads,
                         int32_t rows_per_thread, char* text, int64_t text_cap, int64_t* cubin_bytes) {
  using namespace ssb;
  if (cubin_bytes) *cubin_bytes = 0;
  auto put = [&](const std::string& m) { if (text && text_cap > 0) { snprintf(text, static_cast<size_t>(text_cap), "%s", m.c_str()); } };
  Program prog;
  std::string err;
  int rc = compile_program(nodes, n_nodes, n_inputs, input_types, input_nullable, outputs, n_outputs, predicate, 768, 100 * 1024, 227 * 1024, &prog, &err);
  if (rc != 0) { put(err); return rc; }
  JitRowsShape sh;
  memset(&sh, 0, sizeof(sh));
  sh.n_keys = n_keys; sh.n_aggs = n_aggs; sh.groups = groups;
  jit_rows_tune(&sh);
  if (threads > 0) sh.threads = threads;
  if (rows_per_thread > 0) sh.rows_per_thread = rows_per_thread;
  for (int a = 0; a < n_aggs && a < 16; ++a) {
    sh.fn[a] = aggs[a].fn;
    sh.in_phys[a] = aggs[a].input < 0 ? -1 : phys_of(aggs[a].in_type);
    sh.out_phys[a] = phys_of(aggs[a].out_type);
    sh.out[a] = aggs[a].input < 0 ? -1 : n_keys + aggs[a].input;
  }
  const std::string src = jit_rows_source(prog, sh, &err);
  if (src.empty()) { put(err); return SSB_ERROR_NOT_IMPLEMENTED; }
  std::vector<char> cubin;
  std::string log;
  rc = jit_compile(src, &cubin, &log);
  if (rc != 0) { put(log + "\n---- source ----\n" + src); return rc; }
  if (cubin_bytes) *cubin_bytes = static_cast<int64_t>(cubin.size());
  if (const char* dump = getenv("SSB200_JIT_DUMP")) {
    if (FILE* f = fopen(dump, "wb")) { fwrite(cubin.data(), 1, cubin.size(), f); fclose(f); }
  }
  put(src);
  return 0;
}

}  // extern "C"
