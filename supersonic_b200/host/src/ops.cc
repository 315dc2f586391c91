// ops.cc -- Operation factories and GPU cursors (see include/supersonic/cursor.h).
#include <stdio.h>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <map>
#include <set>
#include <mutex>
#include <thread>

#include "internal.h"
#include "narrow.h"

namespace supersonic {

using namespace internal;   // NOLINT

// ------------------------------------------------------------------ BasicOperation
BasicOperation::~BasicOperation() {
  for (size_t i = 0; i < children_.size(); ++i) delete children_[i];
}
void BasicOperation::SetBufferAllocator(BufferAllocator* allocator, bool cascade_to_children) {
  allocator_ = allocator;
  if (cascade_to_children) {
    for (size_t i = 0; i < children_.size(); ++i) children_[i]->SetBufferAllocator(allocator, true);
  }
}
void BasicOperation::SetBufferAllocatorWhereUnset(BufferAllocator* allocator, bool cascade_to_children) {
  if (allocator_ == NULL) allocator_ = allocator;
  if (cascade_to_children) {
    for (size_t i = 0; i < children_.size(); ++i) children_[i]->SetBufferAllocatorWhereUnset(allocator, true);
  }
}
void BasicOperation::AppendDebugDescription(string* target) const {
  target->append(DebugName());
  target->append("(");
  for (size_t i = 0; i < children_.size(); ++i) {
    if (i) target->append(", ");
    children_[i]->AppendDebugDescription(target);
  }
  target->append(")");
}

namespace {

#define SSB_CALL(session, call, what)                              \
  do {                                                             \
    const int rc_ = (call);                                        \
    if (rc_ != 0) THROW((session)->Error(rc_, what));              \
  } while (0)

// ------------------------------------------------------------------ host view cursor
// cursor/infrastructure/view_cursor.cc:47-75: slices an in-memory view, zero copy.
class ViewCursor : public Cursor {
 public:
  explicit ViewCursor(const View& view) : full_(view), out_(view.schema()), offset_(0), interrupted_(false) {}
  virtual const TupleSchema& schema() const { return full_.schema(); }
  virtual ResultView Next(rowcount_t max_row_count) {
    if (interrupted_) return ResultView::Failure(new Exception(INTERRUPTED, "cursor interrupted"));
    if (offset_ >= full_.row_count()) return ResultView::EOS();
    rowcount_t n = full_.row_count() - offset_;
    if (n > max_row_count) n = max_row_count;
    out_.ResetFromSubRange(full_, offset_, n);
    offset_ += n;
    return ResultView::Success(&out_);
  }
  virtual void Interrupt() { interrupted_ = true; }
  virtual void AppendDebugDescription(string* target) const { target->append("ViewCursor"); }
  virtual CursorId GetCursorId() const { return VIEW; }
  const View& full_view() const { return full_; }
 private:
  View full_, out_;
  rowcount_t offset_;
  volatile bool interrupted_;
};

// ------------------------------------------------------------------ transfer narrowing
// Streaming a host table is bound by PCIe, not by the GPU (51 GB/s against 5 TB/s of kernel
// throughput). 64-bit integer columns whose values all fit 32 bits in a chunk therefore cross the
// bus as 32-bit values: host threads narrow the chunk into pinned staging (checking every value),
// the kernel program of that chunk widens them again with a CAST at the INPUT node. Lossless by
// construction -- a chunk with one value that does not fit is sent as it is.
class HostPool {
 public:
  static HostPool& Get() { static HostPool pool; return pool; }
  int size() const { return static_cast<int>(threads_.size()) + 1; }
  // Runs fn(part, parts) for part in [0, parts) on the pool (the caller works too); returns when all are done.
  void Run(int parts, const std::function<void(int, int)>& fn) {
    if (parts <= 1 || threads_.empty()) { for (int i = 0; i < parts; ++i) fn(i, parts); return; }
    std::lock_guard<std::mutex> one_caller(run_mu_);
    {
      std::unique_lock<std::mutex> lock(mu_);
      fn_ = &fn; parts_ = parts; next_ = 0; pending_ = parts; ++generation_;
    }
    cv_.notify_all();
    Work();
    std::unique_lock<std::mutex> lock(mu_);
    done_.wait(lock, [this] { return pending_ == 0; });
    fn_ = NULL;
  }
 private:
  HostPool() : fn_(NULL), parts_(0), next_(0), pending_(0), generation_(0), stop_(false) {
    int n = static_cast<int>(std::thread::hardware_concurrency());
    // one process per GPU (torchrun): the ranks of a box share its cores
    if (const char* env = getenv("LOCAL_WORLD_SIZE")) { const int ranks = atoi(env); if (ranks > 1) n /= ranks; }
    if (const char* env = getenv("SSB200_HOST_THREADS")) n = atoi(env);
    n = std::max(1, std::min(n, 128));
    for (int i = 1; i < n; ++i) threads_.push_back(std::thread([this] { Loop(); }));
  }
  ~HostPool() {
    { std::unique_lock<std::mutex> lock(mu_); stop_ = true; }
    cv_.notify_all();
    for (size_t i = 0; i < threads_.size(); ++i) threads_[i].join();
  }
  void Loop() {
    unsigned long long seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lock(mu_);
        cv_.wait(lock, [&] { return stop_ || generation_ != seen; });
        if (stop_) return;
        seen = generation_;
      }
      Work();
    }
  }
  void Work() {
    for (;;) {
      int part;
      const std::function<void(int, int)>* fn;
      int parts;
      {
        std::unique_lock<std::mutex> lock(mu_);
        if (fn_ == NULL || next_ >= parts_) return;
        part = next_++; fn = fn_; parts = parts_;
      }
      (*fn)(part, parts);
      std::unique_lock<std::mutex> lock(mu_);
      if (--pending_ == 0) done_.notify_all();
    }
  }
  std::vector<std::thread> threads_;
  std::mutex mu_, run_mu_;
  std::condition_variable cv_, done_;
  const std::function<void(int, int)>* fn_;
  int parts_, next_, pending_;
  unsigned long long generation_;
  bool stop_;
};

// Set once a cursor found that this host narrows slower than the bus moves the unnarrowed chunk.
std::atomic<bool>& NarrowingUnprofitable() { static std::atomic<bool> flag(false); return flag; }

// dst[i] = int32(src[i]) for i < rows; false when some value does not fit (dst is then garbage).
bool NarrowInt64Column(const int64* src, int32* dst, rowcount_t rows) {
  HostPool& pool = HostPool::Get();
  const int parts = static_cast<int>(std::min<rowcount_t>(static_cast<rowcount_t>(pool.size()), rows / 65536 + 1));
  std::atomic<int> misfit(0);
  static const bool streaming_stores = getenv("SSB200_NARROW_NT") != NULL && atoi(getenv("SSB200_NARROW_NT")) != 0;
  pool.Run(parts, [&](int part, int n_parts) {
    const rowcount_t begin = rows * part / n_parts, end = rows * (part + 1) / n_parts;
    int64 bad = 0;
#if defined(__x86_64__)
    if (streaming_stores) {
      // the staging buffer is only read by the DMA engine: the stores may skip the cache
      for (rowcount_t i = begin; i < end; ++i) {
        const int64 v = src[i];
        const int32 n = static_cast<int32>(v);
        _mm_stream_si32(dst + i, n);
        bad |= v ^ static_cast<int64>(n);
      }
      _mm_sfence();
    } else
#endif
    {
      bad = narrow::Range(reinterpret_cast<const int64_t*>(src), reinterpret_cast<int32_t*>(dst),
                          static_cast<size_t>(begin), static_cast<size_t>(end));
    }
    if (bad != 0) misfit.store(1, std::memory_order_relaxed);
  });
  return misfit.load(std::memory_order_relaxed) == 0;
}

// The lowered program with the INPUT nodes in `narrow_mask` (bit k = k-th used input) declared INT32
// and widened back to their type by a CAST.
void NarrowedProgram(const vector<ssb_expr_node>& nodes, const vector<int32_t>& outs, int pred, uint32_t narrow_mask,
                     vector<ssb_expr_node>* new_nodes, vector<int32_t>* new_outs, int* new_pred) {
  vector<int> index(nodes.size(), -1);
  new_nodes->clear();
  for (size_t i = 0; i < nodes.size(); ++i) {
    ssb_expr_node nd = nodes[i];
    if (nd.op == SSB_OP_INPUT) {
      if ((narrow_mask >> nd.arg[0]) & 1u) {
        ssb_expr_node in = nd;
        in.out_type = SSB_INT32;
        new_nodes->push_back(in);
        ssb_expr_node cast;
        memset(&cast, 0, sizeof(cast));
        cast.op = SSB_OP_CAST;
        cast.out_type = nd.out_type;
        cast.arg[0] = static_cast<int32_t>(new_nodes->size()) - 1;
        cast.arg[1] = cast.arg[2] = -1;
        new_nodes->push_back(cast);
      } else {
        new_nodes->push_back(nd);
      }
    } else {
      if (nd.op != SSB_OP_CONST) {
        for (int a = 0; a < 3; ++a) if (nd.arg[a] >= 0) nd.arg[a] = index[nd.arg[a]];
      }
      new_nodes->push_back(nd);
    }
    index[i] = static_cast<int>(new_nodes->size()) - 1;
  }
  new_outs->clear();
  for (size_t j = 0; j < outs.size(); ++j) new_outs->push_back(index[outs[j]]);
  *new_pred = pred >= 0 ? index[pred] : -1;
}

// ------------------------------------------------------------------ row-wise cursor
// One fused kernel for a chain of ScanView / Compute / Filter / Project
// (replaces ComputeCursor::Next, FilterCursor::Next, ProjectCursor::Next).
// Hands a shared cursor to code that takes ownership of a Cursor* (CursorTransformer::Transform).
class SharedCursor : public Cursor {
 public:
  explicit SharedCursor(const std::shared_ptr<Cursor>& c) : c_(c) {}
  virtual const TupleSchema& schema() const { return c_->schema(); }
  virtual ResultView Next(rowcount_t max_row_count) { return c_->Next(max_row_count); }
  virtual void Interrupt() { c_->Interrupt(); }
  virtual bool IsWaitingOnBarrierSupported() const { return c_->IsWaitingOnBarrierSupported(); }
  virtual void ApplyToChildren(CursorTransformer* transformer) { c_->ApplyToChildren(transformer); }
  virtual void AppendDebugDescription(string* target) const { c_->AppendDebugDescription(target); }
  virtual CursorId GetCursorId() const { return c_->GetCursorId(); }
 private:
  std::shared_ptr<Cursor> c_;
};

class RowwiseCursor : public GpuCursor {
 public:
  RowwiseCursor(const RowwisePlan& plan, BufferAllocator* allocator, CursorId id)
      : GpuCursor(plan.schema, allocator, "RowwiseCursor"), plan_(plan), id_(id), mode_(UNDECIDED),
        out_view_(plan.schema), next_first_(0), active_(-1), served_(0) {
    lanes_[0].Reset();
    lanes_[1].Reset();
  }
  virtual ~RowwiseCursor() {
    for (int i = 0; i < 2; ++i) lanes_[i].Free();
  }
  virtual CursorId GetCursorId() const { return id_; }
  // The row-wise operators below this cursor are fused into it: its only child cursor is the non-row-wise source
  // of the chain, if there is one (a scanned view has none).
  virtual void ApplyToChildren(CursorTransformer* transformer) {
    if (plan_.source) plan_.source.reset(transformer->Transform(new SharedCursor(plan_.source)));
  }

  // Host-resident inputs are streamed: the view is cut into chunks that alternate between two
  // contexts (streams), so that the H2D copy of chunk i+1, the kernel of chunk i and the D2H
  // copy of chunk i-1 overlap. Device-resident inputs and parents that are GPU cursors use
  // Produce() instead (whole shard, no host round trip).
  virtual ResultView Next(rowcount_t max_row_count) {
    if (mode_ == UNDECIDED) {
      FailureOrVoid d = DecideMode();
      if (d.is_failure()) return ResultView::Failure(d.release_exception());
    }
    if (mode_ == WHOLE) return GpuCursor::Next(max_row_count);
    if (interrupted()) return ResultView::Failure(new Exception(INTERRUPTED, "cursor interrupted"));
    for (;;) {
      if (active_ >= 0 && served_ < static_cast<rowcount_t>(lanes_[active_].kept)) {
        Lane& l = lanes_[active_];
        rowcount_t n = static_cast<rowcount_t>(l.kept) - served_;
        if (n > max_row_count) n = max_row_count;
        for (int j = 0; j < out_view_.column_count(); ++j) {
          const size_t w = out_view_.column(j).type_info().size();
          const bool nullable = out_view_.schema().attribute(j).is_nullable();
          out_view_.mutable_column(j)->Reset(static_cast<char*>(l.h_out[j]) + served_ * w,
                                             nullable ? static_cast<bool*>(l.h_out_nulls[j]) + served_ : NULL);
        }
        out_view_.set_row_count(n);
        served_ += n;
        return ResultView::Success(&out_view_);
      }
      // the active lane is drained: give it the next chunk and switch to the other lane
      if (active_ >= 0) {
        lanes_[active_].busy = false;
        FailureOrVoid l = LaunchNext(active_);
        if (l.is_failure()) return ResultView::Failure(l.release_exception());
        active_ ^= 1;
      } else {
        FailureOrVoid l0 = LaunchNext(0);
        if (l0.is_failure()) return ResultView::Failure(l0.release_exception());
        FailureOrVoid l1 = LaunchNext(1);
        if (l1.is_failure()) return ResultView::Failure(l1.release_exception());
        active_ = 0;
      }
      Lane& l = lanes_[active_];
      if (!l.busy) return ResultView::EOS();
      const int rc = ssb_ctx_sync(l.ctx);
      if (rc != 0) return ResultView::Failure(Session::ErrorOn(l.ctx, rc, "row-wise pipeline"));
      const int frc = ssb_program_check_failure(l.launched ? l.launched : l.program);
      if (frc != 0) return ResultView::Failure(Session::ErrorOn(l.ctx, frc, "expression evaluation"));
      l.kept = *l.h_count;
      if (!l.fetched) {
        FailureOrVoid q = QueueOutputs(&l, static_cast<rowcount_t>(l.kept));
        if (q.is_failure()) return ResultView::Failure(q.release_exception());
        const int rc2 = ssb_ctx_sync(l.ctx);
        if (rc2 != 0) return ResultView::Failure(Session::ErrorOn(l.ctx, rc2, "row-wise pipeline (download)"));
        l.fetched = true;
      }
      served_ = 0;
    }
  }

 protected:
  // ---- variable-length columns ------------------------------------------------------------------
  static void CollectNodes(const NodePtr& n, std::set<const ExprNode*>* seen, vector<const ExprNode*>* order) {
    if (!n || !seen->insert(n.get()).second) return;
    for (size_t i = 0; i < n->args.size(); ++i) CollectNodes(n->args[i], seen, order);
    order->push_back(n.get());
  }
  void PlanNodes(vector<const ExprNode*>* order) const {
    std::set<const ExprNode*> seen;
    for (size_t j = 0; j < plan_.outputs.size(); ++j) CollectNodes(plan_.outputs[j], &seen, order);
    CollectNodes(plan_.predicate, &seen, order);
  }
  bool PlanHasStrings() const {
    vector<const ExprNode*> nodes;
    PlanNodes(&nodes);
    for (size_t i = 0; i < nodes.size(); ++i) if (IsVariableLength(nodes[i]->type)) return true;
    return false;
  }
  // Copies the DAG under `n`, replacing the nodes in `replace`.
  static NodePtr Rewrite(const NodePtr& n, const std::map<const ExprNode*, NodePtr>& replace, std::map<const ExprNode*, NodePtr>* done) {
    if (!n) return n;
    std::map<const ExprNode*, NodePtr>::const_iterator r = replace.find(n.get());
    if (r != replace.end()) return r->second;
    std::map<const ExprNode*, NodePtr>::iterator d = done->find(n.get());
    if (d != done->end()) return d->second;
    bool changed = false;
    vector<NodePtr> args;
    for (size_t i = 0; i < n->args.size(); ++i) {
      args.push_back(Rewrite(n->args[i], replace, done));
      changed = changed || args.back().get() != n->args[i].get();
    }
    NodePtr out = n;
    if (changed) {
      std::shared_ptr<ExprNode> c(new ExprNode(*n));
      c->args = args;
      out = c;
    }
    (*done)[n.get()] = out;
    return out;
  }

  // A plan that touches STRING / BINARY values. The kernels see codes: value-producing nodes of variable-length type
  // may only be input columns, literals and NULL (anything else -- CONCAT, SUBSTRING, IF over strings ... -- is not on
  // this path); the columns and literals that meet in a comparison are re-encoded against one merged dictionary, the
  // literals become INT64 constants (their codes), and a pass-through output column inherits its input's dictionary.
  FailureOrVoid RunWithStrings(DeviceTable* result) {
    FailureOr<Session*> sr = Session::Get();
    PROPAGATE_ON_FAILURE(sr);
    Session* s = sr.get();
    vector<const ExprNode*> nodes;
    PlanNodes(&nodes);
    std::map<int, int> needed;
    vector<const ExprNode*> literals;
    std::set<int> compared_cols;
    bool compares_literal = false;
    for (size_t i = 0; i < nodes.size(); ++i) {
      const ExprNode* n = nodes[i];
      if (n->op == SSB_OP_INPUT) needed[n->input] = 1;
      if (IsVariableLength(n->type)) {
        if (n->op == SSB_OP_CONST) { if (!(n->flags & SSB_NODE_NULL)) literals.push_back(n); }
        else if (n->op != SSB_OP_INPUT) {
          THROW(new Exception(ERROR_NOT_IMPLEMENTED, "expressions that compute STRING / BINARY values are not on the B200 hot path "
                                                     "(columns, literals and comparisons are): " + n->name));
        }
        continue;
      }
      bool var_args = false;
      for (size_t a = 0; a < n->args.size(); ++a) var_args = var_args || IsVariableLength(n->args[a]->type);
      if (!var_args) continue;
      switch (n->op) {
        case SSB_OP_EQ: case SSB_OP_NE: case SSB_OP_LT: case SSB_OP_LE:
          for (size_t a = 0; a < n->args.size(); ++a) {
            if (n->args[a]->op == SSB_OP_INPUT) compared_cols.insert(n->args[a]->input);
            else if (!(n->args[a]->flags & SSB_NODE_NULL)) compares_literal = true;
          }
          break;
        case SSB_OP_IS_NULL: break;
        default:
          THROW(new Exception(ERROR_NOT_IMPLEMENTED, "this operator over STRING / BINARY arguments is not on the B200 hot path: " + n->name));
      }
    }
    // the base columns
    DeviceTable base;
    std::unique_ptr<Block> keepalive;
    int64 rows = 0;
    std::map<int, DeviceColumnRef> col_of;
    PROPAGATE_ON_FAILURE(FetchBase(needed, &base, &keepalive, &rows, &col_of));
    // one dictionary for everything that is compared
    std::map<const ExprNode*, NodePtr> replace;
    if (!compared_cols.empty() || compares_literal) {
      vector<DeviceColumnRef*> cols;
      vector<int64> col_rows;
      for (std::set<int>::iterator it = compared_cols.begin(); it != compared_cols.end(); ++it) { cols.push_back(&col_of[*it]); col_rows.push_back(rows); }
      vector<string> texts;
      for (size_t i = 0; i < literals.size(); ++i) texts.push_back(literals[i]->text);
      vector<int64> codes;
      PROPAGATE_ON_FAILURE(UnifyDictionaries(s, cols, col_rows, texts, &codes));
      for (size_t i = 0; i < literals.size(); ++i) {
        std::shared_ptr<ExprNode> c(new ExprNode(*literals[i]));
        c->type = INT64;
        c->imm.u64 = 0;
        c->imm.i64 = i < codes.size() ? codes[i] : 0;
        replace[literals[i]] = c;
      }
    } else {
      for (size_t i = 0; i < literals.size(); ++i) {   // a literal that is only passed through: not on this path
        THROW(new Exception(ERROR_NOT_IMPLEMENTED, "STRING / BINARY literals as result columns are not on the B200 hot path"));
      }
    }
    std::map<const ExprNode*, NodePtr> done;
    vector<NodePtr> outputs;
    for (size_t j = 0; j < plan_.outputs.size(); ++j) outputs.push_back(Rewrite(plan_.outputs[j], replace, &done));
    const NodePtr predicate = Rewrite(plan_.predicate, replace, &done);
    const size_t n_out = outputs.size();
    for (size_t group = n_out > 0 ? n_out : 1;; group = (group + 1) / 2) {
      vector<std::unique_ptr<DeviceProgram> > programs;
      Exception* error = NULL;
      for (size_t first = 0; first < (n_out > 0 ? n_out : 1) && error == NULL; first += group) {
        vector<NodePtr> outs;
        for (size_t j = first; j < n_out && j < first + group; ++j) outs.push_back(outputs[j]);
        FailureOrOwned<DeviceProgram> created = DeviceProgram::Create(plan_.base_schema, outs, predicate);
        if (created.is_failure()) error = created.release_exception();
        else programs.push_back(std::unique_ptr<DeviceProgram>(created.release()));
      }
      if (error == NULL) {
        PROPAGATE_ON_FAILURE(RunOver(programs, group, rows, col_of, result));
        break;
      }
      if (error->return_code() != ERROR_NOT_IMPLEMENTED || group <= 1) return Failure(error);
      delete error;
    }
    for (size_t j = 0; j < plan_.outputs.size(); ++j) {
      if (!IsVariableLength(plan_.outputs[j]->type)) continue;
      if (plan_.outputs[j]->op == SSB_OP_INPUT) {
        result->columns[j].dict = col_of[plan_.outputs[j]->input].dict;
      } else {   // the NULL literal: every row is NULL, any dictionary will do
        result->columns[j].dict.reset(new DeviceDict);
      }
    }
    return Success();
  }

  virtual FailureOrVoid Run(DeviceTable* result) {
    if (PlanHasStrings()) return RunWithStrings(result);
    // One kernel evaluates the predicate and every output column. Plans too wide for one
    // CTA's shared memory are split into column groups that share the predicate.
    const size_t n_out = plan_.outputs.size();
    for (size_t group = n_out > 0 ? n_out : 1;; group = (group + 1) / 2) {
      vector<std::unique_ptr<DeviceProgram> > programs;
      Exception* error = NULL;
      for (size_t first = 0; first < (n_out > 0 ? n_out : 1) && error == NULL; first += group) {
        vector<NodePtr> outs;
        for (size_t j = first; j < n_out && j < first + group; ++j) outs.push_back(plan_.outputs[j]);
        FailureOrOwned<DeviceProgram> created = DeviceProgram::Create(plan_.base_schema, outs, plan_.predicate);
        if (created.is_failure()) error = created.release_exception();
        else programs.push_back(std::unique_ptr<DeviceProgram>(created.release()));
      }
      if (error == NULL) return RunPrograms(programs, group, result);
      if (error->return_code() != ERROR_NOT_IMPLEMENTED || group <= 1) return Failure(error);
      delete error;
    }
  }

  // The base columns `needed` (schema positions), uploaded or referenced once.
  FailureOrVoid FetchBase(const std::map<int, int>& needed, DeviceTable* base, std::unique_ptr<Block>* keepalive, int64* rows,
                          std::map<int, DeviceColumnRef>* col_of) {
    if (plan_.source) {
      PROPAGATE_ON_FAILURE(MaterializeOnDevice(plan_.source.get(), base, keepalive));
      *rows = base->rows;
      for (std::map<int, int>::const_iterator it = needed.begin(); it != needed.end(); ++it) (*col_of)[it->first] = base->columns[it->first];
    } else {
      *rows = static_cast<int64>(plan_.base.row_count());
      vector<int> cols;
      for (std::map<int, int>::const_iterator it = needed.begin(); it != needed.end(); ++it) cols.push_back(it->first);
      PROPAGATE_ON_FAILURE(UploadColumns(plan_.base, cols, 0, plan_.base.row_count(), base));
      for (size_t k = 0; k < cols.size(); ++k) (*col_of)[cols[k]] = base->columns[k];
    }
    return Success();
  }

  FailureOrVoid RunPrograms(const vector<std::unique_ptr<DeviceProgram> >& programs, size_t group,
                            DeviceTable* result) {
    std::map<int, int> needed;
    for (size_t g = 0; g < programs.size(); ++g) {
      for (size_t k = 0; k < programs[g]->used_inputs().size(); ++k) needed[programs[g]->used_inputs()[k]] = 1;
    }
    DeviceTable base;
    std::unique_ptr<Block> keepalive;
    int64 rows = 0;
    std::map<int, DeviceColumnRef> col_of;
    PROPAGATE_ON_FAILURE(FetchBase(needed, &base, &keepalive, &rows, &col_of));
    return RunOver(programs, group, rows, col_of, result);
  }

  FailureOrVoid RunOver(const vector<std::unique_ptr<DeviceProgram> >& programs, size_t group, int64 rows,
                        std::map<int, DeviceColumnRef>& col_of, DeviceTable* result) {
    FailureOr<Session*> s = Session::Get();
    PROPAGATE_ON_FAILURE(s);
    PROPAGATE_ON_FAILURE(result->Allocate(plan_.schema, rows, /* force_nulls = */ true));
    for (size_t j = 0; j < result->columns.size(); ++j) {   // columns the program proves NOT NULL are never written
      SSB_CALL(s.get(), ssb_memset(s.get()->ctx(), result->columns[j].col.nulls, 0,
                                   static_cast<size_t>((rows + 31) / 32 + 1) * 4), "memset");
    }
    int64 kept = 0;
    for (size_t g = 0; g < programs.size(); ++g) {
      vector<ssb_column> ic, oc;
      for (size_t k = 0; k < programs[g]->used_inputs().size(); ++k) {
        const int idx = programs[g]->used_inputs()[k];
        ssb_column c = col_of[idx].col;
        // a GPU child hands every column over with a bitmap; a NOT_NULLABLE attribute never sets a bit
        if (!plan_.base_schema.attribute(idx).is_nullable()) c.nulls = NULL;
        ic.push_back(c);
      }
      for (size_t j = g * group; j < result->columns.size() && j < (g + 1) * group; ++j) oc.push_back(result->columns[j].col);
      FailureOr<int64> n = programs[g]->Run(ic, rows, oc);
      PROPAGATE_ON_FAILURE(n);
      kept = n.get();
    }
    result->rows = kept;
    return Success();
  }
 private:
  enum Mode { UNDECIDED, WHOLE, STREAM };
  struct Lane {
    ssb_ctx* ctx;
    ssb_program* program;
    std::map<uint32_t, ssb_program*> narrowed;   // program variants by narrow mask (transfer narrowing)
    ssb_program* launched;                       // the variant the lane's current chunk ran
    vector<void*> h_narrow;                      // pinned staging of the narrowed chunk per used input (or NULL)
    vector<void*> d_in, d_in_nulls, d_out, d_out_nulls, h_out, h_out_nulls;
    vector<std::pair<int, std::pair<void*, size_t> > > owned;   // (kind, (ptr, granted)) from the MemoryPool
    void* d_bools;
    int64_t* d_count;
    int64_t* h_count;
    int64 kept;
    bool busy;
    bool fetched;   // the chunk's output rows are (being) copied to the host
    void Reset() { ctx = NULL; program = NULL; launched = NULL; d_bools = NULL; d_count = NULL; h_count = NULL; kept = 0; busy = false; fetched = false; }
    void Free() {
      if (ctx == NULL) return;
      ssb_ctx_sync(ctx);
      if (program) ssb_program_destroy(program);
      for (std::map<uint32_t, ssb_program*>::iterator it = narrowed.begin(); it != narrowed.end(); ++it) ssb_program_destroy(it->second);
      narrowed.clear();
      h_narrow.clear();
      for (size_t i = 0; i < owned.size(); ++i) {
        MemoryPool::Release(static_cast<MemoryPool::Kind>(owned[i].first), owned[i].second.first, owned[i].second.second);
      }
      owned.clear();
      d_in.clear(); d_in_nulls.clear(); d_out.clear(); d_out_nulls.clear(); h_out.clear(); h_out_nulls.clear();
      Reset();
    }
  };

#define LANE_CALL(lane, call, what)                                              \
  do {                                                                           \
    const int rc_ = (call);                                                      \
    if (rc_ != 0) THROW(Session::ErrorOn((lane).ctx, rc_, what));                \
  } while (0)

  static FailureOr<void*> Pooled(Lane* l, MemoryPool::Kind kind, size_t bytes) {
    size_t granted = 0;
    FailureOr<void*> p = MemoryPool::Acquire(kind, bytes, &granted);
    PROPAGATE_ON_FAILURE(p);
    l->owned.push_back(std::make_pair(static_cast<int>(kind), std::make_pair(p.get(), granted)));
    return p;
  }
#define POOLED(var, lane, kind, bytes)                                \
  void* var = NULL;                                                   \
  {                                                                   \
    FailureOr<void*> r_ = Pooled(&(lane), MemoryPool::kind, (bytes)); \
    PROPAGATE_ON_FAILURE(r_);                                         \
    var = r_.get();                                                   \
  }

  FailureOrVoid DecideMode() {
    mode_ = WHOLE;
    if (plan_.source) return Success();
    if (PlanHasStrings()) return Success();   // variable-length columns: one dictionary per column, whole-table path
    const rowcount_t rows = plan_.base.row_count();
    // rows per chunk: large tables stream in 16M-row chunks (the per-chunk synchronisation costs less: 1.97 -> 2.15 G rows/s
    // at 256M rows against 4M-row chunks, tools/e2e_chunks.sh), small ones keep at least ~32 chunks in the pipeline
    rowcount_t chunk = rows / 32;
    if (chunk < (4u << 20)) chunk = 4u << 20;
    if (chunk > (16u << 20)) chunk = 16u << 20;
    if (const char* env = getenv("SSB200_CHUNK_ROWS")) chunk = static_cast<rowcount_t>(atoll(env));
    chunk = (chunk / 1024) * 1024;
    if (chunk < 1024 || rows <= chunk) return Success();
    vector<int32_t>& outs = outs_;
    int& pred = pred_;
    LowerProgram(plan_.outputs, plan_.predicate, &nodes_, &used_, &outs, &pred);
    for (size_t k = 0; k < used_.size(); ++k) {
      if (IsDevicePointer(plan_.base.column(used_[k]).data().raw())) return Success();
    }
    narrowing_ = (getenv("SSB200_NARROW_TRANSFERS") == NULL || atoi(getenv("SSB200_NARROW_TRANSFERS")) != 0) &&
                 !NarrowingUnprofitable().load(std::memory_order_relaxed);
    narrow_skip_.assign(used_.size(), 0);
    FailureOr<Session*> sr = Session::Get();
    PROPAGATE_ON_FAILURE(sr);
    vector<int32_t> types, nullable;
    for (size_t k = 0; k < used_.size(); ++k) {
      types.push_back(plan_.base_schema.attribute(used_[k]).type());
      nullable.push_back(plan_.base_schema.attribute(used_[k]).is_nullable() ? 1 : 0);
    }
    int32_t dummy = 0;
    for (int i = 0; i < 2; ++i) {
      Lane& l = lanes_[i];
      FailureOr<ssb_ctx*> c = sr.get()->lane(i);
      PROPAGATE_ON_FAILURE(c);
      l.ctx = c.get();
      const int rc = ssb_program_create(l.ctx, nodes_.data(), static_cast<int32_t>(nodes_.size()),
                                        static_cast<int32_t>(used_.size()), types.empty() ? &dummy : types.data(),
                                        nullable.empty() ? &dummy : nullable.data(), outs.empty() ? &dummy : outs.data(),
                                        static_cast<int32_t>(outs.size()), pred, &l.program);
      if (rc == SSB_ERROR_NOT_IMPLEMENTED) { l.program = NULL; return Success(); }   // too wide: whole-shard path splits it
      LANE_CALL(l, rc, "expression compilation");
      for (size_t k = 0; k < used_.size(); ++k) {
        POOLED(d, l, DEVICE, chunk * plan_.base.column(used_[k]).type_info().size() + 256);
        POOLED(dn, l, DEVICE, chunk / 8 + 256);
        l.d_in.push_back(d);
        l.d_in_nulls.push_back(dn);
        const DataType t = plan_.base_schema.attribute(used_[k]).type();
        void* staging = NULL;
        if (narrowing_ && k < 32 && (t == INT64 || t == DATETIME)) {
          POOLED(hn, l, PINNED, chunk * 4 + 256);
          staging = hn;
        }
        l.h_narrow.push_back(staging);
      }
      for (int j = 0; j < plan_.schema.attribute_count(); ++j) {
        const size_t w = GetTypeInfo(plan_.schema.attribute(j).type()).size();
        POOLED(d, l, DEVICE, chunk * w + 256);
        POOLED(dn, l, DEVICE, chunk / 8 + 256);
        POOLED(h, l, PINNED, chunk * w + 256);
        POOLED(hn, l, PINNED, chunk + 256);
        memset(hn, 0, chunk + 256);
        l.d_out.push_back(d); l.d_out_nulls.push_back(dn); l.h_out.push_back(h); l.h_out_nulls.push_back(hn);
      }
      POOLED(db, l, DEVICE, chunk + 256);
      l.d_bools = db;
      POOLED(dc, l, DEVICE, 64);
      POOLED(hc, l, PINNED, 64);
      l.d_count = static_cast<int64_t*>(dc);
      l.h_count = static_cast<int64_t*>(hc);
    }
    chunk_ = chunk;
    mode_ = STREAM;
    return Success();
  }

  // Queues H2D + kernel + D2H of the next chunk on the lane's stream (asynchronous).
  FailureOrVoid LaunchNext(int lane) {
    Lane& l = lanes_[lane];
    const rowcount_t total = plan_.base.row_count();
    if (next_first_ >= total) return Success();
    const rowcount_t first = next_first_;
    const rowcount_t rows = total - first < chunk_ ? total - first : chunk_;
    next_first_ += rows;
    vector<ssb_column> ic(used_.size()), oc(l.d_out.size());
    uint32_t narrow_mask = 0;
    double narrow_seconds = 0;
    size_t chunk_bytes = 0;
    for (size_t k = 0; k < used_.size(); ++k) {
      const Column& src = plan_.base.column(used_[k]);
      const size_t w = src.type_info().size();
      const char* host = static_cast<const char*>(src.data().raw()) + first * w;
      ic[k].data = l.d_in[k];
      ic[k].dtype = src.attribute().type();
      ic[k].reserved = 0;
      ic[k].nulls = NULL;
      // transfer narrowing: a column that did not fit is not tried again for the next 16 chunks
      bool narrowed = false;
      chunk_bytes += rows * w;
      if (narrowing_ && l.h_narrow[k] != NULL && narrow_skip_[k] == 0) {
        const std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
        narrowed = NarrowInt64Column(reinterpret_cast<const int64*>(host), static_cast<int32*>(l.h_narrow[k]), rows);
        narrow_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (!narrowed) narrow_skip_[k] = 16;
      } else if (narrow_skip_[k] > 0) {
        --narrow_skip_[k];
      }
      if (narrowed) {
        LANE_CALL(l, ssb_memcpy_h2d(l.ctx, l.d_in[k], l.h_narrow[k], rows * 4), "upload (narrowed)");
        ic[k].dtype = INT32;
        narrow_mask |= 1u << k;
      } else {
        LANE_CALL(l, ssb_memcpy_h2d(l.ctx, l.d_in[k], host, rows * w), "upload");
      }
      if (src.is_null() != NULL) {
        LANE_CALL(l, ssb_memcpy_h2d(l.ctx, l.d_bools, src.is_null() + first, rows), "upload nulls");
        LANE_CALL(l, ssb_nulls_pack(l.ctx, static_cast<const uint8_t*>(l.d_bools), static_cast<int64_t>(rows),
                                    static_cast<uint32_t*>(l.d_in_nulls[k])), "null pack");
        ic[k].nulls = static_cast<uint32_t*>(l.d_in_nulls[k]);
      }
    }
    for (size_t j = 0; j < oc.size(); ++j) {
      oc[j].data = l.d_out[j];
      oc[j].nulls = static_cast<uint32_t*>(l.d_out_nulls[j]);
      oc[j].dtype = plan_.schema.attribute(static_cast<int>(j)).type();
      oc[j].reserved = 0;
    }
    ssb_column dummy;
    memset(&dummy, 0, sizeof(dummy));
    // Narrowing pays while the host narrows a chunk faster than the bus would move it unnarrowed
    // (the two overlap across the lanes); on a host with few cores it does not, and is switched off.
    if (narrowing_ && narrow_seconds > 0) {
      narrow_total_seconds_ += narrow_seconds;
      narrow_total_bytes_ += static_cast<double>(chunk_bytes);
      if (++narrow_chunks_ >= 4 && narrow_total_seconds_ > narrow_total_bytes_ / 50e9) {
        narrowing_ = false;
        NarrowingUnprofitable().store(true, std::memory_order_relaxed);   // later cursors of this process do not probe again
      }
    }
    ssb_program* program = l.program;
    if (narrow_mask != 0) {
      std::map<uint32_t, ssb_program*>::iterator it = l.narrowed.find(narrow_mask);
      if (it == l.narrowed.end()) {
        vector<ssb_expr_node> nodes;
        vector<int32_t> outs, types, nullable;
        int pred = -1;
        NarrowedProgram(nodes_, outs_, pred_, narrow_mask, &nodes, &outs, &pred);
        for (size_t k = 0; k < used_.size(); ++k) {
          types.push_back(((narrow_mask >> k) & 1u) ? static_cast<int32_t>(INT32) : static_cast<int32_t>(plan_.base_schema.attribute(used_[k]).type()));
          nullable.push_back(plan_.base_schema.attribute(used_[k]).is_nullable() ? 1 : 0);
        }
        int32_t dummy32 = 0;
        ssb_program* variant = NULL;
        LANE_CALL(l, ssb_program_create(l.ctx, nodes.data(), static_cast<int32_t>(nodes.size()), static_cast<int32_t>(used_.size()),
                                        types.data(), nullable.data(), outs.empty() ? &dummy32 : outs.data(),
                                        static_cast<int32_t>(outs.size()), pred, &variant), "expression compilation (narrowed inputs)");
        it = l.narrowed.insert(std::make_pair(narrow_mask, variant)).first;
      }
      program = it->second;
    }
    l.launched = program;
    LANE_CALL(l, ssb_program_run(program, ic.empty() ? &dummy : ic.data(), static_cast<int64_t>(rows),
                                 oc.empty() ? &dummy : oc.data(), l.d_count), "expression evaluation");
    LANE_CALL(l, ssb_memcpy_d2h(l.ctx, l.h_count, l.d_count, sizeof(int64_t)), "download");
    // Without a predicate every row is kept: the outputs can follow the kernel at once. With one, the
    // kept-row count is not known on the host yet; Next() fetches exactly the kept rows once it is
    // (copying the chunk's capacity instead moved twice the bytes at selectivity 0.5).
    l.fetched = false;
    if (!plan_.predicate) {
      PROPAGATE_ON_FAILURE(QueueOutputs(&l, rows));
      l.fetched = true;
    }
    l.busy = true;
    return Success();
  }

  // Queues the D2H copies of the first `kept` rows of every output column of the lane's chunk.
  FailureOrVoid QueueOutputs(Lane* lane, rowcount_t kept) {
    Lane& l = *lane;
    if (kept == 0) return Success();
    for (size_t j = 0; j < l.d_out.size(); ++j) {
      const size_t w = GetTypeInfo(plan_.schema.attribute(static_cast<int>(j)).type()).size();
      LANE_CALL(l, ssb_memcpy_d2h(l.ctx, l.h_out[j], l.d_out[j], kept * w), "download");
      if (plan_.schema.attribute(static_cast<int>(j)).is_nullable() &&
          ssb_program_output_nullable(l.program, static_cast<int32_t>(j))) {
        LANE_CALL(l, ssb_nulls_unpack(l.ctx, static_cast<const uint32_t*>(l.d_out_nulls[j]), static_cast<int64_t>(kept),
                                      static_cast<uint8_t*>(l.d_bools)), "null unpack");
        LANE_CALL(l, ssb_memcpy_d2h(l.ctx, l.h_out_nulls[j], l.d_bools, kept), "download nulls");
      }
    }
    return Success();
  }
#undef LANE_CALL
#undef POOLED

  RowwisePlan plan_;
  CursorId id_;
  Mode mode_;
  View out_view_;
  vector<ssb_expr_node> nodes_;
  vector<int> used_;
  vector<int32_t> outs_;
  int pred_ = -1;
  bool narrowing_ = false;
  vector<int> narrow_skip_;
  double narrow_total_seconds_ = 0, narrow_total_bytes_ = 0;
  int narrow_chunks_ = 0;
  Lane lanes_[2];
  rowcount_t chunk_, next_first_;
  int active_;
  rowcount_t served_;
};

FailureOrOwned<Cursor> CreateRowwiseCursor(const Operation* op, BufferAllocator* allocator, CursorId id) {
  RowwisePlan plan;
  PROPAGATE_ON_FAILURE(DescribeAny(op, &plan));
  return Success(static_cast<Cursor*>(new RowwiseCursor(plan, allocator, id)));
}

// ------------------------------------------------------------------ ScanView
class ScanViewOperation : public BasicOperation {
 public:
  explicit ScanViewOperation(const View& view) : view_(view) {}
  virtual FailureOrOwned<Cursor> CreateCursor() const { return Success(static_cast<Cursor*>(new ViewCursor(view_))); }
  virtual bool DescribeRowwise(RowwisePlan* plan, Exception**) const {
    plan->base = view_;
    plan->source.reset();
    plan->base_schema = view_.schema();
    plan->schema = view_.schema();
    plan->outputs.clear();
    for (int i = 0; i < view_.schema().attribute_count(); ++i) plan->outputs.push_back(MakeInputNode(view_.schema(), i));
    plan->predicate.reset();
    return true;
  }
 protected:
  virtual string DebugName() const { return "ScanView"; }
 private:
  View view_;
};

// ------------------------------------------------------------------ Compute
class ComputeOperation : public BasicOperation {
 public:
  ComputeOperation(const Expression* computation, Operation* child) : BasicOperation(child), computation_(computation) {}
  virtual FailureOrOwned<Cursor> CreateCursor() const { return CreateRowwiseCursor(this, buffer_allocator(), COMPUTE); }
  virtual bool DescribeRowwise(RowwisePlan* plan, Exception** error) const {
    FailureOrVoid d = DescribeAny(child(), plan);
    if (d.is_failure()) { *error = d.release_exception(); return true; }
    FailureOrOwned<BoundExpression> bound = computation_->DoBind(plan->schema, buffer_allocator(), Cursor::kDefaultRowCount);
    if (bound.is_failure()) { *error = bound.release_exception(); return true; }
    vector<NodePtr> outs;
    // a Compute above a Filter evaluates only the rows the Filter kept (signaling operators: GuardSignaling)
    for (int i = 0; i < bound->column_count(); ++i) {
      outs.push_back(GuardSignaling(Substitute(bound->node(i), plan->outputs), plan->outputs, plan->predicate));
    }
    plan->outputs = outs;
    plan->schema = bound->result_schema();
    return true;
  }
 protected:
  virtual string DebugName() const { return "Compute"; }
 private:
  std::unique_ptr<const Expression> computation_;
};

// ------------------------------------------------------------------ Filter
class FilterOperation : public BasicOperation {
 public:
  FilterOperation(const Expression* predicate, const SingleSourceProjector* projector, Operation* child)
      : BasicOperation(child), predicate_(predicate), projector_(projector) {}
  virtual FailureOrOwned<Cursor> CreateCursor() const { return CreateRowwiseCursor(this, buffer_allocator(), FILTER); }
  virtual bool DescribeRowwise(RowwisePlan* plan, Exception** error) const {
    FailureOrVoid d = DescribeAny(child(), plan);
    if (d.is_failure()) { *error = d.release_exception(); return true; }
    FailureOrOwned<BoundExpression> bound = predicate_->DoBind(plan->schema, buffer_allocator(), Cursor::kDefaultRowCount);
    if (bound.is_failure()) { *error = bound.release_exception(); return true; }
    // cursor/core/filter.cc:79-87
    if (bound->column_count() != 1) {
      *error = new Exception(ERROR_ATTRIBUTE_COUNT_MISMATCH, "Predicate has to return exactly one column in (" +
                                                                 bound->result_schema().GetHumanReadableSpecification() + ")");
      return true;
    }
    if (bound->result_schema().attribute(0).type() != BOOL) {
      *error = new Exception(ERROR_ATTRIBUTE_TYPE_MISMATCH, "Predicate has to return column of type BOOL in (" +
                                                                bound->result_schema().GetHumanReadableSpecification() + ")");
      return true;
    }
    FailureOrOwned<const BoundSingleSourceProjector> proj = projector_->Bind(plan->schema);
    if (proj.is_failure()) { *error = proj.release_exception(); return true; }
    NodePtr pred = GuardSignaling(Substitute(bound->node(0), plan->outputs), plan->outputs, plan->predicate);
    plan->predicate = plan->predicate ? MakeBinaryLogic(SSB_OP_AND, plan->predicate, pred) : pred;
    vector<NodePtr> outs;
    for (int i = 0; i < proj->result_schema().attribute_count(); ++i) {
      outs.push_back(plan->outputs[proj->source_attribute_position(i)]);
    }
    plan->outputs = outs;
    plan->schema = proj->result_schema();
    return true;
  }
 protected:
  virtual string DebugName() const { return "Filter"; }
 private:
  std::unique_ptr<const Expression> predicate_;
  std::unique_ptr<const SingleSourceProjector> projector_;
};

// ------------------------------------------------------------------ Project
class ProjectOperation : public BasicOperation {
 public:
  ProjectOperation(const SingleSourceProjector* projector, Operation* child) : BasicOperation(child), projector_(projector) {}
  virtual FailureOrOwned<Cursor> CreateCursor() const { return CreateRowwiseCursor(this, buffer_allocator(), PROJECT); }
  virtual bool DescribeRowwise(RowwisePlan* plan, Exception** error) const {
    FailureOrVoid d = DescribeAny(child(), plan);
    if (d.is_failure()) { *error = d.release_exception(); return true; }
    FailureOrOwned<const BoundSingleSourceProjector> proj = projector_->Bind(plan->schema);
    if (proj.is_failure()) { *error = proj.release_exception(); return true; }
    vector<NodePtr> outs;
    for (int i = 0; i < proj->result_schema().attribute_count(); ++i) {
      outs.push_back(plan->outputs[proj->source_attribute_position(i)]);
    }
    plan->outputs = outs;
    plan->schema = proj->result_schema();
    return true;
  }
 protected:
  virtual string DebugName() const { return "Project"; }
 private:
  std::unique_ptr<const SingleSourceProjector> projector_;
};

// ------------------------------------------------------------------ GroupAggregate
struct BoundAggregation {
  ssb_agg_spec spec;
  int input_position;   // child column, -1 for COUNT(*)
  bool distinct;        // COUNT / SUM over the distinct non-NULL values of the input per group
};

bool IsNumericType(DataType t) { return GetTypeInfo(t).is_numeric(); }
}  // namespace
struct Aggregator::Impl {
  vector<BoundAggregation> aggs;
};
namespace {

// cursor/core/aggregator.cc:63-152, column_aggregator.cc:520-566: result types and nullability.
FailureOrVoid BindAggregations(const AggregationSpecification& spec, const TupleSchema& child,
                               vector<BoundAggregation>* out, TupleSchema* result_schema) {
  for (int i = 0; i < spec.size(); ++i) {
    const AggregationSpecification::Element& e = spec.aggregation(i);
    BoundAggregation b;
    memset(&b.spec, 0, sizeof(b.spec));
    const Aggregation fn = e.aggregation_operator();
    // column_aggregator.cc:333-433 (DistinctAggregator): duplicates of the input value inside a group count once.
    // MIN / MAX do not change under DISTINCT; COUNT and SUM take the two-phase form of GroupCursor::RunDistinct.
    b.distinct = e.is_distinct() && (fn == COUNT || fn == SUM);
    if (e.is_distinct() && (fn == FIRST || fn == LAST)) {
      THROW(new Exception(ERROR_NOT_IMPLEMENTED, "DISTINCT FIRST / LAST aggregations are not on the B200 hot path"));
    }
    if (fn == CONCAT) THROW(new Exception(ERROR_NOT_IMPLEMENTED, "CONCAT needs STRING columns (SURVEY 8f)"));
    b.spec.fn = fn;
    b.input_position = -1;
    DataType in_type = INT64;
    bool in_nullable = false;
    if (e.input().empty()) {
      if (fn != COUNT) THROW(new Exception(ERROR_ATTRIBUTE_MISSING, "Only COUNT may have an empty input attribute name"));
      b.distinct = false;
    } else {
      b.input_position = child.LookupAttributePosition(e.input());
      if (b.input_position < 0) {
        THROW(new Exception(ERROR_ATTRIBUTE_MISSING, "No attribute '" + e.input() + "' in the schema: (" +
                                                         child.GetHumanReadableSpecification() + ")"));
      }
      in_type = child.attribute(b.input_position).type();
      in_nullable = child.attribute(b.input_position).is_nullable();
    }
    DataType out_type;
    Nullability out_null = NULLABLE;
    if (fn == COUNT) {
      out_type = e.output_type_specified() ? e.output_type() : UINT64;
      out_null = NOT_NULLABLE;
      if (!GetTypeInfo(out_type).is_integer()) {
        THROW(new Exception(ERROR_INVALID_ARGUMENT_TYPE, "Aggregation not supported. Count can not store result in output column of type " + DataType_Name(out_type) + "."));
      }
    } else {
      out_type = e.output_type_specified() ? e.output_type() : in_type;
      if (fn == SUM && !IsNumericType(in_type)) {
        THROW(new Exception(ERROR_INVALID_ARGUMENT_TYPE, "Aggregation not supported. Aggregation function SUM not defined for types " +
                                                             DataType_Name(in_type) + " and " + DataType_Name(out_type) + "."));
      }
      if (in_type == BINARY) {   // column_aggregator.cc: no aggregator but COUNT is instantiated for BINARY
        THROW(new Exception(ERROR_INVALID_ARGUMENT_TYPE, "Aggregation not supported. Aggregation function " + Aggregation_Name(fn) +
                                                             " not defined for types " + DataType_Name(in_type) + " and " +
                                                             DataType_Name(out_type) + "."));
      }
      if (out_type != in_type) {
        if (!(IsNumericType(in_type) && IsNumericType(out_type))) {
          THROW(new Exception(ERROR_INVALID_ARGUMENT_TYPE,
                              "Aggregation not supported. Aggregation function " + Aggregation_Name(fn) +
                                  " not defined for types " + DataType_Name(in_type) + " and " + DataType_Name(out_type) + "."));
        }
      }
    }
    b.spec.input = -1;   // filled by the cursor (index into the values array)
    // MIN / MAX / FIRST / LAST / COUNT over STRING / BINARY run on the column's order-preserving codes
    b.spec.in_type = DeviceType(in_type);
    b.spec.out_type = DeviceType(out_type);
    b.spec.in_nullable = in_nullable ? 1 : 0;
    if (!result_schema->add_attribute(Attribute(e.output(), out_type, out_null))) {
      THROW(new Exception(ERROR_ATTRIBUTE_EXISTS, "Duplicate attribute name '" + e.output() + "' in result schema"));
    }
    out->push_back(b);
  }
  return Success();
}

class GroupCursor : public GpuCursor {
 public:
  GroupCursor(const TupleSchema& schema, BufferAllocator* allocator, Cursor* child, const vector<int>& keys,
              const vector<BoundAggregation>& aggs, size_t estimated_groups, bool scalar)
      : GpuCursor(schema, allocator, scalar ? "ScalarAggregateCursor" : "GroupAggregateCursor"), child_(child),
        keys_(keys), aggs_(aggs), estimated_groups_(estimated_groups), scalar_(scalar), group_(NULL),
        quota_(std::numeric_limits<size_t>::max()), budgeted_(false) {}
  // The reference's memory contract for the result block (see cursor.h at GroupAggregate).
  void SetResultBudget(size_t memory_quota) { quota_ = memory_quota; budgeted_ = true; }
  static size_t ResultRowBytes(const TupleSchema& schema) {
    size_t bytes = 0;
    for (int i = 0; i < schema.attribute_count(); ++i) {
      bytes += GetTypeInfo(schema.attribute(i).type()).size() + (schema.attribute(i).is_nullable() ? 1 : 0);
    }
    return bytes ? bytes : 1;
  }
  // The child is a chain of row-wise operators (Filter / Compute / Project over a scan): its
  // plan is evaluated inside the aggregation kernel (ssb_group_update_program), nothing is
  // materialised between the two operators.
  void FuseWith(const RowwisePlan& plan) { fused_.reset(new RowwisePlan(plan)); }
  virtual void DropFusion() { fused_.reset(); }
  virtual ~GroupCursor() { if (group_) ssb_group_destroy(group_); }
  virtual CursorId GetCursorId() const { return scalar_ ? SCALAR_AGGREGATE : GROUP_AGGREGATE; }
  virtual void Interrupt() { GpuCursor::Interrupt(); child_->Interrupt(); }
  // cursor/base/cursor.h:210 (basic_cursor.h: every child is handed to the transformer and replaced by its result).
  // A transformed child is no longer known to be a GPU cursor: its rows then arrive through Next().
  virtual void ApplyToChildren(CursorTransformer* transformer) { child_.reset(transformer->Transform(child_.release())); DropFusion(); }
 protected:
  // Returns true when the fused form ran; false = not applicable (plan too wide for one program).
  FailureOr<bool> RunFused(Session* s, DeviceTable* result) {
    const RowwisePlan& plan = *fused_;
    // variable-length columns: the row-wise child resolves dictionaries itself (RowwiseCursor::RunWithStrings)
    for (int i = 0; i < plan.base_schema.attribute_count(); ++i) if (IsVariableLength(plan.base_schema.attribute(i).type())) return Success(false);
    for (int i = 0; i < plan.schema.attribute_count(); ++i) if (IsVariableLength(plan.schema.attribute(i).type())) return Success(false);
    // program outputs: the key columns, then the distinct aggregate inputs in order of first use
    vector<NodePtr> outs;
    for (size_t k = 0; k < keys_.size(); ++k) outs.push_back(plan.outputs[keys_[k]]);
    vector<int> value_pos;
    vector<ssb_agg_spec> specs;
    for (size_t i = 0; i < aggs_.size(); ++i) {
      ssb_agg_spec sp = aggs_[i].spec;
      if (aggs_[i].input_position >= 0) {
        size_t v = 0;
        while (v < value_pos.size() && value_pos[v] != aggs_[i].input_position) ++v;
        if (v == value_pos.size()) { value_pos.push_back(aggs_[i].input_position); outs.push_back(plan.outputs[aggs_[i].input_position]); }
        sp.input = static_cast<int32_t>(v);
      }
      specs.push_back(sp);
    }
    FailureOrOwned<DeviceProgram> created = DeviceProgram::Create(plan.base_schema, outs, plan.predicate);
    if (created.is_failure()) {
      std::unique_ptr<Exception> e(created.release_exception());
      if (e->return_code() == ERROR_NOT_IMPLEMENTED) return Success(false);
      return Failure(e.release());
    }
    std::unique_ptr<DeviceProgram> program(created.release());
    // the base columns the program reads
    DeviceTable base;
    std::unique_ptr<Block> keepalive;
    int64 rows = 0;
    vector<ssb_column> ic;
    if (plan.source) {
      PROPAGATE_ON_FAILURE(MaterializeOnDevice(plan.source.get(), &base, &keepalive));
      rows = base.rows;
      for (size_t k = 0; k < program->used_inputs().size(); ++k) ic.push_back(base.columns[program->used_inputs()[k]].col);
    } else {
      rows = static_cast<int64>(plan.base.row_count());
    }
    vector<int32_t> key_types, key_nullable;
    for (size_t k = 0; k < keys_.size(); ++k) {
      const Attribute& a = plan.schema.attribute(keys_[k]);
      key_types.push_back(a.type());
      key_nullable.push_back((a.is_nullable() || ssb_program_output_nullable(program->handle(), static_cast<int32_t>(k))) ? 1 : 0);
    }
    for (size_t i = 0; i < specs.size(); ++i) {
      if (specs[i].input >= 0 && ssb_program_output_nullable(program->handle(), static_cast<int32_t>(keys_.size()) + specs[i].input)) {
        specs[i].in_nullable = 1;
      }
    }
    int32_t dummy = 0;
    ssb_column dummy_col;
    memset(&dummy_col, 0, sizeof(dummy_col));
    const int64 expected = estimated_groups_ ? static_cast<int64>(estimated_groups_) : 0;
    SSB_CALL(s, ssb_group_create(s->ctx(), static_cast<int32_t>(keys_.size()), key_types.empty() ? &dummy : key_types.data(),
                                 key_nullable.empty() ? &dummy : key_nullable.data(), static_cast<int32_t>(specs.size()),
                                 specs.data(), expected, &group_), "group-by setup");
    // A scanned host view is fed in chunks: the table accumulates across calls, so device memory is
    // bounded by one chunk of the input columns and inputs larger than HBM aggregate as well
    // (the reference spills for that: aggregate_groups.cc:490-1107). Device-resident views and GPU
    // children are handed over whole.
    int64 chunk = rows;
    if (!plan.source) {
      chunk = static_cast<int64>(1) << 26;
      if (const char* env = getenv("SSB200_GROUP_CHUNK_ROWS")) chunk = atoll(env);
      chunk = std::max<int64>(1024, (chunk / 1024) * 1024);   // bitmap words never straddle two chunks
    }
    int64 offset = 0;
    do {
      const int64 n = std::min<int64>(chunk, rows - offset);
      if (!plan.source) {
        ic.clear();
        PROPAGATE_ON_FAILURE(UploadColumns(plan.base, program->used_inputs(), static_cast<rowcount_t>(offset),
                                           static_cast<rowcount_t>(n), &base));
        for (size_t k = 0; k < base.columns.size(); ++k) ic.push_back(base.columns[k].col);
      }
      for (size_t k = 0; k < ic.size(); ++k) {   // a NOT_NULLABLE attribute never sets a bit of the bitmap a GPU child hands over
        if (!plan.base_schema.attribute(program->used_inputs()[k]).is_nullable()) ic[k].nulls = NULL;
      }
      SSB_CALL(s, ssb_group_update_program(group_, program->handle(), ic.empty() ? &dummy_col : ic.data(), n),
               "fused group-by");
      offset += n;
    } while (offset < rows);
    PROPAGATE_ON_FAILURE(Finish(s, specs, result));
    return Success(true);
  }

  // `in`: the aggregated table when it carries variable-length columns (their dictionaries pass to the result)
  FailureOrVoid Finish(Session* s, const vector<ssb_agg_spec>& specs, DeviceTable* result, const DeviceTable* in = NULL) {
    int64_t n_groups = 0;
    vector<ssb_column> kout(keys_.size() ? keys_.size() : 1), aout(specs.size() ? specs.size() : 1);
    SSB_CALL(s, ssb_group_finalize(group_, &n_groups, kout.data(), aout.data()), "group-by finalize");
    if (budgeted_) {
      const size_t quota = std::min(quota_, allocator()->Available());
      const size_t rows_allowed = std::max<size_t>(estimated_groups_, quota / ResultRowBytes(schema()));
      if (static_cast<size_t>(n_groups) > rows_allowed) {
        char buf[200];
        snprintf(buf, sizeof(buf), "Can't allocate the result block of GroupAggregate: %lld groups, %zu rows fit the memory quota",
                 static_cast<long long>(n_groups), rows_allowed);
        THROW(new Exception(ERROR_MEMORY_EXCEEDED, buf));
      }
    }
    result->schema = schema();
    result->columns.clear();
    for (size_t k = 0; k < keys_.size(); ++k) {
      DeviceColumnRef c;
      c.col = kout[k];
      if (in != NULL) c.dict = in->columns[keys_[k]].dict;
      result->columns.push_back(c);
    }
    for (size_t i = 0; i < specs.size(); ++i) {
      DeviceColumnRef c;
      c.col = aout[i];
      // MIN / MAX / FIRST / LAST of a variable-length column: a code of the input's dictionary
      if (in != NULL && aggs_[i].input_position >= 0 && IsVariableLength(schema().attribute(static_cast<int>(keys_.size() + i)).type())) {
        c.dict = in->columns[aggs_[i].input_position].dict;
      }
      result->columns.push_back(c);
    }
    result->rows = n_groups;
    return Success();
  }

  // DISTINCT aggregates (COUNT / SUM over the distinct non-NULL input values of a group; the reference keeps a hash set
  // per aggregate and group, column_aggregator.cc:333-433). Here in passes over the same hash-aggregation kernels:
  //   pass 0         the non-distinct aggregates, grouped by the keys;
  //   per input x    (a) the distinct (keys, x) combinations = a group-by on keys + x, (b) the DISTINCT aggregates over x
  //                  grouped by the keys, fed with those combinations;
  // every pass's result is merged into the final table as partial aggregates (ssb_group_merge; the aggregates a pass
  // does not compute travel as NULL partials), which also lines the passes up by key, NULL keys included.
  struct ScopedGroup {
    ssb_group* g;
    ScopedGroup() : g(NULL) {}
    ~ScopedGroup() { if (g) ssb_group_destroy(g); }
  };
  FailureOrVoid MergePass(Session* s, const vector<int32_t>& key_types, const vector<int32_t>& key_nullable,
                          const vector<ssb_column>& key_cols, const vector<size_t>& which, const vector<ssb_agg_spec>& pass_specs,
                          const vector<ssb_column>& value_cols, int64 rows, const vector<ssb_agg_spec>& all_specs) {
    int32_t dummy = 0;
    ssb_column dummy_col;
    memset(&dummy_col, 0, sizeof(dummy_col));
    ScopedGroup pass;
    SSB_CALL(s, ssb_group_create(s->ctx(), static_cast<int32_t>(keys_.size()), key_types.empty() ? &dummy : key_types.data(),
                                 key_nullable.empty() ? &dummy : key_nullable.data(), static_cast<int32_t>(pass_specs.size()),
                                 pass_specs.data(), 0, &pass.g), "group-by setup (DISTINCT pass)");
    SSB_CALL(s, ssb_group_update(pass.g, key_cols.empty() ? &dummy_col : key_cols.data(),
                                 value_cols.empty() ? &dummy_col : value_cols.data(), rows), "group-by (DISTINCT pass)");
    int64_t n = 0;
    vector<ssb_column> kout(keys_.size() ? keys_.size() : 1), aout(pass_specs.size() ? pass_specs.size() : 1);
    SSB_CALL(s, ssb_group_finalize(pass.g, &n, kout.data(), aout.data()), "group-by finalize (DISTINCT pass)");
    if (n == 0) return Success();
    // partials of the aggregates this pass does not compute: NULL (COUNT: zero)
    DeviceBuffer zeros, ones;
    PROPAGATE_ON_FAILURE(zeros.Allocate(static_cast<size_t>(n) * 8 + 256));
    PROPAGATE_ON_FAILURE(ones.Allocate(static_cast<size_t>(n / 32 + 2) * 4 + 256));
    SSB_CALL(s, ssb_memset(s->ctx(), zeros.get(), 0, static_cast<size_t>(n) * 8 + 256), "memset");
    SSB_CALL(s, ssb_memset(s->ctx(), ones.get(), 0xff, static_cast<size_t>(n / 32 + 2) * 4 + 256), "memset");
    vector<ssb_column> full(all_specs.size());
    for (size_t a = 0; a < all_specs.size(); ++a) {
      full[a].data = zeros.get();
      full[a].nulls = all_specs[a].fn == SSB_AGG_COUNT ? NULL : static_cast<uint32_t*>(ones.get());
      full[a].dtype = all_specs[a].out_type;
      full[a].reserved = 0;
    }
    for (size_t k = 0; k < which.size(); ++k) full[which[k]] = aout[k];
    SSB_CALL(s, ssb_group_merge(group_, n, kout.data(), full.data()), "group-by merge (DISTINCT pass)");
    SSB_CALL(s, ssb_ctx_sync(s->ctx()), "sync");   // the pass's table and the NULL partials go away with this scope
    return Success();
  }

  FailureOrVoid RunDistinct(Session* s, DeviceTable* result) {
    DeviceTable in;
    std::unique_ptr<Block> keepalive;
    PROPAGATE_ON_FAILURE(MaterializeOnDevice(child_.get(), &in, &keepalive));
    vector<int32_t> key_types, key_nullable;
    vector<ssb_column> key_cols;
    for (size_t k = 0; k < keys_.size(); ++k) {
      const Attribute& a = child_->schema().attribute(keys_[k]);
      key_types.push_back(DeviceType(a.type()));
      key_nullable.push_back(a.is_nullable() ? 1 : 0);
      key_cols.push_back(in.columns[keys_[k]].col);
    }
    // the final table: every aggregate, fed with partials only
    vector<ssb_agg_spec> all_specs;
    for (size_t i = 0; i < aggs_.size(); ++i) {
      ssb_agg_spec sp = aggs_[i].spec;
      sp.input = aggs_[i].input_position >= 0 ? static_cast<int32_t>(i) : -1;
      if (sp.fn != SSB_AGG_COUNT) sp.in_nullable = 1;   // the table receives NULL partials: it tracks which groups saw a value
      all_specs.push_back(sp);
    }
    int32_t dummy = 0;
    SSB_CALL(s, ssb_group_create(s->ctx(), static_cast<int32_t>(keys_.size()), key_types.empty() ? &dummy : key_types.data(),
                                 key_nullable.empty() ? &dummy : key_nullable.data(), static_cast<int32_t>(all_specs.size()),
                                 all_specs.data(), estimated_groups_ ? static_cast<int64>(estimated_groups_) : 0, &group_), "group-by setup");
    // pass 0: the non-distinct aggregates (with none, a COUNT(*) nobody reads keeps groups alive whose DISTINCT inputs are all NULL)
    {
      vector<size_t> which;
      vector<ssb_agg_spec> specs;
      vector<ssb_column> values;
      for (size_t i = 0; i < aggs_.size(); ++i) {
        if (aggs_[i].distinct) continue;
        ssb_agg_spec sp = aggs_[i].spec;
        if (aggs_[i].input_position >= 0) { sp.input = static_cast<int32_t>(values.size()); values.push_back(in.columns[aggs_[i].input_position].col); }
        which.push_back(i);
        specs.push_back(sp);
      }
      if (!specs.empty()) PROPAGATE_ON_FAILURE(MergePass(s, key_types, key_nullable, key_cols, which, specs, values, in.rows, all_specs));
    }
    // one pair of passes per DISTINCT input column
    std::set<int> inputs;
    for (size_t i = 0; i < aggs_.size(); ++i) if (aggs_[i].distinct) inputs.insert(aggs_[i].input_position);
    for (std::set<int>::const_iterator it = inputs.begin(); it != inputs.end(); ++it) {
      const int px = *it;
      const Attribute& xa = child_->schema().attribute(px);
      // (a) the distinct (keys, x) combinations
      vector<int32_t> kt2(key_types), kn2(key_nullable);
      vector<ssb_column> kc2(key_cols);
      kt2.push_back(DeviceType(xa.type()));
      kn2.push_back(xa.is_nullable() ? 1 : 0);
      kc2.push_back(in.columns[px].col);
      if (kt2.size() > 8) THROW(new Exception(ERROR_NOT_IMPLEMENTED, "DISTINCT aggregation with eight group-by columns"));
      ssb_agg_spec star;
      memset(&star, 0, sizeof(star));
      star.fn = SSB_AGG_COUNT; star.input = -1; star.in_type = SSB_INT64; star.out_type = SSB_UINT64;
      ScopedGroup combos;
      ssb_column dummy_col;
      memset(&dummy_col, 0, sizeof(dummy_col));
      SSB_CALL(s, ssb_group_create(s->ctx(), static_cast<int32_t>(kt2.size()), kt2.data(), kn2.data(), 1, &star, 0, &combos.g), "group-by setup (DISTINCT values)");
      SSB_CALL(s, ssb_group_update(combos.g, kc2.data(), &dummy_col, in.rows), "group-by (DISTINCT values)");
      int64_t n = 0;
      vector<ssb_column> kout(kt2.size()), aout(1);
      SSB_CALL(s, ssb_group_finalize(combos.g, &n, kout.data(), aout.data()), "group-by finalize (DISTINCT values)");
      // (b) the DISTINCT aggregates of this input over the combinations
      vector<size_t> which;
      vector<ssb_agg_spec> specs;
      for (size_t i = 0; i < aggs_.size(); ++i) {
        if (!aggs_[i].distinct || aggs_[i].input_position != px) continue;
        ssb_agg_spec sp = aggs_[i].spec;
        sp.input = 0;
        which.push_back(i);
        specs.push_back(sp);
      }
      vector<ssb_column> kcols(kout.begin(), kout.begin() + static_cast<long>(keys_.size()));
      vector<ssb_column> values(1, kout[keys_.size()]);
      PROPAGATE_ON_FAILURE(MergePass(s, key_types, key_nullable, kcols, which, specs, values, n, all_specs));
    }
    return Finish(s, all_specs, result, &in);
  }

  virtual FailureOrVoid Run(DeviceTable* result) {
    FailureOr<Session*> sr = Session::Get();
    PROPAGATE_ON_FAILURE(sr);
    Session* s = sr.get();
    for (size_t i = 0; i < aggs_.size(); ++i) if (aggs_[i].distinct) return RunDistinct(s, result);
    if (fused_ && !aggs_.empty()) {
      FailureOr<bool> fused = RunFused(s, result);
      PROPAGATE_ON_FAILURE(fused);
      if (fused.get()) return Success();
    }
    DeviceTable in;
    std::unique_ptr<Block> keepalive;
    PROPAGATE_ON_FAILURE(MaterializeOnDevice(child_.get(), &in, &keepalive));
    vector<int32_t> key_types, key_nullable;
    vector<ssb_column> key_cols, value_cols;
    for (size_t k = 0; k < keys_.size(); ++k) {
      const Attribute& a = child_->schema().attribute(keys_[k]);
      key_types.push_back(DeviceType(a.type()));
      key_nullable.push_back(a.is_nullable() ? 1 : 0);
      key_cols.push_back(in.columns[keys_[k]].col);
    }
    vector<ssb_agg_spec> specs;
    for (size_t i = 0; i < aggs_.size(); ++i) {
      ssb_agg_spec sp = aggs_[i].spec;
      if (aggs_[i].input_position >= 0) {
        sp.input = static_cast<int32_t>(value_cols.size());
        value_cols.push_back(in.columns[aggs_[i].input_position].col);
      }
      specs.push_back(sp);
    }
    int32_t dummy = 0;
    ssb_column dummy_col;
    memset(&dummy_col, 0, sizeof(dummy_col));
    int64 expected = estimated_groups_ ? static_cast<int64>(estimated_groups_) : 0;
    SSB_CALL(s, ssb_group_create(s->ctx(), static_cast<int32_t>(keys_.size()), key_types.empty() ? &dummy : key_types.data(),
                                 key_nullable.empty() ? &dummy : key_nullable.data(), static_cast<int32_t>(specs.size()),
                                 specs.data(), expected, &group_), "group-by setup");
    SSB_CALL(s, ssb_group_update(group_, key_cols.empty() ? &dummy_col : key_cols.data(),
                                 value_cols.empty() ? &dummy_col : value_cols.data(), in.rows), "group-by");
    return Finish(s, specs, result, &in);
  }
 private:
  std::unique_ptr<RowwisePlan> fused_;
  std::unique_ptr<Cursor> child_;
  vector<int> keys_;
  vector<BoundAggregation> aggs_;
  size_t estimated_groups_;
  bool scalar_;
  ssb_group* group_;
  size_t quota_;
  bool budgeted_;
};

FailureOrVoid GatherColumns(Session* s, const DeviceTable& src, const vector<int>& positions, const int64_t* d_idx,
                            int64 n, bool force_nulls, DeviceTable* out, size_t first_out);

// ------------------------------------------------------------------ AggregateClusters
// cursor/core/aggregate_clusters.cc:233-433: the input is clustered by the key columns; every run of consecutive
// rows with equal keys becomes one output row, in input order (a key that comes back later is a new row). Here:
// cluster ids from one compare-with-predecessor pass and a scan (ssb_cluster_ids), the aggregates by cluster id in
// the hash table of GroupAggregate, the key values gathered from the first row of each cluster, rows ordered by id.
class ClustersCursor : public GpuCursor {
 public:
  ClustersCursor(const TupleSchema& schema, BufferAllocator* allocator, Cursor* child, const vector<int>& keys,
                 const vector<BoundAggregation>& aggs)
      : GpuCursor(schema, allocator, "AggregateClustersCursor"), child_(child), keys_(keys), aggs_(aggs) {}
  virtual CursorId GetCursorId() const { return AGGREGATE_CLUSTERS; }
  virtual void Interrupt() { GpuCursor::Interrupt(); child_->Interrupt(); }
  // cursor/base/cursor.h:210 (basic_cursor.h: every child is handed to the transformer and replaced by its result).
  // A transformed child is no longer known to be a GPU cursor: its rows then arrive through Next().
  virtual void ApplyToChildren(CursorTransformer* transformer) { child_.reset(transformer->Transform(child_.release())); DropFusion(); }
 protected:
  virtual FailureOrVoid Run(DeviceTable* result) {
    FailureOr<Session*> sr = Session::Get();
    PROPAGATE_ON_FAILURE(sr);
    Session* s = sr.get();
    DeviceTable in;
    std::unique_ptr<Block> keepalive;
    PROPAGATE_ON_FAILURE(MaterializeOnDevice(child_.get(), &in, &keepalive));
    const int64 rows = in.rows;
    vector<ssb_column> key_cols;
    for (size_t k = 0; k < keys_.size(); ++k) key_cols.push_back(in.columns[keys_[k]].col);
    DeviceBuffer ids, starts;
    PROPAGATE_ON_FAILURE(ids.Allocate(static_cast<size_t>(rows) * 8 + 128));
    PROPAGATE_ON_FAILURE(starts.Allocate(static_cast<size_t>(rows) * 8 + 128));
    ssb_column dummy_col;
    memset(&dummy_col, 0, sizeof(dummy_col));
    int64_t clusters = 0;
    SSB_CALL(s, ssb_cluster_ids(s->ctx(), static_cast<int32_t>(key_cols.size()), key_cols.empty() ? &dummy_col : key_cols.data(), rows,
                                static_cast<int64_t*>(ids.get()), static_cast<int64_t*>(starts.get()), &clusters), "cluster ids");
    PROPAGATE_ON_FAILURE(result->Allocate(schema(), clusters, /* force_nulls = */ true));
    result->rows = clusters;
    if (clusters == 0) return Success();
    // the aggregates, keyed by the cluster id
    vector<ssb_agg_spec> specs;
    vector<ssb_column> value_cols;
    for (size_t i = 0; i < aggs_.size(); ++i) {
      ssb_agg_spec sp = aggs_[i].spec;
      if (aggs_[i].input_position >= 0) {
        sp.input = static_cast<int32_t>(value_cols.size());
        value_cols.push_back(in.columns[aggs_[i].input_position].col);
      }
      specs.push_back(sp);
    }
    int32_t key_type = INT64, key_nullable = 0;
    ssb_group* g = NULL;
    SSB_CALL(s, ssb_group_create(s->ctx(), 1, &key_type, &key_nullable, static_cast<int32_t>(specs.size()), specs.data(), clusters, &g),
             "cluster aggregation setup");
    struct Guard { ssb_group* g; ~Guard() { if (g) ssb_group_destroy(g); } } guard = {g};
    ssb_column id_col;
    id_col.data = ids.get(); id_col.nulls = NULL; id_col.dtype = INT64; id_col.reserved = 0;
    SSB_CALL(s, ssb_group_update(g, &id_col, value_cols.empty() ? &dummy_col : value_cols.data(), rows), "cluster aggregation");
    int64_t n_groups = 0;
    ssb_column kout;
    vector<ssb_column> aout(specs.size() ? specs.size() : 1);
    SSB_CALL(s, ssb_group_finalize(g, &n_groups, &kout, aout.data()), "cluster aggregation finalize");
    if (n_groups != clusters) THROW(new Exception(ERROR_UNKNOWN_ERROR, "internal: cluster count mismatch"));
    // order by cluster id
    DeviceBuffer perm;
    PROPAGATE_ON_FAILURE(perm.Allocate(static_cast<size_t>(clusters) * 8 + 128));
    int32_t asc = 0;
    SSB_CALL(s, ssb_sort_permutation(s->ctx(), 1, &kout, &asc, clusters, static_cast<int64_t*>(perm.get())), "cluster order");
    for (size_t k = 0; k < keys_.size(); ++k) {
      const vector<int> pos(1, keys_[k]);
      PROPAGATE_ON_FAILURE(GatherColumns(s, in, pos, static_cast<const int64_t*>(starts.get()), clusters, false, result, k));
      if (in.columns[keys_[k]].col.nulls == NULL) {
        SSB_CALL(s, ssb_memset(s->ctx(), result->columns[k].col.nulls, 0, static_cast<size_t>((clusters + 31) / 32 + 1) * 4), "memset");
      }
    }
    for (size_t i = 0; i < specs.size(); ++i) {
      ssb_column dst = result->columns[keys_.size() + i].col;
      if (aout[i].nulls == NULL) {
        SSB_CALL(s, ssb_memset(s->ctx(), dst.nulls, 0, static_cast<size_t>((clusters + 31) / 32 + 1) * 4), "memset");
        dst.nulls = NULL;
      }
      SSB_CALL(s, ssb_gather(s->ctx(), &aout[i], static_cast<const int64_t*>(perm.get()), clusters, &dst), "gather");
      if (aggs_[i].input_position >= 0 && IsVariableLength(schema().attribute(static_cast<int>(keys_.size() + i)).type())) {
        result->columns[keys_.size() + i].dict = in.columns[aggs_[i].input_position].dict;
      }
    }
    SSB_CALL(s, ssb_ctx_sync(s->ctx()), "sync");
    return Success();
  }
 private:
  std::unique_ptr<Cursor> child_;
  vector<int> keys_;
  vector<BoundAggregation> aggs_;
};

class AggregateClustersOperation : public BasicOperation {
 public:
  AggregateClustersOperation(const SingleSourceProjector* clustered_by, const AggregationSpecification* aggregation, Operation* child)
      : BasicOperation(child), clustered_by_(clustered_by), aggregation_(aggregation) {}
  virtual FailureOrOwned<Cursor> CreateCursor() const {
    FailureOrOwned<Cursor> child_cursor = child()->CreateCursor();
    PROPAGATE_ON_FAILURE(child_cursor);
    const TupleSchema& cs = child_cursor->schema();
    TupleSchema result;
    vector<int> keys;
    FailureOrOwned<const BoundSingleSourceProjector> proj = clustered_by_->Bind(cs);
    PROPAGATE_ON_FAILURE(proj);
    for (int i = 0; i < proj->result_schema().attribute_count(); ++i) {
      keys.push_back(proj->source_attribute_position(i));
      result.add_attribute(proj->result_schema().attribute(i));
    }
    vector<BoundAggregation> aggs;
    PROPAGATE_ON_FAILURE(BindAggregations(*aggregation_, cs, &aggs, &result));
    for (size_t i = 0; i < aggs.size(); ++i) {
      if (aggs[i].distinct) THROW(new Exception(ERROR_NOT_IMPLEMENTED, "DISTINCT aggregations in AggregateClusters are not on the B200 hot path"));
    }
    return Success(static_cast<Cursor*>(new ClustersCursor(result, buffer_allocator(), child_cursor.release(), keys, aggs)));
  }
 protected:
  virtual string DebugName() const { return "AggregateClusters"; }
 private:
  std::unique_ptr<const SingleSourceProjector> clustered_by_;
  std::unique_ptr<const AggregationSpecification> aggregation_;
};

// ------------------------------------------------------------------ MergeUnionAll
// cursor/core/merge_union_all.cc:127-300: merges inputs that are each sorted by `sort_order` into one sorted
// stream (a priority queue over the inputs' current rows; ties between inputs come out in no defined order).
// Here: the inputs are concatenated in HBM and sorted by the key columns with the stable radix sort -- rows with
// equal keys keep input order, then row order. The result column is nullable if any input's column is (:296-311).
class ConcatCursor : public GpuCursor {
 public:
  ConcatCursor(const TupleSchema& schema, BufferAllocator* allocator, vector<Cursor*>* inputs)
      : GpuCursor(schema, allocator, "MergeUnionAllInputs") {
    for (size_t i = 0; i < inputs->size(); ++i) inputs_.push_back(std::unique_ptr<Cursor>((*inputs)[i]));
    inputs->clear();
  }
  virtual CursorId GetCursorId() const { return MERGE_UNION_ALL; }
  virtual void Interrupt() { GpuCursor::Interrupt(); for (size_t i = 0; i < inputs_.size(); ++i) inputs_[i]->Interrupt(); }
  virtual void ApplyToChildren(CursorTransformer* transformer) {
    for (size_t i = 0; i < inputs_.size(); ++i) inputs_[i].reset(transformer->Transform(inputs_[i].release()));
  }
 protected:
  virtual FailureOrVoid Run(DeviceTable* result) {
    FailureOr<Session*> sr = Session::Get();
    PROPAGATE_ON_FAILURE(sr);
    Session* s = sr.get();
    vector<DeviceTable> parts(inputs_.size());
    vector<std::unique_ptr<Block> > keep(inputs_.size());
    int64 total = 0;
    for (size_t i = 0; i < inputs_.size(); ++i) {
      PROPAGATE_ON_FAILURE(MaterializeOnDevice(inputs_[i].get(), &parts[i], &keep[i]));
      total += parts[i].rows;
    }
    PROPAGATE_ON_FAILURE(result->Allocate(schema(), total, /* force_nulls = */ true));
    result->rows = total;
    DeviceBuffer bytes, part_bytes;
    PROPAGATE_ON_FAILURE(bytes.Allocate(static_cast<size_t>(total) + 128));
    for (int c = 0; c < schema().attribute_count(); ++c) {
      const size_t w = DeviceWidth(schema().attribute(c).type());
      if (IsVariableLength(schema().attribute(c).type())) {   // one dictionary for the column of every input
        vector<DeviceColumnRef*> cols;
        vector<int64> col_rows;
        for (size_t i = 0; i < parts.size(); ++i) { cols.push_back(&parts[i].columns[c]); col_rows.push_back(parts[i].rows); }
        PROPAGATE_ON_FAILURE(UnifyDictionaries(s, cols, col_rows, vector<string>(), NULL));
        if (!parts.empty()) result->columns[c].dict = parts[0].columns[c].dict;
      }
      int64 at = 0;
      bool any_nulls = false;
      for (size_t i = 0; i < parts.size(); ++i) any_nulls = any_nulls || parts[i].columns[c].col.nulls != NULL;
      for (size_t i = 0; i < parts.size(); ++i) {
        const int64 n = parts[i].rows;
        if (n == 0) continue;
        SSB_CALL(s, ssb_memcpy_d2d(s->ctx(), static_cast<char*>(result->columns[c].col.data) + at * w, parts[i].columns[c].col.data,
                                   static_cast<size_t>(n) * w), "concatenate");
        if (any_nulls) {   // is_null bits do not concatenate at arbitrary row offsets: through one byte per row
          char* dst = static_cast<char*>(bytes.get()) + at;
          if (parts[i].columns[c].col.nulls != NULL) {
            SSB_CALL(s, ssb_nulls_unpack(s->ctx(), parts[i].columns[c].col.nulls, n, reinterpret_cast<uint8_t*>(dst)), "null unpack");
          } else {
            SSB_CALL(s, ssb_memset(s->ctx(), dst, 0, static_cast<size_t>(n)), "memset");
          }
        }
        at += n;
      }
      if (any_nulls && total > 0) {
        SSB_CALL(s, ssb_nulls_pack(s->ctx(), static_cast<const uint8_t*>(bytes.get()), total, result->columns[c].col.nulls), "null pack");
      } else {
        SSB_CALL(s, ssb_memset(s->ctx(), result->columns[c].col.nulls, 0, static_cast<size_t>((total + 31) / 32 + 1) * 4), "memset");
      }
    }
    SSB_CALL(s, ssb_ctx_sync(s->ctx()), "sync");
    return Success();
  }
 private:
  vector<std::unique_ptr<Cursor> > inputs_;
};

class GroupAggregateOperation : public BasicOperation {
 public:
  GroupAggregateOperation(const SingleSourceProjector* group_by, AggregationSpecification* aggregation,
                          GroupAggregateOptions* options, Operation* child, bool best_effort = false)
      : BasicOperation(child), group_by_(group_by), aggregation_(aggregation), options_(options), best_effort_(best_effort) {}
  virtual FailureOrOwned<Cursor> CreateCursor() const {
    FailureOrOwned<Cursor> child_cursor = child()->CreateCursor();
    PROPAGATE_ON_FAILURE(child_cursor);
    const TupleSchema& cs = child_cursor->schema();
    TupleSchema result;
    vector<int> keys;
    if (group_by_) {
      FailureOrOwned<const BoundSingleSourceProjector> proj = group_by_->Bind(cs);
      PROPAGATE_ON_FAILURE(proj);
      for (int i = 0; i < proj->result_schema().attribute_count(); ++i) {
        keys.push_back(proj->source_attribute_position(i));
        result.add_attribute(proj->result_schema().attribute(i));
      }
    }
    vector<BoundAggregation> aggs;
    PROPAGATE_ON_FAILURE(BindAggregations(*aggregation_, cs, &aggs, &result));
    const size_t est = options_ ? options_->estimated_result_row_count() : 16;
    // aggregate_groups.cc:465-469: the aggregator's first block comes out of the operation's allocator at bind time
    if (group_by_ != NULL && buffer_allocator()->Available() / GroupCursor::ResultRowBytes(result) < std::max<size_t>(est, 1)) {
      THROW(new Exception(ERROR_MEMORY_EXCEEDED, "Can't allocate the result block of GroupAggregate (allocator exhausted)"));
    }
    GroupCursor* cursor = new GroupCursor(result, buffer_allocator(), child_cursor.release(), keys, aggs, est, group_by_ == NULL);
    if (group_by_ != NULL && !best_effort_) {
      cursor->SetResultBudget(options_ ? options_->memory_quota() : std::numeric_limits<size_t>::max());
    }
    // A row-wise child that filters or computes is evaluated inside the aggregation kernel.
    RowwisePlan plan;
    Exception* describe_error = NULL;
    if (getenv("SSB200_HOST_FUSE_GROUP") == NULL || atoi(getenv("SSB200_HOST_FUSE_GROUP")) != 0) {
      if (child()->DescribeRowwise(&plan, &describe_error) && describe_error == NULL) {
        bool nontrivial = static_cast<bool>(plan.predicate);
        for (size_t k = 0; k < keys.size(); ++k) nontrivial = nontrivial || plan.outputs[keys[k]]->op != SSB_OP_INPUT;
        for (size_t i = 0; i < aggs.size(); ++i) {
          if (aggs[i].input_position >= 0) nontrivial = nontrivial || plan.outputs[aggs[i].input_position]->op != SSB_OP_INPUT;
        }
        for (size_t i = 0; i < aggs.size(); ++i) nontrivial = nontrivial && !aggs[i].distinct;   // DISTINCT runs in passes over the materialised child
        if (nontrivial) cursor->FuseWith(plan);
      }
      delete describe_error;
    }
    return Success(static_cast<Cursor*>(cursor));
  }
 protected:
  virtual string DebugName() const { return group_by_ ? "GroupAggregate" : "ScalarAggregate"; }
 private:
  std::unique_ptr<const SingleSourceProjector> group_by_;
  std::unique_ptr<AggregationSpecification> aggregation_;
  std::unique_ptr<GroupAggregateOptions> options_;
  bool best_effort_;
};

// ------------------------------------------------------------------ gather helper
FailureOrVoid GatherColumns(Session* s, const DeviceTable& src, const vector<int>& positions, const int64_t* d_idx,
                            int64 n, bool force_nulls, DeviceTable* out, size_t first_out) {
  for (size_t k = 0; k < positions.size(); ++k) {
    const DeviceColumnRef& from = src.columns[positions[k]];
    DeviceColumnRef& to = out->columns[first_out + k];
    if (to.col.nulls == NULL && (from.col.nulls != NULL || force_nulls)) {
      THROW(new Exception(ERROR_UNKNOWN_ERROR, "internal: gather target lacks a null bitmap"));
    }
    ssb_column dst = to.col;
    if (from.col.nulls == NULL && !force_nulls) dst.nulls = NULL;
    SSB_CALL(s, ssb_gather(s->ctx(), &from.col, d_idx, n, &dst), "gather");
    to.dict = from.dict;   // variable-length columns: the gathered codes keep their dictionary
  }
  return Success();
}

// ------------------------------------------------------------------ HashJoin
class HashJoinCursor : public GpuCursor {
 public:
  HashJoinCursor(const TupleSchema& schema, BufferAllocator* allocator, Cursor* lhs, Cursor* rhs, JoinType join_type,
                 KeyUniqueness uniqueness, const vector<int>& lhs_keys, const vector<int>& rhs_keys,
                 const BoundMultiSourceProjector* projector)
      : GpuCursor(schema, allocator, "HashJoinCursor"), lhs_(lhs), rhs_(rhs), join_type_(join_type),
        uniqueness_(uniqueness), lhs_keys_(lhs_keys), rhs_keys_(rhs_keys), projector_(projector), join_(NULL) {}
  virtual ~HashJoinCursor() { if (join_) ssb_join_destroy(join_); }
  virtual CursorId GetCursorId() const { return HASH_JOIN; }
  virtual void Interrupt() { GpuCursor::Interrupt(); lhs_->Interrupt(); rhs_->Interrupt(); }
  virtual void ApplyToChildren(CursorTransformer* transformer) {
    lhs_.reset(transformer->Transform(lhs_.release()));
    rhs_.reset(transformer->Transform(rhs_.release()));
  }
 protected:
  virtual FailureOrVoid Run(DeviceTable* result) {
    FailureOr<Session*> sr = Session::Get();
    PROPAGATE_ON_FAILURE(sr);
    Session* s = sr.get();
    std::unique_ptr<Block> keep_l, keep_r;
    // build side first (hash_join.cc:406-420), then the probe side
    PROPAGATE_ON_FAILURE(MaterializeOnDevice(rhs_.get(), &rhs_table_, &keep_r));
    PROPAGATE_ON_FAILURE(MaterializeOnDevice(lhs_.get(), &lhs_table_, &keep_l));
    // hash_join.cc:713-726: the reference binds any join type and refuses the others at the first lookup,
    // i.e. once the probe side has produced a row
    if (join_type_ != INNER && join_type_ != LEFT_OUTER && lhs_table_.rows > 0) {
      THROW(new Exception(ERROR_NOT_IMPLEMENTED, "Unsupported join_type in hash_join: " + JoinType_Name(join_type_)));
    }
    // variable-length keys: both sides' codes must come from one dictionary
    for (size_t k = 0; k < rhs_keys_.size() && k < lhs_keys_.size(); ++k) {
      if (!IsVariableLength(rhs_table_.schema.attribute(rhs_keys_[k]).type())) continue;
      vector<DeviceColumnRef*> cols;
      cols.push_back(&lhs_table_.columns[lhs_keys_[k]]);
      cols.push_back(&rhs_table_.columns[rhs_keys_[k]]);
      vector<int64> col_rows;
      col_rows.push_back(lhs_table_.rows);
      col_rows.push_back(rhs_table_.rows);
      PROPAGATE_ON_FAILURE(UnifyDictionaries(s, cols, col_rows, vector<string>(), NULL));
    }
    vector<ssb_column> rk, lk;
    for (size_t k = 0; k < rhs_keys_.size(); ++k) rk.push_back(rhs_table_.columns[rhs_keys_[k]].col);
    for (size_t k = 0; k < lhs_keys_.size(); ++k) lk.push_back(lhs_table_.columns[lhs_keys_[k]].col);
    SSB_CALL(s, ssb_join_build(s->ctx(), static_cast<int32_t>(rk.size()), rk.data(), rhs_table_.rows,
                               uniqueness_ == UNIQUE ? SSB_KEYS_UNIQUE : SSB_KEYS_NOT_UNIQUE, &join_), "hash join build");
    int64_t n_pairs = 0;
    const int64_t* d_l = NULL;
    const int64_t* d_r = NULL;
    SSB_CALL(s, ssb_join_probe(join_, lk.data(), lhs_table_.rows,
                               join_type_ == LEFT_OUTER ? SSB_JOIN_LEFT_OUTER : SSB_JOIN_INNER, &n_pairs, &d_l, &d_r),
             "hash join probe");
    PROPAGATE_ON_FAILURE(result->Allocate(schema(), n_pairs, /* force_nulls = */ true));
    for (int i = 0; i < projector_->result_schema().attribute_count(); ++i) {
      const int src = projector_->source_index(i);
      const vector<int> pos(1, projector_->source_attribute_position(i));
      const bool outer_side = (src == 1 && join_type_ == LEFT_OUTER);
      PROPAGATE_ON_FAILURE(GatherColumns(s, src == 0 ? lhs_table_ : rhs_table_, pos, src == 0 ? d_l : d_r, n_pairs,
                                         outer_side, result, static_cast<size_t>(i)));
      // columns gathered without nulls: clear the bitmap so the download reads zeros
      const DeviceColumnRef& from = (src == 0 ? lhs_table_ : rhs_table_).columns[pos[0]];
      if (from.col.nulls == NULL && !outer_side) {
        SSB_CALL(s, ssb_memset(s->ctx(), result->columns[i].col.nulls, 0, static_cast<size_t>((n_pairs + 31) / 32 + 1) * 4), "memset");
      }
    }
    result->rows = n_pairs;
    SSB_CALL(s, ssb_ctx_sync(s->ctx()), "sync");
    return Success();
  }
 private:
  std::unique_ptr<Cursor> lhs_, rhs_;
  JoinType join_type_;
  KeyUniqueness uniqueness_;
  vector<int> lhs_keys_, rhs_keys_;
  std::unique_ptr<const BoundMultiSourceProjector> projector_;
  ssb_join* join_;
  DeviceTable lhs_table_, rhs_table_;
};

// ------------------------------------------------------------------ Sort
class SortCursor : public GpuCursor {
 public:
  // limit < 0: every row; otherwise the first `limit` rows of the order (ExtendedSort, sort.cc:1010-1016)
  SortCursor(const TupleSchema& schema, BufferAllocator* allocator, Cursor* child,
             const vector<std::pair<int, ColumnOrder> >& keys, const vector<int>& projected, int64 limit = -1)
      : GpuCursor(schema, allocator, "SortCursor"), child_(child), keys_(keys), projected_(projected), limit_(limit) {}
  virtual CursorId GetCursorId() const { return SORT; }
  virtual void Interrupt() { GpuCursor::Interrupt(); child_->Interrupt(); }
  // cursor/base/cursor.h:210 (basic_cursor.h: every child is handed to the transformer and replaced by its result).
  // A transformed child is no longer known to be a GPU cursor: its rows then arrive through Next().
  virtual void ApplyToChildren(CursorTransformer* transformer) { child_.reset(transformer->Transform(child_.release())); DropFusion(); }
 protected:
  virtual FailureOrVoid Run(DeviceTable* result) {
    FailureOr<Session*> sr = Session::Get();
    PROPAGATE_ON_FAILURE(sr);
    Session* s = sr.get();
    DeviceTable in;
    std::unique_ptr<Block> keepalive;
    PROPAGATE_ON_FAILURE(MaterializeOnDevice(child_.get(), &in, &keepalive));
    vector<ssb_column> kc;
    vector<int32_t> desc;
    for (size_t k = 0; k < keys_.size(); ++k) {
      kc.push_back(in.columns[keys_[k].first].col);
      desc.push_back(keys_[k].second == DESCENDING ? 1 : 0);
    }
    DeviceBuffer perm;
    PROPAGATE_ON_FAILURE(perm.Allocate(static_cast<size_t>(in.rows) * 8 + 128));
    ssb_column dummy_col;
    memset(&dummy_col, 0, sizeof(dummy_col));
    int32_t dummy = 0;
    SSB_CALL(s, ssb_sort_permutation(s->ctx(), static_cast<int32_t>(kc.size()), kc.empty() ? &dummy_col : kc.data(),
                                     desc.empty() ? &dummy : desc.data(), in.rows, static_cast<int64_t*>(perm.get())),
             "sort");
    const int64_t out_rows = (limit_ >= 0 && limit_ < static_cast<int64>(in.rows)) ? static_cast<int64_t>(limit_) : static_cast<int64_t>(in.rows);
    PROPAGATE_ON_FAILURE(result->Allocate(schema(), out_rows, /* force_nulls = */ true));
    for (size_t i = 0; i < projected_.size(); ++i) {
      const vector<int> pos(1, projected_[i]);
      PROPAGATE_ON_FAILURE(GatherColumns(s, in, pos, static_cast<const int64_t*>(perm.get()), out_rows, false, result, i));
      if (in.columns[projected_[i]].col.nulls == NULL) {
        SSB_CALL(s, ssb_memset(s->ctx(), result->columns[i].col.nulls, 0, static_cast<size_t>((out_rows + 31) / 32 + 1) * 4), "memset");
      }
    }
    result->rows = out_rows;
    SSB_CALL(s, ssb_ctx_sync(s->ctx()), "sync");
    return Success();
  }
 private:
  std::unique_ptr<Cursor> child_;
  vector<std::pair<int, ColumnOrder> > keys_;
  vector<int> projected_;
  int64 limit_;
};

// ------------------------------------------------------------------ ScanViewWithSelection
// scan_view.cc: a cursor over view[selection[i]]. The view is uploaded once, the selection vector is the index
// list of one gather per column (ssb_gather; a row may be selected any number of times).
class SelectionCursor : public GpuCursor {
 public:
  SelectionCursor(const View& view, rowcount_t row_count, const rowid_t* selection, BufferAllocator* allocator)
      : GpuCursor(view.schema(), allocator, "ScanViewWithSelection"), source_(new ViewCursor(view)),
        row_count_(row_count), selection_(selection) {}
  virtual CursorId GetCursorId() const { return VIEW; }
 protected:
  virtual FailureOrVoid Run(DeviceTable* result) {
    FailureOr<Session*> sr = Session::Get();
    PROPAGATE_ON_FAILURE(sr);
    Session* s = sr.get();
    DeviceTable in;
    std::unique_ptr<Block> keepalive;
    PROPAGATE_ON_FAILURE(MaterializeOnDevice(source_.get(), &in, &keepalive));
    const int64_t n = static_cast<int64_t>(row_count_);
    for (int64_t i = 0; i < n; ++i) {
      if (selection_[i] < 0 || selection_[i] >= static_cast<rowid_t>(in.rows)) {
        THROW(new Exception(ERROR_INVALID_ARGUMENT_VALUE, "selection vector points outside the view"));
      }
    }
    DeviceBuffer ids;
    PROPAGATE_ON_FAILURE(ids.Allocate(static_cast<size_t>(n) * 8 + 128));
    if (n > 0) SSB_CALL(s, ssb_memcpy_h2d(s->ctx(), ids.get(), selection_, static_cast<size_t>(n) * 8), "upload");
    PROPAGATE_ON_FAILURE(result->Allocate(schema(), n, /* force_nulls = */ true));
    for (int i = 0; i < schema().attribute_count(); ++i) {
      const vector<int> pos(1, i);
      PROPAGATE_ON_FAILURE(GatherColumns(s, in, pos, static_cast<const int64_t*>(ids.get()), n, false, result, static_cast<size_t>(i)));
      if (in.columns[i].col.nulls == NULL) {
        SSB_CALL(s, ssb_memset(s->ctx(), result->columns[i].col.nulls, 0, static_cast<size_t>((n + 31) / 32 + 1) * 4), "memset");
      }
    }
    result->rows = n;
    SSB_CALL(s, ssb_ctx_sync(s->ctx()), "sync");
    return Success();
  }
 private:
  std::unique_ptr<Cursor> source_;
  rowcount_t row_count_;
  const rowid_t* selection_;
};

class SelectionOperation : public BasicOperation {
 public:
  SelectionOperation(const View& view, rowcount_t row_count, const rowid_t* selection)
      : view_(view), row_count_(row_count), selection_(selection) {}
  virtual FailureOrOwned<Cursor> CreateCursor() const {
    return Success(static_cast<Cursor*>(new SelectionCursor(view_, row_count_, selection_, buffer_allocator())));
  }
 protected:
  virtual string DebugName() const { return "ScanViewWithSelection"; }
 private:
  View view_;
  rowcount_t row_count_;
  const rowid_t* selection_;
};

class SortOperation : public BasicOperation {
 public:
  SortOperation(const SortOrder* order, const SingleSourceProjector* projector, Operation* child)
      : BasicOperation(child), order_(order), projector_(projector) {}
  virtual FailureOrOwned<Cursor> CreateCursor() const {
    FailureOrOwned<Cursor> child_cursor = child()->CreateCursor();
    PROPAGATE_ON_FAILURE(child_cursor);
    const TupleSchema& cs = child_cursor->schema();
    vector<std::pair<int, ColumnOrder> > keys;
    PROPAGATE_ON_FAILURE(order_->Bind(cs, &keys));
    vector<int> projected;
    TupleSchema result;
    if (projector_) {
      FailureOrOwned<const BoundSingleSourceProjector> proj = projector_->Bind(cs);
      PROPAGATE_ON_FAILURE(proj);
      result = proj->result_schema();
      for (int i = 0; i < result.attribute_count(); ++i) projected.push_back(proj->source_attribute_position(i));
    } else {
      result = cs;
      for (int i = 0; i < cs.attribute_count(); ++i) projected.push_back(i);
    }
    return Success(static_cast<Cursor*>(new SortCursor(result, buffer_allocator(), child_cursor.release(), keys, projected)));
  }
 protected:
  virtual string DebugName() const { return "Sort"; }
 private:
  std::unique_ptr<const SortOrder> order_;
  std::unique_ptr<const SingleSourceProjector> projector_;
};

// cursor/core/merge_union_all.cc:296-311
FailureOr<TupleSchema> MergeResultSchema(const vector<Cursor*>& inputs) {
  TupleSchema result;
  const TupleSchema& first = inputs[0]->schema();
  for (size_t j = 1; j < inputs.size(); ++j) {
    const TupleSchema& other = inputs[j]->schema();
    bool same = other.attribute_count() == first.attribute_count();
    for (int i = 0; same && i < first.attribute_count(); ++i) same = other.attribute(i).type() == first.attribute(i).type();
    if (!same) {   // the reference CHECK-fails here (EqualByType)
      THROW(new Exception(ERROR_ATTRIBUTE_TYPE_MISMATCH, "MergeUnionAll inputs differ in column count or types: (" +
                                                             first.GetHumanReadableSpecification() + ") vs (" +
                                                             other.GetHumanReadableSpecification() + ")"));
    }
  }
  for (int i = 0; i < first.attribute_count(); ++i) {
    bool nullable = false;
    for (size_t j = 0; j < inputs.size(); ++j) nullable = nullable || inputs[j]->schema().attribute(i).is_nullable();
    result.add_attribute(Attribute(first.attribute(i).name(), first.attribute(i).type(), nullable ? NULLABLE : NOT_NULLABLE));
  }
  return Success(result);
}

FailureOrOwned<Cursor> MakeMergeUnionAll(const vector<std::pair<int, ColumnOrder> >& keys, const TupleSchema& schema,
                                         vector<Cursor*>* inputs, BufferAllocator* allocator) {
  vector<int> all;
  for (int i = 0; i < schema.attribute_count(); ++i) all.push_back(i);
  Cursor* concat = new ConcatCursor(schema, allocator, inputs);
  return Success(static_cast<Cursor*>(new SortCursor(schema, allocator, concat, keys, all)));
}

class MergeUnionAllOperation : public BasicOperation {
 public:
  MergeUnionAllOperation(const SortOrder* order, const vector<Operation*>& inputs) : BasicOperation(inputs), order_(order) {}
  virtual FailureOrOwned<Cursor> CreateCursor() const {
    vector<Cursor*> inputs;
    struct Deleter { vector<Cursor*>* v; ~Deleter() { for (size_t i = 0; i < v->size(); ++i) delete (*v)[i]; } } deleter = {&inputs};
    for (size_t i = 0; i < children_count(); ++i) {
      FailureOrOwned<Cursor> c = child_at(i)->CreateCursor();
      PROPAGATE_ON_FAILURE(c);
      inputs.push_back(c.release());
    }
    FailureOr<TupleSchema> schema = MergeResultSchema(inputs);
    PROPAGATE_ON_FAILURE(schema);
    vector<std::pair<int, ColumnOrder> > keys;
    PROPAGATE_ON_FAILURE(order_->Bind(schema.get(), &keys));
    return MakeMergeUnionAll(keys, schema.get(), &inputs, buffer_allocator());
  }
 protected:
  virtual string DebugName() const { return "MergeUnionAll"; }
 private:
  std::unique_ptr<const SortOrder> order_;
};

// ------------------------------------------------------------------ Limit / Coalesce
// cursor/core/limit.cc: skips `offset` rows of the child, then passes at most `limit` rows on. No computation: the views the
// child hands out are sliced (a GPU child's result is already on the host when Next() returns it).
class LimitCursor : public Cursor {
 public:
  LimitCursor(rowcount_t offset, rowcount_t limit, Cursor* child)
      : child_(child), view_(child->schema()), to_skip_(offset), to_pass_(limit) {}
  virtual const TupleSchema& schema() const { return child_->schema(); }
  virtual ResultView Next(rowcount_t max_row_count) {
    for (;;) {
      if (to_pass_ == 0) return ResultView::EOS();
      ResultView r = child_->Next(to_skip_ > 0 ? max_row_count : std::min(max_row_count, to_pass_));
      if (!r.has_data()) return r;
      rowcount_t n = r.view().row_count(), first = 0;
      if (to_skip_ > 0) {
        const rowcount_t skipped = std::min(to_skip_, n);
        to_skip_ -= skipped;
        first = skipped;
        n -= skipped;
        if (n == 0) continue;
      }
      n = std::min(n, to_pass_);
      to_pass_ -= n;
      view_.ResetFromSubRange(r.view(), first, n);
      return ResultView::Success(&view_);
    }
  }
  virtual void Interrupt() { child_->Interrupt(); }
  virtual bool IsWaitingOnBarrierSupported() const { return child_->IsWaitingOnBarrierSupported(); }
  virtual void ApplyToChildren(CursorTransformer* transformer) { child_.reset(transformer->Transform(child_.release())); }
  virtual void AppendDebugDescription(string* target) const {
    target->append("LimitCursor(");
    child_->AppendDebugDescription(target);
    target->append(")");
  }
 private:
  std::unique_ptr<Cursor> child_;
  View view_;
  rowcount_t to_skip_, to_pass_;
};

class LimitOperation : public BasicOperation {
 public:
  LimitOperation(rowcount_t offset, rowcount_t limit, Operation* child) : BasicOperation(child), offset_(offset), limit_(limit) {}
  virtual FailureOrOwned<Cursor> CreateCursor() const {
    FailureOrOwned<Cursor> c = child()->CreateCursor();
    PROPAGATE_ON_FAILURE(c);
    return Success(static_cast<Cursor*>(new LimitCursor(offset_, limit_, c.release())));
  }
 protected:
  virtual string DebugName() const { return "Limit"; }
 private:
  rowcount_t offset_, limit_;
};

// cursor/core/coalesce.cc: the columns of several inputs side by side (distinct attribute names); the stream ends when the
// shortest input does. On the device the result references the inputs' columns: nothing is copied.
class CoalesceCursor : public GpuCursor {
 public:
  CoalesceCursor(const TupleSchema& schema, BufferAllocator* allocator, vector<Cursor*>* inputs)
      : GpuCursor(schema, allocator, "CoalesceCursor") {
    for (size_t i = 0; i < inputs->size(); ++i) inputs_.push_back(std::unique_ptr<Cursor>((*inputs)[i]));
    inputs->clear();
  }
  virtual void Interrupt() { GpuCursor::Interrupt(); for (size_t i = 0; i < inputs_.size(); ++i) inputs_[i]->Interrupt(); }
  virtual void ApplyToChildren(CursorTransformer* transformer) {
    for (size_t i = 0; i < inputs_.size(); ++i) inputs_[i].reset(transformer->Transform(inputs_[i].release()));
  }
 protected:
  virtual FailureOrVoid Run(DeviceTable* result) {
    parts_.resize(inputs_.size());
    keep_.resize(inputs_.size());
    result->schema = schema();
    result->columns.clear();
    result->rows = 0;
    for (size_t i = 0; i < inputs_.size(); ++i) {
      PROPAGATE_ON_FAILURE(MaterializeOnDevice(inputs_[i].get(), &parts_[i], &keep_[i]));
      for (size_t c = 0; c < parts_[i].columns.size(); ++c) result->columns.push_back(parts_[i].columns[c]);
      // coalesce.cc: the stream ends with its shortest input
      result->rows = i == 0 ? parts_[i].rows : std::min(result->rows, parts_[i].rows);
    }
    return Success();
  }
 private:
  vector<std::unique_ptr<Cursor> > inputs_;
  vector<DeviceTable> parts_;                  // keep the inputs' columns alive
  vector<std::unique_ptr<Block> > keep_;
};

FailureOr<TupleSchema> CoalescedSchema(const vector<Cursor*>& inputs) {
  TupleSchema schema;
  for (size_t i = 0; i < inputs.size(); ++i) {
    for (int c = 0; c < inputs[i]->schema().attribute_count(); ++c) {
      if (!schema.add_attribute(inputs[i]->schema().attribute(c))) {
        THROW(new Exception(ERROR_ATTRIBUTE_EXISTS, "Can't coalesce, ambiguous attribute name: " + inputs[i]->schema().attribute(c).name()));
      }
    }
  }
  return Success(schema);
}

class CoalesceOperation : public BasicOperation {
 public:
  explicit CoalesceOperation(const vector<Operation*>& children) : BasicOperation(children) {}
  virtual FailureOrOwned<Cursor> CreateCursor() const {
    vector<Cursor*> inputs;
    struct Deleter { vector<Cursor*>* v; ~Deleter() { for (size_t i = 0; i < v->size(); ++i) delete (*v)[i]; } } deleter = {&inputs};
    for (size_t i = 0; i < children_count(); ++i) {
      FailureOrOwned<Cursor> c = child_at(i)->CreateCursor();
      PROPAGATE_ON_FAILURE(c);
      inputs.push_back(c.release());
    }
    FailureOr<TupleSchema> schema = CoalescedSchema(inputs);
    PROPAGATE_ON_FAILURE(schema);
    return Success(static_cast<Cursor*>(new CoalesceCursor(schema.get(), buffer_allocator(), &inputs)));
  }
 protected:
  virtual string DebugName() const { return "Coalesce"; }
};

}  // namespace

Operation* Limit(rowcount_t offset, rowcount_t limit, Operation* child) { return new LimitOperation(offset, limit, child); }
Cursor* BoundLimit(rowcount_t offset, rowcount_t limit, Cursor* child) { return new LimitCursor(offset, limit, child); }
Operation* Coalesce(const vector<Operation*>& children) { return new CoalesceOperation(children); }
FailureOrOwned<Cursor> BoundCoalesce(const vector<Cursor*>& children) {
  vector<Cursor*> inputs(children);
  FailureOr<TupleSchema> schema = CoalescedSchema(inputs);
  if (schema.is_failure()) {
    for (size_t i = 0; i < inputs.size(); ++i) delete inputs[i];
    return Failure(schema.release_exception());
  }
  return Success(static_cast<Cursor*>(new CoalesceCursor(schema.get(), HeapBufferAllocator::Get(), &inputs)));
}

Operation* MergeUnionAll(const SortOrder* sort_order, const vector<Operation*>& inputs) {
  std::unique_ptr<const SortOrder> order(sort_order);
  if (inputs.empty()) return ScanView(View(TupleSchema()));   // merge_union_all.cc:352-356: Generate(0)
  if (inputs.size() == 1) return inputs[0];
  return new MergeUnionAllOperation(order.release(), inputs);
}

FailureOrOwned<Cursor> BoundMergeUnionAll(const BoundSortOrder* sort_order, vector<Cursor*> inputs, BufferAllocator* buffer_allocator) {
  std::unique_ptr<const BoundSortOrder> order(sort_order);
  if (inputs.empty()) THROW(new Exception(ERROR_INVALID_ARGUMENT_VALUE, "BoundMergeUnionAll needs at least one input"));
  FailureOr<TupleSchema> schema = MergeResultSchema(inputs);
  if (schema.is_failure()) {
    for (size_t i = 0; i < inputs.size(); ++i) delete inputs[i];
    return Failure(schema.release_exception());
  }
  vector<std::pair<int, ColumnOrder> > keys;
  for (int i = 0; i < order->schema().attribute_count(); ++i) {
    keys.push_back(std::make_pair(order->projector().source_attribute_position(i), order->column_order(i)));
  }
  return MakeMergeUnionAll(keys, schema.get(), &inputs, buffer_allocator ? buffer_allocator : HeapBufferAllocator::Get());
}

Operation* AggregateClusters(const SingleSourceProjector* clustered_by_columns, const AggregationSpecification* aggregation,
                             Operation* child) {
  return new AggregateClustersOperation(clustered_by_columns, aggregation, child);
}
Operation* AggregateClustersWithSpecifiedOutputBlockSize(const SingleSourceProjector* clustered_by_columns,
                                                         const AggregationSpecification* aggregation, rowcount_t,
                                                         Operation* child) {
  return new AggregateClustersOperation(clustered_by_columns, aggregation, child);   // the whole result is one device table
}

FailureOrOwned<Cursor> BoundAggregateClusters(const BoundSingleSourceProjector* group_by, Aggregator* aggregator,
                                              BufferAllocator* allocator, Cursor* child) {
  std::unique_ptr<const BoundSingleSourceProjector> proj(group_by);
  std::unique_ptr<Aggregator> agg(aggregator);
  std::unique_ptr<Cursor> child_cursor(child);
  TupleSchema result;
  vector<int> keys;
  for (int i = 0; i < proj->result_schema().attribute_count(); ++i) {
    keys.push_back(proj->source_attribute_position(i));
    result.add_attribute(proj->result_schema().attribute(i));
  }
  for (int i = 0; i < agg->schema().attribute_count(); ++i) {
    if (!result.add_attribute(agg->schema().attribute(i))) {
      THROW(new Exception(ERROR_ATTRIBUTE_EXISTS, "Duplicate attribute name '" + agg->schema().attribute(i).name() + "' in result schema"));
    }
  }
  return Success(static_cast<Cursor*>(new ClustersCursor(result, allocator ? allocator : HeapBufferAllocator::Get(), child_cursor.release(),
                                                         keys, agg->impl()->aggs)));
}

namespace internal {

FailureOrVoid DescribeAny(const Operation* op, RowwisePlan* plan) {
  Exception* error = NULL;
  if (op->DescribeRowwise(plan, &error)) {
    if (error != NULL) return Failure(error);
    return Success();
  }
  FailureOrOwned<Cursor> c = op->CreateCursor();
  PROPAGATE_ON_FAILURE(c);
  plan->source.reset(c.release());
  plan->base_schema = plan->source->schema();
  plan->base = View(plan->base_schema);
  plan->schema = plan->base_schema;
  plan->outputs.clear();
  for (int i = 0; i < plan->schema.attribute_count(); ++i) plan->outputs.push_back(MakeInputNode(plan->schema, i));
  plan->predicate.reset();
  return Success();
}

}  // namespace internal

// ------------------------------------------------------------------ Bound* factories
// The cursors above already work on bound objects (expression nodes over the child's schema,
// column positions, bound aggregations); the Bound* factories of the reference hand exactly
// those over, so each one only repackages its arguments.
Cursor* BoundScanView(const View& view) { return new ViewCursor(view); }

namespace {
RowwisePlan PlanOverCursor(Cursor* child) {
  RowwisePlan plan;
  plan.source.reset(child);
  plan.base_schema = child->schema();
  plan.base = View(plan.base_schema);
  plan.schema = plan.base_schema;
  for (int i = 0; i < plan.schema.attribute_count(); ++i) plan.outputs.push_back(MakeInputNode(plan.schema, i));
  return plan;
}
}  // namespace

FailureOrOwned<Cursor> BoundCompute(BoundExpressionTree* computation, BufferAllocator* allocator, rowcount_t,
                                    Cursor* child) {
  std::unique_ptr<BoundExpressionTree> tree(computation);
  RowwisePlan plan = PlanOverCursor(child);
  plan.outputs.clear();
  for (int i = 0; i < tree->root()->column_count(); ++i) {
    plan.outputs.push_back(GuardSignaling(tree->root()->node(i), vector<NodePtr>(), NodePtr()));
  }
  plan.schema = tree->result_schema();
  return Success(static_cast<Cursor*>(new RowwiseCursor(plan, allocator, COMPUTE)));
}

FailureOrOwned<Cursor> BoundFilter(BoundExpressionTree* predicate, const BoundSingleSourceProjector* projector,
                                   BufferAllocator* buffer_allocator, Cursor* child_cursor) {
  std::unique_ptr<BoundExpressionTree> tree(predicate);
  std::unique_ptr<const BoundSingleSourceProjector> proj(projector);
  std::unique_ptr<Cursor> child(child_cursor);
  // cursor/core/filter.cc:79-87
  if (tree->root()->column_count() != 1) {
    THROW(new Exception(ERROR_ATTRIBUTE_COUNT_MISMATCH, "Predicate has to return exactly one column in (" +
                                                            tree->result_schema().GetHumanReadableSpecification() + ")"));
  }
  if (tree->result_schema().attribute(0).type() != BOOL) {
    THROW(new Exception(ERROR_ATTRIBUTE_TYPE_MISMATCH, "Predicate has to return a BOOL column in (" +
                                                           tree->result_schema().GetHumanReadableSpecification() + ")"));
  }
  RowwisePlan plan = PlanOverCursor(child.release());
  vector<NodePtr> outs;
  for (int i = 0; i < proj->result_schema().attribute_count(); ++i) outs.push_back(plan.outputs[proj->source_attribute_position(i)]);
  plan.outputs = outs;
  plan.schema = proj->result_schema();
  plan.predicate = GuardSignaling(tree->root()->node(0), vector<NodePtr>(), NodePtr());
  return Success(static_cast<Cursor*>(new RowwiseCursor(plan, buffer_allocator, FILTER)));
}

Cursor* BoundProject(const BoundSingleSourceProjector* projector, Cursor* child) {
  std::unique_ptr<const BoundSingleSourceProjector> proj(projector);
  RowwisePlan plan = PlanOverCursor(child);
  vector<NodePtr> outs;
  for (int i = 0; i < proj->result_schema().attribute_count(); ++i) outs.push_back(plan.outputs[proj->source_attribute_position(i)]);
  plan.outputs = outs;
  plan.schema = proj->result_schema();
  return new RowwiseCursor(plan, HeapBufferAllocator::Get(), PROJECT);
}

Aggregator::~Aggregator() { delete impl_; }

FailureOrOwned<Aggregator> Aggregator::Create(const AggregationSpecification& aggregation_specification,
                                              const TupleSchema& input_schema, BufferAllocator*,
                                              rowcount_t result_initial_row_capacity) {
  std::unique_ptr<Aggregator> a(new Aggregator);
  a->impl_ = new Impl;
  a->capacity_ = result_initial_row_capacity;
  PROPAGATE_ON_FAILURE(BindAggregations(aggregation_specification, input_schema, &a->impl_->aggs, &a->schema_));
  return Success(a.release());
}

FailureOrOwned<Cursor> BoundGroupAggregate(const BoundSingleSourceProjector* group_by, Aggregator* aggregator,
                                           BufferAllocator* allocator, BufferAllocator* original_allocator, bool,
                                           Cursor* child) {
  std::unique_ptr<const BoundSingleSourceProjector> proj(group_by);
  std::unique_ptr<Aggregator> agg(aggregator);
  std::unique_ptr<Cursor> child_cursor(child);
  // the reference takes ownership of `allocator` unless it is the original one (aggregate.h:258-260)
  std::unique_ptr<BufferAllocator> owned(allocator != original_allocator && allocator != HeapBufferAllocator::Get() ? allocator : NULL);
  TupleSchema result;
  vector<int> keys;
  for (int i = 0; i < proj->result_schema().attribute_count(); ++i) {
    keys.push_back(proj->source_attribute_position(i));
    result.add_attribute(proj->result_schema().attribute(i));
  }
  for (int i = 0; i < agg->schema().attribute_count(); ++i) {
    if (!result.add_attribute(agg->schema().attribute(i))) {
      THROW(new Exception(ERROR_ATTRIBUTE_EXISTS, "Duplicate attribute name '" + agg->schema().attribute(i).name() + "' in result schema"));
    }
  }
  BufferAllocator* use = original_allocator ? original_allocator : HeapBufferAllocator::Get();
  return Success(static_cast<Cursor*>(new GroupCursor(result, use, child_cursor.release(), keys, agg->impl()->aggs,
                                                      static_cast<size_t>(agg->initial_row_capacity()), false)));
}

// aggregate.h:309-336: the reference's aggregation for inputs and DISTINCT sets larger than memory (it sorts and spills
// to temporary files under `temporary_directory_prefix`). Here the groups are aggregated in HBM whatever the quota says
// (host tables stream through in chunks), DISTINCT aggregates run as passes (GroupCursor::RunDistinct), nothing spills:
// the result is the exact aggregation, every key once.
Operation* HybridGroupAggregate(const SingleSourceProjector* group_by_columns, const AggregationSpecification* aggregation_specification,
                                size_t, StringPiece, Operation* child) {
  return new GroupAggregateOperation(group_by_columns, const_cast<AggregationSpecification*>(aggregation_specification), NULL, child,
                                     /* no result budget */ true);
}
FailureOrOwned<Cursor> BoundHybridGroupAggregate(const SingleSourceProjector* group_by_columns,
                                                 const AggregationSpecification& aggregation_specification, StringPiece,
                                                 BufferAllocator* allocator, size_t, const HybridGroupDebugOptions* debug_options,
                                                 Cursor* child) {
  (void)debug_options;
  std::unique_ptr<const SingleSourceProjector> group_by(group_by_columns);
  std::unique_ptr<Cursor> child_cursor(child);
  FailureOrOwned<const BoundSingleSourceProjector> proj = group_by->Bind(child_cursor->schema());
  PROPAGATE_ON_FAILURE(proj);
  TupleSchema result;
  vector<int> keys;
  for (int i = 0; i < proj->result_schema().attribute_count(); ++i) {
    keys.push_back(proj->source_attribute_position(i));
    result.add_attribute(proj->result_schema().attribute(i));
  }
  vector<BoundAggregation> aggs;
  PROPAGATE_ON_FAILURE(BindAggregations(aggregation_specification, child_cursor->schema(), &aggs, &result));
  return Success(static_cast<Cursor*>(new GroupCursor(result, allocator ? allocator : HeapBufferAllocator::Get(), child_cursor.release(),
                                                      keys, aggs, 0, false)));
}

Cursor* BoundScalarAggregate(Aggregator* aggregator, Cursor* child) {
  std::unique_ptr<Aggregator> agg(aggregator);
  return new GroupCursor(agg->schema(), HeapBufferAllocator::Get(), child, vector<int>(), agg->impl()->aggs, 0, true);
}

FailureOrOwned<const BoundSortOrder> SortOrder::Bind(const TupleSchema& source_schema) const {
  vector<std::pair<int, ColumnOrder> > keys;
  PROPAGATE_ON_FAILURE(Bind(source_schema, &keys));
  std::unique_ptr<BoundSingleSourceProjector> proj(new BoundSingleSourceProjector(source_schema));
  vector<ColumnOrder> orders;
  for (size_t i = 0; i < keys.size(); ++i) {
    if (!proj->Add(keys[i].first)) THROW(new Exception(ERROR_ATTRIBUTE_EXISTS, "Duplicate attribute in the sort order"));
    orders.push_back(keys[i].second);
  }
  return Success(static_cast<const BoundSortOrder*>(new BoundSortOrder(proj.release(), orders)));
}

FailureOrOwned<Cursor> BoundSort(const BoundSortOrder* sort_order, const BoundSingleSourceProjector* result_projector,
                                 size_t, StringPiece, BufferAllocator* allocator, Cursor* child_cursor) {
  std::unique_ptr<const BoundSortOrder> order(sort_order);
  std::unique_ptr<const BoundSingleSourceProjector> proj(result_projector);
  std::unique_ptr<Cursor> child(child_cursor);
  vector<std::pair<int, ColumnOrder> > keys;
  for (int i = 0; i < order->schema().attribute_count(); ++i) {
    keys.push_back(std::make_pair(order->projector().source_attribute_position(i), order->column_order(i)));
  }
  vector<int> projected;
  TupleSchema result;
  if (proj) {
    result = proj->result_schema();
    for (int i = 0; i < result.attribute_count(); ++i) projected.push_back(proj->source_attribute_position(i));
  } else {
    result = child->schema();
    for (int i = 0; i < result.attribute_count(); ++i) projected.push_back(i);
  }
  return Success(static_cast<Cursor*>(new SortCursor(result, allocator, child.release(), keys, projected)));
}

Operation* BestEffortGroupAggregate(const SingleSourceProjector* group_by, AggregationSpecification* aggregation,
                                    GroupAggregateOptions* options, Operation* child) {
  return new GroupAggregateOperation(group_by, aggregation, options, child, /* best effort */ true);
}
Operation* SortWithTempDirPrefix(const SortOrder* sort_order, const SingleSourceProjector* result_projector,
                                 size_t memory_limit, StringPiece, Operation* child) {
  return Sort(sort_order, result_projector, memory_limit, child);
}

// ------------------------------------------------------------------ factories
Operation* ScanView(const View& view) { return new ScanViewOperation(view); }
Operation* ScanViewWithSelection(const View& view, const rowcount_t row_count, const rowid_t* selection_vector, rowcount_t) {
  return new SelectionOperation(view, row_count, selection_vector);
}
FailureOrOwned<Cursor> BoundScanViewWithSelection(const View& view, const rowcount_t row_count, const rowid_t* selection_vector,
                                                  BufferAllocator* allocator, rowcount_t) {
  return Success(static_cast<Cursor*>(new SelectionCursor(view, row_count, selection_vector, allocator)));
}
Operation* Compute(const Expression* computation, Operation* child) { return new ComputeOperation(computation, child); }
Operation* Filter(const Expression* predicate, const SingleSourceProjector* projector, Operation* child) {
  return new FilterOperation(predicate, projector, child);
}
Operation* Project(const SingleSourceProjector* projector, Operation* child) { return new ProjectOperation(projector, child); }
Operation* GroupAggregate(const SingleSourceProjector* group_by, AggregationSpecification* aggregation,
                          GroupAggregateOptions* options, Operation* child) {
  return new GroupAggregateOperation(group_by, aggregation, options, child);
}
Operation* ScalarAggregate(AggregationSpecification* aggregation, Operation* child) {
  return new GroupAggregateOperation(NULL, aggregation, NULL, child);
}
Operation* Sort(const SortOrder* sort_order, const SingleSourceProjector* result_projector, size_t, Operation* child) {
  return new SortOperation(sort_order, result_projector, child);
}

// sort.cc:857-1017. The reference layers Compute (upper-cased copies of case-insensitive STRING keys), Sort and
// Limit; without STRING columns that is a sort by attribute name whose cursor returns the first `limit` rows.
FailureOrOwned<Cursor> BoundExtendedSort(const ExtendedSortSpecification* sort_specification,
                                         const BoundSingleSourceProjector* result_projector, size_t, StringPiece,
                                         BufferAllocator* allocator, rowcount_t, Cursor* child_cursor) {
  std::unique_ptr<const ExtendedSortSpecification> spec(sort_specification);
  std::unique_ptr<const BoundSingleSourceProjector> proj(result_projector);
  std::unique_ptr<Cursor> child(child_cursor);
  const TupleSchema& cs = child->schema();
  vector<std::pair<int, ColumnOrder> > keys;
  vector<string> seen;
  for (int i = 0; i < spec->keys_size(); ++i) {
    const string& name = spec->keys(i).attribute_name();
    if (std::find(seen.begin(), seen.end(), name) != seen.end()) {
      THROW(new Exception(ERROR_INVALID_ARGUMENT_VALUE, "Duplicate case sensitive key: " + name + " column in schema (" +
                                                            cs.GetHumanReadableSpecification() + ")"));
    }
    seen.push_back(name);
    const int pos = cs.LookupAttributePosition(name);
    if (pos < 0) {   // the reference CHECK-fails here (TupleSchema::LookupAttribute)
      THROW(new Exception(ERROR_ATTRIBUTE_MISSING, "No attribute '" + name + "' in schema (" + cs.GetHumanReadableSpecification() + ")"));
    }
    const DataType t = cs.attribute(pos).type();
    if (t == STRING && !spec->keys(i).case_sensitive()) {   // sort.cc:886-931 sorts by an upper-cased copy of the key
      THROW(new Exception(ERROR_NOT_IMPLEMENTED, "case-insensitive STRING sort keys are not on the B200 hot path (SURVEY 8f)"));
    }
    keys.push_back(std::make_pair(pos, spec->keys(i).column_order()));
  }
  vector<int> projected;
  TupleSchema result;
  if (proj) {
    result = proj->result_schema();
    for (int i = 0; i < result.attribute_count(); ++i) projected.push_back(proj->source_attribute_position(i));
  } else {
    result = cs;
    for (int i = 0; i < cs.attribute_count(); ++i) projected.push_back(i);
  }
  const int64 limit = spec->has_limit() ? static_cast<int64>(std::min<uint64>(spec->limit(), static_cast<uint64>(1) << 62)) : -1;
  return Success(static_cast<Cursor*>(new SortCursor(result, allocator, child.release(), keys, projected, limit)));
}

namespace {
class ExtendedSortOperation : public BasicOperation {
 public:
  ExtendedSortOperation(const ExtendedSortSpecification* spec, const SingleSourceProjector* projector, Operation* child)
      : BasicOperation(child), spec_(spec), projector_(projector) {}
  virtual FailureOrOwned<Cursor> CreateCursor() const {
    FailureOrOwned<Cursor> child_cursor = child()->CreateCursor();
    PROPAGATE_ON_FAILURE(child_cursor);
    std::unique_ptr<Cursor> cursor(child_cursor.release());
    std::unique_ptr<const BoundSingleSourceProjector> bound;
    if (projector_) {
      FailureOrOwned<const BoundSingleSourceProjector> proj = projector_->Bind(cursor->schema());
      PROPAGATE_ON_FAILURE(proj);
      bound.reset(proj.release());
    }
    return BoundExtendedSort(new ExtendedSortSpecification(*spec_), bound.release(), 0, "", buffer_allocator(),
                             Cursor::kDefaultRowCount, cursor.release());
  }
 protected:
  virtual string DebugName() const { return "ExtendedSort"; }
 private:
  std::unique_ptr<const ExtendedSortSpecification> spec_;
  std::unique_ptr<const SingleSourceProjector> projector_;
};
}  // namespace

Operation* ExtendedSort(const ExtendedSortSpecification* specification, const SingleSourceProjector* result_projector,
                        size_t, Operation* child) {
  return new ExtendedSortOperation(specification, result_projector, child);
}

SortOrder::~SortOrder() { for (size_t i = 0; i < keys_.size(); ++i) delete keys_[i].first; }
FailureOrVoid SortOrder::Bind(const TupleSchema& schema, vector<std::pair<int, ColumnOrder> >* keys) const {
  // the keys form one projection (sort.cc:74-98, a CompoundSingleSourceProjector): a column named twice is
  // a duplicate attribute of its result schema
  vector<string> seen;
  for (size_t i = 0; i < keys_.size(); ++i) {
    FailureOrOwned<const BoundSingleSourceProjector> p = keys_[i].first->Bind(schema);
    PROPAGATE_ON_FAILURE(p);
    for (int c = 0; c < p->result_schema().attribute_count(); ++c) {
      const string& name = p->result_schema().attribute(c).name();
      if (std::find(seen.begin(), seen.end(), name) != seen.end()) {
        THROW(new Exception(ERROR_ATTRIBUTE_EXISTS, "Duplicate attribute name '" + name + "' in result schema"));
      }
      seen.push_back(name);
      keys->push_back(std::make_pair(p->source_attribute_position(c), keys_[i].second));
    }
  }
  return Success();
}

// ------------------------------------------------------------------ HashJoinOperation
HashJoinOperation::HashJoinOperation(JoinType join_type, const SingleSourceProjector* lhs_key_selector,
                                     const SingleSourceProjector* rhs_key_selector,
                                     const MultiSourceProjector* result_projector, KeyUniqueness rhs_key_uniqueness,
                                     Operation* lhs_child, Operation* rhs_child)
    : BasicOperation(lhs_child, rhs_child), join_type_(join_type), lhs_key_selector_(lhs_key_selector),
      rhs_key_selector_(rhs_key_selector), result_projector_(result_projector),
      rhs_key_uniqueness_(rhs_key_uniqueness) {}
HashJoinOperation::~HashJoinOperation() {}

FailureOrOwned<Cursor> HashJoinOperation::CreateCursor() const {
  FailureOrOwned<Cursor> lhs = child_at(0)->CreateCursor();
  PROPAGATE_ON_FAILURE(lhs);
  FailureOrOwned<Cursor> rhs = child_at(1)->CreateCursor();
  PROPAGATE_ON_FAILURE(rhs);
  FailureOrOwned<const BoundSingleSourceProjector> lk = lhs_key_selector_->Bind(lhs->schema());
  PROPAGATE_ON_FAILURE(lk);
  FailureOrOwned<const BoundSingleSourceProjector> rk = rhs_key_selector_->Bind(rhs->schema());
  PROPAGATE_ON_FAILURE(rk);
  {
    // key types must agree, except that integer keys of different widths / signedness are
    // compared by value (operators::Equal, row_hash_set.cc key comparators)
    bool ok = lk->result_schema().attribute_count() == rk->result_schema().attribute_count();
    for (int i = 0; ok && i < lk->result_schema().attribute_count(); ++i) {
      const DataType a = lk->result_schema().attribute(i).type(), b = rk->result_schema().attribute(i).type();
      ok = a == b || (GetTypeInfo(a).is_integer() && GetTypeInfo(b).is_integer());
    }
    if (!ok) {
      THROW(new Exception(ERROR_ATTRIBUTE_TYPE_MISMATCH, "Hash join key columns differ in number or type: (" +
                                                             lk->result_schema().GetHumanReadableSpecification() + ") vs (" +
                                                             rk->result_schema().GetHumanReadableSpecification() + ")"));
    }
  }
  vector<int> lkeys, rkeys;
  for (int i = 0; i < lk->result_schema().attribute_count(); ++i) {
    lkeys.push_back(lk->source_attribute_position(i));
    rkeys.push_back(rk->source_attribute_position(i));
  }
  // LEFT_OUTER: the rhs columns become nullable in the result (hash_join.h:37-38)
  TupleSchema rhs_schema;
  for (int i = 0; i < rhs->schema().attribute_count(); ++i) {
    const Attribute& a = rhs->schema().attribute(i);
    rhs_schema.add_attribute(Attribute(a.name(), a.type(), join_type_ == LEFT_OUTER ? NULLABLE : a.nullability()));
  }
  vector<const TupleSchema*> sources;
  sources.push_back(&lhs->schema());
  sources.push_back(&rhs_schema);
  FailureOrOwned<const BoundMultiSourceProjector> proj = result_projector_->Bind(sources);
  PROPAGATE_ON_FAILURE(proj);
  const TupleSchema result = proj->result_schema();
  return Success(static_cast<Cursor*>(new HashJoinCursor(result, buffer_allocator(), lhs.release(), rhs.release(), join_type_,
                                                         rhs_key_uniqueness_, lkeys, rkeys, proj.release())));
}

// ------------------------------------------------------------------ Generate (cursor/core/generate.cc)
Operation* Generate(rowcount_t count) {
  View v((TupleSchema()));
  v.set_row_count(count);
  return ScanView(v);
}
FailureOrOwned<Cursor> BoundGenerate(rowcount_t count) {
  View v((TupleSchema()));
  v.set_row_count(count);
  return Success(BoundScanView(v));
}

// ------------------------------------------------------------------ Table
Table::Table(const TupleSchema& schema, BufferAllocator* allocator)
    : block_(new Block(schema, allocator)), view_(schema), arena_(allocator, 4096, 1 << 20) {}
Table::~Table() {}
bool Table::ReserveRowCapacity(rowcount_t needed) {
  if (needed <= block_->row_capacity()) return true;
  rowcount_t cap = block_->row_capacity() ? block_->row_capacity() : 16;
  while (cap < needed) cap *= 2;
  const rowcount_t rows = view_.row_count();
  if (!block_->Reallocate(cap)) return false;
  view_.ResetFromSubRange(block_->view(), 0, rows);
  return true;
}
rowid_t Table::AddRow() {
  if (!ReserveRowCapacity(view_.row_count() + 1)) return -1;
  const rowcount_t r = view_.row_count();
  view_.ResetFromSubRange(block_->view(), 0, r + 1);
  return static_cast<rowid_t>(r);
}
rowcount_t Table::AppendView(const View& view) {
  const rowcount_t rows = view_.row_count(), n = view.row_count();
  if (!ReserveRowCapacity(rows + n)) return 0;
  const rowcount_t copied = ViewCopier(schema(), /* deep copy */ true).Copy(n, view, rows, block_.get());
  view_.ResetFromSubRange(block_->view(), 0, rows + copied);
  return copied;
}
FailureOrOwned<Cursor> Table::CreateCursor() const { return Success(static_cast<Cursor*>(new ViewCursor(view_))); }

TableRowWriter& TableRowWriter::AddRow() {
  row_ = table_->AddRow();
  col_ = 0;
  if (row_ < 0) ok_ = false;
  return *this;
}
void TableRowWriter::CheckSuccess() const {
  if (!ok_) { fprintf(stderr, "FATAL: TableRowWriter failed (out of memory)\n"); abort(); }
}

}  // namespace supersonic
