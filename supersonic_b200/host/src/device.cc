// device.cc -- bridge between the supersonic.h mirror and the C ABI of libssb200.so.
#include <stdio.h>

#include <map>
#include <mutex>

#include "internal.h"

namespace supersonic {
namespace internal {

// ------------------------------------------------------------------ session
FailureOr<Session*> Session::Get() {
  static std::mutex mu;
  static Session* session = NULL;
  std::lock_guard<std::mutex> lock(mu);
  if (session == NULL) {
    int device = 0;
    if (const char* env = getenv("SSB200_DEVICE")) device = atoi(env);
    else if (const char* lr = getenv("LOCAL_RANK")) device = atoi(lr);
    ssb_ctx* ctx = NULL;
    const int rc = ssb_ctx_create(device, &ctx);
    if (rc != 0 || ctx == NULL) {
      THROW(new Exception(ERROR_GENERAL_IO_ERROR,
                          "supersonic-b200: no usable B200 (sm_100a) device; this implementation has no CPU path"));
    }
    session = new Session(ctx, device);
  }
  return Success(session);
}

void Session::SyncAll() {
  ssb_ctx_sync(ctx_);
  for (int i = 0; i < 2; ++i) if (lanes_[i] != NULL) ssb_ctx_sync(lanes_[i]);
}

FailureOr<ssb_ctx*> Session::lane(int i) {
  if (lanes_[i] == NULL) {
    ssb_ctx* c = NULL;
    if (ssb_ctx_create(device_, &c) != 0 || c == NULL) {
      THROW(new Exception(ERROR_GENERAL_IO_ERROR, "supersonic-b200: cannot create a copy/compute lane context"));
    }
    lanes_[i] = c;
  }
  ssb_ctx* c = lanes_[i];
  return Success(c);
}

Exception* Session::Error(int code, const char* what) const { return ErrorOn(ctx_, code, what); }

Exception* Session::ErrorOn(ssb_ctx* ctx, int code, const char* what) {
  ReturnCode rc = ERROR_UNKNOWN_ERROR;
  switch (code) {
    case SSB_ERROR_MEMORY_EXCEEDED: rc = ERROR_MEMORY_EXCEEDED; break;
    case SSB_ERROR_NOT_IMPLEMENTED: rc = ERROR_NOT_IMPLEMENTED; break;
    case SSB_ERROR_EVALUATION_ERROR: rc = ERROR_EVALUATION_ERROR; break;
    case SSB_ERROR_INVALID_ARGUMENT_TYPE: rc = ERROR_INVALID_ARGUMENT_TYPE; break;
    case SSB_ERROR_INVALID_ARGUMENT_VALUE: rc = ERROR_INVALID_ARGUMENT_VALUE; break;
    default: break;
  }
  return new Exception(rc, string(what) + ": " + ssb_last_error(ctx));
}

#define SSB_CALL(session, call, what)                              \
  do {                                                             \
    const int rc_ = (call);                                        \
    if (rc_ != 0) THROW((session)->Error(rc_, what));              \
  } while (0)

// ------------------------------------------------------------------ device memory
namespace {
struct PoolState {
  std::mutex mu;
  std::multimap<size_t, void*> free_blocks[2];
  // a block was released since the session's streams were last drained: work queued on the releasing cursor's stream
  // may still touch it, so the next reuse waits for the streams first (Release did not order against them)
  bool dirty[2];
  PoolState() { dirty[0] = dirty[1] = false; }
};
PoolState* pool_state() {
  static PoolState* p = new PoolState;
  return p;
}
}  // namespace

FailureOr<void*> MemoryPool::Acquire(Kind kind, size_t bytes, size_t* granted) {
  const size_t granule = 1u << 20;
  const size_t want = ((bytes ? bytes : 1) + granule - 1) / granule * granule;
  PoolState* ps = pool_state();
  {
    std::lock_guard<std::mutex> lock(ps->mu);
    std::multimap<size_t, void*>::iterator it = ps->free_blocks[kind].lower_bound(want);
    if (it != ps->free_blocks[kind].end() && it->first <= want * 2 + granule) {
      void* p = it->second;
      *granted = it->first;
      ps->free_blocks[kind].erase(it);
      const bool drain = ps->dirty[kind];
      ps->dirty[kind] = false;
      if (drain) {
        FailureOr<Session*> session = Session::Get();
        if (session.is_success()) session.get()->SyncAll();
      }
      return Success(p);
    }
  }
  FailureOr<Session*> s = Session::Get();
  PROPAGATE_ON_FAILURE(s);
  void* p = NULL;
  int rc = kind == DEVICE ? ssb_malloc(s.get()->ctx(), want, &p) : ssb_malloc_host(s.get()->ctx(), want, &p);
  if (rc != 0) {
    // give cached blocks back to the driver and retry once
    {
      std::lock_guard<std::mutex> lock(ps->mu);
      for (std::multimap<size_t, void*>::iterator it = ps->free_blocks[kind].begin(); it != ps->free_blocks[kind].end(); ++it) {
        if (kind == DEVICE) ssb_free(s.get()->ctx(), it->second); else ssb_free_host(s.get()->ctx(), it->second);
      }
      ps->free_blocks[kind].clear();
    }
    rc = kind == DEVICE ? ssb_malloc(s.get()->ctx(), want, &p) : ssb_malloc_host(s.get()->ctx(), want, &p);
    if (rc != 0) THROW(s.get()->Error(rc, kind == DEVICE ? "device allocation" : "pinned allocation"));
  }
  *granted = want;
  return Success(p);
}

void MemoryPool::Release(Kind kind, void* ptr, size_t granted) {
  if (ptr == NULL) return;
  PoolState* ps = pool_state();
  std::lock_guard<std::mutex> lock(ps->mu);
  ps->free_blocks[kind].insert(std::make_pair(granted, ptr));
  ps->dirty[kind] = true;
}

FailureOrVoid DeviceBuffer::Allocate(size_t bytes) {
  Free();
  FailureOr<void*> p = MemoryPool::Acquire(MemoryPool::DEVICE, bytes, &granted_);
  PROPAGATE_ON_FAILURE(p);
  ptr_ = p.get();
  bytes_ = bytes;
  return Success();
}
void DeviceBuffer::Free() {
  if (ptr_ != NULL) {
    MemoryPool::Release(MemoryPool::DEVICE, ptr_, granted_);
    ptr_ = NULL;
    bytes_ = 0;
    granted_ = 0;
  }
}

bool IsDevicePointer(const void* p) { return ssb_pointer_is_device(p) != 0; }

static size_t BitmapBytes(int64 rows) { return static_cast<size_t>((rows + 31) / 32 + 1) * 4 + 128; }

FailureOrVoid DeviceTable::Allocate(const TupleSchema& s, int64 capacity, bool force_nulls) {
  schema = s;
  columns.clear();
  columns.resize(s.attribute_count());
  rows = 0;
  for (int i = 0; i < s.attribute_count(); ++i) {
    const Attribute& a = s.attribute(i);
    DeviceColumnRef& c = columns[i];
    c.data.reset(new DeviceBuffer);
    PROPAGATE_ON_FAILURE(c.data->Allocate(static_cast<size_t>(capacity) * DeviceWidth(a.type()) + 128));
    c.col.data = c.data->get();
    c.col.dtype = DeviceType(a.type());   // variable-length attributes: INT64 codes (the dictionary is attached by the operator)
    c.col.nulls = NULL;
    c.col.reserved = 0;
    if (a.is_nullable() || force_nulls) {
      c.nulls.reset(new DeviceBuffer);
      PROPAGATE_ON_FAILURE(c.nulls->Allocate(BitmapBytes(capacity)));
      c.col.nulls = static_cast<uint32_t*>(c.nulls->get());
    }
  }
  return Success();
}

FailureOrVoid DeviceTable::Download(Block* block, rowcount_t block_offset) const {
  FailureOr<Session*> sr = Session::Get();
  PROPAGATE_ON_FAILURE(sr);
  Session* s = sr.get();
  DeviceBuffer bools;
  for (size_t i = 0; i < columns.size(); ++i) {
    const DataType type = schema.attribute(static_cast<int>(i)).type();
    const size_t w = GetTypeInfo(type).size();
    if (IsVariableLength(type)) {
      PROPAGATE_ON_FAILURE(DownloadStringColumn(s, columns[i], rows,
                                                static_cast<StringPiece*>(block->mutable_data(static_cast<int>(i))) + block_offset, block));
    } else if (rows > 0) {
      SSB_CALL(s, ssb_memcpy_d2h(s->ctx(), static_cast<char*>(block->mutable_data(static_cast<int>(i))) + block_offset * w,
                                 columns[i].col.data, static_cast<size_t>(rows) * w), "download");
    }
    bool* hn = block->mutable_is_null(static_cast<int>(i));
    if (hn != NULL && rows > 0) {
      if (columns[i].col.nulls != NULL) {
        if (bools.size() < static_cast<size_t>(rows)) PROPAGATE_ON_FAILURE(bools.Allocate(static_cast<size_t>(rows)));
        SSB_CALL(s, ssb_nulls_unpack(s->ctx(), columns[i].col.nulls, rows, static_cast<uint8_t*>(bools.get())), "null unpack");
        SSB_CALL(s, ssb_memcpy_d2h(s->ctx(), hn + block_offset, bools.get(), static_cast<size_t>(rows)), "download nulls");
        SSB_CALL(s, ssb_ctx_sync(s->ctx()), "sync");   // `bools` is reused by the next column
      } else {
        memset(hn + block_offset, 0, static_cast<size_t>(rows));
      }
    }
  }
  SSB_CALL(s, ssb_ctx_sync(s->ctx()), "sync");
  return Success();
}

FailureOrVoid UploadColumns(const View& view, const vector<int>& cols, rowcount_t offset, rowcount_t rows,
                            DeviceTable* out) {
  FailureOr<Session*> sr = Session::Get();
  PROPAGATE_ON_FAILURE(sr);
  Session* s = sr.get();
  out->columns.clear();
  out->columns.resize(cols.size());
  out->rows = static_cast<int64>(rows);
  TupleSchema schema;
  DeviceBuffer bools;
  for (size_t k = 0; k < cols.size(); ++k) {
    const Column& src = view.column(cols[k]);
    const Attribute& a = src.attribute();
    schema.add_attribute(a);
    DeviceColumnRef& c = out->columns[k];
    const size_t w = src.type_info().size();
    c.col.dtype = DeviceType(a.type());
    c.col.reserved = 0;
    c.col.nulls = NULL;
    const char* p = static_cast<const char*>(src.data().raw()) + offset * w;
    if (IsVariableLength(a.type())) {
      // StringPiece cells (host memory): packed, copied and ranked into codes + dictionary
      PROPAGATE_ON_FAILURE(UploadStringColumn(s, reinterpret_cast<const StringPiece*>(p), src.is_null() ? src.is_null() + offset : NULL,
                                              rows, &c));
    } else {
      if (IsDevicePointer(src.data().raw())) {
        if (src.is_null() != NULL) {
          THROW(new Exception(ERROR_NOT_IMPLEMENTED, "device-resident input columns must not carry an is_null vector"));
        }
        c.col.data = const_cast<char*>(p);
        continue;
      }
      c.data.reset(new DeviceBuffer);
      PROPAGATE_ON_FAILURE(c.data->Allocate(static_cast<size_t>(rows) * w + 128));
      c.col.data = c.data->get();
      if (rows > 0) SSB_CALL(s, ssb_memcpy_h2d(s->ctx(), c.col.data, p, static_cast<size_t>(rows) * w), "upload");
    }
    if (src.is_null() != NULL && rows > 0) {
      c.nulls.reset(new DeviceBuffer);
      PROPAGATE_ON_FAILURE(c.nulls->Allocate(BitmapBytes(static_cast<int64>(rows))));
      c.col.nulls = static_cast<uint32_t*>(c.nulls->get());
      if (bools.size() < rows) PROPAGATE_ON_FAILURE(bools.Allocate(rows));
      SSB_CALL(s, ssb_memcpy_h2d(s->ctx(), bools.get(), src.is_null() + offset, rows), "upload nulls");
      SSB_CALL(s, ssb_nulls_pack(s->ctx(), static_cast<const uint8_t*>(bools.get()), static_cast<int64>(rows), c.col.nulls), "null pack");
      SSB_CALL(s, ssb_ctx_sync(s->ctx()), "sync");
    }
  }
  out->schema = schema;
  SSB_CALL(s, ssb_ctx_sync(s->ctx()), "sync");
  return Success();
}

// ------------------------------------------------------------------ program lowering
namespace {
struct Lowering {
  vector<ssb_expr_node> nodes;
  std::map<const ExprNode*, int> index;
  std::map<int, int> input_node;  // schema column -> program node of its INPUT
  vector<int> used;

  int Lower(const NodePtr& n) {
    std::map<const ExprNode*, int>::iterator it = index.find(n.get());
    if (it != index.end()) return it->second;
    ssb_expr_node out;
    memset(&out, 0, sizeof(out));
    out.op = n->op;
    out.out_type = DeviceType(n->type);   // variable-length values travel as INT64 codes (constants are resolved by the cursor)
    out.arg[0] = out.arg[1] = out.arg[2] = -1;
    out.flags = n->flags;
    memcpy(&out.imm, &n->imm, sizeof(out.imm));
    if (n->op == SSB_OP_INPUT) {
      // several INPUT nodes of one column lower to one program node
      std::map<int, int>::iterator m = input_node.find(n->input);
      if (m != input_node.end()) { index[n.get()] = m->second; return m->second; }
      out.arg[0] = static_cast<int>(used.size());
      used.push_back(n->input);
      input_node[n->input] = static_cast<int>(nodes.size());
    } else {
      for (size_t i = 0; i < n->args.size() && i < 3; ++i) out.arg[i] = Lower(n->args[i]);
    }
    const int id = static_cast<int>(nodes.size());
    nodes.push_back(out);
    index[n.get()] = id;
    return id;
  }
};
}  // namespace

void LowerProgram(const vector<NodePtr>& outputs, const NodePtr& predicate, vector<ssb_expr_node>* nodes,
                  vector<int>* used, vector<int32_t>* outs, int* pred) {
  Lowering low;
  outs->clear();
  for (size_t j = 0; j < outputs.size(); ++j) outs->push_back(low.Lower(outputs[j]));
  *pred = predicate ? low.Lower(predicate) : -1;
  *nodes = low.nodes;
  *used = low.used;
}

DeviceProgram::~DeviceProgram() { if (prog_) ssb_program_destroy(prog_); }

FailureOrOwned<DeviceProgram> DeviceProgram::Create(const TupleSchema& input_schema, const vector<NodePtr>& outputs,
                                                    const NodePtr& predicate) {
  FailureOr<Session*> sr = Session::Get();
  PROPAGATE_ON_FAILURE(sr);
  Session* s = sr.get();
  Lowering low;
  vector<int32_t> outs;
  for (size_t j = 0; j < outputs.size(); ++j) outs.push_back(low.Lower(outputs[j]));
  const int pred = predicate ? low.Lower(predicate) : -1;
  vector<int32_t> types, nullable;
  for (size_t k = 0; k < low.used.size(); ++k) {
    types.push_back(DeviceType(input_schema.attribute(low.used[k]).type()));
    nullable.push_back(input_schema.attribute(low.used[k]).is_nullable() ? 1 : 0);
  }
  std::unique_ptr<DeviceProgram> p(new DeviceProgram);
  int32_t dummy = 0;
  SSB_CALL(s, ssb_program_create(s->ctx(), low.nodes.data(), static_cast<int32_t>(low.nodes.size()),
                                 static_cast<int32_t>(low.used.size()), types.empty() ? &dummy : types.data(),
                                 nullable.empty() ? &dummy : nullable.data(), outs.empty() ? &dummy : outs.data(),
                                 static_cast<int32_t>(outs.size()), pred, &p->prog_),
           "expression compilation");
  p->used_ = low.used;
  p->has_pred_ = pred >= 0;
  p->n_out_ = static_cast<int>(outs.size());
  return Success(p.release());
}

FailureOr<int64> DeviceProgram::Run(const vector<ssb_column>& inputs, int64 rows, const vector<ssb_column>& outputs) {
  FailureOr<Session*> sr = Session::Get();
  PROPAGATE_ON_FAILURE(sr);
  Session* s = sr.get();
  int64_t out_rows = 0;
  ssb_column dummy;
  memset(&dummy, 0, sizeof(dummy));
  SSB_CALL(s, ssb_program_run_sync(prog_, inputs.empty() ? &dummy : inputs.data(), rows,
                                   outputs.empty() ? &dummy : outputs.data(), &out_rows),
           "expression evaluation");
  return Success(static_cast<int64>(out_rows));
}

// ------------------------------------------------------------------ GpuCursor
FailureOr<const DeviceTable*> GpuCursor::Produce() {
  if (!produced_) {
    if (interrupted()) THROW(new Exception(INTERRUPTED, "cursor interrupted"));
    PROPAGATE_ON_FAILURE(Run(&result_));
    produced_ = true;
  }
  const DeviceTable* t = &result_;
  return Success(t);
}

ResultView GpuCursor::Next(rowcount_t max_row_count) {
  if (interrupted()) return ResultView::Failure(new Exception(INTERRUPTED, "cursor interrupted"));
  if (!host_) {
    FailureOr<const DeviceTable*> t = Produce();
    if (t.is_failure()) return ResultView::Failure(t.release_exception());
    host_.reset(new Block(schema_, allocator_));
    if (!host_->Reallocate(static_cast<rowcount_t>(result_.rows))) {
      return ResultView::Failure(new Exception(ERROR_MEMORY_EXCEEDED, "cannot allocate the host copy of the result"));
    }
    FailureOrVoid d = result_.Download(host_.get());
    if (d.is_failure()) return ResultView::Failure(d.release_exception());
    offset_ = 0;
  }
  const rowcount_t total = static_cast<rowcount_t>(result_.rows);
  if (offset_ >= total) return ResultView::EOS();
  rowcount_t n = total - offset_;
  if (n > max_row_count) n = max_row_count;
  view_.ResetFromSubRange(host_->view(), offset_, n);
  // NOT_NULLABLE columns report no is_null vector (block.h:131-135)
  offset_ += n;
  return ResultView::Success(&view_);
}

FailureOrVoid MaterializeOnDevice(Cursor* child, DeviceTable* out, std::unique_ptr<Block>* host_keepalive) {
  if (GpuCursor* g = dynamic_cast<GpuCursor*>(child)) {
    FailureOr<const DeviceTable*> t = g->Produce();
    PROPAGATE_ON_FAILURE(t);
    *out = *t.get();   // shares the buffers
    return Success();
  }
  // a foreign cursor: drain it to a host block, then upload
  std::unique_ptr<Block> block(new Block(child->schema(), HeapBufferAllocator::Get()));
  rowcount_t rows = 0, cap = 0;
  for (;;) {
    ResultView rv = child->Next(1 << 20);
    if (rv.is_failure()) return Failure(rv.release_exception());
    if (!rv.has_data()) {
      if (rv.is_eos()) break;
      THROW(new Exception(ERROR_NOT_IMPLEMENTED, "WAITING_ON_BARRIER inputs are not supported by GPU operators"));
    }
    const View& v = rv.view();
    if (rows + v.row_count() > cap) {
      cap = (rows + v.row_count()) * 2;
      if (!block->Reallocate(cap)) THROW(new Exception(ERROR_MEMORY_EXCEEDED, "cannot buffer the input of a GPU operator"));
    }
    for (int c = 0; c < v.column_count(); ++c) {
      const size_t w = v.column(c).type_info().size();
      if (IsVariableLength(v.column(c).attribute().type())) {
        // deep copy: the child may reuse the memory its cells point into on the next call
        const StringPiece* src = static_cast<const StringPiece*>(v.column(c).data().raw());
        StringPiece* dst = static_cast<StringPiece*>(block->mutable_data(c)) + rows;
        size_t total = 0;
        for (rowcount_t i = 0; i < v.row_count(); ++i) if (!v.column(c).is_null() || !v.column(c).is_null()[i]) total += src[i].size();
        std::shared_ptr<vector<char> > store(new vector<char>(total + 1));
        size_t at = 0;
        for (rowcount_t i = 0; i < v.row_count(); ++i) {
          if (v.column(c).is_null() && v.column(c).is_null()[i]) { dst[i] = StringPiece(); continue; }
          if (src[i].size() > 0) memcpy(store->data() + at, src[i].data(), src[i].size());
          dst[i] = StringPiece(store->data() + at, src[i].size());
          at += src[i].size();
        }
        block->KeepAlive(store);
      } else {
        memcpy(static_cast<char*>(block->mutable_data(c)) + rows * w, v.column(c).data().raw(), v.row_count() * w);
      }
      if (bool* hn = block->mutable_is_null(c)) {
        if (v.column(c).is_null()) memcpy(hn + rows, v.column(c).is_null(), v.row_count());
        else memset(hn + rows, 0, v.row_count());
      }
    }
    rows += v.row_count();
  }
  vector<int> cols;
  for (int c = 0; c < child->schema().attribute_count(); ++c) cols.push_back(c);
  View v(block->view());
  v.set_row_count(rows);
  PROPAGATE_ON_FAILURE(UploadColumns(v, cols, 0, rows, out));
  if (host_keepalive) host_keepalive->reset(block.release());
  return Success();
}

}  // namespace internal

// ------------------------------------------------------------------ BoundExpression
BoundExpression::BoundExpression(const TupleSchema& input_schema, const TupleSchema& result_schema, const vector<internal::NodePtr>& nodes)
    : input_schema_(input_schema), result_schema_(result_schema), nodes_(nodes), row_capacity_(Cursor::kDefaultRowCount),
      view_(result_schema) {}
BoundExpression::~BoundExpression() {}

// One kernel for the whole DAG over the rows of `input` (expression.h:60-66). Rows flagged in a column's skip vector
// are not part of the result: they come back NULL where the column is NULLABLE, unspecified otherwise, and the
// signaling operators cannot be told to spare them (the fused kernel has no per-row skip input) -- which is why the
// cursors state skips as guards instead (expression.cc GuardSignaling) and never call this.
EvaluationResult BoundExpression::DoEvaluate(const View& input, const BoolView& skip_vectors) {
  using namespace internal;   // NOLINT
  if (!program_) {
    vector<NodePtr> outs;
    for (size_t i = 0; i < nodes_.size(); ++i) outs.push_back(GuardSignaling(nodes_[i], vector<NodePtr>(), NodePtr()));
    FailureOrOwned<DeviceProgram> p = DeviceProgram::Create(input_schema_, outs, NodePtr());
    PROPAGATE_ON_FAILURE(p);
    program_.reset(p.release());
  }
  DeviceTable in;
  PROPAGATE_ON_FAILURE(UploadColumns(input, program_->used_inputs(), 0, input.row_count(), &in));
  DeviceTable out;
  PROPAGATE_ON_FAILURE(out.Allocate(result_schema_, static_cast<int64>(input.row_count())));
  vector<ssb_column> ic, oc;
  for (size_t i = 0; i < in.columns.size(); ++i) ic.push_back(in.columns[i].col);
  for (size_t i = 0; i < out.columns.size(); ++i) oc.push_back(out.columns[i].col);
  FailureOr<int64> n = program_->Run(ic, static_cast<int64>(input.row_count()), oc);
  PROPAGATE_ON_FAILURE(n);
  out.rows = n.get();
  if (!result_block_) result_block_.reset(new Block(result_schema_, HeapBufferAllocator::Get()));
  if (result_block_->row_capacity() < input.row_count() && !result_block_->Reallocate(input.row_count())) {
    THROW(new Exception(ERROR_MEMORY_EXCEEDED, "DoEvaluate: cannot allocate the result block"));
  }
  PROPAGATE_ON_FAILURE(out.Download(result_block_.get()));
  for (int c = 0; c < result_schema_.attribute_count() && c < skip_vectors.column_count(); ++c) {
    bool* skip = skip_vectors.column(c);
    bool* is_null = result_block_->mutable_is_null(c);
    if (skip == NULL) continue;
    for (rowcount_t i = 0; i < input.row_count(); ++i) {
      if (is_null != NULL) { is_null[i] = is_null[i] || skip[i]; skip[i] = is_null[i]; }
    }
  }
  view_.ResetFromSubRange(result_block_->view(), 0, input.row_count());
  const View& rv = view_;
  return Success(rv);
}

// ------------------------------------------------------------------ BoundExpressionTree
BoundExpressionTree::BoundExpressionTree(BoundExpression* root, BufferAllocator* allocator, rowcount_t max_row_count)
    : root_(root), allocator_(allocator), max_row_count_(max_row_count), result_view_(root->result_schema()) {}
BoundExpressionTree::~BoundExpressionTree() {}

EvaluationResult BoundExpressionTree::Evaluate(const View& input) {
  using namespace internal;   // NOLINT
  if (input.row_count() > max_row_count_) {
    THROW(new Exception(ERROR_TOO_MANY_ROWS, "Evaluate: the view is larger than the expression's row capacity"));
  }
  if (!program_) {
    vector<NodePtr> outs;
    for (int i = 0; i < root_->column_count(); ++i) outs.push_back(GuardSignaling(root_->node(i), vector<NodePtr>(), NodePtr()));
    FailureOrOwned<DeviceProgram> p = DeviceProgram::Create(root_->input_schema(), outs, NodePtr());
    PROPAGATE_ON_FAILURE(p);
    program_.reset(p.release());
  }
  DeviceTable in;
  PROPAGATE_ON_FAILURE(UploadColumns(input, program_->used_inputs(), 0, input.row_count(), &in));
  DeviceTable out;
  PROPAGATE_ON_FAILURE(out.Allocate(root_->result_schema(), static_cast<int64>(input.row_count())));
  vector<ssb_column> ic, oc;
  for (size_t i = 0; i < in.columns.size(); ++i) ic.push_back(in.columns[i].col);
  for (size_t i = 0; i < out.columns.size(); ++i) oc.push_back(out.columns[i].col);
  FailureOr<int64> n = program_->Run(ic, static_cast<int64>(input.row_count()), oc);
  PROPAGATE_ON_FAILURE(n);
  out.rows = n.get();
  if (!result_block_) result_block_.reset(new Block(root_->result_schema(), allocator_));
  if (result_block_->row_capacity() < input.row_count() && !result_block_->Reallocate(input.row_count())) {
    THROW(new Exception(ERROR_MEMORY_EXCEEDED, "Evaluate: cannot allocate the result block"));
  }
  PROPAGATE_ON_FAILURE(out.Download(result_block_.get()));
  result_view_.ResetFromSubRange(result_block_->view(), 0, input.row_count());
  const View& rv = result_view_;
  return Success(rv);
}

}  // namespace supersonic
