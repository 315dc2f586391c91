// supersonic/cursor.h -- Cursor / Operation and the operator factories of the hot path.
//
// Mirrors cursor/base/cursor.h:42-229, cursor/base/operation.h:35-82 and the factory headers
// cursor/core/{scan_view,compute,filter,project,aggregate,hash_join,sort}.h,
// cursor/infrastructure/{table,ordering}.h. Factories take ownership of the expressions,
// projectors, specifications and child operations passed to them (compute.h:31, filter.h:34,
// aggregate.h:252-262); allocators are borrowed.
//
// Every operator here runs on the GPU through libssb200.so. Chains of row-wise operators
// (ScanView, Compute, Filter, Project) over one source are fused into a single kernel launch
// per chunk; the others hand device-resident columns to one another without a host round
// trip. Cursor::Next still returns host-readable Views, as the contract demands
// (cursor.h:150-186).
#ifndef SUPERSONIC_B200_HOST_CURSOR_H_
#define SUPERSONIC_B200_HOST_CURSOR_H_

#include "supersonic/base.h"
#include "supersonic/expression.h"
#include "supersonic/projector.h"

namespace supersonic {

// cursor/base/cursor.h:42-122
class ResultView {
 public:
  static ResultView Success(const View* view) { return ResultView(view, OK, NULL); }
  static ResultView EOS() { return ResultView(NULL, END_OF_INPUT, NULL); }
  static ResultView BOS() { return ResultView(NULL, BEFORE_INPUT, NULL); }
  static ResultView WaitingOnBarrier() { return ResultView(NULL, WAITING_ON_BARRIER, NULL); }
  static ResultView Failure(Exception* e) { return ResultView(NULL, e->return_code(), e); }
  ResultView(const ResultView& o) : view_(o.view_), status_(o.status_), exception_(const_cast<ResultView&>(o).exception_.release()) {}
  bool has_data() const { return view_ != NULL; }
  bool is_eos() const { return status_ == END_OF_INPUT; }
  bool is_bos() const { return status_ == BEFORE_INPUT; }
  bool is_waiting_on_barrier() const { return status_ == WAITING_ON_BARRIER; }
  bool is_failure() const { return !!exception_; }
  bool is_done() const { return is_eos() || is_failure(); }
  const View& view() const { return *view_; }
  const Exception& exception() const { return *exception_; }
  Exception* release_exception() { return exception_.release(); }
 private:
  ResultView(const View* v, ReturnCode s, Exception* e) : view_(v), status_(s), exception_(e) {}
  void operator=(const ResultView&);
  const View* view_;
  ReturnCode status_;
  std::unique_ptr<Exception> exception_;
};

class CursorTransformer;

// cursor/base/cursor.h:131-229
class Cursor {
 public:
  static const rowcount_t kDefaultRowCount = 1024;
  virtual ~Cursor() {}
  virtual const TupleSchema& schema() const = 0;
  // 1..max_row_count rows, EOS, or a failure. The view stays valid until the next call.
  virtual ResultView Next(rowcount_t max_row_count) = 0;
  // Callable from another thread; non-blocking.
  virtual void Interrupt() = 0;
  virtual bool IsWaitingOnBarrierSupported() const { return false; }
  virtual void ApplyToChildren(CursorTransformer* transformer) { (void)transformer; }
  virtual void AppendDebugDescription(string* target) const = 0;
  virtual CursorId GetCursorId() const { return UNKNOWN_ID; }
 protected:
  Cursor() {}
 private:
  Cursor(const Cursor&);
  void operator=(const Cursor&);
};

class CursorTransformer {
 public:
  virtual ~CursorTransformer() {}
  virtual Cursor* Transform(Cursor* cursor) = 0;
};

namespace internal { struct RowwisePlan; class OperationImpl; }

// cursor/base/operation.h:35-82
class Operation {
 public:
  virtual ~Operation() {}
  virtual void SetBufferAllocator(BufferAllocator* allocator, bool cascade_to_children) = 0;
  virtual void SetBufferAllocatorWhereUnset(BufferAllocator* allocator, bool cascade_to_children) = 0;
  // May be called repeatedly; the operation must outlive its cursors.
  virtual FailureOrOwned<Cursor> CreateCursor() const = 0;
  virtual void AppendDebugDescription(string* target) const = 0;
  string DebugDescription() const { string s; AppendDebugDescription(&s); return s; }
  // Internal: row-wise operators describe themselves so that chains fuse into one kernel.
  // Returns false when the operation is not a row-wise transform of a single scan.
  virtual bool DescribeRowwise(internal::RowwisePlan* plan, Exception** error) const {
    (void)plan; (void)error; return false;
  }
 protected:
  Operation() {}
 private:
  Operation(const Operation&);
  void operator=(const Operation&);
};

// cursor/infrastructure/basic_operation.h:47-181
class BasicOperation : public Operation {
 public:
  virtual ~BasicOperation();
  virtual void SetBufferAllocator(BufferAllocator* allocator, bool cascade_to_children);
  virtual void SetBufferAllocatorWhereUnset(BufferAllocator* allocator, bool cascade_to_children);
  virtual void AppendDebugDescription(string* target) const;
 protected:
  BasicOperation() : allocator_(NULL) {}
  explicit BasicOperation(Operation* child) : allocator_(NULL) { children_.push_back(child); }
  BasicOperation(Operation* child1, Operation* child2) : allocator_(NULL) {
    children_.push_back(child1);
    children_.push_back(child2);
  }
  explicit BasicOperation(const vector<Operation*>& children) : allocator_(NULL), children_(children) {}
  BufferAllocator* buffer_allocator() const { return allocator_ ? allocator_ : HeapBufferAllocator::Get(); }
  Operation* child() const { return children_[0]; }
  Operation* child_at(size_t i) const { return children_[i]; }
  size_t children_count() const { return children_.size(); }
  virtual string DebugName() const = 0;
 private:
  BufferAllocator* allocator_;
  vector<Operation*> children_;
};

// ---- sources ---------------------------------------------------------------------------
// cursor/core/scan_view.h:35. The view's column pointers may be host (pageable or pinned) or
// device memory; they must stay valid while cursors created from the operation live.
Operation* ScanView(const View& view);
// Rows view[selection_vector[i]], i < row_count (scan_view.h:37-46); the vector is borrowed.
Operation* ScanViewWithSelection(const View& view, const rowcount_t row_count, const rowid_t* selection_vector,
                                 rowcount_t buffer_row_capacity);

// cursor/infrastructure/table.h:49-172 (host-memory table; the subset plan code uses)
class Table {
 public:
  Table(const TupleSchema& schema, BufferAllocator* allocator);
  ~Table();
  const TupleSchema& schema() const { return block_->schema(); }
  const View& view() const { return view_; }
  rowcount_t row_count() const { return view_.row_count(); }
  rowcount_t row_capacity() const { return block_->row_capacity(); }
  bool ReserveRowCapacity(rowcount_t needed);
  // Appends the rows of `view` (deep copy); returns the number of rows appended.
  rowcount_t AppendView(const View& view);
  // Appends one uninitialised row; returns its id or -1 when out of memory.
  rowid_t AddRow();
  // Variable-length values are copied into the table's arena (table.h:120-127, 319-334).
  template <DataType type>
  void Set(int col, rowid_t row, const typename TypeTraits<type>::cpp_type& value) {
    static_cast<typename TypeTraits<type>::cpp_type*>(block_->mutable_data(col))[row] = Owned(value);
    if (block_->mutable_is_null(col)) block_->mutable_is_null(col)[row] = false;
  }
  void SetNull(int col, rowid_t row) { block_->mutable_is_null(col)[row] = true; }
  void Clear() { view_.set_row_count(0); }
  FailureOrOwned<Cursor> CreateCursor() const;
 private:
  template <typename T> const T& Owned(const T& v) { return v; }
  StringPiece Owned(const StringPiece& v) { return StringPiece(arena_.AddStringPieceContent(v), v.size()); }
  std::unique_ptr<Block> block_;
  View view_;
  Arena arena_;
};

class TableRowWriter {
 public:
  explicit TableRowWriter(Table* table) : table_(table), row_(-1), col_(0), ok_(true) {}
  TableRowWriter& AddRow();
  TableRowWriter& Int32(int32 v) { return Put<INT32>(v); }
  TableRowWriter& Int64(int64 v) { return Put<INT64>(v); }
  TableRowWriter& Uint32(uint32 v) { return Put<UINT32>(v); }
  TableRowWriter& Uint64(uint64 v) { return Put<UINT64>(v); }
  TableRowWriter& Float(float v) { return Put<FLOAT>(v); }
  TableRowWriter& Double(double v) { return Put<DOUBLE>(v); }
  TableRowWriter& Bool(bool v) { return Put<BOOL>(v); }
  TableRowWriter& Date(int32 v) { return Put<DATE>(v); }
  TableRowWriter& Datetime(int64 v) { return Put<DATETIME>(v); }
  TableRowWriter& String(const StringPiece& v) { return Put<STRING>(v); }
  TableRowWriter& Binary(const StringPiece& v) { return Put<BINARY>(v); }
  TableRowWriter& Null() { if (ok_) table_->SetNull(col_++, row_); return *this; }
  bool success() const { return ok_; }
  void CheckSuccess() const;
 private:
  template <DataType type> TableRowWriter& Put(const typename TypeTraits<type>::cpp_type& v) {
    if (ok_) table_->Set<type>(col_++, row_, v);
    return *this;
  }
  Table* table_;
  rowid_t row_;
  int col_;
  bool ok_;
};

// cursor/core/generate.h:32-35: `count` rows of the empty schema (the input of constant-only Compute plans)
Operation* Generate(rowcount_t count);
FailureOrOwned<Cursor> BoundGenerate(rowcount_t count);

// cursor/core/limit.h:27-30: rows [offset, offset + limit) of the child, in the child's order
Operation* Limit(rowcount_t offset, rowcount_t limit, Operation* child);
Cursor* BoundLimit(rowcount_t offset, rowcount_t limit, Cursor* child);
// cursor/core/coalesce.h:30-34: the columns of the children side by side (equal row counts, distinct attribute names)
Operation* Coalesce(const vector<Operation*>& children);
FailureOrOwned<Cursor> BoundCoalesce(const vector<Cursor*>& children);

// ---- row-wise operators ------------------------------------------------------------------
Operation* Compute(const Expression* computation, Operation* child);                  // compute.h:32
Operation* Filter(const Expression* predicate, const SingleSourceProjector* projector,
                  Operation* child);                                                  // filter.h:35
Operation* Project(const SingleSourceProjector* projector, Operation* child);         // project.h:30

// ---- aggregation (cursor/core/aggregate.h:47-345) -----------------------------------------
class AggregationSpecification {
 public:
  class Element {
   public:
    Element(Aggregation aggregation, const StringPiece& input_name, const StringPiece& output_name, bool distinct)
        : aggregation_(aggregation), input_name_(input_name.as_string()), output_name_(output_name.as_string()),
          output_type_(INT32), output_type_specified_(false), distinct_(distinct) {}
    Element(Aggregation aggregation, const StringPiece& input_name, const StringPiece& output_name,
            DataType output_type, bool distinct)
        : aggregation_(aggregation), input_name_(input_name.as_string()), output_name_(output_name.as_string()),
          output_type_(output_type), output_type_specified_(true), distinct_(distinct) {}
    const Aggregation& aggregation_operator() const { return aggregation_; }
    const string& input() const { return input_name_; }
    const string& output() const { return output_name_; }
    bool output_type_specified() const { return output_type_specified_; }
    DataType output_type() const { return output_type_; }
    bool is_distinct() const { return distinct_; }
   private:
    Aggregation aggregation_;
    string input_name_, output_name_;
    DataType output_type_;
    bool output_type_specified_, distinct_;
  };
  AggregationSpecification() {}
  AggregationSpecification* AddAggregation(Aggregation aggregation, const StringPiece& input_name,
                                           const StringPiece& output_name) {
    return add(Element(aggregation, input_name, output_name, false));
  }
  AggregationSpecification* AddDistinctAggregation(Aggregation aggregation, const StringPiece& input_name,
                                                   const StringPiece& output_name) {
    return add(Element(aggregation, input_name, output_name, true));
  }
  AggregationSpecification* AddAggregationWithDefinedOutputType(Aggregation aggregation, const StringPiece& input_name,
                                                                const StringPiece& output_name, DataType output_type) {
    return add(Element(aggregation, input_name, output_name, output_type, false));
  }
  AggregationSpecification* add(const Element& e) { aggregations_.push_back(e); return this; }
  int size() const { return static_cast<int>(aggregations_.size()); }
  const Element& aggregation(int i) const { return aggregations_[i]; }
 private:
  vector<Element> aggregations_;
};

class GroupAggregateOptions {
 public:
  // aggregate.h:162-167: 16 result rows are allocated up front; the block grows as groups arrive
  GroupAggregateOptions() : memory_quota_(std::numeric_limits<size_t>::max()), enforce_quota_(false),
                            estimated_result_row_count_(16) {}
  GroupAggregateOptions* set_memory_quota(size_t q) { memory_quota_ = q; return this; }
  GroupAggregateOptions* set_enforce_quota(bool e) { enforce_quota_ = e; return this; }
  GroupAggregateOptions* set_estimated_result_row_count(size_t n) { estimated_result_row_count_ = n; return this; }
  size_t memory_quota() const { return memory_quota_; }
  bool enforce_quota() const { return enforce_quota_; }
  size_t estimated_result_row_count() const { return estimated_result_row_count_; }
 private:
  size_t memory_quota_;
  bool enforce_quota_;
  size_t estimated_result_row_count_;
};

Operation* GroupAggregate(const SingleSourceProjector* group_by, AggregationSpecification* aggregation,
                          GroupAggregateOptions* options, Operation* child);          // aggregate.h:224
// Memory contract (aggregate_groups.cc:452-480, 490-1107): the result block lives under a soft MemoryLimit of
// options->memory_quota() on top of the operation's allocator; it starts with estimated_result_row_count() rows
// and fails with ERROR_MEMORY_EXCEEDED when it cannot grow. Here the groups are aggregated in HBM; the same
// budget is applied to the result they form: more groups than max(estimated rows, quota / bytes per result row)
// -- quota = min(memory_quota, allocator->Available()) -- fail the cursor with ERROR_MEMORY_EXCEEDED, and an
// allocator that cannot hold the initial block (estimated rows) fails CreateCursor, as in the reference.
// aggregate.h:230-250: BestEffortGroupAggregate may emit partial results (a key in several rows) under memory
// pressure in the reference; here the aggregation always completes in HBM and every key comes out once, which
// that contract allows, so the budget does not apply to it.
Operation* BestEffortGroupAggregate(const SingleSourceProjector* group_by, AggregationSpecification* aggregation,
                                    GroupAggregateOptions* options, Operation* child);
// aggregate.h:309-336: exact aggregation (DISTINCT included) of inputs larger than memory; the reference spills sorted
// runs under temporary_directory_prefix, this implementation aggregates in HBM and ignores quota and directory.
class HybridGroupDebugOptions;
Operation* HybridGroupAggregate(const SingleSourceProjector* group_by_columns, const AggregationSpecification* aggregation_specification,
                                size_t memory_quota, StringPiece temporary_directory_prefix, Operation* child);
Operation* ScalarAggregate(AggregationSpecification* aggregation, Operation* child);  // aggregate.h:341
// aggregate.h:277-307: aggregates an input that is CLUSTERED by the key columns (rows with equal keys are
// consecutive; a key that comes back later is a new cluster), output in cluster order. On the GPU: cluster ids from
// one compare-with-predecessor pass and a scan (ssb_cluster_ids), aggregation by cluster id, ordered by it.
Operation* AggregateClusters(const SingleSourceProjector* clustered_by_columns, const AggregationSpecification* aggregation,
                             Operation* child);
Operation* AggregateClustersWithSpecifiedOutputBlockSize(const SingleSourceProjector* clustered_by_columns,
                                                         const AggregationSpecification* aggregation,
                                                         rowcount_t block_size, Operation* child);

// cursor/core/aggregator.h:37-112: the bound form of an AggregationSpecification (result schema
// and per-aggregate types / nullability, aggregator.cc:63-152). The accumulators themselves live
// in the GPU hash table of the cursor that receives the Aggregator.
class Aggregator {
 public:
  static FailureOrOwned<Aggregator> Create(const AggregationSpecification& aggregation_specification,
                                           const TupleSchema& input_schema, BufferAllocator* allocator,
                                           rowcount_t result_initial_row_capacity);
  ~Aggregator();
  const TupleSchema& schema() const { return schema_; }
  struct Impl;
  const Impl* impl() const { return impl_; }
  rowcount_t initial_row_capacity() const { return capacity_; }
 private:
  Aggregator() : impl_(NULL), capacity_(0) {}
  TupleSchema schema_;
  Impl* impl_;
  rowcount_t capacity_;
};

// ---- cursors created directly from bound objects (the Bound* factories of cursor/core/*.h).
// Ownership as in the reference: the bound objects and the child are taken over, allocators are not.
Cursor* BoundScanView(const View& view);                                                          // scan_view.h:52
FailureOrOwned<Cursor> BoundScanViewWithSelection(const View& view, const rowcount_t row_count,
                                                  const rowid_t* selection_vector, BufferAllocator* allocator,
                                                  rowcount_t buffer_row_capacity);                  // scan_view.h:59
FailureOrOwned<Cursor> BoundCompute(BoundExpressionTree* computation, BufferAllocator* allocator,
                                    rowcount_t max_row_count, Cursor* child);                     // compute.h:36
FailureOrOwned<Cursor> BoundFilter(BoundExpressionTree* predicate, const BoundSingleSourceProjector* projector,
                                   BufferAllocator* buffer_allocator, Cursor* child_cursor);     // filter.h:43
Cursor* BoundProject(const BoundSingleSourceProjector* projector, Cursor* child);                 // project.h:34
FailureOrOwned<Cursor> BoundGroupAggregate(const BoundSingleSourceProjector* group_by, Aggregator* aggregator,
                                           BufferAllocator* allocator, BufferAllocator* original_allocator,
                                           bool best_effort, Cursor* child);                      // aggregate.h:254
FailureOrOwned<Cursor> BoundHybridGroupAggregate(const SingleSourceProjector* group_by_columns,
                                                 const AggregationSpecification& aggregation_specification,
                                                 StringPiece temporary_directory_prefix, BufferAllocator* allocator, size_t memory_quota,
                                                 const HybridGroupDebugOptions* debug_options, Cursor* child);
Cursor* BoundScalarAggregate(Aggregator* aggregator, Cursor* child);                              // aggregate.h:345
FailureOrOwned<Cursor> BoundAggregateClusters(const BoundSingleSourceProjector* group_by, Aggregator* aggregator,
                                              BufferAllocator* allocator, Cursor* child);          // aggregate.h:291

// ---- hash join (cursor/core/hash_join.h:35-69) ---------------------------------------------
class HashJoinOperation : public BasicOperation {
 public:
  HashJoinOperation(JoinType join_type, const SingleSourceProjector* lhs_key_selector,
                    const SingleSourceProjector* rhs_key_selector,
                    const MultiSourceProjector* result_projector, KeyUniqueness rhs_key_uniqueness,
                    Operation* lhs_child, Operation* rhs_child);
  virtual ~HashJoinOperation();
  virtual FailureOrOwned<Cursor> CreateCursor() const;
 protected:
  virtual string DebugName() const { return "HashJoinOperation"; }
 private:
  const JoinType join_type_;
  std::unique_ptr<const SingleSourceProjector> lhs_key_selector_, rhs_key_selector_;
  std::unique_ptr<const MultiSourceProjector> result_projector_;
  const KeyUniqueness rhs_key_uniqueness_;
};

// ---- sort (cursor/infrastructure/ordering.h:103-137, cursor/core/sort.h:89-131) -------------
class BoundSortOrder;
class SortOrder {
 public:
  SortOrder() {}
  ~SortOrder();
  SortOrder* add(const SingleSourceProjector* projector, ColumnOrder column_order) {
    keys_.push_back(std::make_pair(projector, column_order));
    return this;
  }
  SortOrder* OrderByAttributeAt(int position, ColumnOrder order) { return add(ProjectAttributeAt(position), order); }
  SortOrder* OrderByNamedAttribute(const StringPiece& name, ColumnOrder order) { return add(ProjectNamedAttribute(name), order); }
  // Resolves to (source column, order) pairs, most significant first.
  FailureOrVoid Bind(const TupleSchema& schema, vector<std::pair<int, ColumnOrder> >* keys) const;
  // ordering.h:127: the reference's bound form.
  FailureOrOwned<const BoundSortOrder> Bind(const TupleSchema& source_schema) const;
 private:
  vector<std::pair<const SingleSourceProjector*, ColumnOrder> > keys_;
};

// cursor/infrastructure/ordering.h:48-101
class BoundSortOrder {
 public:
  BoundSortOrder(const BoundSingleSourceProjector* projector, const vector<ColumnOrder>& column_order)
      : projector_(projector), column_order_(column_order) {}
  explicit BoundSortOrder(const BoundSingleSourceProjector* projector)
      : projector_(projector), column_order_(projector->result_schema().attribute_count(), ASCENDING) {}
  const TupleSchema& schema() const { return projector_->result_schema(); }
  ColumnOrder column_order(int i) const { return column_order_[i]; }
  const BoundSingleSourceProjector& projector() const { return *projector_; }
 private:
  std::unique_ptr<const BoundSingleSourceProjector> projector_;
  vector<ColumnOrder> column_order_;
};

// supersonic/proto/specification.proto:12-30 (the generated message's accessors).
class ExtendedSortSpecification {
 public:
  class Key {
   public:
    Key() : column_order_(ASCENDING), case_sensitive_(true), has_case_sensitive_(false) {}
    const string& attribute_name() const { return attribute_name_; }
    ColumnOrder column_order() const { return column_order_; }
    bool case_sensitive() const { return case_sensitive_; }
    bool has_case_sensitive() const { return has_case_sensitive_; }
    void set_attribute_name(const string& v) { attribute_name_ = v; }
    void set_column_order(ColumnOrder v) { column_order_ = v; }
    void set_case_sensitive(bool v) { case_sensitive_ = v; has_case_sensitive_ = true; }
   private:
    string attribute_name_;
    ColumnOrder column_order_;
    bool case_sensitive_, has_case_sensitive_;
  };
  ExtendedSortSpecification() : limit_(0), has_limit_(false) {}
  int keys_size() const { return static_cast<int>(keys_.size()); }
  const Key& keys(int i) const { return keys_[i]; }
  Key* mutable_keys(int i) { return &keys_[i]; }
  Key* add_keys() { keys_.push_back(Key()); return &keys_.back(); }
  bool has_limit() const { return has_limit_; }
  uint64 limit() const { return limit_; }
  void set_limit(uint64 v) { limit_ = v; has_limit_ = true; }
  void CopyFrom(const ExtendedSortSpecification& o) { *this = o; }
 private:
  vector<Key> keys_;
  uint64 limit_;
  bool has_limit_;
};

// memory_limit is accepted for API compatibility; the whole input is sorted in HBM.
Operation* Sort(const SortOrder* sort_order, const SingleSourceProjector* result_projector,
                size_t memory_limit, Operation* child);
Operation* SortWithTempDirPrefix(const SortOrder* sort_order, const SingleSourceProjector* result_projector,
                                 size_t memory_limit, StringPiece temporary_directory_prefix, Operation* child);  // sort.h:98
FailureOrOwned<Cursor> BoundSort(const BoundSortOrder* sort_order, const BoundSingleSourceProjector* result_projector,
                                 size_t memory_limit, StringPiece temporary_directory_prefix, BufferAllocator* allocator,
                                 Cursor* child_cursor);                                            // sort.h:114
// Sort by attribute names with an optional row limit (sort.h:103-131, sort.cc:857-1017); case
// insensitivity concerns STRING keys only, which are not on this path.
// cursor/core/merge_union_all.h: merges inputs that are each sorted by `sort_order` (same column types) into one
// sorted stream; no inputs gives an empty zero-column operation, one input is returned as it is.
Operation* MergeUnionAll(const SortOrder* sort_order, const vector<Operation*>& inputs);
FailureOrOwned<Cursor> BoundMergeUnionAll(const BoundSortOrder* sort_order, vector<Cursor*> inputs,
                                          BufferAllocator* buffer_allocator);
Operation* ExtendedSort(const ExtendedSortSpecification* specification, const SingleSourceProjector* result_projector,
                        size_t memory_limit, Operation* child);
FailureOrOwned<Cursor> BoundExtendedSort(const ExtendedSortSpecification* sort_specification,
                                         const BoundSingleSourceProjector* result_projector, size_t memory_quota,
                                         StringPiece temporary_directory_prefix, BufferAllocator* allocator,
                                         rowcount_t max_row_count, Cursor* child);

}  // namespace supersonic
#endif  // SUPERSONIC_B200_HOST_CURSOR_H_
