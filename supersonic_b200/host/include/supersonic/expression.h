// supersonic/expression.h -- Expression / BoundExpression / BoundExpressionTree and the
// expression factories of the hot path.
//
// Mirrors expression/base/expression.h:46-225 and the factory headers
// expression/core/{arithmetic,comparison,elementary,projecting}_expressions.h and
// expression/infrastructure/terminal_expressions.h. Binding performs the reference's type
// promotion (expression/templated/bound_expression_factory.cc:44-123,
// expression/core/comparison_bound_expressions.cc:587-636) and produces result names in the
// reference's formats (expression/vector/expression_traits.h:1209-1453). A bound expression is
// a typed DAG that the cursors hand, whole, to libssb200.so: one kernel evaluates it.
#ifndef SUPERSONIC_B200_HOST_EXPRESSION_H_
#define SUPERSONIC_B200_HOST_EXPRESSION_H_

#include <set>

#include "supersonic/base.h"
#include "supersonic/projector.h"

namespace supersonic {

namespace internal {
// One node of a bound (typed) expression DAG over the columns of an input schema.
struct ExprNode {
  int op;                 // SSB_OP_* (include/supersonic_b200.h)
  DataType type;
  bool nullable;
  bool constant;          // no INPUT below this node
  string name;            // the reference's description of this sub-expression
  vector<std::shared_ptr<const ExprNode> > args;
  int input;              // SSB_OP_INPUT: column index in the input schema
  int flags;              // SSB_NODE_*
  union { int64 i64; uint64 u64; double f64; float f32; int32 i32; uint32 u32; bool b; } imm;
  string text;            // SSB_OP_CONST of type STRING / BINARY: the bytes
  ExprNode() : op(0), type(INT32), nullable(false), constant(false), input(-1), flags(0) { imm.u64 = 0; }
};
typedef std::shared_ptr<const ExprNode> NodePtr;
class DeviceProgram;
}  // namespace internal

typedef FailureOrReference<const View> EvaluationResult;

// The result of binding: result_schema() describes the columns, node(i) computes column i
// from the columns of the schema the expression was bound to.
// expression/base/expression.h:46-93. The reference evaluates a bound tree node by node through DoEvaluate; here a
// bound expression is a DAG the cursors compile into one kernel, and DoEvaluate is the same kernel run for one view
// (rows whose skip flag is set come back NULL / unspecified, as the skip-vector contract says). Subclasses may
// override the virtuals; every factory below returns this class.
class BoundExpression {
 public:
  BoundExpression(const TupleSchema& input_schema, const TupleSchema& result_schema,
                  const vector<internal::NodePtr>& nodes);
  virtual ~BoundExpression();
  const TupleSchema& result_schema() const { return result_schema_; }
  const TupleSchema& input_schema() const { return input_schema_; }
  int column_count() const { return static_cast<int>(nodes_.size()); }
  const internal::NodePtr& node(int i) const { return nodes_[i]; }
  // One skip vector per result column (BoolView of result_schema().attribute_count() columns; a NULL column = skip
  // nothing). On return the skip vector of a NULLABLE column also flags the rows whose result is NULL.
  virtual EvaluationResult DoEvaluate(const View& input, const BoolView& skip_vectors);
  virtual rowcount_t row_capacity() const { return row_capacity_; }
  void set_row_capacity(rowcount_t n) { row_capacity_ = n; }
  virtual bool is_constant() const;
  std::set<string> referred_attribute_names() const;
  virtual void CollectReferredAttributeNames(std::set<string>* referred_attribute_names) const;
  // Names of the input attributes the expression reads, in schema order.
  void CollectReferredAttributeNames(vector<string>* names) const;
 private:
  BoundExpression(const BoundExpression&);
  void operator=(const BoundExpression&);
  TupleSchema input_schema_, result_schema_;
  vector<internal::NodePtr> nodes_;
  rowcount_t row_capacity_;
  std::unique_ptr<internal::DeviceProgram> program_;
  std::unique_ptr<Block> result_block_;
  View view_;
};

class BoundExpressionList {
 public:
  BoundExpressionList() {}
  ~BoundExpressionList() { for (size_t i = 0; i < list_.size(); ++i) delete list_[i]; }
  BoundExpressionList* add(BoundExpression* e) { list_.push_back(e); return this; }
  int size() const { return static_cast<int>(list_.size()); }
  BoundExpression* get(int i) const { return list_[i]; }
  BoundExpression* release(int i) { BoundExpression* e = list_[i]; list_[i] = NULL; return e; }
 private:
  vector<BoundExpression*> list_;
};

// expression/base/expression.h:96-137. Evaluate() runs the expression on the GPU for a host
// view (upload, one kernel, download); the cursors bypass it and keep data in HBM.
class BoundExpressionTree {
 public:
  BoundExpressionTree(BoundExpression* root, BufferAllocator* allocator, rowcount_t max_row_count);
  ~BoundExpressionTree();
  const TupleSchema& result_schema() const { return root_->result_schema(); }
  rowcount_t row_capacity() const { return max_row_count_; }
  bool is_constant() const { return root_->is_constant(); }
  EvaluationResult Evaluate(const View& input);
  const BoundExpression* root() const { return root_.get(); }
  void CollectReferredAttributeNames(vector<string>* names) const { root_->CollectReferredAttributeNames(names); }
 private:
  std::unique_ptr<BoundExpression> root_;
  BufferAllocator* allocator_;
  rowcount_t max_row_count_;
  std::unique_ptr<Block> result_block_;
  View result_view_;
  std::unique_ptr<internal::DeviceProgram> program_;
};

// expression.h:140-144
FailureOrOwned<BoundExpressionTree> CreateBoundExpressionTree(BoundExpression* expression, BufferAllocator* allocator,
                                                              rowcount_t max_row_count);

class Expression {
 public:
  virtual ~Expression() {}
  // expression.cc:84-94
  FailureOrOwned<BoundExpressionTree> Bind(const TupleSchema& input_schema, BufferAllocator* allocator,
                                           rowcount_t max_row_count) const;
  virtual FailureOrOwned<BoundExpression> DoBind(const TupleSchema& input_schema, BufferAllocator* allocator,
                                                 rowcount_t max_row_count) const = 0;
  virtual string ToString(bool verbose) const = 0;
 protected:
  Expression() {}
};

class ExpressionList {
 public:
  ExpressionList() {}
  ~ExpressionList() { for (size_t i = 0; i < list_.size(); ++i) delete list_[i]; }
  ExpressionList* add(const Expression* e) { list_.push_back(e); return this; }
  int size() const { return static_cast<int>(list_.size()); }
  const Expression* get(int i) const { return list_[i]; }
  string ToString(bool verbose) const;
 private:
  vector<const Expression*> list_;
};

// ---- terminal expressions (expression/infrastructure/terminal_expressions.h:36-71) -------
const Expression* ConstInt32(const int32& value);
const Expression* ConstInt64(const int64& value);
const Expression* ConstUint32(const uint32& value);
const Expression* ConstUint64(const uint64& value);
const Expression* ConstFloat(const float& value);
const Expression* ConstDouble(const double& value);
const Expression* ConstBool(const bool& value);
const Expression* ConstDate(const int32& value);
const Expression* ConstDateTime(const int64& value);
const Expression* ConstString(const StringPiece& value);
const Expression* ConstBinary(const StringPiece& value);
const Expression* Null(DataType type);
const Expression* Sequence();

// ---- projecting expressions (expression/core/projecting_expressions.h:52-120) ------------
const Expression* NamedAttribute(const string& name);
const Expression* AttributeAt(int position);
const Expression* Alias(const string& new_name, const Expression* argument);
const Expression* InputAttributeProjection(const SingleSourceProjector* projector);

class CompoundExpression : public Expression {
 public:
  CompoundExpression() {}
  virtual ~CompoundExpression();
  CompoundExpression* Add(const Expression* argument);
  CompoundExpression* AddAs(const StringPiece& alias, const Expression* argument);
  CompoundExpression* AddAsMulti(const vector<string>& aliases, const Expression* argument);
  virtual FailureOrOwned<BoundExpression> DoBind(const TupleSchema& input_schema, BufferAllocator* allocator,
                                                 rowcount_t max_row_count) const;
  virtual string ToString(bool verbose) const;
 private:
  struct Entry { vector<string> aliases; const Expression* expression; };
  vector<Entry> entries_;
};

// ---- arithmetic (expression/core/arithmetic_expressions.h:31-101) ------------------------
const Expression* Plus(const Expression* const a, const Expression* const b);
const Expression* Minus(const Expression* const a, const Expression* const b);
const Expression* Multiply(const Expression* const a, const Expression* const b);
const Expression* Divide(const Expression* const a, const Expression* const b);   // = DivideSignaling
const Expression* DivideSignaling(const Expression* const a, const Expression* const b);
const Expression* DivideNulling(const Expression* const a, const Expression* const b);
const Expression* DivideQuiet(const Expression* const a, const Expression* const b);
const Expression* CppDivide(const Expression* const a, const Expression* const b);  // = CppDivideSignaling
const Expression* CppDivideSignaling(const Expression* const a, const Expression* const b);
const Expression* CppDivideNulling(const Expression* const a, const Expression* const b);
const Expression* Modulus(const Expression* const a, const Expression* const b);   // = ModulusSignaling
const Expression* ModulusSignaling(const Expression* const a, const Expression* const b);
const Expression* ModulusNulling(const Expression* const a, const Expression* const b);
const Expression* Negate(const Expression* const a);

// ---- comparisons (expression/core/comparison_expressions.h:34-88) -------------------------
const Expression* Equal(const Expression* const a, const Expression* const b);
const Expression* NotEqual(const Expression* const a, const Expression* const b);
const Expression* Less(const Expression* const a, const Expression* const b);
const Expression* LessOrEqual(const Expression* const a, const Expression* const b);
const Expression* Greater(const Expression* const a, const Expression* const b);
const Expression* GreaterOrEqual(const Expression* const a, const Expression* const b);
const Expression* IsOdd(const Expression* const arg);
const Expression* IsEven(const Expression* const arg);
const Expression* In(const Expression* const needle_expression, const ExpressionList* haystack_arguments);

// ---- logic and control (expression/core/elementary_expressions.h:31-120) -----------------
const Expression* CastTo(DataType to_type, const Expression* const source);
// Whitespace at either end is accepted; invalid input: garbage (quiet) or NULL (nulling). Constants fold at bind time;
// parsing a STRING column is refused with ERROR_NOT_IMPLEMENTED (see host/src/expression.cc).
const Expression* ParseStringQuiet(DataType to_type, const Expression* const source);
const Expression* ParseStringNulling(DataType to_type, const Expression* const source);
const Expression* And(const Expression* const a, const Expression* const b);
const Expression* Or(const Expression* const a, const Expression* const b);
const Expression* AndNot(const Expression* const left, const Expression* const right);
const Expression* Xor(const Expression* const a, const Expression* const b);
const Expression* Not(const Expression* const e);
const Expression* IsNull(const Expression* const e);
const Expression* IfNull(const Expression* const e, const Expression* const substitute);
const Expression* If(const Expression* const condition, const Expression* const then,
                     const Expression* const otherwise);
const Expression* NullingIf(const Expression* const condition, const Expression* const then,
                            const Expression* const otherwise);
const Expression* Case(const ExpressionList* const arguments);
const Expression* BitwiseAnd(const Expression* a, const Expression* b);
const Expression* BitwiseOr(const Expression* a, const Expression* b);
const Expression* BitwiseXor(const Expression* a, const Expression* b);
const Expression* BitwiseAndNot(const Expression* a, const Expression* b);
const Expression* BitwiseNot(const Expression* argument);
const Expression* ShiftLeft(const Expression* argument, const Expression* shift);
const Expression* ShiftRight(const Expression* argument, const Expression* shift);

// ---- the bound factories: the same binding rules over children that are already bound. They take ownership of
// their arguments; every child must have been bound to the same input schema (constants fit any).
// expression/infrastructure/terminal_bound_expressions.h:35-93
FailureOrOwned<BoundExpression> BoundNull(DataType type, BufferAllocator* allocator, rowcount_t max_row_count);
FailureOrOwned<BoundExpression> BoundConstInt32(const int32& value, BufferAllocator* allocator, rowcount_t max_row_count);
FailureOrOwned<BoundExpression> BoundConstInt64(const int64& value, BufferAllocator* allocator, rowcount_t max_row_count);
FailureOrOwned<BoundExpression> BoundConstUInt32(const uint32& value, BufferAllocator* allocator, rowcount_t max_row_count);
FailureOrOwned<BoundExpression> BoundConstUInt64(const uint64& value, BufferAllocator* allocator, rowcount_t max_row_count);
FailureOrOwned<BoundExpression> BoundConstFloat(const float& value, BufferAllocator* allocator, rowcount_t max_row_count);
FailureOrOwned<BoundExpression> BoundConstDouble(const double& value, BufferAllocator* allocator, rowcount_t max_row_count);
FailureOrOwned<BoundExpression> BoundConstBool(const bool& value, BufferAllocator* allocator, rowcount_t max_row_count);
FailureOrOwned<BoundExpression> BoundConstDate(const int32& value, BufferAllocator* allocator, rowcount_t max_row_count);
FailureOrOwned<BoundExpression> BoundConstDateTime(const int64& value, BufferAllocator* allocator, rowcount_t max_row_count);
FailureOrOwned<BoundExpression> BoundConstString(const StringPiece& value, BufferAllocator* allocator, rowcount_t max_row_count);
FailureOrOwned<BoundExpression> BoundConstBinary(const StringPiece& value, BufferAllocator* allocator, rowcount_t max_row_count);
// expression/core/projecting_bound_expressions.h:40-76
FailureOrOwned<BoundExpression> BoundInputAttributeProjection(const TupleSchema& schema, const SingleSourceProjector& projector);
FailureOrOwned<BoundExpression> BoundAttributeAt(const TupleSchema& schema, size_t position);
FailureOrOwned<BoundExpression> BoundNamedAttribute(const TupleSchema& schema, const string& name);
FailureOrOwned<BoundExpression> BoundAlias(const string& new_name, BoundExpression* argument, BufferAllocator* allocator,
                                           rowcount_t max_row_count);
FailureOrOwned<BoundExpression> BoundCompoundExpression(BoundExpressionList* expressions);
FailureOrOwned<BoundExpression> BoundRenameCompoundExpression(const vector<string>& names, BoundExpressionList* expressions);
// expression/core/{arithmetic,comparison,elementary}_bound_expressions.h
#define SSB200_BOUND1(NAME) \
  FailureOrOwned<BoundExpression> NAME(BoundExpression* source, BufferAllocator* allocator, rowcount_t max_row_count);
#define SSB200_BOUND2(NAME) \
  FailureOrOwned<BoundExpression> NAME(BoundExpression* left, BoundExpression* right, BufferAllocator* allocator, rowcount_t max_row_count);
SSB200_BOUND1(BoundNegate) SSB200_BOUND1(BoundIsOdd) SSB200_BOUND1(BoundIsEven) SSB200_BOUND1(BoundNot) SSB200_BOUND1(BoundIsNull)
SSB200_BOUND1(BoundBitwiseNot)
SSB200_BOUND2(BoundPlus) SSB200_BOUND2(BoundMinus) SSB200_BOUND2(BoundMultiply) SSB200_BOUND2(BoundDivideSignaling)
SSB200_BOUND2(BoundDivideNulling) SSB200_BOUND2(BoundDivideQuiet) SSB200_BOUND2(BoundCppDivideSignaling)
SSB200_BOUND2(BoundCppDivideNulling) SSB200_BOUND2(BoundModulusSignaling) SSB200_BOUND2(BoundModulusNulling)
SSB200_BOUND2(BoundEqual) SSB200_BOUND2(BoundNotEqual) SSB200_BOUND2(BoundLess) SSB200_BOUND2(BoundLessOrEqual)
SSB200_BOUND2(BoundGreater) SSB200_BOUND2(BoundGreaterOrEqual) SSB200_BOUND2(BoundOr) SSB200_BOUND2(BoundAnd)
SSB200_BOUND2(BoundAndNot) SSB200_BOUND2(BoundXor) SSB200_BOUND2(BoundIfNull) SSB200_BOUND2(BoundBitwiseAnd)
SSB200_BOUND2(BoundBitwiseAndNot) SSB200_BOUND2(BoundBitwiseOr) SSB200_BOUND2(BoundBitwiseXor) SSB200_BOUND2(BoundShiftLeft)
SSB200_BOUND2(BoundShiftRight)
#undef SSB200_BOUND1
#undef SSB200_BOUND2
FailureOrOwned<BoundExpression> BoundCastTo(DataType to_type, BoundExpression* source, BufferAllocator* allocator,
                                            rowcount_t max_row_count);
FailureOrOwned<BoundExpression> BoundIf(BoundExpression* condition, BoundExpression* then, BoundExpression* otherwise,
                                        BufferAllocator* allocator, rowcount_t max_row_count);
FailureOrOwned<BoundExpression> BoundIfNulling(BoundExpression* condition, BoundExpression* if_true, BoundExpression* if_false,
                                               BufferAllocator* allocator, rowcount_t max_row_count);

// expression/infrastructure/basic_bound_expression.h:236-244: resolves an expression that is constant after binding
// into its value (*is_null is set when it evaluates to NULL). Does not take ownership.
namespace internal {
FailureOrVoid ConstantExpressionValue(const Expression& expression, DataType type, void* value, string* text, bool* is_null);
template <typename Hold> struct ConstantHolder {
  Hold value;
  ConstantHolder() : value() {}
  void* raw() { return &value; }
  void take(const string&) {}
};
template <> struct ConstantHolder<string> {
  string value;
  void* raw() { return NULL; }
  void take(const string& text) { value = text; }
};
}  // namespace internal
template <DataType data_type>
FailureOr<typename TypeTraits<data_type>::hold_type> GetConstantExpressionValue(const Expression& expression, bool* is_null) {
  internal::ConstantHolder<typename TypeTraits<data_type>::hold_type> holder;
  string text;
  FailureOrVoid r = internal::ConstantExpressionValue(expression, data_type, holder.raw(), &text, is_null);
  if (r.is_failure()) return Failure(r.release_exception());
  holder.take(text);
  return Success(holder.value);
}

}  // namespace supersonic
#endif  // SUPERSONIC_B200_HOST_EXPRESSION_H_
