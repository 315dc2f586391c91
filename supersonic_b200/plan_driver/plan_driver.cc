// plan_driver.cc -- a plain client of "supersonic/supersonic.h" (see ssplan.h).
//
// Nothing in this file knows which implementation it is linked against: the
// include path decides (the reference tree, or supersonic_b200/host/include).
//
// Plan grammar (S-expressions; atoms are bare words, numbers or "quoted"):
//   operation :=
//     (scan N)                                   ScanView(tables[N])
//     (compute EXPR OP)                          Compute
//     (filter EXPR PROJ OP)                      Filter
//     (project PROJ OP)                          Project
//     (group PROJ (aggs AGG...) OP)              GroupAggregate
//     (scalar_agg (aggs AGG...) OP)              ScalarAggregate
//     (hash_join INNER|LEFT_OUTER PROJ PROJ MPROJ UNIQUE|NOT_UNIQUE OP OP)
//     (sort (order (NAME ASC|DESC)...) PROJ OP)  Sort
//     (merge_union_all (order (NAME ASC|DESC)...) OP...)   MergeUnionAll over inputs sorted by that order
//     (aggregate_clusters PROJ (aggs AGG...) OP)           AggregateClusters over an input clustered by PROJ
//   The bound_* forms build the same cursors through the Bound* factories (BoundCompute,
//   BoundFilter, BoundProject, BoundScanView, BoundGroupAggregate, BoundScalarAggregate, BoundSort):
//   the child is created first, the expression / projector / aggregation is bound against its schema.
//     (bound_scan N) (bound_compute EXPR OP) (bound_filter EXPR PROJ OP) (bound_project PROJ OP)
//     (bound_group PROJ (aggs AGG...) OP) (bound_scalar_agg (aggs AGG...) OP)
//     (bound_sort (order (NAME ASC|DESC)...) PROJ OP)
//   (evaluate EXPR N [CAPACITY [SLICE]]) is not a cursor plan: EXPR is bound to table N with Expression::Bind and
//   BoundExpressionTree::Evaluate is called on successive slices of SLICE (default CAPACITY) rows
//   (expression.cc:41-94); CAPACITY 0 = the table's row count (one call).
//   AGG   := (SUM|MIN|MAX|COUNT|FIRST|LAST in out [TYPE]) | (distinct FN in out)
//   PROJ  := (all) | (all PREFIX) | (named N...) | (at I...) | (rename (N A)...) | (cat PROJ...)
//   MPROJ := (multi (SRC PROJ)...)
//   EXPR  := (col NAME) | (at I) | (i32 V) (i64 V) (u32 V) (u64 V) (f32 V) (f64 V)
//          | (bool V) (date V) (datetime V) | (null TYPE) | (sequence)
//          | (OP1 EXPR) | (OP2 EXPR EXPR) | (if EXPR EXPR EXPR) | (nulling_if EXPR EXPR EXPR)
//          | (cast TYPE EXPR) | (as NAME EXPR) | (case EXPR...) | (in EXPR EXPR...)
//          | (compound EXPR|(as NAME EXPR)...)
#include "ssplan.h"

#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>

#include <deque>
#include <memory>
#include <string>
#include <vector>

#include "supersonic/supersonic.h"
#include "supersonic/cursor/core/aggregator.h"
#include "supersonic/cursor/core/merge_union_all.h"
#include "supersonic/cursor/infrastructure/file_io.h"
#include "supersonic/cursor/infrastructure/writer.h"
#include "supersonic/utils/file.h"
#include "supersonic/expression/core/arithmetic_bound_expressions.h"
#include "supersonic/expression/core/comparison_bound_expressions.h"
#include "supersonic/expression/core/elementary_bound_expressions.h"
#include "supersonic/expression/core/projecting_bound_expressions.h"
#include "supersonic/expression/infrastructure/terminal_bound_expressions.h"
#include "supersonic/proto/specification.pb.h"

namespace {

using namespace supersonic;  // NOLINT

// ------------------------------------------------------------------ s-expr
struct Sx {
  bool atom;
  std::string text;
  std::vector<Sx> kids;
  Sx() : atom(false) {}
};

struct ParseError {
  std::string msg;
};

class SxParser {
 public:
  explicit SxParser(const char* s) : p_(s) {}
  Sx Parse() {
    Sx r = ParseOne();
    Skip();
    if (*p_) throw ParseError{"trailing characters in plan"};
    return r;
  }

 private:
  void Skip() {
    while (*p_ == ' ' || *p_ == '\n' || *p_ == '\t' || *p_ == '\r') ++p_;
  }
  Sx ParseOne() {
    Skip();
    Sx r;
    if (*p_ == '(') {
      ++p_;
      for (;;) {
        Skip();
        if (!*p_) throw ParseError{"unbalanced '('"};
        if (*p_ == ')') { ++p_; break; }
        r.kids.push_back(ParseOne());
      }
      return r;
    }
    r.atom = true;
    if (*p_ == '"') {
      ++p_;
      while (*p_ && *p_ != '"') r.text.push_back(*p_++);
      if (*p_ != '"') throw ParseError{"unterminated string"};
      ++p_;
      return r;
    }
    while (*p_ && *p_ != ' ' && *p_ != '(' && *p_ != ')' && *p_ != '\n' && *p_ != '\t') {
      r.text.push_back(*p_++);
    }
    if (r.text.empty()) throw ParseError{"empty atom"};
    return r;
  }
  const char* p_;
};

const std::string& Head(const Sx& s) {
  if (s.atom || s.kids.empty() || !s.kids[0].atom) throw ParseError{"expected (head ...)"};
  return s.kids[0].text;
}

const std::string& Atom(const Sx& s) {
  if (!s.atom) throw ParseError{"expected atom"};
  return s.text;
}

void Arity(const Sx& s, size_t n) {
  if (s.kids.size() != n + 1) {
    throw ParseError{"wrong number of arguments for '" + Head(s) + "'"};
  }
}

DataType ParseType(const std::string& t) {
  if (t == "INT32") return INT32;
  if (t == "INT64") return INT64;
  if (t == "UINT32") return UINT32;
  if (t == "UINT64") return UINT64;
  if (t == "FLOAT") return FLOAT;
  if (t == "DOUBLE") return DOUBLE;
  if (t == "BOOL") return BOOL;
  if (t == "DATE") return DATE;
  if (t == "DATETIME") return DATETIME;
  if (t == "STRING") return STRING;
  if (t == "BINARY") return BINARY;
  throw ParseError{"unknown type '" + t + "'"};
}

// ------------------------------------------------------------------ builders
const Expression* BuildExpr(const Sx& s);

typedef const Expression* (*Unary)(const Expression*);
typedef const Expression* (*Binary)(const Expression*, const Expression*);

Unary FindUnary(const std::string& h) {
  if (h == "negate") return &Negate;
  if (h == "not") return &Not;
  if (h == "is_null") return &IsNull;
  if (h == "bitwise_not") return &BitwiseNot;
  if (h == "is_odd") return &IsOdd;
  if (h == "is_even") return &IsEven;
  return NULL;
}

Binary FindBinary(const std::string& h) {
  if (h == "plus") return &Plus;
  if (h == "minus") return &Minus;
  if (h == "multiply") return &Multiply;
  if (h == "divide_signaling") return &DivideSignaling;
  if (h == "divide_nulling") return &DivideNulling;
  if (h == "divide_quiet") return &DivideQuiet;
  if (h == "cpp_divide_signaling") return &CppDivideSignaling;
  if (h == "cpp_divide_nulling") return &CppDivideNulling;
  if (h == "modulus_signaling") return &ModulusSignaling;
  if (h == "modulus_nulling") return &ModulusNulling;
  if (h == "equal") return &Equal;
  if (h == "not_equal") return &NotEqual;
  if (h == "less") return &Less;
  if (h == "less_or_equal") return &LessOrEqual;
  if (h == "greater") return &Greater;
  if (h == "greater_or_equal") return &GreaterOrEqual;
  if (h == "and") return &And;
  if (h == "or") return &Or;
  if (h == "and_not") return &AndNot;
  if (h == "xor") return &Xor;
  if (h == "bitwise_and") return &BitwiseAnd;
  if (h == "bitwise_or") return &BitwiseOr;
  if (h == "bitwise_xor") return &BitwiseXor;
  if (h == "bitwise_and_not") return &BitwiseAndNot;
  if (h == "shift_left") return &ShiftLeft;
  if (h == "shift_right") return &ShiftRight;
  if (h == "if_null") return &IfNull;
  return NULL;
}

const Expression* BuildExpr(const Sx& s) {
  const std::string& h = Head(s);
  if (h == "col") { Arity(s, 1); return NamedAttribute(Atom(s.kids[1])); }
  if (h == "at") { Arity(s, 1); return AttributeAt(atoi(Atom(s.kids[1]).c_str())); }
  if (h == "i32") { Arity(s, 1); return ConstInt32(static_cast<int32>(strtoll(Atom(s.kids[1]).c_str(), NULL, 0))); }
  if (h == "i64") { Arity(s, 1); return ConstInt64(strtoll(Atom(s.kids[1]).c_str(), NULL, 0)); }
  if (h == "u32") { Arity(s, 1); return ConstUint32(static_cast<uint32>(strtoull(Atom(s.kids[1]).c_str(), NULL, 0))); }
  if (h == "u64") { Arity(s, 1); return ConstUint64(strtoull(Atom(s.kids[1]).c_str(), NULL, 0)); }
  if (h == "f32") { Arity(s, 1); return ConstFloat(strtof(Atom(s.kids[1]).c_str(), NULL)); }
  if (h == "f64") { Arity(s, 1); return ConstDouble(strtod(Atom(s.kids[1]).c_str(), NULL)); }
  if (h == "bool") { Arity(s, 1); return ConstBool(Atom(s.kids[1]) == "true" || Atom(s.kids[1]) == "1"); }
  if (h == "date") { Arity(s, 1); return ConstDate(static_cast<int32>(strtoll(Atom(s.kids[1]).c_str(), NULL, 0))); }
  if (h == "datetime") { Arity(s, 1); return ConstDateTime(strtoll(Atom(s.kids[1]).c_str(), NULL, 0)); }
  if (h == "str") { Arity(s, 1); return ConstString(Atom(s.kids[1])); }
  if (h == "bin") { Arity(s, 1); return ConstBinary(Atom(s.kids[1])); }
  if (h == "null") { Arity(s, 1); return Null(ParseType(Atom(s.kids[1]))); }
  if (h == "sequence") { Arity(s, 0); return Sequence(); }
  if (h == "cast") { Arity(s, 2); return CastTo(ParseType(Atom(s.kids[1])), BuildExpr(s.kids[2])); }
  if (h == "parse_string_nulling") { Arity(s, 2); return ParseStringNulling(ParseType(Atom(s.kids[1])), BuildExpr(s.kids[2])); }
  if (h == "parse_string_quiet") { Arity(s, 2); return ParseStringQuiet(ParseType(Atom(s.kids[1])), BuildExpr(s.kids[2])); }
  if (h == "as") { Arity(s, 2); return Alias(Atom(s.kids[1]), BuildExpr(s.kids[2])); }
  if (h == "if") {
    Arity(s, 3);
    return If(BuildExpr(s.kids[1]), BuildExpr(s.kids[2]), BuildExpr(s.kids[3]));
  }
  if (h == "nulling_if") {
    Arity(s, 3);
    return NullingIf(BuildExpr(s.kids[1]), BuildExpr(s.kids[2]), BuildExpr(s.kids[3]));
  }
  if (h == "case") {
    std::unique_ptr<ExpressionList> list(new ExpressionList);
    for (size_t i = 1; i < s.kids.size(); ++i) list->add(BuildExpr(s.kids[i]));
    return Case(list.release());
  }
  if (h == "in") {
    if (s.kids.size() < 3) throw ParseError{"in needs a needle and a haystack"};
    const Expression* needle = BuildExpr(s.kids[1]);
    std::unique_ptr<ExpressionList> list(new ExpressionList);
    for (size_t i = 2; i < s.kids.size(); ++i) list->add(BuildExpr(s.kids[i]));
    return In(needle, list.release());
  }
  if (h == "compound") {
    std::unique_ptr<CompoundExpression> c(new CompoundExpression);
    for (size_t i = 1; i < s.kids.size(); ++i) {
      const Sx& k = s.kids[i];
      if (Head(k) == "as") {
        Arity(k, 2);
        c->AddAs(Atom(k.kids[1]), BuildExpr(k.kids[2]));
      } else {
        c->Add(BuildExpr(k));
      }
    }
    return c.release();
  }
  if (Unary u = FindUnary(h)) { Arity(s, 1); return u(BuildExpr(s.kids[1])); }
  if (Binary b = FindBinary(h)) {
    Arity(s, 2);
    const Expression* l = BuildExpr(s.kids[1]);
    return b(l, BuildExpr(s.kids[2]));
  }
  throw ParseError{"unknown expression '" + h + "'"};
}

const SingleSourceProjector* BuildProjector(const Sx& s) {
  const std::string& h = Head(s);
  if (h == "all") {
    if (s.kids.size() == 1) return ProjectAllAttributes();
    Arity(s, 1);
    return ProjectAllAttributes(Atom(s.kids[1]));
  }
  if (h == "named" || h == "at" || h == "rename" || h == "cat") {
    std::unique_ptr<CompoundSingleSourceProjector> c(new CompoundSingleSourceProjector);
    for (size_t i = 1; i < s.kids.size(); ++i) {
      const Sx& k = s.kids[i];
      if (h == "named") {
        c->add(ProjectNamedAttribute(Atom(k)));
      } else if (h == "at") {
        c->add(ProjectAttributeAt(atoi(Atom(k).c_str())));
      } else if (h == "rename") {
        if (k.atom || k.kids.size() != 2) throw ParseError{"rename takes (NAME ALIAS) pairs"};
        c->add(ProjectNamedAttributeAs(Atom(k.kids[0]), Atom(k.kids[1])));
      } else {
        c->add(BuildProjector(k));
      }
    }
    return c.release();
  }
  throw ParseError{"unknown projector '" + h + "'"};
}

const MultiSourceProjector* BuildMultiProjector(const Sx& s) {
  if (Head(s) != "multi") throw ParseError{"expected (multi (SRC PROJ)...)"};
  std::unique_ptr<CompoundMultiSourceProjector> c(new CompoundMultiSourceProjector);
  for (size_t i = 1; i < s.kids.size(); ++i) {
    const Sx& k = s.kids[i];
    if (k.atom || k.kids.size() != 2) throw ParseError{"multi takes (SRC PROJ) pairs"};
    c->add(atoi(Atom(k.kids[0]).c_str()), BuildProjector(k.kids[1]));
  }
  return c.release();
}

Aggregation ParseAggregation(const std::string& a) {
  if (a == "SUM") return SUM;
  if (a == "MIN") return MIN;
  if (a == "MAX") return MAX;
  if (a == "COUNT") return COUNT;
  if (a == "FIRST") return FIRST;
  if (a == "LAST") return LAST;
  throw ParseError{"unknown aggregation '" + a + "'"};
}

AggregationSpecification* BuildAggs(const Sx& s) {
  if (Head(s) != "aggs") throw ParseError{"expected (aggs ...)"};
  std::unique_ptr<AggregationSpecification> spec(new AggregationSpecification);
  for (size_t i = 1; i < s.kids.size(); ++i) {
    const Sx& k = s.kids[i];
    const std::string& h = Head(k);
    if (h == "distinct") {
      Arity(k, 3);
      spec->AddDistinctAggregation(ParseAggregation(Atom(k.kids[1])), Atom(k.kids[2]),
                                   Atom(k.kids[3]));
    } else if (k.kids.size() == 4) {
      spec->AddAggregationWithDefinedOutputType(ParseAggregation(h), Atom(k.kids[1]),
                                                Atom(k.kids[2]), ParseType(Atom(k.kids[3])));
    } else {
      Arity(k, 2);
      spec->AddAggregation(ParseAggregation(h), Atom(k.kids[1]), Atom(k.kids[2]));
    }
  }
  return spec.release();
}

struct Inputs {
  std::vector<View> views;
};

SortOrder* BuildOrder(const Sx& o) {
  std::unique_ptr<SortOrder> order(new SortOrder);
  for (size_t i = 1; i < o.kids.size(); ++i) {
    const Sx& k = o.kids[i];
    if (k.atom || k.kids.size() != 2) throw ParseError{"order takes (NAME ASC|DESC) pairs"};
    const std::string& d = Atom(k.kids[1]);
    if (d != "ASC" && d != "DESC") throw ParseError{"order direction must be ASC or DESC"};
    order->add(ProjectNamedAttribute(Atom(k.kids[0])), d == "ASC" ? ASCENDING : DESCENDING);
  }
  return order.release();
}

Operation* BuildOp(const Sx& s, const Inputs& in) {
  const std::string& h = Head(s);
  if (h == "scan") {
    Arity(s, 1);
    size_t n = static_cast<size_t>(atoi(Atom(s.kids[1]).c_str()));
    if (n >= in.views.size()) throw ParseError{"scan: no such table"};
    return ScanView(in.views[n]);
  }
  if (h == "scan_selection") {
    // (scan_selection N (ids ROW...)): scan_view.h:43; the vector lives as long as the process (a test tool)
    Arity(s, 2);
    size_t n = static_cast<size_t>(atoi(Atom(s.kids[1]).c_str()));
    if (n >= in.views.size()) throw ParseError{"scan_selection: no such table"};
    if (Head(s.kids[2]) != "ids") throw ParseError{"expected (ids ...)"};
    static std::deque<std::vector<rowid_t> > kept;
    kept.push_back(std::vector<rowid_t>());
    kept.back().reserve(s.kids[2].kids.size() + 1);   // never a NULL pointer: the reference CHECKs it (scan_view.cc:122)
    for (size_t i = 1; i < s.kids[2].kids.size(); ++i) kept.back().push_back(static_cast<rowid_t>(atoll(Atom(s.kids[2].kids[i]).c_str())));
    return ScanViewWithSelection(in.views[n], static_cast<rowcount_t>(kept.back().size()), kept.back().data(), 1024);
  }
  if (h == "compute") {
    Arity(s, 2);
    const Expression* e = BuildExpr(s.kids[1]);
    return Compute(e, BuildOp(s.kids[2], in));
  }
  if (h == "filter") {
    Arity(s, 3);
    const Expression* e = BuildExpr(s.kids[1]);
    const SingleSourceProjector* p = BuildProjector(s.kids[2]);
    return Filter(e, p, BuildOp(s.kids[3], in));
  }
  if (h == "project") {
    Arity(s, 2);
    const SingleSourceProjector* p = BuildProjector(s.kids[1]);
    return Project(p, BuildOp(s.kids[2], in));
  }
  if (h == "group") {
    Arity(s, 3);
    const SingleSourceProjector* p = BuildProjector(s.kids[1]);
    AggregationSpecification* a = BuildAggs(s.kids[2]);
    return GroupAggregate(p, a, NULL, BuildOp(s.kids[3], in));
  }
  if (h == "group_opts") {
    // (group_opts <memory quota | none> <estimated result rows | none> <best effort: 0 | 1> <allocator quota | none> proj aggs child)
    Arity(s, 7);
    GroupAggregateOptions* options = new GroupAggregateOptions();
    if (Atom(s.kids[1]) != "none") options->set_memory_quota(static_cast<size_t>(strtoull(Atom(s.kids[1]).c_str(), NULL, 10)));
    if (Atom(s.kids[2]) != "none") options->set_estimated_result_row_count(static_cast<size_t>(strtoull(Atom(s.kids[2]).c_str(), NULL, 10)));
    const bool best_effort = Atom(s.kids[3]) == "1";
    const SingleSourceProjector* p = BuildProjector(s.kids[5]);
    AggregationSpecification* a = BuildAggs(s.kids[6]);
    Operation* child = BuildOp(s.kids[7], in);
    Operation* op = best_effort ? BestEffortGroupAggregate(p, a, options, child) : GroupAggregate(p, a, options, child);
    if (Atom(s.kids[4]) != "none") {
      // lives as long as the process: operations keep a bare pointer to their allocator
      op->SetBufferAllocator(new MemoryLimit(static_cast<size_t>(strtoull(Atom(s.kids[4]).c_str(), NULL, 10)), HeapBufferAllocator::Get()), false);
    }
    return op;
  }
  if (h == "limit") {
    // (limit OFFSET LIMIT child): limit.h:27
    Arity(s, 3);
    return Limit(static_cast<rowcount_t>(strtoull(Atom(s.kids[1]).c_str(), NULL, 10)),
                 static_cast<rowcount_t>(strtoull(Atom(s.kids[2]).c_str(), NULL, 10)), BuildOp(s.kids[3], in));
  }
  if (h == "coalesce") {
    // (coalesce child child ...): coalesce.h:30
    std::vector<Operation*> children;
    for (size_t i = 1; i < s.kids.size(); ++i) children.push_back(BuildOp(s.kids[i], in));
    return Coalesce(children);
  }
  if (h == "hybrid_group") {
    // (hybrid_group <memory quota> proj aggs child): aggregate.h:320
    Arity(s, 4);
    const size_t quota = static_cast<size_t>(strtoull(Atom(s.kids[1]).c_str(), NULL, 10));
    const SingleSourceProjector* p = BuildProjector(s.kids[2]);
    AggregationSpecification* a = BuildAggs(s.kids[3]);
    mkdir("/tmp/ssplan_hybrid", 0700);   // the reference spills under this directory
    return HybridGroupAggregate(p, a, quota, "/tmp/ssplan_hybrid", BuildOp(s.kids[4], in));
  }
  if (h == "scalar_agg") {
    Arity(s, 2);
    AggregationSpecification* a = BuildAggs(s.kids[1]);
    return ScalarAggregate(a, BuildOp(s.kids[2], in));
  }
  if (h == "hash_join") {
    Arity(s, 7);
    const std::string& jt = Atom(s.kids[1]);
    JoinType join_type;
    if (jt == "INNER") join_type = INNER;
    else if (jt == "LEFT_OUTER") join_type = LEFT_OUTER;
    else if (jt == "RIGHT_OUTER") join_type = RIGHT_OUTER;
    else if (jt == "FULL_OUTER") join_type = FULL_OUTER;
    else throw ParseError{"unknown join type"};
    const SingleSourceProjector* lk = BuildProjector(s.kids[2]);
    const SingleSourceProjector* rk = BuildProjector(s.kids[3]);
    const MultiSourceProjector* rp = BuildMultiProjector(s.kids[4]);
    const std::string& u = Atom(s.kids[5]);
    if (u != "UNIQUE" && u != "NOT_UNIQUE") throw ParseError{"unknown key uniqueness"};
    Operation* l = BuildOp(s.kids[6], in);
    Operation* r = BuildOp(s.kids[7], in);
    return new HashJoinOperation(join_type, lk, rk, rp, u == "UNIQUE" ? UNIQUE : NOT_UNIQUE, l, r);
  }
  if (h == "merge_union_all") {
    if (s.kids.size() < 2 || Head(s.kids[1]) != "order") throw ParseError{"expected (merge_union_all (order ...) OP...)"};
    std::unique_ptr<SortOrder> order(BuildOrder(s.kids[1]));
    std::vector<Operation*> inputs;
    for (size_t i = 2; i < s.kids.size(); ++i) inputs.push_back(BuildOp(s.kids[i], in));
    return MergeUnionAll(order.release(), inputs);
  }
  if (h == "aggregate_clusters") {
    Arity(s, 3);
    const SingleSourceProjector* p = BuildProjector(s.kids[1]);
    AggregationSpecification* a = BuildAggs(s.kids[2]);
    return AggregateClusters(p, a, BuildOp(s.kids[3], in));
  }
  if (h == "sort") {
    Arity(s, 3);
    if (Head(s.kids[1]) != "order") throw ParseError{"expected (order ...)"};
    std::unique_ptr<SortOrder> order(new SortOrder);
    for (size_t i = 1; i < s.kids[1].kids.size(); ++i) {
      const Sx& k = s.kids[1].kids[i];
      if (k.atom || k.kids.size() != 2) throw ParseError{"order takes (NAME ASC|DESC) pairs"};
      const std::string& d = Atom(k.kids[1]);
      if (d != "ASC" && d != "DESC") throw ParseError{"order direction must be ASC or DESC"};
      order->add(ProjectNamedAttribute(Atom(k.kids[0])), d == "ASC" ? ASCENDING : DESCENDING);
    }
    const SingleSourceProjector* p = BuildProjector(s.kids[2]);
    return Sort(order.release(), p, static_cast<size_t>(1) << 40, BuildOp(s.kids[3], in));
  }
  if (h == "extended_sort") {
    // (extended_sort (order (NAME ASC|DESC)...) LIMIT|none PROJECTOR CHILD): sort.h:103-107
    Arity(s, 4);
    if (Head(s.kids[1]) != "order") throw ParseError{"expected (order ...)"};
    std::unique_ptr<ExtendedSortSpecification> spec(new ExtendedSortSpecification);
    for (size_t i = 1; i < s.kids[1].kids.size(); ++i) {
      const Sx& k = s.kids[1].kids[i];
      if (k.atom || k.kids.size() != 2) throw ParseError{"order takes (NAME ASC|DESC) pairs"};
      const std::string& d = Atom(k.kids[1]);
      if (d != "ASC" && d != "DESC") throw ParseError{"order direction must be ASC or DESC"};
      ExtendedSortSpecification::Key* key = spec->add_keys();
      key->set_attribute_name(Atom(k.kids[0]));
      key->set_column_order(d == "ASC" ? ASCENDING : DESCENDING);
    }
    const std::string& lim = Atom(s.kids[2]);
    if (lim != "none") spec->set_limit(static_cast<uint64_t>(strtoull(lim.c_str(), NULL, 10)));
    const SingleSourceProjector* p = BuildProjector(s.kids[3]);
    return ExtendedSort(spec.release(), p, static_cast<size_t>(1) << 40, BuildOp(s.kids[4], in));
  }
  throw ParseError{"unknown operation '" + h + "'"};
}

// A failure raised while cursors are created bottom-up through the Bound* factories.
struct BindError {
  int code;
  std::string msg;
};

struct Keep {
  std::vector<std::unique_ptr<Operation> > ops;   // operations outlive the cursors created from them
};

template <typename T>
T* Take(FailureOrOwned<T> r) {
  if (r.is_failure()) throw BindError{r.exception().return_code(), r.exception().message()};
  return r.release();
}

// The same expression grammar as BuildExpr, built bottom-up through the BOUND factories (BoundNamedAttribute,
// BoundConst*, BoundPlus, BoundLess, BoundIf ...: expression/core/*_bound_expressions.h,
// expression/infrastructure/terminal_bound_expressions.h) against `schema`.
BoundExpression* BuildBoundExpr(const Sx& s, const TupleSchema& schema) {
  const std::string& h = Head(s);
  BufferAllocator* heap = HeapBufferAllocator::Get();
  const rowcount_t cap = Cursor::kDefaultRowCount;
  if (h == "col") { Arity(s, 1); return Take(BoundNamedAttribute(schema, Atom(s.kids[1]))); }
  if (h == "at") { Arity(s, 1); return Take(BoundAttributeAt(schema, static_cast<size_t>(atoi(Atom(s.kids[1]).c_str())))); }
  if (h == "i32") { Arity(s, 1); return Take(BoundConstInt32(static_cast<int32>(strtoll(Atom(s.kids[1]).c_str(), NULL, 0)), heap, cap)); }
  if (h == "i64") { Arity(s, 1); return Take(BoundConstInt64(strtoll(Atom(s.kids[1]).c_str(), NULL, 0), heap, cap)); }
  if (h == "u32") { Arity(s, 1); return Take(BoundConstUInt32(static_cast<uint32>(strtoull(Atom(s.kids[1]).c_str(), NULL, 0)), heap, cap)); }
  if (h == "u64") { Arity(s, 1); return Take(BoundConstUInt64(strtoull(Atom(s.kids[1]).c_str(), NULL, 0), heap, cap)); }
  if (h == "f32") { Arity(s, 1); return Take(BoundConstFloat(strtof(Atom(s.kids[1]).c_str(), NULL), heap, cap)); }
  if (h == "f64") { Arity(s, 1); return Take(BoundConstDouble(strtod(Atom(s.kids[1]).c_str(), NULL), heap, cap)); }
  if (h == "bool") { Arity(s, 1); return Take(BoundConstBool(Atom(s.kids[1]) == "true" || Atom(s.kids[1]) == "1", heap, cap)); }
  if (h == "str") { Arity(s, 1); return Take(BoundConstString(Atom(s.kids[1]), heap, cap)); }
  if (h == "null") { Arity(s, 1); return Take(BoundNull(ParseType(Atom(s.kids[1])), heap, cap)); }
  if (h == "cast") { Arity(s, 2); return Take(BoundCastTo(ParseType(Atom(s.kids[1])), BuildBoundExpr(s.kids[2], schema), heap, cap)); }
  if (h == "as") { Arity(s, 2); return Take(BoundAlias(Atom(s.kids[1]), BuildBoundExpr(s.kids[2], schema), heap, cap)); }
  if (h == "if" || h == "nulling_if") {
    Arity(s, 3);
    BoundExpression* c = BuildBoundExpr(s.kids[1], schema);
    BoundExpression* a = BuildBoundExpr(s.kids[2], schema);
    BoundExpression* b = BuildBoundExpr(s.kids[3], schema);
    return h == "if" ? Take(BoundIf(c, a, b, heap, cap)) : Take(BoundIfNulling(c, a, b, heap, cap));
  }
  if (h == "compound") {
    std::unique_ptr<BoundExpressionList> list(new BoundExpressionList);
    for (size_t i = 1; i < s.kids.size(); ++i) list->add(BuildBoundExpr(s.kids[i], schema));
    return Take(BoundCompoundExpression(list.release()));
  }
  typedef FailureOrOwned<BoundExpression> (*B1)(BoundExpression*, BufferAllocator*, rowcount_t);
  typedef FailureOrOwned<BoundExpression> (*B2)(BoundExpression*, BoundExpression*, BufferAllocator*, rowcount_t);
  static const struct { const char* name; B1 fn; } unary[] = {
      {"negate", &BoundNegate}, {"not", &BoundNot}, {"is_null", &BoundIsNull}, {"bitwise_not", &BoundBitwiseNot},
      {"is_odd", &BoundIsOdd}, {"is_even", &BoundIsEven}};
  static const struct { const char* name; B2 fn; } binary[] = {
      {"plus", &BoundPlus}, {"minus", &BoundMinus}, {"multiply", &BoundMultiply}, {"divide_signaling", &BoundDivideSignaling},
      {"divide_nulling", &BoundDivideNulling}, {"divide_quiet", &BoundDivideQuiet}, {"cpp_divide_signaling", &BoundCppDivideSignaling},
      {"cpp_divide_nulling", &BoundCppDivideNulling}, {"modulus_signaling", &BoundModulusSignaling},
      {"modulus_nulling", &BoundModulusNulling}, {"equal", &BoundEqual}, {"not_equal", &BoundNotEqual}, {"less", &BoundLess},
      {"less_or_equal", &BoundLessOrEqual}, {"greater", &BoundGreater}, {"greater_or_equal", &BoundGreaterOrEqual},
      {"and", &BoundAnd}, {"or", &BoundOr}, {"and_not", &BoundAndNot}, {"xor", &BoundXor}, {"bitwise_and", &BoundBitwiseAnd},
      {"bitwise_or", &BoundBitwiseOr}, {"bitwise_xor", &BoundBitwiseXor}, {"bitwise_and_not", &BoundBitwiseAndNot},
      {"shift_left", &BoundShiftLeft}, {"shift_right", &BoundShiftRight}, {"if_null", &BoundIfNull}};
  for (size_t i = 0; i < sizeof(unary) / sizeof(unary[0]); ++i) {
    if (h == unary[i].name) { Arity(s, 1); return Take(unary[i].fn(BuildBoundExpr(s.kids[1], schema), heap, cap)); }
  }
  for (size_t i = 0; i < sizeof(binary) / sizeof(binary[0]); ++i) {
    if (h == binary[i].name) {
      Arity(s, 2);
      BoundExpression* a = BuildBoundExpr(s.kids[1], schema);
      BoundExpression* b = BuildBoundExpr(s.kids[2], schema);
      return Take(binary[i].fn(a, b, heap, cap));
    }
  }
  throw ParseError{"unknown bound expression '" + h + "'"};
}

Cursor* BuildCursor(const Sx& s, const Inputs& in, Keep* keep) {
  const std::string& h = Head(s);
  BufferAllocator* heap = HeapBufferAllocator::Get();
  // bound_bx_compute / bound_bx_filter: as bound_compute / bound_filter, but the expression is assembled from the
  // bound factories (BuildBoundExpr) and wrapped with CreateBoundExpressionTree
  if (h == "bound_bx_compute") {
    Arity(s, 2);
    std::unique_ptr<Cursor> child(BuildCursor(s.kids[2], in, keep));
    BoundExpressionTree* tree = Take(CreateBoundExpressionTree(BuildBoundExpr(s.kids[1], child->schema()), heap, Cursor::kDefaultRowCount));
    return Take(BoundCompute(tree, heap, Cursor::kDefaultRowCount, child.release()));
  }
  if (h == "bound_bx_filter") {
    Arity(s, 3);
    std::unique_ptr<const SingleSourceProjector> p(BuildProjector(s.kids[2]));
    std::unique_ptr<Cursor> child(BuildCursor(s.kids[3], in, keep));
    std::unique_ptr<BoundExpressionTree> tree(
        Take(CreateBoundExpressionTree(BuildBoundExpr(s.kids[1], child->schema()), heap, Cursor::kDefaultRowCount)));
    const BoundSingleSourceProjector* bp = Take(p->Bind(child->schema()));
    return Take(BoundFilter(tree.release(), bp, heap, child.release()));
  }
  if (h == "bound_file_read") {
    // (bound_file_read PATH N): FileInput over a file holding rows of table N's schema, as the source of a cursor tree
    Arity(s, 2);
    size_t n = static_cast<size_t>(atoi(Atom(s.kids[2]).c_str()));
    if (n >= in.views.size()) throw ParseError{"bound_file_read: no such table"};
    if (!File::Exists(Atom(s.kids[1]))) throw BindError{ERROR_GENERAL_IO_ERROR, "no such file: " + Atom(s.kids[1])};
    return Take(FileInput(in.views[n].schema(), File::OpenOrDie(Atom(s.kids[1]), "r"), false, heap));
  }
  if (h == "bound_scan") {
    Arity(s, 1);
    size_t n = static_cast<size_t>(atoi(Atom(s.kids[1]).c_str()));
    if (n >= in.views.size()) throw ParseError{"scan: no such table"};
    return BoundScanView(in.views[n]);
  }
  if (h == "bound_compute") {
    Arity(s, 2);
    std::unique_ptr<const Expression> e(BuildExpr(s.kids[1]));
    std::unique_ptr<Cursor> child(BuildCursor(s.kids[2], in, keep));
    BoundExpressionTree* tree = Take(e->Bind(child->schema(), heap, Cursor::kDefaultRowCount));
    return Take(BoundCompute(tree, heap, Cursor::kDefaultRowCount, child.release()));
  }
  if (h == "bound_filter") {
    Arity(s, 3);
    std::unique_ptr<const Expression> e(BuildExpr(s.kids[1]));
    std::unique_ptr<const SingleSourceProjector> p(BuildProjector(s.kids[2]));
    std::unique_ptr<Cursor> child(BuildCursor(s.kids[3], in, keep));
    BoundExpressionTree* tree = Take(e->Bind(child->schema(), heap, Cursor::kDefaultRowCount));
    std::unique_ptr<BoundExpressionTree> tree_owner(tree);
    const BoundSingleSourceProjector* bp = Take(p->Bind(child->schema()));
    return Take(BoundFilter(tree_owner.release(), bp, heap, child.release()));
  }
  if (h == "bound_project") {
    Arity(s, 2);
    std::unique_ptr<const SingleSourceProjector> p(BuildProjector(s.kids[1]));
    std::unique_ptr<Cursor> child(BuildCursor(s.kids[2], in, keep));
    const BoundSingleSourceProjector* bp = Take(p->Bind(child->schema()));
    return BoundProject(bp, child.release());
  }
  if (h == "bound_group" || h == "bound_scalar_agg") {
    const bool scalar = h == "bound_scalar_agg";
    Arity(s, scalar ? 2 : 3);
    std::unique_ptr<const SingleSourceProjector> p(scalar ? NULL : BuildProjector(s.kids[1]));
    std::unique_ptr<AggregationSpecification> a(BuildAggs(s.kids[scalar ? 1 : 2]));
    std::unique_ptr<Cursor> child(BuildCursor(s.kids[scalar ? 2 : 3], in, keep));
    if (scalar) {
      Aggregator* agg = Take(Aggregator::Create(*a, child->schema(), heap, 1));
      return BoundScalarAggregate(agg, child.release());
    }
    std::unique_ptr<BufferAllocator> limit(new MemoryLimit(static_cast<size_t>(1) << 40, heap));
    std::unique_ptr<const BoundSingleSourceProjector> bp(Take(p->Bind(child->schema())));
    Aggregator* agg = Take(Aggregator::Create(*a, child->schema(), limit.get(), 16));
    return Take(BoundGroupAggregate(bp.release(), agg, limit.release(), heap, false, child.release()));
  }
  if (h == "bound_aggregate_clusters") {
    Arity(s, 3);
    std::unique_ptr<const SingleSourceProjector> p(BuildProjector(s.kids[1]));
    std::unique_ptr<AggregationSpecification> a(BuildAggs(s.kids[2]));
    std::unique_ptr<Cursor> child(BuildCursor(s.kids[3], in, keep));
    std::unique_ptr<const BoundSingleSourceProjector> bp(Take(p->Bind(child->schema())));
    Aggregator* agg = Take(Aggregator::Create(*a, child->schema(), heap, 16));
    return Take(BoundAggregateClusters(bp.release(), agg, heap, child.release()));
  }
  if (h == "bound_merge_union_all") {
    if (s.kids.size() < 3 || Head(s.kids[1]) != "order") throw ParseError{"expected (bound_merge_union_all (order ...) OP...)"};
    std::unique_ptr<SortOrder> order(BuildOrder(s.kids[1]));
    std::vector<Cursor*> inputs;
    for (size_t i = 2; i < s.kids.size(); ++i) inputs.push_back(BuildCursor(s.kids[i], in, keep));
    std::unique_ptr<const BoundSortOrder> bo(Take(order->Bind(inputs[0]->schema())));
    return Take(BoundMergeUnionAll(bo.release(), inputs, heap));
  }
  if (h == "bound_sort") {
    Arity(s, 3);
    if (Head(s.kids[1]) != "order") throw ParseError{"expected (order ...)"};
    std::unique_ptr<SortOrder> order(new SortOrder);
    for (size_t i = 1; i < s.kids[1].kids.size(); ++i) {
      const Sx& k = s.kids[1].kids[i];
      if (k.atom || k.kids.size() != 2) throw ParseError{"order takes (NAME ASC|DESC) pairs"};
      order->add(ProjectNamedAttribute(Atom(k.kids[0])), Atom(k.kids[1]) == "ASC" ? ASCENDING : DESCENDING);
    }
    std::unique_ptr<const SingleSourceProjector> p(BuildProjector(s.kids[2]));
    std::unique_ptr<Cursor> child(BuildCursor(s.kids[3], in, keep));
    std::unique_ptr<const BoundSortOrder> bo(Take(order->Bind(child->schema())));
    const BoundSingleSourceProjector* bp = Take(p->Bind(child->schema()));
    return Take(BoundSort(bo.release(), bp, static_cast<size_t>(1) << 40, "", heap, child.release()));
  }
  // an unbound operation below a bound one: create its cursor the usual way
  keep->ops.push_back(std::unique_ptr<Operation>(BuildOp(s, in)));
  return Take(keep->ops.back()->CreateCursor());
}

double WallNow() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + ts.tv_nsec * 1e-9;
}

}  // namespace

// A cursor that forwards everything to the cursor it owns (written against the public Cursor interface only).
class PassThroughCursor : public Cursor {
 public:
  explicit PassThroughCursor(Cursor* c) : c_(c) {}
  virtual const TupleSchema& schema() const { return c_->schema(); }
  virtual ResultView Next(rowcount_t max_row_count) { return c_->Next(max_row_count); }
  virtual void Interrupt() { c_->Interrupt(); }
  virtual bool IsWaitingOnBarrierSupported() const { return c_->IsWaitingOnBarrierSupported(); }
  virtual void ApplyToChildren(CursorTransformer* transformer) { c_->ApplyToChildren(transformer); }
  virtual void AppendDebugDescription(std::string* target) const { c_->AppendDebugDescription(target); }
  virtual CursorId GetCursorId() const { return c_->GetCursorId(); }
 private:
  std::unique_ptr<Cursor> c_;
};
class PassThroughTransformer : public CursorTransformer {
 public:
  PassThroughTransformer() : wrapped_(0) {}
  virtual Cursor* Transform(Cursor* cursor) { ++wrapped_; return new PassThroughCursor(cursor); }
  int64_t wrapped() const { return wrapped_; }
 private:
  int64_t wrapped_;
};

struct ssplan_result {
  int64_t spied_children;
  int code;
  std::string error;
  struct Col {
    std::string name;
    int dtype;
    int nullable;
    size_t width;
    bool saw_nulls;
    bool varlen;               // STRING / BINARY: `data` holds one int64 length per row, `bytes` the cells end to end
    std::vector<char> data;
    std::vector<char> bytes;
    std::vector<uint8_t> is_null;
  };
  std::vector<Col> cols;
  int64_t rows;
  double create_s, drain_s;
  int64_t next_calls;
  ssplan_result() : spied_children(0), code(0), rows(0), create_s(0), drain_s(0), next_calls(0) {}
};

namespace {
void DescribeColumns(ssplan_result* r, const TupleSchema& schema) {
  r->cols.resize(schema.attribute_count());
  for (int c = 0; c < schema.attribute_count(); ++c) {
    const Attribute& a = schema.attribute(c);
    r->cols[c].name = a.name();
    r->cols[c].dtype = a.type();
    r->cols[c].nullable = a.is_nullable() ? 1 : 0;
    r->cols[c].width = GetTypeInfo(a.type()).size();
    r->cols[c].saw_nulls = false;
    r->cols[c].varlen = a.type() == STRING || a.type() == BINARY;
  }
}

void AppendView(ssplan_result* r, const View& v, int32_t flags) {
  const size_t n = v.row_count();
  if (!(flags & SSPLAN_DISCARD)) {
    for (size_t c = 0; c < r->cols.size(); ++c) {
      ssplan_result::Col& col = r->cols[c];
      const char* src = static_cast<const char*>(v.column(c).data().raw());
      const bool* nulls = v.column(c).is_null();
      if (col.varlen) {   // deep copy: the cells point into memory the cursor may reuse
        const StringPiece* cells = reinterpret_cast<const StringPiece*>(src);
        for (size_t i = 0; i < n; ++i) {
          const int64_t len = (nulls != NULL && nulls[i]) ? 0 : static_cast<int64_t>(cells[i].size());
          const char* lp = reinterpret_cast<const char*>(&len);
          col.data.insert(col.data.end(), lp, lp + sizeof(len));
          if (len > 0) col.bytes.insert(col.bytes.end(), cells[i].data(), cells[i].data() + len);
        }
      } else {
        col.data.insert(col.data.end(), src, src + n * col.width);
      }
      if (nulls != NULL) {
        if (!col.saw_nulls) {
          col.is_null.assign(r->rows, 0);
          col.saw_nulls = true;
        }
        const uint8_t* nb = reinterpret_cast<const uint8_t*>(nulls);
        col.is_null.insert(col.is_null.end(), nb, nb + n);
      } else if (col.saw_nulls) {
        col.is_null.insert(col.is_null.end(), n, 0);
      }
    }
  }
  r->rows += n;
}

// (evaluate EXPR N [CAPACITY]): Expression::Bind + BoundExpressionTree::Evaluate over slices of the table.
int RunEvaluate(const Sx& sx, const Inputs& in, int32_t flags, ssplan_result* r) {
  if (sx.kids.size() < 3 || sx.kids.size() > 5) throw ParseError{"evaluate takes EXPR TABLE [CAPACITY [SLICE]]"};
  std::unique_ptr<const Expression> e(BuildExpr(sx.kids[1]));
  const size_t n = static_cast<size_t>(atoi(Atom(sx.kids[2]).c_str()));
  if (n >= in.views.size()) throw ParseError{"evaluate: no such table"};
  const View& table = in.views[n];
  rowcount_t capacity = sx.kids.size() >= 4 ? static_cast<rowcount_t>(atoll(Atom(sx.kids[3]).c_str())) : 0;
  if (capacity == 0) capacity = table.row_count() > 0 ? table.row_count() : 1;
  // SLICE > CAPACITY asks Evaluate for more rows than the tree was bound for (ERROR_TOO_MANY_ROWS, expression.cc:57-66)
  const rowcount_t step = sx.kids.size() == 5 ? static_cast<rowcount_t>(atoll(Atom(sx.kids[4]).c_str())) : capacity;
  if (step == 0) throw ParseError{"evaluate: SLICE must be positive"};
  double t0 = WallNow();
  FailureOrOwned<BoundExpressionTree> bound = e->Bind(table.schema(), HeapBufferAllocator::Get(), capacity);
  r->create_s = WallNow() - t0;
  if (bound.is_failure()) {
    r->code = bound.exception().return_code();
    r->error = bound.exception().message();
    return r->code;
  }
  std::unique_ptr<BoundExpressionTree> tree(bound.release());
  DescribeColumns(r, tree->result_schema());
  if (r->code != 0 || (flags & SSPLAN_BIND_ONLY)) return r->code;
  t0 = WallNow();
  View slice(table.schema());
  for (rowcount_t first = 0; first < table.row_count() || (first == 0 && table.row_count() == 0); first += step) {
    const rowcount_t rows = table.row_count() - first < step ? table.row_count() - first : step;
    slice.ResetFromSubRange(table, first, rows);
    EvaluationResult result = tree->Evaluate(slice);
    ++r->next_calls;
    if (result.is_failure()) {
      r->code = result.exception().return_code();
      r->error = result.exception().message();
      break;
    }
    AppendView(r, result.get(), flags);
    if (table.row_count() == 0) break;
  }
  r->drain_s = WallNow() - t0;
  return r->code;
}
// (bx_evaluate EXPR N): the expression assembled from the bound factories, evaluated through the virtual
// BoundExpression::DoEvaluate(view, skip vectors) (expression/base/expression.h:60-66) with nothing skipped. At most
// Cursor::kDefaultRowCount rows (the capacity the factories are given).
int RunBoundEvaluate(const Sx& sx, const Inputs& in, int32_t flags, ssplan_result* r) {
  Arity(sx, 2);
  const size_t n = static_cast<size_t>(atoi(Atom(sx.kids[2]).c_str()));
  if (n >= in.views.size()) throw ParseError{"bx_evaluate: no such table"};
  const View& table = in.views[n];
  if (table.row_count() > Cursor::kDefaultRowCount) throw ParseError{"bx_evaluate: more rows than the expression's capacity"};
  std::unique_ptr<BoundExpression> e;
  try {
    e.reset(BuildBoundExpr(sx.kids[1], table.schema()));
  } catch (const BindError& b) {
    r->code = b.code;
    r->error = b.msg;
    return r->code;
  }
  DescribeColumns(r, e->result_schema());
  if (r->code != 0 || (flags & SSPLAN_BIND_ONLY)) return r->code;
  const int cols = e->result_schema().attribute_count();
  BoolView skip(cols);
  std::vector<std::unique_ptr<bool[]> > store;
  for (int c = 0; c < cols; ++c) {
    store.push_back(std::unique_ptr<bool[]>(new bool[table.row_count() + 1]()));
    skip.ResetColumn(c, store.back().get());
  }
  skip.set_row_count(table.row_count());
  EvaluationResult result = e->DoEvaluate(table, skip);
  ++r->next_calls;
  if (result.is_failure()) {
    r->code = result.exception().return_code();
    r->error = result.exception().message();
    return r->code;
  }
  AppendView(r, result.get(), flags);
  return r->code;
}
}  // namespace

extern "C" {

int ssplan_run(const char* plan, int32_t ntables, const ssplan_table* tables,
               int64_t next_max_rows, int32_t flags, ssplan_result** out) {
  ssplan_result* r = new ssplan_result;
  *out = r;
  Inputs in;
  for (int t = 0; t < ntables; ++t) {
    TupleSchema schema;
    for (int c = 0; c < tables[t].ncols; ++c) {
      const ssplan_column& col = tables[t].cols[c];
      if (!schema.add_attribute(Attribute(col.name, static_cast<DataType>(col.dtype),
                                          col.nullable ? NULLABLE : NOT_NULLABLE))) {
        r->code = ERROR_ATTRIBUTE_EXISTS;
        r->error = "duplicate input column name";
        return r->code;
      }
    }
    View v(schema);
    v.set_row_count(tables[t].rows);
    for (int c = 0; c < tables[t].ncols; ++c) {
      const ssplan_column& col = tables[t].cols[c];
      v.mutable_column(c)->Reset(col.data, reinterpret_cast<const bool*>(col.is_null));
    }
    in.views.push_back(v);
  }

  std::unique_ptr<Operation> op;
  Keep keep;
  std::unique_ptr<Cursor> cursor;
  double t0 = WallNow();
  try {
    Sx sx = SxParser(plan).Parse();
    if (Head(sx) == "evaluate") return RunEvaluate(sx, in, flags, r);
    if (Head(sx) == "bx_evaluate") return RunBoundEvaluate(sx, in, flags, r);
    if (Head(sx) == "file_write" || Head(sx) == "file_read") {
      // (file_write PATH <operation>): the operation's rows are written in the reference's block format
      // (cursor/infrastructure/file_io.h: WriteCursor into FileOutput), then the file is scanned back (FileInput);
      // (file_read PATH N): scans a file that holds rows of table N's schema.
      Arity(sx, 2);
      const std::string path = Atom(sx.kids[1]);
      TupleSchema schema;
      if (Head(sx) == "file_write") {
        std::unique_ptr<Operation> source(BuildOp(sx.kids[2], in));
        FailureOrOwned<Cursor> created = source->CreateCursor();
        if (created.is_failure()) throw BindError{created.exception().return_code(), created.exception().message()};
        schema = created->schema();
        std::unique_ptr<Sink> sink(FileOutput(File::OpenOrDie(path, "w"), TAKE_OWNERSHIP));
        FailureOrVoid written = WriteCursor(created.release(), sink.get());
        FailureOrVoid finalized = sink->Finalize();   // closes the file
        if (written.is_failure()) throw BindError{written.exception().return_code(), written.exception().message()};
        if (finalized.is_failure()) throw BindError{finalized.exception().return_code(), finalized.exception().message()};
      } else {
        const size_t n = static_cast<size_t>(atoi(Atom(sx.kids[2]).c_str()));
        if (n >= in.views.size()) throw ParseError{"file_read: no such table"};
        schema = in.views[n].schema();
      }
      if (!File::Exists(path)) throw BindError{ERROR_GENERAL_IO_ERROR, "no such file: " + path};
      FailureOrOwned<Cursor> scan = FileInput(schema, File::OpenOrDie(path, "r"), false, HeapBufferAllocator::Get());
      if (scan.is_failure()) throw BindError{scan.exception().return_code(), scan.exception().message()};
      cursor.reset(scan.release());
    } else if (Head(sx).compare(0, 6, "bound_") == 0) {
      t0 = WallNow();
      cursor.reset(BuildCursor(sx, in, &keep));
      r->create_s = WallNow() - t0;
    } else {
      op.reset(BuildOp(sx, in));
    }
  } catch (const ParseError& e) {
    r->code = ERROR_BAD_PROTO;
    r->error = "plan parse error: " + e.msg;
    return r->code;
  } catch (const BindError& e) {
    r->code = e.code;
    r->error = e.msg;
    return r->code;
  }

  if (!cursor) {
    t0 = WallNow();
    FailureOrOwned<Cursor> created = op->CreateCursor();
    r->create_s = WallNow() - t0;
    if (created.is_failure()) {
      r->code = created.exception().return_code();
      r->error = created.exception().message();
      return r->code;
    }
    cursor.reset(created.release());
  }

  DescribeColumns(r, cursor->schema());
  if (r->code != 0) return r->code;

  if (flags & SSPLAN_BIND_ONLY) return r->code;

  // SSPLAN_SPY: the cursor tree's seams, exercised the way the reference's tests do with their spy cursors
  // (cursor/core/spy.h, e.g. aggregate_clusters_test.cc:84-103): every child of the root is handed to a
  // transformer that wraps it in a pass-through cursor, then the root itself is wrapped; the result must not change.
  int64_t spied_children = 0;
  if (flags & SSPLAN_SPY) {
    PassThroughTransformer spy;
    cursor->ApplyToChildren(&spy);
    spied_children = spy.wrapped();
    cursor.reset(spy.Transform(cursor.release()));
  }
  r->spied_children = spied_children;

  const rowcount_t max_rows =
      next_max_rows > 0 ? static_cast<rowcount_t>(next_max_rows) : Cursor::kDefaultRowCount;
  t0 = WallNow();
  for (;;) {
    ResultView rv = cursor->Next(max_rows);
    ++r->next_calls;
    if (rv.is_failure()) {
      r->code = rv.exception().return_code();
      r->error = rv.exception().message();
      break;
    }
    if (!rv.has_data()) {
      if (rv.is_eos()) break;
      r->code = WAITING_ON_BARRIER;
      r->error = "unexpected WAITING_ON_BARRIER";
      break;
    }
    AppendView(r, rv.view(), flags);
  }
  r->drain_s = WallNow() - t0;
  return r->code;
}

int ssplan_result_code(const ssplan_result* r) { return r->code; }
const char* ssplan_result_error(const ssplan_result* r) { return r->error.c_str(); }
int32_t ssplan_result_ncols(const ssplan_result* r) { return static_cast<int32_t>(r->cols.size()); }
int64_t ssplan_result_rows(const ssplan_result* r) { return r->rows; }
const char* ssplan_result_col_name(const ssplan_result* r, int32_t i) { return r->cols[i].name.c_str(); }
int32_t ssplan_result_col_dtype(const ssplan_result* r, int32_t i) { return r->cols[i].dtype; }
int32_t ssplan_result_col_nullable(const ssplan_result* r, int32_t i) { return r->cols[i].nullable; }
const void* ssplan_result_col_data(const ssplan_result* r, int32_t i) { return r->cols[i].data.data(); }
const char* ssplan_result_col_bytes(const ssplan_result* r, int32_t i) { return r->cols[i].bytes.data(); }
const uint8_t* ssplan_result_col_is_null(const ssplan_result* r, int32_t i) {
  return r->cols[i].saw_nulls ? r->cols[i].is_null.data() : NULL;
}
double ssplan_result_create_seconds(const ssplan_result* r) { return r->create_s; }
double ssplan_result_drain_seconds(const ssplan_result* r) { return r->drain_s; }
int64_t ssplan_result_next_calls(const ssplan_result* r) { return r->next_calls; }
int64_t ssplan_result_spied_children(const ssplan_result* r) { return r->spied_children; }
void ssplan_result_free(ssplan_result* r) { delete r; }

#ifndef SSPLAN_IMPL_NAME
#define SSPLAN_IMPL_NAME "reference"
#endif
const char* ssplan_impl(void) { return SSPLAN_IMPL_NAME; }

}  // extern "C"
