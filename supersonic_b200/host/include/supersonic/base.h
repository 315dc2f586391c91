// supersonic/base.h -- data model and error convention of the B200 implementation.
//
// A from-scratch mirror of the public surface that plan code touches in the reference's
// L1 layer: integral types, the proto enums, Exception / FailureOr*,
// DataType traits, Attribute / TupleSchema, Column / View / Block and BufferAllocator.
// Same names, argument meaning and error behaviour as the reference headers cited at each
// class; the implementation is new.
#ifndef SUPERSONIC_B200_HOST_BASE_H_
#define SUPERSONIC_B200_HOST_BASE_H_

#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <utility>
#include <vector>

// supersonic/utils/integral_types.h:24-43: the integral types are GLOBAL typedefs in the reference (client
// code such as test/guide/primer.cc writes plain `int32`), 64-bit ones are (unsigned) long long.
typedef int int32;
typedef long long int64;
typedef unsigned int uint32;
typedef unsigned long long uint64;

// supersonic/utils/strings/stringpiece.h (the subset plan code uses); global, as in the reference
class StringPiece {
 public:
  StringPiece() : ptr_(NULL), len_(0) {}
  StringPiece(const char* s) : ptr_(s), len_(s ? strlen(s) : 0) {}            // NOLINT
  StringPiece(const std::string& s) : ptr_(s.data()), len_(s.size()) {}       // NOLINT
  StringPiece(const char* s, size_t n) : ptr_(s), len_(n) {}
  const char* data() const { return ptr_; }
  size_t size() const { return len_; }
  size_t length() const { return len_; }
  bool empty() const { return len_ == 0; }
  std::string as_string() const { return ptr_ ? std::string(ptr_, len_) : std::string(); }
  std::string ToString() const { return as_string(); }
  bool operator==(const StringPiece& o) const {
    return len_ == o.len_ && (len_ == 0 || memcmp(ptr_, o.ptr_, len_) == 0);
  }
  bool operator!=(const StringPiece& o) const { return !(*this == o); }
  bool operator<(const StringPiece& o) const {
    const int r = memcmp(ptr_, o.ptr_, len_ < o.len_ ? len_ : o.len_);
    return r < 0 || (r == 0 && len_ < o.len_);
  }
 private:
  const char* ptr_;
  size_t len_;
};

// utils/strings/stringpiece.h:373: a StringPiece prints as its bytes
inline std::ostream& operator<<(std::ostream& o, const StringPiece& piece) {
  if (piece.size() > 0) o.write(piece.data(), static_cast<std::streamsize>(piece.size()));
  return o;
}

// supersonic/utils/logging-inl.h: the CHECK family client code uses (test/guide/join.cc). A failed check prints
// the streamed message and aborts, as LOG(FATAL) does.
namespace supersonic_b200_logging {
class CheckFailure {
 public:
  CheckFailure(const char* file, int line, const char* what) { std::cerr << file << ":" << line << ": Check failed: " << what << " "; }
  ~CheckFailure() { std::cerr << std::endl; abort(); }
  std::ostream& stream() { return std::cerr; }
};
struct Voidify { void operator&(std::ostream&) {} };
}  // namespace supersonic_b200_logging
#ifndef CHECK
#define CHECK(condition) \
  (condition) ? (void)0 : ::supersonic_b200_logging::Voidify() & ::supersonic_b200_logging::CheckFailure(__FILE__, __LINE__, #condition).stream()
#define CHECK_EQ(a, b) CHECK((a) == (b))
#define CHECK_NE(a, b) CHECK((a) != (b))
#define CHECK_LT(a, b) CHECK((a) < (b))
#define CHECK_LE(a, b) CHECK((a) <= (b))
#define CHECK_GT(a, b) CHECK((a) > (b))
#define CHECK_GE(a, b) CHECK((a) >= (b))
#define CHECK_NOTNULL(p) (p)
#define DCHECK(condition) CHECK(condition)
#define DCHECK_EQ(a, b) CHECK_EQ(a, b)
#endif

// utils/basictypes.h:14-17 (global, as in the reference): who closes a File handed to a sink
enum Ownership { DO_NOT_TAKE_OWNERSHIP, TAKE_OWNERSHIP };

namespace supersonic {

using std::string;
using std::vector;
using ::int32;
using ::int64;
using ::uint32;
using ::uint64;
using ::StringPiece;

// ---- enums of supersonic/proto/supersonic.proto:15-118 (same names and numbers) ---------
enum DataType {
  INT32 = 1, INT64 = 2, UINT32 = 8, UINT64 = 3, FLOAT = 9, DOUBLE = 5, BOOL = 6, DATE = 10,
  DATETIME = 4, STRING = 0, BINARY = 7, ENUM = 13, DATA_TYPE = 11
};
enum ReturnCode {
  OK = 0, END_OF_INPUT = 1, BEFORE_INPUT = 2, WAITING_ON_BARRIER = 3,
  ERROR_UNKNOWN_ERROR = 100, ERROR_GENERAL_IO_ERROR = 101, ERROR_MEMORY_EXCEEDED = 102,
  ERROR_NOT_IMPLEMENTED = 103, ERROR_EVALUATION_ERROR = 104, ERROR_BAD_PROTO = 105,
  ERROR_TEMP_FILE_CREATION_ERROR = 106,
  ERROR_TOO_FEW_ROWS = 301, ERROR_TOO_MANY_ROWS = 302, ERROR_TOO_FEW_COLUMNS = 303,
  ERROR_TOO_MANY_COLUMNS = 304, ERROR_DUPLICATED_UNIQUE_KEY = 305,
  ERROR_GENERIC_SCHEMA_ERROR = 400, ERROR_ATTRIBUTE_COUNT_MISMATCH = 401,
  ERROR_ATTRIBUTE_TYPE_MISMATCH = 402, ERROR_ATTRIBUTE_MISSING = 403,
  ERROR_ATTRIBUTE_EXISTS = 404, ERROR_INVALID_ARGUMENT_TYPE = 405,
  ERROR_ATTRIBUTE_IS_NULLABLE = 406, ERROR_INVALID_ARGUMENT_VALUE = 407,
  ERROR_ATTRIBUTE_AMBIGUOUS = 408,
  ERROR_DUPLICATE_ENUM_VALUE_NUMBER = 420, ERROR_DUPLICATE_ENUM_VALUE_NAME = 421,
  ERROR_UNDEFINED_ENUM_VALUE_NUMBER = 422, ERROR_UNDEFINED_ENUM_VALUE_NAME = 423,
  ERROR_FOREIGN_KEY_INVALID = 501, INTERRUPTED = 1000
};
enum Nullability { NOT_NULLABLE = 0, NULLABLE = 1 };
enum Aggregation { SUM = 0, MIN = 1, MAX = 2, COUNT = 3, CONCAT = 4, FIRST = 5, LAST = 6 };
enum ColumnOrder { ASCENDING = 0, DESCENDING = 1 };
enum JoinType { INNER = 0, LEFT_OUTER = 1, RIGHT_OUTER = 2, FULL_OUTER = 3 };
enum KeyUniqueness { NOT_UNIQUE = 0, UNIQUE = 1 };
// supersonic/cursor/proto/cursors.proto:13-62 (the ids this implementation reports)
enum CursorId {
  FILE_INPUT = 3, VIEW = 7, AGGREGATE_CLUSTERS = 8, COMPUTE = 15, FILTER = 16, GROUP_AGGREGATE = 20, HASH_JOIN = 21, MERGE_UNION_ALL = 24, PROJECT = 27,
  SCALAR_AGGREGATE = 30, SORT = 32, UNKNOWN_ID = 99
};

const string& DataType_Name(DataType t);
const string& ReturnCode_Name(ReturnCode c);
const string& Aggregation_Name(Aggregation a);
const string& JoinType_Name(JoinType j);

typedef uint64 rowcount_t;   // base/infrastructure/types.h:252-256
typedef int64 rowid_t;

// ---- Exception (base/exception/exception.h:53-137) ---------------------------------------
class Exception {
 public:
  Exception(const ReturnCode code, const string& message) : code_(code), message_(message) {}
  ReturnCode return_code() const { return code_; }
  const string& message() const { return message_; }
  void set_message(const string& message) { message_ = message; }
  Exception* Clone() const { return new Exception(*this); }
  string ToString() const { return ReturnCode_Name(code_) + ": " + message_; }
  string PrintStackTrace() const { return ToString() + "\n" + trace_; }
  Exception* AddStackTraceElement(const StringPiece& function, const StringPiece& filename, int line,
                                  const StringPiece& context);
 private:
  ReturnCode code_;
  string message_;
  string trace_;
};

// ---- FailureOr family (base/exception/result.h:43-126, utils/exception/failureor.h) -----
namespace result_internal {
struct FailureTag { Exception* exception; };
struct VoidSuccessTag {};
// Points at the argument of Success(); consumed inside the same full expression.
template <typename T> struct RefTag { T* ref; };
}  // namespace result_internal

inline result_internal::FailureTag Failure(Exception* e) { return result_internal::FailureTag{e}; }
inline result_internal::VoidSuccessTag Success() { return result_internal::VoidSuccessTag(); }
template <typename T> result_internal::RefTag<const T> Success(const T& v) { return result_internal::RefTag<const T>{&v}; }
template <typename T> result_internal::RefTag<T> Success(T& v) { return result_internal::RefTag<T>{&v}; }   // NOLINT

class FailureOrVoid {
 public:
  FailureOrVoid(result_internal::FailureTag f) : exception_(f.exception) {}   // NOLINT
  FailureOrVoid(result_internal::VoidSuccessTag) {}                           // NOLINT
  bool is_success() const { return !exception_; }
  bool is_failure() const { return !!exception_; }
  const Exception& exception() const { return *exception_; }
  Exception* release_exception() { return exception_.release(); }
  void mark_checked() const {}
 private:
  std::unique_ptr<Exception> exception_;
};

template <typename T>
class FailureOr {
 public:
  FailureOr(result_internal::FailureTag f) : exception_(f.exception), value_() {}   // NOLINT
  template <typename U> FailureOr(const result_internal::RefTag<U>& v) : value_(*v.ref) {}     // NOLINT
  bool is_success() const { return !exception_; }
  bool is_failure() const { return !!exception_; }
  const T& get() const { return value_; }
  const Exception& exception() const { return *exception_; }
  Exception* release_exception() { return exception_.release(); }
 private:
  std::unique_ptr<Exception> exception_;
  T value_;
};

template <typename T>
class FailureOrReference {
 public:
  FailureOrReference(result_internal::FailureTag f) : exception_(f.exception), ref_(NULL) {}  // NOLINT
  template <typename U> FailureOrReference(const result_internal::RefTag<U>& v) : ref_(v.ref) {}  // NOLINT
  bool is_success() const { return !exception_; }
  bool is_failure() const { return !!exception_; }
  T& get() const { return *ref_; }
  const Exception& exception() const { return *exception_; }
  Exception* release_exception() { return exception_.release(); }
 private:
  std::unique_ptr<Exception> exception_;
  T* ref_;
};

// Owns its result; copying transfers ownership (as the reference's propagators do).
template <typename T>
class FailureOrOwned {
 public:
  FailureOrOwned(result_internal::FailureTag f) : exception_(f.exception) {}   // NOLINT
  template <typename U> FailureOrOwned(const result_internal::RefTag<U>& v) : value_(*v.ref) {}     // NOLINT
  FailureOrOwned(const FailureOrOwned& o)
      : exception_(const_cast<FailureOrOwned&>(o).exception_.release()),
        value_(const_cast<FailureOrOwned&>(o).value_.release()) {}
  bool is_success() const { return !exception_; }
  bool is_failure() const { return !!exception_; }
  T* get() const { return value_.get(); }
  T* release() { return value_.release(); }
  T* operator->() const { return value_.get(); }
  T& operator*() const { return *value_; }
  const Exception& exception() const { return *exception_; }
  Exception* release_exception() { return exception_.release(); }
 private:
  void operator=(const FailureOrOwned&);
  std::unique_ptr<Exception> exception_;
  std::unique_ptr<T> value_;
};

void DieOnFailure(const Exception& e);
template <typename T> T* SucceedOrDie(FailureOrOwned<T> r) {
  if (r.is_failure()) DieOnFailure(r.exception());
  return r.release();
}
template <typename T> T SucceedOrDie(FailureOr<T> r) {
  if (r.is_failure()) DieOnFailure(r.exception());
  return r.get();
}
inline void SucceedOrDie(FailureOrVoid r) { if (r.is_failure()) DieOnFailure(r.exception()); }

// base/exception/exception_macros.h:42-82
#define THROW(exception_ptr)                                                              \
  return ::supersonic::Failure((exception_ptr)->AddStackTraceElement(__FUNCTION__, __FILE__, \
                                                                     __LINE__, "(thrown here)"))
// `result` is evaluated exactly once (it may be a call: a second evaluation would run the call
// again and, if that succeeded, dereference a NULL exception).
#define PROPAGATE_ON_FAILURE(result)                                                       \
  do {                                                                                     \
    auto&& propagated_result_ = (result);                                                  \
    if (propagated_result_.is_failure()) {                                                 \
      return ::supersonic::Failure(propagated_result_.release_exception()                  \
                                       ->AddStackTraceElement(__FUNCTION__, __FILE__,      \
                                                              __LINE__, #result));         \
    }                                                                                      \
  } while (0)

// ---- DataType traits (base/infrastructure/types.h:70-249) ---------------------------------
template <DataType type> struct TypeTraits;
#define SSB200_TYPE_TRAITS(DT, CPP)                                      \
  template <> struct TypeTraits<DT> { typedef CPP cpp_type; typedef CPP hold_type; static const DataType type = DT; }
SSB200_TYPE_TRAITS(INT32, int32);
SSB200_TYPE_TRAITS(INT64, int64);
SSB200_TYPE_TRAITS(UINT32, uint32);
SSB200_TYPE_TRAITS(UINT64, uint64);
SSB200_TYPE_TRAITS(FLOAT, float);
SSB200_TYPE_TRAITS(DOUBLE, double);
SSB200_TYPE_TRAITS(BOOL, bool);
SSB200_TYPE_TRAITS(DATE, int32);
SSB200_TYPE_TRAITS(DATETIME, int64);
SSB200_TYPE_TRAITS(ENUM, int32);
SSB200_TYPE_TRAITS(DATA_TYPE, DataType);
#undef SSB200_TYPE_TRAITS
// variable-length values are held (owned) as std::string: types.h:226-249
template <> struct TypeTraits<STRING> { typedef StringPiece cpp_type; typedef string hold_type; static const DataType type = STRING; };
template <> struct TypeTraits<BINARY> { typedef StringPiece cpp_type; typedef string hold_type; static const DataType type = BINARY; };

class TypeInfo {
 public:
  TypeInfo(DataType t, const char* name, size_t size, bool numeric, bool integer, bool fp, bool varlen)
      : type_(t), name_(name), size_(size), numeric_(numeric), integer_(integer), fp_(fp), varlen_(varlen) {}
  DataType type() const { return type_; }
  const string& name() const { return name_; }
  size_t size() const { return size_; }
  int log2_size() const { int l = 0; while ((size_t(1) << l) < size_) ++l; return l; }
  bool is_numeric() const { return numeric_; }
  bool is_integer() const { return integer_; }
  bool is_floating_point() const { return fp_; }
  bool is_variable_length() const { return varlen_; }
 private:
  DataType type_; string name_; size_t size_; bool numeric_, integer_, fp_, varlen_;
};
const TypeInfo& GetTypeInfo(DataType type);

// ---- Attribute / TupleSchema (base/infrastructure/tuple_schema.h) ------------------------
class Attribute {
 public:
  Attribute(const string& name, const DataType type, const Nullability nullability)
      : name_(name), type_(type), nullability_(nullability) {}
  const string& name() const { return name_; }
  DataType type() const { return type_; }
  Nullability nullability() const { return nullability_; }
  bool is_nullable() const { return nullability_ == NULLABLE; }
 private:
  string name_;
  DataType type_;
  Nullability nullability_;
};

class TupleSchema {
 public:
  TupleSchema() {}
  int attribute_count() const { return static_cast<int>(attributes_.size()); }
  const Attribute& attribute(const int position) const { return attributes_[position]; }
  // false (and no change) when an attribute of that name already exists
  bool add_attribute(const Attribute& attribute);
  // -1 when missing
  int LookupAttributePosition(const string& attribute_name) const;
  const Attribute& LookupAttribute(const string& name) const { return attributes_[LookupAttributePosition(name)]; }
  static TupleSchema Singleton(const string& name, const DataType type, Nullability nullability);
  static bool AreEqual(const TupleSchema& a, const TupleSchema& b, bool check_names);
  static bool CanMerge(const TupleSchema& a, const TupleSchema& b);
  static TupleSchema Merge(const TupleSchema& a, const TupleSchema& b);
  static FailureOr<TupleSchema> TryMerge(const TupleSchema& a, const TupleSchema& b);
  bool EqualByType(const TupleSchema& other) const;
  string GetHumanReadableSpecification() const;
 private:
  vector<Attribute> attributes_;
  std::map<string, int> positions_;
};

// ---- BufferAllocator (base/memory/memory.h:58-236, the seam kept API-compatible) ----------
class BufferAllocator;
class Buffer {
 public:
  ~Buffer();
  void* data() const { return data_; }
  size_t size() const { return size_; }
 private:
  friend class BufferAllocator;
  Buffer(void* data, size_t size, BufferAllocator* a) : data_(data), size_(size), allocator_(a) {}
  void* data_;
  size_t size_;
  BufferAllocator* allocator_;
};

class BufferAllocator {
 public:
  virtual ~BufferAllocator() {}
  // NULL when the request cannot be granted; zero-byte requests succeed (memory.h:112-116).
  Buffer* Allocate(size_t requested) { return BestEffortAllocate(requested, requested); }
  Buffer* BestEffortAllocate(size_t requested, size_t minimal);
  bool Reallocate(size_t requested, Buffer* buffer) { return BestEffortReallocate(requested, requested, buffer); }
  bool BestEffortReallocate(size_t requested, size_t minimal, Buffer* buffer);
  virtual size_t Available() const { return std::numeric_limits<size_t>::max(); }
 protected:
  BufferAllocator() {}
  // Grants a size in [minimal, requested] or 0 (= refuse). Default: everything.
  virtual size_t Grant(size_t requested, size_t minimal) { (void)minimal; return requested; }
  virtual void Release(size_t bytes) { (void)bytes; }
 private:
  friend class Buffer;
  friend class MemoryLimit;
};

class HeapBufferAllocator : public BufferAllocator {
 public:
  static HeapBufferAllocator* Get();
};

// Hard quota on top of a delegate (memory.h MemoryLimit): the fault-injection tool of the
// reference's tests.
class MemoryLimit : public BufferAllocator {
 public:
  explicit MemoryLimit(size_t quota, BufferAllocator* delegate = HeapBufferAllocator::Get())
      : quota_(quota), used_(0), delegate_(delegate) {}
  virtual size_t Available() const { return quota_ > used_ ? quota_ - used_ : 0; }
  size_t GetUsage() const { return used_; }
  size_t GetQuota() const { return quota_; }
 protected:
  virtual size_t Grant(size_t requested, size_t minimal);
  virtual void Release(size_t bytes);
 private:
  size_t quota_, used_;
  BufferAllocator* delegate_;
};

// ---- Column / View / Block (base/infrastructure/block.h:55-489) ---------------------------
typedef bool* bool_ptr;
typedef const bool* bool_const_ptr;

// base/infrastructure/bit_pointers.h:541-583 (the boolean flavour: one bool per row): a set of skip / is_null vectors.
class BoolView {
 public:
  explicit BoolView(size_t column_count) : columns_(column_count, static_cast<bool_ptr>(NULL)), row_count_(0) {}
  explicit BoolView(bool_ptr data) : columns_(1, data), row_count_(0) {}
  int column_count() const { return static_cast<int>(columns_.size()); }
  rowcount_t row_count() const { return row_count_; }
  bool_ptr column(int i) const { return columns_[i]; }
  void ResetColumn(int i, bool_ptr data) { columns_[i] = data; }
  void set_row_count(rowcount_t n) { row_count_ = n; }
 private:
  vector<bool_ptr> columns_;
  rowcount_t row_count_;
};

class VariantConstPointer {
 public:
  VariantConstPointer() : p_(NULL) {}
  VariantConstPointer(const void* p) : p_(p) {}   // NOLINT
  const void* raw() const { return p_; }
  bool is_null() const { return p_ == NULL; }
  template <DataType type> const typename TypeTraits<type>::cpp_type* as() const {
    return static_cast<const typename TypeTraits<type>::cpp_type*>(p_);
  }
  VariantConstPointer offset(rowcount_t rows, const TypeInfo& info) const {
    return VariantConstPointer(static_cast<const char*>(p_) + rows * info.size());
  }
 private:
  const void* p_;
};

class Column {
 public:
  Column() : attribute_(NULL), info_(NULL), data_(NULL), is_null_(NULL) {}
  const Attribute& attribute() const { return *attribute_; }
  const TypeInfo& type_info() const { return *info_; }
  VariantConstPointer data() const { return VariantConstPointer(data_); }
  template <DataType type> const typename TypeTraits<type>::cpp_type* typed_data() const {
    return static_cast<const typename TypeTraits<type>::cpp_type*>(data_);
  }
  // NULL = no NULLs in this view (block.h:117-121)
  bool_const_ptr is_null() const { return is_null_; }
  void Reset(const void* data, bool_const_ptr is_null) { data_ = data; is_null_ = is_null; }
  void ResetFrom(const Column& other) { data_ = other.data_; is_null_ = other.is_null_; }
  void ResetFromPlusOffset(const Column& other, rowcount_t offset) {
    data_ = static_cast<const char*>(other.data_) + offset * info_->size();
    is_null_ = other.is_null_ ? other.is_null_ + offset : NULL;
  }
  void ResetIsNull(bool_const_ptr is_null) { is_null_ = is_null; }
 private:
  friend class View;
  const Attribute* attribute_;
  const TypeInfo* info_;
  const void* data_;
  bool_const_ptr is_null_;
};

class View {
 public:
  explicit View(const TupleSchema& schema);
  View(const View& other);
  View(const View& other, rowcount_t offset, rowcount_t row_count);
  View& operator=(const View& other);
  const TupleSchema& schema() const { return schema_; }
  int column_count() const { return static_cast<int>(columns_.size()); }
  rowcount_t row_count() const { return row_count_; }
  void set_row_count(rowcount_t n) { row_count_ = n; }
  const Column& column(int i) const { return columns_[i]; }
  Column* mutable_column(int i) { return &columns_[i]; }
  void ResetFrom(const View& other);
  void ResetFromSubRange(const View& other, rowcount_t offset, rowcount_t row_count);
  void Advance(rowcount_t offset);
 private:
  void Bind();
  TupleSchema schema_;
  vector<Column> columns_;
  rowcount_t row_count_;
};

// Owns host storage for `row_capacity` rows of `schema` (one typed array + one bool per row
// for NULLABLE attributes), obtained from a BufferAllocator.
class Block {
 public:
  Block(const TupleSchema& schema, BufferAllocator* allocator);
  ~Block();
  const TupleSchema& schema() const { return view_.schema(); }
  int column_count() const { return view_.column_count(); }
  rowcount_t row_capacity() const { return capacity_; }
  // false on allocation failure (the previous contents are kept)
  bool Reallocate(rowcount_t new_capacity);
  const View& view() const { return view_; }
  void* mutable_data(int column) { return data_[column] ? data_[column]->data() : NULL; }
  bool* mutable_is_null(int column) { return nulls_[column] ? static_cast<bool*>(nulls_[column]->data()) : NULL; }
  bool is_nullable(int column) const { return schema().attribute(column).is_nullable(); }
  // Variable-length cells (StringPiece) point into storage the block keeps alive: the arena of block.h:259-281.
  void KeepAlive(const std::shared_ptr<const void>& storage) { storage_.push_back(storage); }
 private:
  Block(const Block&);
  void operator=(const Block&);
  BufferAllocator* allocator_;
  View view_;
  rowcount_t capacity_;
  vector<Buffer*> data_;
  vector<Buffer*> nulls_;
  vector<std::shared_ptr<const void> > storage_;
};

// base/memory/arena.h:48-110: bump allocator for variable-length values; pointers stay valid until Reset().
class Arena {
 public:
  Arena(BufferAllocator* const buffer_allocator, size_t initial_buffer_size, size_t max_buffer_size);
  Arena(size_t initial_buffer_size, size_t max_buffer_size);
  ~Arena();
  // Copies the bytes of `value` into the arena; NULL when the allocator refuses.
  const char* AddStringPieceContent(const StringPiece& value);
  void* AllocateBytes(const size_t size);
  void Reset();
  size_t memory_footprint() const { return footprint_; }
 private:
  Arena(const Arena&);
  void operator=(const Arena&);
  bool AddComponent(size_t at_least);
  BufferAllocator* allocator_;
  size_t next_size_, max_size_, footprint_;
  vector<Buffer*> buffers_;
  char* cursor_;
  size_t left_;
};

// base/infrastructure/copy_column.h:30-36
enum RowSelectorType { NO_SELECTOR = 0, INPUT_SELECTOR = 1 };

// base/infrastructure/view_copier.h:89-106: copies row_count rows of input_view into output_block at
// output_offset (the block must have the capacity). deep_copy: variable-length values are copied into storage
// the block keeps alive, otherwise the cells keep pointing at the source's bytes. Returns the rows copied.
class BoundSingleSourceProjector;
class ViewCopier {
 public:
  ViewCopier(const TupleSchema& schema, bool deep_copy);
  ViewCopier(const BoundSingleSourceProjector* projector, bool deep_copy);
  rowcount_t Copy(const rowcount_t row_count, const View& input_view, const rowcount_t output_offset, Block* output_block) const;
 private:
  TupleSchema schema_;
  vector<int> source_;   // input column of output column i
  bool deep_copy_;
};

}  // namespace supersonic
#endif  // SUPERSONIC_B200_HOST_BASE_H_
