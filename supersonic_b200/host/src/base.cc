// base.cc -- data model of the host layer (see include/supersonic/base.h).
#include "supersonic/base.h"

#include <stdio.h>

#include <mutex>

namespace supersonic {

namespace {
struct NameTable {
  std::map<int, string> names;
  string empty;
  const string& Get(int v) const {
    std::map<int, string>::const_iterator it = names.find(v);
    return it == names.end() ? empty : it->second;
  }
};
#define N(x) t->names[x] = #x
const NameTable& DataTypeNames() {
  static NameTable* t = [] { NameTable* t = new NameTable;
    N(INT32); N(INT64); N(UINT32); N(UINT64); N(FLOAT); N(DOUBLE); N(BOOL); N(DATE); N(DATETIME);
    N(STRING); N(BINARY); N(ENUM); N(DATA_TYPE); return t; }();
  return *t;
}
const NameTable& ReturnCodeNames() {
  static NameTable* t = [] { NameTable* t = new NameTable;
    N(OK); N(END_OF_INPUT); N(BEFORE_INPUT); N(WAITING_ON_BARRIER); N(ERROR_UNKNOWN_ERROR);
    N(ERROR_GENERAL_IO_ERROR); N(ERROR_MEMORY_EXCEEDED); N(ERROR_NOT_IMPLEMENTED);
    N(ERROR_EVALUATION_ERROR); N(ERROR_BAD_PROTO); N(ERROR_TEMP_FILE_CREATION_ERROR);
    N(ERROR_TOO_FEW_ROWS); N(ERROR_TOO_MANY_ROWS); N(ERROR_TOO_FEW_COLUMNS); N(ERROR_TOO_MANY_COLUMNS);
    N(ERROR_DUPLICATED_UNIQUE_KEY); N(ERROR_GENERIC_SCHEMA_ERROR); N(ERROR_ATTRIBUTE_COUNT_MISMATCH);
    N(ERROR_ATTRIBUTE_TYPE_MISMATCH); N(ERROR_ATTRIBUTE_MISSING); N(ERROR_ATTRIBUTE_EXISTS);
    N(ERROR_INVALID_ARGUMENT_TYPE); N(ERROR_ATTRIBUTE_IS_NULLABLE); N(ERROR_INVALID_ARGUMENT_VALUE);
    N(ERROR_ATTRIBUTE_AMBIGUOUS); N(ERROR_DUPLICATE_ENUM_VALUE_NUMBER); N(ERROR_DUPLICATE_ENUM_VALUE_NAME);
    N(ERROR_UNDEFINED_ENUM_VALUE_NUMBER); N(ERROR_UNDEFINED_ENUM_VALUE_NAME); N(ERROR_FOREIGN_KEY_INVALID);
    N(INTERRUPTED); return t; }();
  return *t;
}
const NameTable& AggregationNames() {
  static NameTable* t = [] { NameTable* t = new NameTable;
    N(SUM); N(MIN); N(MAX); N(COUNT); N(CONCAT); N(FIRST); N(LAST); return t; }();
  return *t;
}
const NameTable& JoinTypeNames() {
  static NameTable* t = [] { NameTable* t = new NameTable;
    N(INNER); N(LEFT_OUTER); N(RIGHT_OUTER); N(FULL_OUTER); return t; }();
  return *t;
}
#undef N
}  // namespace

const string& DataType_Name(DataType t) { return DataTypeNames().Get(t); }
const string& ReturnCode_Name(ReturnCode c) { return ReturnCodeNames().Get(c); }
const string& Aggregation_Name(Aggregation a) { return AggregationNames().Get(a); }
const string& JoinType_Name(JoinType j) { return JoinTypeNames().Get(j); }

Exception* Exception::AddStackTraceElement(const StringPiece& function, const StringPiece& filename,
                                           int line, const StringPiece& context) {
  char buf[32];
  snprintf(buf, sizeof(buf), "%d", line);
  trace_ += "    at " + function.as_string() + "(" + filename.as_string() + ":" + buf + ") " +
            context.as_string() + "\n";
  return this;
}

void DieOnFailure(const Exception& e) {
  fprintf(stderr, "FATAL: %s\n", e.PrintStackTrace().c_str());
  abort();
}

const TypeInfo& GetTypeInfo(DataType type) {
  static std::map<int, TypeInfo>* infos = [] {
    std::map<int, TypeInfo>* m = new std::map<int, TypeInfo>;
    m->insert(std::make_pair(INT32, TypeInfo(INT32, "INT32", 4, true, true, false, false)));
    m->insert(std::make_pair(INT64, TypeInfo(INT64, "INT64", 8, true, true, false, false)));
    m->insert(std::make_pair(UINT32, TypeInfo(UINT32, "UINT32", 4, true, true, false, false)));
    m->insert(std::make_pair(UINT64, TypeInfo(UINT64, "UINT64", 8, true, true, false, false)));
    m->insert(std::make_pair(FLOAT, TypeInfo(FLOAT, "FLOAT", 4, true, false, true, false)));
    m->insert(std::make_pair(DOUBLE, TypeInfo(DOUBLE, "DOUBLE", 8, true, false, true, false)));
    m->insert(std::make_pair(BOOL, TypeInfo(BOOL, "BOOL", 1, false, false, false, false)));
    m->insert(std::make_pair(DATE, TypeInfo(DATE, "DATE", 4, false, false, false, false)));
    m->insert(std::make_pair(DATETIME, TypeInfo(DATETIME, "DATETIME", 8, false, false, false, false)));
    m->insert(std::make_pair(ENUM, TypeInfo(ENUM, "ENUM", 4, false, false, false, false)));
    m->insert(std::make_pair(DATA_TYPE, TypeInfo(DATA_TYPE, "DATA_TYPE", 4, false, false, false, false)));
    m->insert(std::make_pair(STRING, TypeInfo(STRING, "STRING", sizeof(StringPiece), false, false, false, true)));
    m->insert(std::make_pair(BINARY, TypeInfo(BINARY, "BINARY", sizeof(StringPiece), false, false, false, true)));
    return m;
  }();
  return infos->find(type)->second;
}

// ---- TupleSchema
bool TupleSchema::add_attribute(const Attribute& attribute) {
  if (positions_.count(attribute.name())) return false;
  positions_[attribute.name()] = static_cast<int>(attributes_.size());
  attributes_.push_back(attribute);
  return true;
}
int TupleSchema::LookupAttributePosition(const string& attribute_name) const {
  std::map<string, int>::const_iterator it = positions_.find(attribute_name);
  return it == positions_.end() ? -1 : it->second;
}
TupleSchema TupleSchema::Singleton(const string& name, const DataType type, Nullability nullability) {
  TupleSchema s;
  s.add_attribute(Attribute(name, type, nullability));
  return s;
}
bool TupleSchema::AreEqual(const TupleSchema& a, const TupleSchema& b, bool check_names) {
  if (a.attribute_count() != b.attribute_count()) return false;
  for (int i = 0; i < a.attribute_count(); ++i) {
    const Attribute& x = a.attribute(i);
    const Attribute& y = b.attribute(i);
    if (x.type() != y.type() || x.nullability() != y.nullability()) return false;
    if (check_names && x.name() != y.name()) return false;
  }
  return true;
}
bool TupleSchema::CanMerge(const TupleSchema& a, const TupleSchema& b) {
  for (int i = 0; i < b.attribute_count(); ++i) {
    if (a.LookupAttributePosition(b.attribute(i).name()) >= 0) return false;
  }
  return true;
}
TupleSchema TupleSchema::Merge(const TupleSchema& a, const TupleSchema& b) {
  TupleSchema r(a);
  for (int i = 0; i < b.attribute_count(); ++i) r.add_attribute(b.attribute(i));
  return r;
}
FailureOr<TupleSchema> TupleSchema::TryMerge(const TupleSchema& a, const TupleSchema& b) {
  TupleSchema r(a);
  for (int i = 0; i < b.attribute_count(); ++i) {
    if (!r.add_attribute(b.attribute(i))) {
      THROW(new Exception(ERROR_ATTRIBUTE_EXISTS,
                          "Can't merge schemas, ambiguous attribute name: " + b.attribute(i).name()));
    }
  }
  return Success(r);
}
bool TupleSchema::EqualByType(const TupleSchema& other) const { return AreEqual(*this, other, false); }
string TupleSchema::GetHumanReadableSpecification() const {
  string s;
  for (int i = 0; i < attribute_count(); ++i) {
    if (i) s += ", ";
    s += attribute(i).name() + ": " + GetTypeInfo(attribute(i).type()).name();
    if (!attribute(i).is_nullable()) s += " NOT NULL";
  }
  return s;
}

// ---- BufferAllocator
Buffer::~Buffer() {
  free(data_);
  allocator_->Release(size_);
}
Buffer* BufferAllocator::BestEffortAllocate(size_t requested, size_t minimal) {
  const size_t granted = Grant(requested, minimal);
  if (granted < minimal || (granted == 0 && requested > 0)) { if (granted) Release(granted); return NULL; }
  void* p = malloc(granted ? granted : 16);
  if (p == NULL) { Release(granted); return NULL; }
  return new Buffer(p, granted, this);
}
bool BufferAllocator::BestEffortReallocate(size_t requested, size_t minimal, Buffer* buffer) {
  if (requested <= buffer->size_) {
    Release(buffer->size_ - requested);
    buffer->size_ = requested;
    return true;
  }
  const size_t extra = Grant(requested - buffer->size_, minimal > buffer->size_ ? minimal - buffer->size_ : 0);
  const size_t total = buffer->size_ + extra;
  if (total < minimal || extra == 0) { if (extra) Release(extra); return false; }
  void* p = realloc(buffer->data_, total);
  if (p == NULL) { Release(extra); return false; }
  buffer->data_ = p;
  buffer->size_ = total;
  return true;
}
HeapBufferAllocator* HeapBufferAllocator::Get() {
  static HeapBufferAllocator* a = new HeapBufferAllocator;
  return a;
}
size_t MemoryLimit::Grant(size_t requested, size_t minimal) {
  const size_t avail = Available();
  size_t want = requested <= avail ? requested : (minimal <= avail ? avail : 0);
  if (want == 0 && requested > 0) return 0;
  const size_t got = delegate_->Grant(want, minimal < want ? minimal : want);
  used_ += got;
  return got;
}
void MemoryLimit::Release(size_t bytes) {
  used_ -= bytes < used_ ? bytes : used_;
  delegate_->Release(bytes);
}

// ---- View
View::View(const TupleSchema& schema) : schema_(schema), columns_(schema.attribute_count()), row_count_(0) { Bind(); }
View::View(const View& other) : schema_(other.schema_), columns_(other.columns_), row_count_(other.row_count_) { Bind(); }
View::View(const View& other, rowcount_t offset, rowcount_t row_count)
    : schema_(other.schema_), columns_(other.columns_), row_count_(row_count) {
  Bind();
  for (size_t i = 0; i < columns_.size(); ++i) columns_[i].ResetFromPlusOffset(other.columns_[i], offset);
}
View& View::operator=(const View& other) {
  schema_ = other.schema_;
  columns_ = other.columns_;
  row_count_ = other.row_count_;
  Bind();
  return *this;
}
void View::Bind() {
  for (size_t i = 0; i < columns_.size(); ++i) {
    columns_[i].attribute_ = &schema_.attribute(static_cast<int>(i));
    columns_[i].info_ = &GetTypeInfo(schema_.attribute(static_cast<int>(i)).type());
  }
}
void View::ResetFrom(const View& other) {
  for (size_t i = 0; i < columns_.size(); ++i) columns_[i].ResetFrom(other.columns_[i]);
  row_count_ = other.row_count_;
}
void View::ResetFromSubRange(const View& other, rowcount_t offset, rowcount_t row_count) {
  for (size_t i = 0; i < columns_.size(); ++i) columns_[i].ResetFromPlusOffset(other.columns_[i], offset);
  row_count_ = row_count;
}
void View::Advance(rowcount_t offset) {
  for (size_t i = 0; i < columns_.size(); ++i) columns_[i].ResetFromPlusOffset(columns_[i], offset);
  row_count_ -= offset;
}

// ---- Block
Block::Block(const TupleSchema& schema, BufferAllocator* allocator)
    : allocator_(allocator), view_(schema), capacity_(0), data_(schema.attribute_count(), NULL),
      nulls_(schema.attribute_count(), NULL) {}
Block::~Block() {
  for (size_t i = 0; i < data_.size(); ++i) { delete data_[i]; delete nulls_[i]; }
}
bool Block::Reallocate(rowcount_t new_capacity) {
  for (int i = 0; i < column_count(); ++i) {
    const size_t bytes = new_capacity * GetTypeInfo(schema().attribute(i).type()).size();
    if (data_[i] == NULL) {
      data_[i] = allocator_->Allocate(bytes);
      if (data_[i] == NULL) return false;
    } else if (!allocator_->Reallocate(bytes, data_[i])) {
      return false;
    }
    if (schema().attribute(i).is_nullable()) {
      if (nulls_[i] == NULL) {
        nulls_[i] = allocator_->Allocate(new_capacity);
        if (nulls_[i] == NULL) return false;
      } else if (!allocator_->Reallocate(new_capacity, nulls_[i])) {
        return false;
      }
    }
    view_.mutable_column(i)->Reset(data_[i]->data(), nulls_[i] ? static_cast<bool*>(nulls_[i]->data()) : NULL);
  }
  capacity_ = new_capacity;
  view_.set_row_count(new_capacity);   // block.h:467,481: a block's view spans its capacity
  return true;
}

// ---- Arena (base/memory/arena.h:48-110)
Arena::Arena(BufferAllocator* const buffer_allocator, size_t initial_buffer_size, size_t max_buffer_size)
    : allocator_(buffer_allocator), next_size_(initial_buffer_size ? initial_buffer_size : 16),
      max_size_(max_buffer_size < initial_buffer_size ? initial_buffer_size : max_buffer_size), footprint_(0), cursor_(NULL), left_(0) {}
Arena::Arena(size_t initial_buffer_size, size_t max_buffer_size)
    : allocator_(HeapBufferAllocator::Get()), next_size_(initial_buffer_size ? initial_buffer_size : 16),
      max_size_(max_buffer_size < initial_buffer_size ? initial_buffer_size : max_buffer_size), footprint_(0), cursor_(NULL), left_(0) {}
Arena::~Arena() { for (size_t i = 0; i < buffers_.size(); ++i) delete buffers_[i]; }
bool Arena::AddComponent(size_t at_least) {
  // buffers double up to max_buffer_size; a single request larger than that gets a buffer of its own
  size_t want = next_size_ < at_least ? at_least : next_size_;
  Buffer* b = allocator_->BestEffortAllocate(want, at_least);
  if (b == NULL) return false;
  buffers_.push_back(b);
  footprint_ += b->size();
  cursor_ = static_cast<char*>(b->data());
  left_ = b->size();
  if (next_size_ < max_size_) next_size_ = next_size_ * 2 > max_size_ ? max_size_ : next_size_ * 2;
  return true;
}
void* Arena::AllocateBytes(const size_t size) {
  if (size > left_ && !AddComponent(size)) return NULL;
  void* p = cursor_;
  cursor_ += size;
  left_ -= size;
  return p;
}
const char* Arena::AddStringPieceContent(const StringPiece& value) {
  char* p = static_cast<char*>(AllocateBytes(value.size()));
  if (p == NULL) return NULL;
  if (value.size() > 0) memcpy(p, value.data(), value.size());
  return p;
}
void Arena::Reset() {
  for (size_t i = 0; i < buffers_.size(); ++i) delete buffers_[i];
  buffers_.clear();
  footprint_ = 0;
  cursor_ = NULL;
  left_ = 0;
}

// ---- ViewCopier (base/infrastructure/view_copier.h:89-106)
ViewCopier::ViewCopier(const TupleSchema& schema, bool deep_copy) : schema_(schema), deep_copy_(deep_copy) {
  for (int i = 0; i < schema.attribute_count(); ++i) source_.push_back(i);
}
rowcount_t ViewCopier::Copy(const rowcount_t row_count, const View& input_view, const rowcount_t output_offset, Block* output_block) const {
  if (output_offset + row_count > output_block->row_capacity()) return 0;
  for (size_t i = 0; i < source_.size(); ++i) {
    const Column& in = input_view.column(source_[i]);
    const DataType type = in.type_info().type();
    const size_t w = in.type_info().size();
    char* dst = static_cast<char*>(output_block->mutable_data(static_cast<int>(i))) + output_offset * w;
    if (deep_copy_ && (type == STRING || type == BINARY)) {
      const StringPiece* cells = static_cast<const StringPiece*>(in.data().raw());
      size_t total = 0;
      for (rowcount_t r = 0; r < row_count; ++r) if (!(in.is_null() && in.is_null()[r])) total += cells[r].size();
      std::shared_ptr<string> bytes(new string());
      bytes->resize(total);
      size_t at = 0;
      StringPiece* out = reinterpret_cast<StringPiece*>(dst);
      for (rowcount_t r = 0; r < row_count; ++r) {
        if (in.is_null() && in.is_null()[r]) { out[r] = StringPiece(); continue; }
        if (cells[r].size() > 0) memcpy(&(*bytes)[at], cells[r].data(), cells[r].size());
        out[r] = StringPiece(bytes->data() + at, cells[r].size());
        at += cells[r].size();
      }
      output_block->KeepAlive(bytes);
    } else {
      memcpy(dst, in.data().raw(), row_count * w);
    }
    if (bool* nulls = output_block->mutable_is_null(static_cast<int>(i))) {
      if (in.is_null()) memcpy(nulls + output_offset, in.is_null(), row_count);
      else memset(nulls + output_offset, 0, row_count);
    }
  }
  return row_count;
}

}  // namespace supersonic
