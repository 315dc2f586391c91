// file_io.cc -- File, Sink / Writer and the reference's block format (cursor/infrastructure/file_io.cc:70-420,
// writer.cc, utils/file.cc). Host code: data enters and leaves the process here; the operators run on the GPU.
#include "supersonic/cursor/infrastructure/file_io.h"

#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <deque>

namespace supersonic {
namespace {

const rowcount_t kMaxChunkRowCount = 8192;   // file_io.cc:70

Exception* OutputError() { return new Exception(ERROR_GENERAL_IO_ERROR, "Writing view to the output file failed."); }
Exception* InputError() { return new Exception(ERROR_GENERAL_IO_ERROR, "Reading cursor's data from the input file failed."); }

FailureOrVoid Put(File* f, const void* data, size_t bytes) {
  if (bytes > 0 && f->Write(data, bytes) != static_cast<int64>(bytes)) THROW(OutputError());
  return Success();
}

FailureOrVoid WriteChunk(const View& view, rowcount_t offset, rowcount_t rows, File* f) {
  const uint64 n = rows;
  PROPAGATE_ON_FAILURE(Put(f, &n, sizeof(n)));
  for (int c = 0; c < view.column_count(); ++c) {
    const Column& col = view.column(c);
    const bool* nulls = col.is_null() ? col.is_null() + offset : NULL;
    if (col.attribute().is_nullable()) {
      // the reference CHECKs that a nullable column carries an is_null array; an absent one means "no NULLs"
      if (nulls != NULL) {
        PROPAGATE_ON_FAILURE(Put(f, nulls, rows));
      } else {
        const vector<char> zeros(rows, 0);
        PROPAGATE_ON_FAILURE(Put(f, zeros.data(), rows));
      }
    }
    const TypeInfo& info = col.type_info();
    if (info.is_variable_length()) {
      const StringPiece* cells = static_cast<const StringPiece*>(col.data().raw()) + offset;
      vector<uint64> lengths(rows);
      size_t total = 0;
      for (rowcount_t r = 0; r < rows; ++r) {
        lengths[r] = (nulls != NULL && nulls[r]) ? 0 : cells[r].size();
        total += lengths[r];
      }
      PROPAGATE_ON_FAILURE(Put(f, lengths.data(), rows * sizeof(uint64)));
      string bytes;
      bytes.reserve(total);
      for (rowcount_t r = 0; r < rows; ++r) if (lengths[r] > 0) bytes.append(cells[r].data(), lengths[r]);
      PROPAGATE_ON_FAILURE(Put(f, bytes.data(), bytes.size()));
    } else {
      PROPAGATE_ON_FAILURE(Put(f, static_cast<const char*>(col.data().raw()) + offset * info.size(), rows * info.size()));
    }
  }
  return Success();
}

class FileSink : public Sink {
 public:
  FileSink(File* f, Ownership ownership) : f_(f), ownership_(ownership) {}
  virtual ~FileSink() {}
  virtual FailureOr<rowcount_t> Write(const View& data) {
    for (rowcount_t offset = 0; offset < data.row_count(); offset += kMaxChunkRowCount) {
      PROPAGATE_ON_FAILURE(WriteChunk(data, offset, std::min<rowcount_t>(kMaxChunkRowCount, data.row_count() - offset), f_));
    }
    return Success(data.row_count());
  }
  virtual FailureOrVoid Finalize() {
    File* f = f_;
    f_ = NULL;
    if (ownership_ == TAKE_OWNERSHIP && f != NULL && !f->Close()) THROW(new Exception(ERROR_GENERAL_IO_ERROR, "Error closing the file."));
    return Success();
  }
 private:
  File* f_;
  Ownership ownership_;
};

class FileInputCursor : public Cursor {
 public:
  FileInputCursor(Block* block, File* f, bool delete_when_done)
      : block_(block), f_(f), delete_when_done_(delete_when_done), view_(block->schema()), pending_(0), first_pending_(0), interrupted_(false) {}
  virtual ~FileInputCursor() {
    if (delete_when_done_) f_->Delete();
    f_->Close();
  }
  virtual const TupleSchema& schema() const { return block_->schema(); }
  virtual void Interrupt() { interrupted_ = true; }
  virtual void AppendDebugDescription(string* target) const { target->append("FileInputCursor"); }
  virtual CursorId GetCursorId() const { return FILE_INPUT; }
  virtual ResultView Next(rowcount_t max_row_count) {
    if (interrupted_) return ResultView::Failure(new Exception(INTERRUPTED, "The cursor was interrupted"));
    if (pending_ == 0) {
      uint64 n = 0;
      const int64 got = f_->Read(&n, sizeof(n));
      if (got == 0 && f_->eof()) return ResultView::EOS();
      if (got != static_cast<int64>(sizeof(n))) return ResultView::Failure(InputError());
      if (n == 0) return ResultView::Failure(new Exception(ERROR_GENERAL_IO_ERROR, "Reading cursor's data from the input file failed. Chunk of size 0."));
      if (n > block_->row_capacity()) {
        return ResultView::Failure(new Exception(ERROR_GENERAL_IO_ERROR, "Reading cursor's data from the input file failed. Input chunk too large."));
      }
      FailureOrVoid r = ReadChunk(static_cast<rowcount_t>(n));
      if (r.is_failure()) return ResultView::Failure(r.release_exception());
      pending_ = static_cast<rowcount_t>(n);
      first_pending_ = 0;
    }
    const rowcount_t rows = std::min(max_row_count, pending_);
    view_.ResetFromSubRange(block_->view(), first_pending_, rows);
    first_pending_ += rows;
    pending_ -= rows;
    return ResultView::Success(&view_);
  }
 private:
  FailureOrVoid Get(void* data, size_t bytes) {
    if (bytes == 0) return Success();
    const int64 got = f_->Read(data, bytes);
    if (got == static_cast<int64>(bytes)) return Success();
    if (got == 0 && f_->eof()) THROW(new Exception(ERROR_GENERAL_IO_ERROR, "Premature END_OF_FILE."));
    THROW(InputError());
  }
  FailureOrVoid ReadChunk(rowcount_t rows) {
    strings_.clear();   // the previous chunk's bytes (block.h ResetArenas)
    for (int c = 0; c < block_->column_count(); ++c) {
      const Attribute& a = block_->schema().attribute(c);
      bool* nulls = block_->mutable_is_null(c);
      if (a.is_nullable()) PROPAGATE_ON_FAILURE(Get(nulls, rows));
      const TypeInfo& info = GetTypeInfo(a.type());
      if (info.is_variable_length()) {
        lengths_.resize(rows);
        PROPAGATE_ON_FAILURE(Get(lengths_.data(), rows * sizeof(uint64)));
        size_t total = 0;
        for (rowcount_t r = 0; r < rows; ++r) total += lengths_[r];
        strings_.push_back(string());
        string& bytes = strings_.back();
        bytes.resize(total);
        PROPAGATE_ON_FAILURE(Get(&bytes[0], total));
        StringPiece* cells = static_cast<StringPiece*>(block_->mutable_data(c));
        size_t at = 0;
        for (rowcount_t r = 0; r < rows; ++r) {
          cells[r] = StringPiece(bytes.data() + at, lengths_[r]);
          at += lengths_[r];
        }
      } else {
        PROPAGATE_ON_FAILURE(Get(block_->mutable_data(c), rows * info.size()));
      }
    }
    return Success();
  }
  std::unique_ptr<Block> block_;
  File* f_;
  bool delete_when_done_;
  View view_;
  rowcount_t pending_, first_pending_;
  bool interrupted_;
  vector<uint64> lengths_;
  std::deque<string> strings_;   // one buffer per variable-length column of the current chunk (a deque: no element moves)
};

}  // namespace

Sink* FileOutput(File* output_file, Ownership file_ownership) { return new FileSink(output_file, file_ownership); }

FailureOrOwned<Cursor> FileInput(const TupleSchema& schema, File* input_file, const bool delete_when_done, BufferAllocator* allocator) {
  std::unique_ptr<Block> block(new Block(schema, allocator));
  if (!block->Reallocate(kMaxChunkRowCount)) {
    input_file->Close();
    THROW(new Exception(ERROR_MEMORY_EXCEEDED, "Block allocation for FileInputCursor failed."));
  }
  return Success(static_cast<Cursor*>(new FileInputCursor(block.release(), input_file, delete_when_done)));
}

// ---- Writer (cursor/infrastructure/writer.cc)
FailureOr<rowcount_t> Writer::Write(Sink* sink, rowcount_t max_row_count) {
  rowcount_t written = 0;
  barrier_ = false;
  while (written < max_row_count && !eos_) {
    if (!has_pending()) {
      ResultView r = cursor_->Next(max_row_count - written);
      if (r.is_failure()) return Failure(r.release_exception());
      if (r.is_eos()) { eos_ = true; break; }
      if (r.is_waiting_on_barrier()) { barrier_ = true; break; }
      if (!r.has_data()) break;
      pending_ = View(r.view());
      if (pending_.row_count() == 0) continue;
    }
    FailureOr<rowcount_t> w = sink->Write(pending_);
    PROPAGATE_ON_FAILURE(w);
    written += w.get();
    if (w.get() >= pending_.row_count()) {
      pending_ = View(TupleSchema());
    } else {
      if (w.get() == 0) break;   // the sink is full
      pending_.Advance(w.get());
    }
  }
  return Success(written);
}

FailureOrVoid WriteCursor(Cursor* cursor, Sink* sink) {
  Writer writer(cursor);
  FailureOr<rowcount_t> r = writer.WriteAll(sink);
  PROPAGATE_ON_FAILURE(r);
  if (writer.is_waiting_on_barrier()) THROW(new Exception(ERROR_UNKNOWN_ERROR, "Writing stumbled on a barrier."));
  if (!writer.is_eos()) {
    char buf[80];
    snprintf(buf, sizeof(buf), "Writing stopped after %llu rows.", static_cast<unsigned long long>(r.get()));
    THROW(new Exception(ERROR_MEMORY_EXCEEDED, buf));
  }
  return Success();
}

}  // namespace supersonic

// ---- File (utils/file.cc: a local file over stdio)
File* File::Create(const std::string& file_name, const std::string& mode) { return new File(file_name, mode); }
File* File::OpenOrDie(const std::string& file_name, const std::string& mode) {
  File* f = Create(file_name, mode);
  if (!f->Open()) {
    fprintf(stderr, "FATAL: cannot open %s in mode %s\n", file_name.c_str(), mode.c_str());
    abort();
  }
  return f;
}
bool File::Exists(const std::string& file) { return access(file.c_str(), F_OK) == 0; }
std::string File::JoinPath(const std::string& dirname, const std::string& basename) {
  if ((!basename.empty() && basename[0] == '/') || dirname.empty()) return basename;
  return dirname[dirname.size() - 1] == '/' ? dirname + basename : dirname + "/" + basename;
}
File::~File() {}
bool File::Exists() const { return access(create_file_name_.c_str(), F_OK) == 0; }
bool File::Open() {
  if (f_ != NULL) return false;
  f_ = fopen(create_file_name_.c_str(), mode_.c_str());
  return f_ != NULL;
}
bool File::Delete() { return unlink(create_file_name_.c_str()) == 0; }
bool File::Close() {
  bool ok = true;
  if (f_ != NULL) { ok = fclose(f_) == 0; f_ = NULL; }
  delete this;
  return ok;
}
int64 File::Read(void* buffer, uint64 length) { return f_ ? static_cast<int64>(fread(buffer, 1, length, f_)) : -1; }
char* File::ReadLine(char* buffer, uint64 max_length) { return f_ ? fgets(buffer, static_cast<int>(max_length), f_) : NULL; }
int64 File::Write(const void* buffer, uint64 length) { return f_ ? static_cast<int64>(fwrite(buffer, 1, length, f_)) : -1; }
bool File::Seek(int64 position) { return f_ != NULL && fseeko(f_, position, SEEK_SET) == 0; }
bool File::eof() { return f_ == NULL || feof(f_) != 0; }
