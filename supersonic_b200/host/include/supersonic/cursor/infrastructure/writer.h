// supersonic/cursor/infrastructure/writer.h:36-88: Sink (something views are written to), Writer (drains a cursor
// into a sink) and WriteCursor.
#ifndef SUPERSONIC_B200_HOST_CURSOR_INFRASTRUCTURE_WRITER_H_
#define SUPERSONIC_B200_HOST_CURSOR_INFRASTRUCTURE_WRITER_H_
#include "supersonic/cursor.h"

namespace supersonic {

class Sink {
 public:
  virtual ~Sink() {}
  // Writes the view; returns the number of rows written (may be fewer than the view holds).
  virtual FailureOr<rowcount_t> Write(const View& data) = 0;
  // Must be called exactly once, after the last Write.
  virtual FailureOrVoid Finalize() = 0;
 protected:
  Sink() {}
 private:
  Sink(const Sink&);
  void operator=(const Sink&);
};

// Takes ownership of the cursor.
class Writer {
 public:
  explicit Writer(Cursor* cursor) : cursor_(cursor), pending_(TupleSchema()), eos_(false), barrier_(false) {}

  // state of the input after the last Write
  bool is_eos() const { return eos_; }
  bool is_waiting_on_barrier() const { return barrier_; }
  const TupleSchema& schema() const { return cursor_->schema(); }
  void Interrupt() { cursor_->Interrupt(); }

  // Writes at most max_row_count rows; fewer when the input ends, waits on a barrier or the sink takes fewer.
  FailureOr<rowcount_t> Write(Sink* sink, rowcount_t max_row_count);
  FailureOr<rowcount_t> WriteAll(Sink* sink) { return Write(sink, std::numeric_limits<rowcount_t>::max()); }

 private:
  bool has_pending() const { return pending_.column_count() == cursor_->schema().attribute_count() && pending_.row_count() > 0; }
  std::unique_ptr<Cursor> cursor_;
  View pending_;           // rows of the last Next() the sink has not taken yet
  bool eos_, barrier_;
};

// Drains the cursor into the sink (writer.cc:69-84). Takes ownership of the cursor; the caller finalizes the sink.
FailureOrVoid WriteCursor(Cursor* cursor, Sink* sink);

}  // namespace supersonic
#endif
