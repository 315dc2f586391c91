// supersonic/cursor/infrastructure/file_io.h:64-77: the reference's block format on disk / on the wire
// (file_io.cc:91-108, 147-190, 277-420), so that tables written by the reference can be scanned here and results
// written here can be read by it. A file is a sequence of chunks of at most 8192 rows:
//   uint64 row_count, then per column: [NULLABLE: one bool per row] + the values (fixed width: the raw array;
//   STRING / BINARY: uint64 length per row (0 for NULL), then the bytes of the non-empty values back to back).
#ifndef SUPERSONIC_B200_HOST_CURSOR_INFRASTRUCTURE_FILE_IO_H_
#define SUPERSONIC_B200_HOST_CURSOR_INFRASTRUCTURE_FILE_IO_H_
#include "supersonic/cursor/infrastructure/writer.h"
#include "supersonic/utils/basictypes.h"
#include "supersonic/utils/file.h"

namespace supersonic {

// TAKE_OWNERSHIP: Finalize() closes the file.
Sink* FileOutput(File* output_file, Ownership file_ownership);
// The cursor closes the file when it is destroyed (and deletes it first when delete_when_done).
FailureOrOwned<Cursor> FileInput(const TupleSchema& schema, File* input_file, const bool delete_when_done, BufferAllocator* allocator);

}  // namespace supersonic
#endif
