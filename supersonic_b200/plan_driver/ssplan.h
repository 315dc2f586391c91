/* ssplan.h -- C entry points of the plan driver.
 *
 * plan_driver.cc is an ordinary *client* of "supersonic/supersonic.h": it parses
 * a small S-expression plan, builds it with the public Supersonic factories
 * (ScanView, Compute, Filter, Project, GroupAggregate, ScalarAggregate,
 * HashJoinOperation, Sort and the Expression factories), drains the resulting
 * Cursor and hands the result columns back through this C interface.
 *
 * The SAME source is compiled twice:
 *   - against /root/reference            -> oracle/_ref/libssref.so   (the oracle)
 *   - against supersonic_b200/host       -> supersonic_b200/lib/libssb200_plan.so
 * so a parity test runs one plan text through both and compares the columns.
 * That is the drop-in claim of this repo, executed.
 */
#ifndef SSPLAN_H_
#define SSPLAN_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* dtype numbers are supersonic::DataType (supersonic/proto/supersonic.proto:15-37). */
typedef struct {
  const char* name;
  int32_t dtype;
  int32_t nullable;          /* 0 = NOT_NULLABLE, 1 = NULLABLE */
  const void* data;          /* host (or, for the B200 build, device) pointer; STRING / BINARY: an array of
                                { const char* ptr; int64 length; } cells (the layout of StringPiece) */
  const uint8_t* is_null;    /* bool per row, or NULL = no nulls in this view */
} ssplan_column;

typedef struct {
  int32_t ncols;
  int64_t rows;
  const ssplan_column* cols;
} ssplan_table;

typedef struct ssplan_result ssplan_result;

enum {
  SSPLAN_DISCARD = 1,        /* drain the cursor but do not materialise rows (timing) */
  SSPLAN_BIND_ONLY = 2,      /* CreateCursor only: report the result schema, never call Next */
  SSPLAN_SPY = 4,            /* Cursor::ApplyToChildren with a pass-through CursorTransformer, then wrap the root too */
};

/* Builds and runs `plan` over `tables`. Always sets *out (free it with
 * ssplan_result_free). Returns the supersonic::ReturnCode (0 = OK).
 * next_max_rows: the max_row_count passed to every Cursor::Next (<=0: 1024,
 * Cursor::kDefaultRowCount). */
int ssplan_run(const char* plan, int32_t ntables, const ssplan_table* tables,
               int64_t next_max_rows, int32_t flags, ssplan_result** out);

int ssplan_result_code(const ssplan_result* r);
const char* ssplan_result_error(const ssplan_result* r);
int32_t ssplan_result_ncols(const ssplan_result* r);
int64_t ssplan_result_rows(const ssplan_result* r);
const char* ssplan_result_col_name(const ssplan_result* r, int32_t i);
int32_t ssplan_result_col_dtype(const ssplan_result* r, int32_t i);
int32_t ssplan_result_col_nullable(const ssplan_result* r, int32_t i);
/* Fixed-width columns: the values. STRING / BINARY columns: one int64 length per row (0 for NULL rows); the cells
 * themselves lie end to end in ssplan_result_col_bytes. */
const void* ssplan_result_col_data(const ssplan_result* r, int32_t i);
const char* ssplan_result_col_bytes(const ssplan_result* r, int32_t i);
/* bool per row; NULL when the column never reported an is_null vector. */
const uint8_t* ssplan_result_col_is_null(const ssplan_result* r, int32_t i);
/* seconds spent in CreateCursor() and in the Next() drain loop (wall clock). */
double ssplan_result_create_seconds(const ssplan_result* r);
double ssplan_result_drain_seconds(const ssplan_result* r);
int64_t ssplan_result_next_calls(const ssplan_result* r);
/* SSPLAN_SPY: how many children of the root cursor were handed to the transformer. */
int64_t ssplan_result_spied_children(const ssplan_result* r);
void ssplan_result_free(ssplan_result* r);

/* "reference" for the oracle build, "b200" for the product build. */
const char* ssplan_impl(void);

#ifdef __cplusplus
}
#endif
#endif  /* SSPLAN_H_ */
