/* supersonic_b200.h -- C ABI of libssb200.so, the B200 (sm_100a) implementation of
 * Supersonic's vectorised columnar execution hot path.
 *
 * The reference (google/supersonic) has no FFI: its seam is the set of C++ factory
 * functions and virtual interfaces of supersonic.h. This
 * header is the thin C boundary *underneath* that seam: every entry point names the
 * reference interface it replaces (file:line under /root/reference/supersonic).
 * The C++ mirror of supersonic.h that calls it lives in supersonic_b200/host/.
 *
 * Conventions
 *  - plain pointers and sizes only; no C++ or torch types
 *  - every function returns a supersonic::ReturnCode value (proto/supersonic.proto:40-82):
 *    0 OK, 102 ERROR_MEMORY_EXCEEDED, 103 ERROR_NOT_IMPLEMENTED, 104 ERROR_EVALUATION_ERROR,
 *    405 ERROR_INVALID_ARGUMENT_TYPE, 407 ERROR_INVALID_ARGUMENT_VALUE, 100 unknown/CUDA error;
 *    ssb_last_error(ctx) holds the message of the last failure on that context
 *  - all device work is ordered on the context's stream; calls that return a count to
 *    the host synchronise that stream, the others are asynchronous
 *  - column data live in HBM as SoA: one typed array per column plus an optional
 *    is_null BITMAP (bit i of 32-bit word i/32 set = row i is NULL). The reference's
 *    in-memory format is one bool per row (base/infrastructure/bit_pointers.h:528-534);
 *    ssb_nulls_pack/unpack convert at the boundary only
 *  - there is no CPU fallback anywhere behind this header
 */
#ifndef SUPERSONIC_B200_H_
#define SUPERSONIC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSB_ABI_VERSION 2

/* supersonic::DataType numbering (proto/supersonic.proto:15-37). Fixed-width types only. */
enum {
  SSB_INT32 = 1, SSB_INT64 = 2, SSB_UINT64 = 3, SSB_DATETIME = 4, SSB_DOUBLE = 5, SSB_BOOL = 6,
  SSB_UINT32 = 8, SSB_FLOAT = 9, SSB_DATE = 10, SSB_ENUM = 13
};

enum {
  SSB_OK = 0, SSB_ERROR_UNKNOWN = 100, SSB_ERROR_MEMORY_EXCEEDED = 102,
  SSB_ERROR_NOT_IMPLEMENTED = 103, SSB_ERROR_EVALUATION_ERROR = 104,
  SSB_ERROR_INVALID_ARGUMENT_TYPE = 405, SSB_ERROR_INVALID_ARGUMENT_VALUE = 407
};

typedef struct ssb_ctx ssb_ctx;

/* A column resident in HBM. `data` must be 16-byte aligned for the TMA path (any
 * cudaMalloc'ed pointer is); other alignments run on the plain-load path.
 * Replaces supersonic::Column (base/infrastructure/block.h:55-192). */
typedef struct {
  void* data;
  uint32_t* nulls;   /* bitmap, or NULL = no NULLs in this column */
  int32_t dtype;     /* SSB_* */
  int32_t reserved;
} ssb_column;

/* ------------------------------------------------------------------ context */
/* One context = one device + one stream + a workspace pool. Plays the role of the
 * BufferAllocator seam (base/memory/memory.h:100-236) for device memory. */
int ssb_ctx_create(int device, ssb_ctx** out);
void ssb_ctx_destroy(ssb_ctx* ctx);
const char* ssb_last_error(const ssb_ctx* ctx);
void* ssb_ctx_stream(ssb_ctx* ctx);              /* cudaStream_t */
int ssb_ctx_sync(ssb_ctx* ctx);
int ssb_abi_version(void);
/* Number of kernels launched by this library on the context since creation. */
int64_t ssb_ctx_launch_count(const ssb_ctx* ctx);
/* Device time (ms, CUDA events on the context stream) of the most recent ssb_program_run /
 * ssb_group_update / ssb_join_probe / ssb_sort_permutation call; valid after a sync. */
int ssb_ctx_last_kernel_ms(ssb_ctx* ctx, float* ms);
int ssb_ctx_enable_timing(ssb_ctx* ctx, int enable);
/* A stopwatch on the context stream (CUDA events): start records, stop records, waits and
 * returns the device time between the two in milliseconds. */
int ssb_ctx_timer_start(ssb_ctx* ctx);
int ssb_ctx_timer_stop(ssb_ctx* ctx, float* ms);

int ssb_malloc(ssb_ctx* ctx, size_t bytes, void** out);
int ssb_free(ssb_ctx* ctx, void* ptr);
int ssb_malloc_host(ssb_ctx* ctx, size_t bytes, void** out);   /* pinned */
int ssb_free_host(ssb_ctx* ctx, void* ptr);
int ssb_memcpy_h2d(ssb_ctx* ctx, void* dst, const void* src, size_t bytes);  /* async */
int ssb_memcpy_d2h(ssb_ctx* ctx, void* dst, const void* src, size_t bytes);  /* async */
int ssb_memcpy_d2d(ssb_ctx* ctx, void* dst, const void* src, size_t bytes);  /* async */
/* Bytes moved by ssb_memcpy_h2d / ssb_memcpy_d2h in this process so far (all contexts). */
void ssb_transfer_bytes(uint64_t* h2d, uint64_t* d2h);
int ssb_memset(ssb_ctx* ctx, void* dst, int value, size_t bytes);            /* async */
/* 1 when `ptr` is device memory, 0 for host memory (pageable or pinned). Lets ScanView accept
 * views over either (cursor/core/scan_view.h:35 takes any readable pointer). */
int ssb_pointer_is_device(const void* ptr);

/* bool-per-row (device) <-> bitmap (device). rows may be any value; the bitmap must
 * hold ceil(rows/32) words. Boundary conversion for bit_pointers.h:528-534. */
int ssb_nulls_pack(ssb_ctx* ctx, const uint8_t* d_bools, int64_t rows, uint32_t* d_bitmap);
int ssb_nulls_unpack(ssb_ctx* ctx, const uint32_t* d_bitmap, int64_t rows, uint8_t* d_bools);

/* Counter-based synthetic column generator, identical on host and device:
 * value(i) = splitmix64(seed ^ (stream * 0x9E3779B97F4A7C15) + i), then
 * kind 0: INT64 uniform in [lo, lo + span)   (span a power of two, or 0 = full 64 bit)
 * kind 1: INT64 u mod span + lo
 * kind 2: DOUBLE (u >> 44) * 2^-10           (exactly summable payload)
 * kind 3: DOUBLE (u >> 11) * 2^-53           (uniform [0,1))
 * kind 4: INT64 (row * lo) mod span          (no randomness: a permutation of [0, span) when
 *                                             gcd(lo, span) = 1; the C4 build key)
 * kind 5: DOUBLE (lo + u mod span) / 16       (dyadic values: products and sums stay exact, C5) */
int ssb_generate(ssb_ctx* ctx, void* d_out, int64_t rows, int64_t first_row, uint64_t seed,
                 uint64_t stream, int kind, int64_t lo, uint64_t span);
void ssb_generate_host(void* out, int64_t rows, int64_t first_row, uint64_t seed,
                       uint64_t stream, int kind, int64_t lo, uint64_t span);

/* ------------------------------------------------------------------ expressions */
/* Bound (fully typed) expression DAG, children before parents. This is what the
 * reference's Bound* factories produce as a tree of BoundExpression objects
 * (expression/templated/abstract_bound_expressions.h:63-230); here the whole tree is
 * handed over so that one kernel evaluates it. All type promotion has already been
 * done by the caller (explicit SSB_OP_CAST nodes), exactly as
 * expression/templated/bound_expression_factory.cc:44-123 inserts casts at bind time. */
enum {
  SSB_OP_INPUT = 1,        /* arg[0] = input column index */
  SSB_OP_CONST = 2,        /* imm; flags & SSB_NODE_NULL = typed NULL constant */
  SSB_OP_CAST = 3,         /* child type -> out_type (operators.h:50-57, C++ conversion) */
  SSB_OP_DATE_TO_DATETIME = 4,  /* operators.h:59-61 */
  SSB_OP_NEGATE = 10,      /* operators.h:63-71 (UINT32/UINT64 -> INT64) */
  SSB_OP_ADD = 11, SSB_OP_SUB = 12, SSB_OP_MUL = 13,   /* operators.h:73-86 */
  SSB_OP_DIV = 14,         /* C++ '/' on the (common) operand type, operators.h:88-91 */
  SSB_OP_MOD = 15,         /* C++ '%'; FLOAT/DOUBLE operands: int64 % int64, operators.h:93-106 */
  SSB_OP_IS_ODD = 16, SSB_OP_IS_EVEN = 17,             /* operators.h:108-128 */
  SSB_OP_EQ = 20, SSB_OP_NE = 21, SSB_OP_LT = 22, SSB_OP_LE = 23, SSB_OP_GT = 24,
  SSB_OP_GE = 25,          /* operators.h:185-294 incl. the mixed signed/unsigned overloads */
  SSB_OP_AND = 30, SSB_OP_OR = 31, SSB_OP_XOR = 32, SSB_OP_AND_NOT = 33, SSB_OP_NOT = 34,
                           /* SQL three-valued logic, elementary_bound_expressions.cc:270-506 */
  SSB_OP_BIT_AND = 40, SSB_OP_BIT_OR = 41, SSB_OP_BIT_XOR = 42, SSB_OP_BIT_AND_NOT = 43,
  SSB_OP_BIT_NOT = 44, SSB_OP_SHL = 45, SSB_OP_SHR = 46,  /* operators.h:150-183 */
  SSB_OP_IS_NULL = 50,     /* elementary_bound_expressions.cc:58-100 */
  SSB_OP_IF_NULL = 51,     /* arg0 unless NULL, else arg1 */
  SSB_OP_IF = 52,          /* arg0 ? arg1 : arg2; NULL condition selects arg2 */
  SSB_OP_NULLING_IF = 53   /* as IF, NULL condition gives NULL */
};

enum {
  SSB_NODE_NULL = 1,            /* CONST: the constant is NULL */
  SSB_NODE_ZERO_NULLS = 2,      /* DIV/MOD: divisor == 0 -> result NULL (…Nulling variants) */
  SSB_NODE_ZERO_FAILS = 4,      /* DIV/MOD: divisor == 0 -> ERROR_EVALUATION_ERROR (…Signaling) */
  SSB_NODE_GUARDED = 8          /* with ZERO_FAILS: arg[2] is a BOOL node; the zero divisor fails only on rows where it is
                                   TRUE and not NULL. The reference evaluates a sub-expression only on the rows its skip
                                   vector leaves (the taken branch of IF / CASE, the undecided side of AND / OR, rows whose
                                   other operand is not NULL, rows a Filter below kept: elementary_bound_expressions.cc:
                                   406-539,541-1050, binary_column_computers.h:137-166); the caller states that set here. */
};

typedef struct {
  int32_t op;          /* SSB_OP_* */
  int32_t out_type;    /* SSB_* type of the node's value */
  int32_t arg[3];      /* child node indices (< own index), -1 = unused */
  int32_t flags;       /* SSB_NODE_* */
  union { int64_t i64; uint64_t u64; double f64; float f32; int32_t i32; uint32_t u32; uint8_t b; } imm;
} ssb_expr_node;

typedef struct ssb_program ssb_program;

/* Compiles an expression DAG into one fused kernel configuration.
 *   outputs[j]   node whose value becomes output column j (Compute / Project;
 *                cursor/core/compute.cc:49-56, project.cc:49-59)
 *   predicate    node index of a BOOL predicate or -1. With a predicate the run is a
 *                Filter (cursor/core/filter.cc:96-230): rows where the predicate is
 *                true and not NULL are kept, in input order.
 * input_nullable[i] != 0 declares that input column i may carry a null bitmap. */
int ssb_program_create(ssb_ctx* ctx, const ssb_expr_node* nodes, int32_t n_nodes,
                       int32_t n_inputs, const int32_t* input_types,
                       const int32_t* input_nullable,
                       const int32_t* outputs, int32_t n_outputs, int32_t predicate,
                       ssb_program** out);
void ssb_program_destroy(ssb_program* prog);
int32_t ssb_program_output_type(const ssb_program* prog, int32_t j);
int32_t ssb_program_output_nullable(const ssb_program* prog, int32_t j);
/* Bytes of HBM traffic the algorithm needs per input row / per output row (roofline). */
int32_t ssb_program_bytes_per_input_row(const ssb_program* prog);
int32_t ssb_program_bytes_per_output_row(const ssb_program* prog);

/* The same validation and planning without a device: type-checks the DAG as the bound expression
 * tree does (expression/infrastructure/bound_expression_tree.h, the nullability rules of
 * expression/templated/bound_expression_factory.h) and lays the kernel's shared memory out for tiles of
 * `tile` rows within `smem_budget` bytes per CTA. For planners and for CPU-side tests; needs no context.
 * Returns 0 or an SSB_ERROR_* code with a message in err (err_len bytes, may be NULL). */
typedef struct {
  int32_t tile, stages, smem_bytes;      /* rows per tile, TMA stages in flight, bytes of shared memory per CTA */
  int32_t n_insn, n_tmp;                 /* instructions of the accumulator machine, temporaries */
  int32_t bytes_per_input_row, bytes_per_output_row;
  int32_t has_signaling;                 /* some node can raise ERROR_EVALUATION_ERROR */
  int32_t n_outputs;
  int32_t out_types[16], out_nullable[16];
} ssb_plan_info;
int ssb_program_plan(const ssb_expr_node* nodes, int32_t n_nodes,
                     int32_t n_inputs, const int32_t* input_types, const int32_t* input_nullable,
                     const int32_t* outputs, int32_t n_outputs, int32_t predicate,
                     int32_t tile, uint32_t smem_budget, ssb_plan_info* info, char* err, int32_t err_len);

/* Runs the program over `rows` rows. outputs[j].data must hold `rows` elements (Filter
 * may keep every row); outputs[j].nulls must be non-NULL (ceil(rows/32)+1 words) when
 * ssb_program_output_nullable(j). d_out_rows (device int64, may be NULL without a
 * predicate) receives the number of rows written. Asynchronous. */
int ssb_program_run(ssb_program* prog, const ssb_column* inputs, int64_t rows,
                    const ssb_column* outputs, int64_t* d_out_rows);
/* Same, then waits and returns the count on the host; 104 if a signaling op failed. */
int ssb_program_run_sync(ssb_program* prog, const ssb_column* inputs, int64_t rows,
                         const ssb_column* outputs, int64_t* out_rows);
/* After a sync: 0, or 104 if any ZERO_FAILS node saw a zero divisor since the last check. */
int ssb_program_check_failure(ssb_program* prog);

/* ------------------------------------------------------------------ group-by */
/* Replaces GroupAggregateCursor / RowHashSet / Aggregator (cursor/core/aggregate_groups.cc:
 * 332-433, cursor/infrastructure/row_hash_set.cc:458-518, cursor/core/aggregator.cc:206-221,
 * column_aggregator.cc:108-226). supersonic::Aggregation numbering (supersonic.proto:91-99). */
enum { SSB_AGG_SUM = 0, SSB_AGG_MIN = 1, SSB_AGG_MAX = 2, SSB_AGG_COUNT = 3,
       SSB_AGG_FIRST = 5, SSB_AGG_LAST = 6 };

typedef struct {
  int32_t fn;          /* SSB_AGG_* */
  int32_t input;       /* index into the `values` array of ssb_group_update, -1 = COUNT(*) */
  int32_t in_type;     /* SSB_* of the input column (ignored for COUNT(*)) */
  int32_t out_type;    /* SSB_* of the result (SUM/MIN/MAX: numeric; COUNT: UINT64 ...) */
  int32_t in_nullable; /* input column may carry a null bitmap */
  int32_t reserved;
} ssb_agg_spec;

typedef struct ssb_group ssb_group;

/* n_keys == 0 gives ScalarAggregate (cursor/core/aggregate_scalar.cc:40-90): one group. */
int ssb_group_create(ssb_ctx* ctx, int32_t n_keys, const int32_t* key_types,
                     const int32_t* key_nullable, int32_t n_aggs, const ssb_agg_spec* aggs,
                     int64_t expected_groups, ssb_group** out);
void ssb_group_destroy(ssb_group* g);
/* Accumulates `rows` rows; may be called repeatedly (row order across calls = call order). */
int ssb_group_update(ssb_group* g, const ssb_column* keys, const ssb_column* values,
                     int64_t rows);
/* GroupAggregate fused with its row-wise child (Filter / Compute over the scan): `prog` is
 * evaluated per row inside the aggregation kernel and nothing is materialised in between. The
 * program's outputs must be, in order, the key columns of `g` followed by its aggregate input
 * columns (ssb_agg_spec.input indexes the latter); its predicate, if any, filters the rows.
 * Same result as ssb_program_run followed by ssb_group_update (GroupAggregateCursor pulling
 * from FilterCursor / ComputeCursor: aggregate_groups.cc:332-433, filter.cc:96-230,
 * compute.cc:49-56). Plans with many groups are materialised slice by slice internally.
 * Returns 104 when a signaling expression failed. Synchronises. */
int ssb_group_update_program(ssb_group* g, ssb_program* prog, const ssb_column* inputs, int64_t rows);
/* Tooling / tests (needs no device): ssb_group_update_program runs calls of 64M rows and more (SSB200_JIT_MIN_ROWS;
 * SSB200_GROUP_JIT=1 always, 0 never) on a kernel compiled at run time for the plan -- the bound expression program,
 * column types and aggregate list become compile-time constants of csrc/jit_rows.h, NVRTC produces the sm_100a
 * cubin (the reference instead instantiates one column primitive per operator and type ahead of time:
 * expression/vector/vector_primitives.h, expression/templated/bound_expression_factory.h). This entry compiles the
 * kernel of one plan the same way and returns the generated source (or the compiler log) in `text`; `groups` =
 * CTA-local group entries (1..8), threads / rows_per_thread 0 = the defaults. */
/* Process-wide counters of the run-time compiled kernels: distinct kernels compiled and loaded, the time that took
 * (NVRTC + cudaLibraryLoadData, milliseconds), and how many times such a kernel was launched. */
void ssb_jit_stats(int64_t* kernels_compiled, double* compile_ms, int64_t* launches);
int ssb_jit_rows_compile(const ssb_expr_node* nodes, int32_t n_nodes, int32_t n_inputs, const int32_t* input_types,
                         const int32_t* input_nullable, const int32_t* outputs, int32_t n_outputs, int32_t predicate,
                         int32_t n_keys, int32_t n_aggs, const ssb_agg_spec* aggs, int32_t groups, int32_t threads,
                         int32_t rows_per_thread, char* text, int64_t text_cap, int64_t* cubin_bytes);
/* Compacts the table into dense result columns owned by `g` (valid until destroy or the
 * next update). Synchronises. A NULL key is a group of its own (row_hash_set.cc:81-90);
 * an aggregate over only-NULL inputs is NULL (column_aggregator.cc:108-125). */
int ssb_group_finalize(ssb_group* g, int64_t* n_groups, ssb_column* key_out, ssb_column* agg_out);
/* Adds the groups of a finalized `src` table (same specification) into `dst`: the merge
 * step of the row-range sharded aggregate. Both on the same device. */
int ssb_group_merge(ssb_group* dst, int64_t n_groups, const ssb_column* key_cols,
                    const ssb_column* agg_cols);

/* AggregateClusters (cursor/core/aggregate_clusters.cc:67-125,233-300): rows with equal keys that are consecutive
 * in the input form a cluster (NULL equals NULL, values by operator==); a key that comes back later starts a new
 * cluster. d_ids[i] = cluster of row i (0, 1, 2, ... in input order), d_starts[c] = first row of cluster c (room
 * for `rows` entries). Aggregating by d_ids with ssb_group_* and ordering by it gives the reference's output.
 * n_keys == 0: one cluster. Synchronises. */
int ssb_cluster_ids(ssb_ctx* ctx, int32_t n_keys, const ssb_column* keys, int64_t rows, int64_t* d_ids,
                    int64_t* d_starts, int64_t* n_clusters);

/* ------------------------------------------------------------------ hash join */
/* Replaces HashIndexOnMaterializedCursor + ResultCursor (cursor/core/hash_join.cc:604-625,
 * 707-831) and RowHashSet/RowHashMultiSet (row_hash_set.cc:424-608). */
enum { SSB_JOIN_INNER = 0, SSB_JOIN_LEFT_OUTER = 1 };
enum { SSB_KEYS_NOT_UNIQUE = 0, SSB_KEYS_UNIQUE = 1,
       SSB_KEYS_COMPACT_TABLE = 0x100 /* or-ed in: size the table rows / 0.6 instead of the next power of two >= 2 x rows */ };

typedef struct ssb_join ssb_join;

/* Builds the index over the rhs key columns (rows with a NULL key column never match,
 * hash_join.cc:67-76,616-617). The key columns must stay valid while the index lives. */
int ssb_join_build(ssb_ctx* ctx, int32_t n_keys, const ssb_column* keys, int64_t rows,
                   int32_t uniqueness, ssb_join** out);
void ssb_join_destroy(ssb_join* j);
/* Probes with the lhs key columns; the result is the list of (lhs row, rhs row) pairs in
 * lhs order and, per lhs row, rhs insertion order (hash_join.cc:793-831). LEFT_OUTER emits
 * rhs row -1 for unmatched lhs rows. Pair buffers are owned by `j` (valid until the next
 * probe). Synchronises to return the count. */
int ssb_join_probe(ssb_join* j, const ssb_column* keys, int64_t rows, int32_t join_type,
                   int64_t* n_pairs, const int64_t** d_lhs_rows, const int64_t** d_rhs_rows);

/* UNIQUE keys: probe and materialise in one kernel. Instead of the two row-id lists, the probe writes the result
 * columns themselves at the pairs' final positions (hash_join.cc:793-831 copies lhs and rhs cells pair by pair as it
 * finds them): out_cols[0 .. n_lhs) = lhs_cols gathered by the probe row, out_cols[n_lhs ..) = rhs_cols gathered by the
 * matched build row (for an attached index: by row_offsets[part] + row, i.e. rows of the concatenated build sides).
 * Source columns must be NOT NULL fixed-width columns; out_cols need room for `rows` rows. LEFT_OUTER: every lhs row
 * comes out, rhs cells of unmatched rows are zero and d_matched[row] (one byte per row, may be NULL) says which
 * matched. *n_rows = rows written. Synchronises. */
int ssb_join_probe_materialize(ssb_join* j, const ssb_column* keys, int64_t rows, int32_t join_type, int32_t n_lhs,
                               const ssb_column* lhs_cols, int32_t n_rhs, const ssb_column* rhs_cols,
                               const ssb_column* out_cols, uint8_t* d_matched, int64_t* n_rows);

/* The replicated form of the sharded join (UNIQUE single-column keys): the build side is
 * hash-partitioned over the ranks (ssb_partition_rows), every rank builds the index of the part it
 * received (ssb_join_build), the tables are all-gathered, and ssb_join_attach_parts makes an index over
 * the `n_parts` gathered tables: a probe looks its key up in the table of the key's part and reports
 * rhs rows as row_offsets[part] + (row inside that part's build input). No probe row ever moves and the
 * output keeps lhs order (hash_join.cc:793-831). ssb_join_table exposes a built table: `capacity`
 * slots of 16 bytes (build with SSB_KEYS_COMPACT_TABLE to keep what travels small). The attached index owns
 * none of the tables. */
int ssb_join_table(const ssb_join* j, const void** d_slots, int64_t* capacity);
int ssb_join_attach_parts(ssb_ctx* ctx, int32_t key_type, int32_t n_parts, const void* const* d_slots,
                          const int64_t* capacities, const int64_t* row_offsets, ssb_join** out);

/* Stable hash partition of rows for the multi-GPU join redistribution (the reference
 * has no counterpart: cursor/core/hash_join.cc runs on one thread). part(row) = high bits of
 * the join key hash scaled to [0, n_parts); integer keys of different widths hash alike, so the
 * two sides of a join agree. Rows with a NULL key column never match (hash_join.cc:67-76) and
 * are assigned to `null_part`; null_part == n_parts sets them aside in an extra part behind the hash
 * parts (h_counts then has n_parts + 1 entries). d_perm[rows] receives the row ids grouped by part,
 * ascending inside each part; h_counts[n_parts] (HOST memory) the rows per part. Synchronises. */
int ssb_partition_rows(ssb_ctx* ctx, int32_t n_keys, const ssb_column* keys, int64_t rows,
                       int32_t n_parts, int32_t null_part, int64_t* d_perm, int64_t* h_counts);

/* dst[i] = src[idx[i]]; idx[i] < 0 -> NULL (base/infrastructure/copy_column.cc:112-127,
 * 200-286). dst.nulls may be NULL when neither src.nulls nor negative indices occur. */
int ssb_gather(ssb_ctx* ctx, const ssb_column* src, const int64_t* d_idx, int64_t n,
               const ssb_column* dst);

/* dst[idx[i]] = src[i] for i < n (idx[i] < 0: skipped; indices must be distinct). The inverse of
 * ssb_gather; the sharded UNIQUE join uses it to put returned matches back at their lhs rows
 * (the reference's ResultCursor writes matches in lhs order as it goes: hash_join.cc:793-831). */
int ssb_scatter(ssb_ctx* ctx, const ssb_column* src, const int64_t* d_idx, int64_t n, const ssb_column* dst);

/* ------------------------------------------------------------------ sort */
/* Replaces SortPermutation / SortTypedColumn (cursor/core/sort.cc:150-322,781-805): writes
 * the permutation that sorts `rows` rows by the key columns (first key most significant;
 * descending[k] != 0 = DESCENDING; ASCENDING puts NULLs first, DESCENDING last,
 * sort.cc:174-238). Stable (the reference is not: ties may differ). */
int ssb_sort_permutation(ssb_ctx* ctx, int32_t n_keys, const ssb_column* keys,
                         const int32_t* descending, int64_t rows, int64_t* d_perm);

/* ------------------------------------------------------------------ STRING / BINARY columns */
/* Replaces the StringPiece-into-arena cells of the reference (base/infrastructure/types.h:53-68, block.h:259-281,
 * base/memory/arena.h:48) and their comparisons (utils/strings/stringpiece.h:268-283: memcmp over the common
 * prefix, the shorter one first; row_hash_set.cc:424-498 hashes the bytes and confirms with ==). On the device a
 * variable-length column is (d_offsets INT64[rows + 1], d_bytes). ssb_string_rank writes dense ORDER-PRESERVING
 * codes: d_codes[i] < d_codes[j] iff string i < string j, equal iff the bytes are equal; GroupAggregate, HashJoin,
 * Sort and the comparison operators over STRING keys then run on INT64 code columns through the entry points
 * above. d_first_rows[c] (room for `rows` entries, or NULL) = a row holding the string with code c: gathering
 * those rows gives the sorted dictionary. max_len = an upper bound of the lengths. Synchronises. */
int ssb_string_rank(ssb_ctx* ctx, const int64_t* d_offsets, const uint8_t* d_bytes, int64_t rows, int64_t max_len,
                    int64_t* d_codes, int64_t* d_first_rows, int64_t* n_distinct);
/* Gather of variable-length cells in two steps (the byte count is needed to allocate the result):
 * d_out_offsets[n + 1] = offsets of the strings d_idx[0..n) laid end to end (d_idx NULL = identity, an index < 0 =
 * an empty cell), *total_bytes = their total length (synchronises); then the bytes themselves (asynchronous).
 * Replaces the deep copy of copy_column.cc:129-198 (ColumnCopier with an arena). */
int ssb_string_gather_offsets(ssb_ctx* ctx, const int64_t* d_offsets, const int64_t* d_idx, int64_t n,
                              int64_t* d_out_offsets, int64_t* total_bytes);
int ssb_string_gather_bytes(ssb_ctx* ctx, const int64_t* d_offsets, const uint8_t* d_bytes, const int64_t* d_idx,
                            int64_t n, const int64_t* d_out_offsets, uint8_t* d_out_bytes);
/* d_dst[i] = d_src[i] + delta for i < n: appends one offsets array to another when two variable-length columns
 * (or dictionaries) are concatenated (the bytes themselves are appended with ssb_memcpy_d2d). Asynchronous. */
int ssb_string_shift_offsets(ssb_ctx* ctx, const int64_t* d_src, int64_t n, int64_t delta, int64_t* d_dst);

/* ------------------------------------------------------------------ multi-GPU */
/* One process per GPU; tables are sharded by contiguous row ranges. Compute / Project / Filter need no
 * exchange (run ssb_program_run on the shard). The operators whose CPU form keeps global state exchange
 * once: GroupAggregate (aggregate_groups.cc:332-433, one hash set over the whole input) and HashJoin
 * (hash_join.cc:406-517, one index over the whole rhs). The reference itself is single-threaded; these
 * entry points are what a sharded HashJoinOperation / GroupAggregate of the C++ layer calls.
 * NCCL (over NVLink / NVSwitch) is bound at run time; every transfer is ordered on the context's stream. */
typedef struct ssb_comm ssb_comm;
#define SSB_COMM_ID_BYTES 128
/* Rank 0 creates an id and hands it to the other ranks by any means (file, socket, MPI, torch.distributed). */
int ssb_comm_unique_id(uint8_t* id /* [SSB_COMM_ID_BYTES] */);
int ssb_comm_create(ssb_ctx* ctx, const uint8_t* id, int32_t world, int32_t rank, ssb_comm** out);
/* Same, with the id passed through a file: rank 0 writes `path`, the others wait for it (timeout_s seconds). */
int ssb_comm_create_file(ssb_ctx* ctx, const char* path, int32_t world, int32_t rank, int32_t timeout_s, ssb_comm** out);
void ssb_comm_destroy(ssb_comm* comm);
int32_t ssb_comm_rank(const ssb_comm* comm);
int32_t ssb_comm_size(const ssb_comm* comm);
/* h_recv[r] = the count rank r holds for this rank in its h_send[this rank] (HOST arrays of `size` entries). */
int ssb_comm_exchange_counts(ssb_comm* comm, const int64_t* h_send, int64_t* h_recv);
/* Ragged all-to-all of n_cols device columns in ONE grouped NCCL operation: column i is cut into `size`
 * consecutive slices of send_rows[r] elements of width[i] bytes; slice r goes to rank r; the slices received
 * from ranks 0..size-1 land consecutively in recv[i] (recv_rows[r] elements each). Asynchronous. */
int ssb_comm_all_to_all(ssb_comm* comm, int32_t n_cols, const void* const* send, void* const* recv,
                        const int32_t* width, const int64_t* send_rows, const int64_t* recv_rows);

/* HashJoin over row-range shards with UNIQUE single-column keys (hash_join.cc:406-517,707-831 keep one index over
 * the whole rhs): the build rows are hash-partitioned and exchanged (one grouped send/recv for key + payload
 * columns), every rank builds the table of one hash part, tables and payload columns are all-gathered, and the
 * local lhs shard probes them in place -- no probe row moves, pairs come out in lhs order (hash_join.cc:793-831).
 * Collective: every rank of `comm` calls build with its shard of the rhs (key: any fixed-width type, may be
 * nullable -- rows with a NULL key never match; payload: NOT NULL fixed-width columns). probe = ssb_join_probe on
 * the local lhs shard; d_rhs_rows index the gathered payload columns (ssb_shard_join_payload: `rows` = rhs rows
 * with a key over all ranks, in (hash part, global insertion) order). */
typedef struct ssb_shard_join ssb_shard_join;
int ssb_shard_join_build(ssb_comm* comm, const ssb_column* key, int32_t n_payload, const ssb_column* payload,
                         int64_t rows, ssb_shard_join** out);
int ssb_shard_join_probe(ssb_shard_join* j, const ssb_column* key, int64_t rows, int32_t join_type, int64_t* n_pairs,
                         const int64_t** d_lhs_rows, const int64_t** d_rhs_rows);
int ssb_shard_join_payload(const ssb_shard_join* j, int32_t i, ssb_column* out, int64_t* rows);
/* Which form the collective build chose (the same on every rank): 0 = hash-partitioned tables (build rows to the owner of
 * their key's hash part, per-part compact tables all-gathered), 1 = dense integer keys (key and payload columns
 * all-gathered as they are, every rank builds the direct index key - min -> row; no table travels). */
int ssb_shard_join_form(const ssb_shard_join* j);
/* ssb_join_probe_materialize on the sharded index: rhs columns are named by their payload index. */
int ssb_shard_join_probe_materialize(ssb_shard_join* j, const ssb_column* keys, int64_t rows, int32_t join_type,
                                     int32_t n_lhs, const ssb_column* lhs_cols, int32_t n_rhs, const int32_t* rhs_payload,
                                     const ssb_column* out_cols, uint8_t* d_matched, int64_t* n_rows);
void ssb_shard_join_destroy(ssb_shard_join* j);

/* The exchange step of a row-range sharded GroupAggregate / ScalarAggregate ("a reduce for global aggregates"):
 * every rank has aggregated its shard into `g`. The partial groups are hash-partitioned by key over the ranks,
 * exchanged (one all-to-all), merged by the rank that owns the key -- a reduce-scatter by key, each rank
 * merges about n_groups rows in total instead of (size - 1) x n_groups -- and the merged ranges are
 * all-gathered, so that every rank ends with the whole result (groups in no particular order, as in
 * the reference: aggregate_groups_test.cc:74-93 sorts before comparing). NULL keys and all-NULL aggregates
 * travel with their is_null flags. Result columns are owned by `g` like those of ssb_group_finalize.
 * Collective: every rank of `comm` must call it with a table of the same specification. Synchronises. */
int ssb_shard_group_merge(ssb_comm* comm, ssb_group* g, int64_t* n_groups, ssb_column* key_out, ssb_column* agg_out);

#ifdef __cplusplus
}
#endif
#endif  /* SUPERSONIC_B200_H_ */
