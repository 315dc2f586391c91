// comm.h -- internal view of the communicator (comm.cu) for the sharded operators.
#ifndef SSB_CSRC_COMM_H_
#define SSB_CSRC_COMM_H_
#include <stdint.h>

#include "../../include/supersonic_b200.h"

namespace ssb {
int comm_world(const ssb_comm* c);
int comm_rank(const ssb_comm* c);
ssb_ctx* comm_ctx(const ssb_comm* c);
// Every rank contributes n (<= 8) int64 values; h_all[r * n + i] = value i of rank r. Synchronises.
int comm_all_gather_counts(ssb_comm* c, const int64_t* h_mine, int n, int64_t* h_all);
// h_recv[r] = what rank r put into its h_send[this rank]. Synchronises.
int comm_exchange_counts(ssb_comm* c, const int64_t* h_send, int64_t* h_recv);
// Ragged all-to-all / all-gather of several columns in one NCCL group (see comm.cu). Asynchronous.
int comm_all_to_all_v(ssb_comm* c, int n_cols, const void* const* send, void* const* recv, const int32_t* width,
                      const int64_t* send_rows, const int64_t* recv_rows);
int comm_all_gather_v(ssb_comm* c, int n_cols, const void* const* send, void* const* recv, const int32_t* width,
                      const int64_t* all_rows);
// join.cu: smallest / largest value (as signed 64-bit containers) and number of non-NULL keys of an integer key
// column: out = {min, max, count}; eligible = 0 when the column's type has no dense form. Synchronises.
int join_key_range(ssb_ctx* ctx, const ssb_column* key, int64_t rows, long long out[3], int* eligible);
// True when a UNIQUE index over `rows` keys spanning [lo, hi] is built as a direct index (JoinTable::dense_rows).
bool join_dense_fits(long long lo, long long hi, long long count, long long rows);
}  // namespace ssb
#endif  // SSB_CSRC_COMM_H_
