// internal.h -- plumbing shared by the host-layer sources: the device session (context of
// libssb200.so), device-resident tables and the lowered form of bound expressions.
#ifndef SUPERSONIC_B200_HOST_INTERNAL_H_
#define SUPERSONIC_B200_HOST_INTERNAL_H_

#include <atomic>

#include "../../../include/supersonic_b200.h"
#include "supersonic/supersonic.h"

namespace supersonic {
namespace internal {

// ---- the fused description of a chain of row-wise operators over one scan ----------------
struct RowwisePlan {
  View base;                       // the scanned view (host or device pointers), or
  std::shared_ptr<Cursor> source;  // a non-row-wise child whose whole output is the base
  TupleSchema base_schema;
  TupleSchema schema;              // schema of the plan's output
  vector<NodePtr> outputs;         // output column j as an expression over the base columns
  NodePtr predicate;               // rows kept iff true and not NULL; NULL = keep all
  RowwisePlan() : base(TupleSchema()) {}
};

// Describes `op` as a row-wise plan: fused when the operation is row-wise itself, otherwise
// the identity over a cursor created from it.
FailureOrVoid DescribeAny(const Operation* op, RowwisePlan* plan);

// Rewrites `node` (bound against a schema whose column i is computed by `inputs[i]`) into an
// expression over the columns `inputs` are expressed in. Shared sub-trees stay shared.
NodePtr Substitute(const NodePtr& node, const vector<NodePtr>& inputs);
// Rewrites `node` so that every signaling operator in it (not below one of the nodes in `below`) fails
// only on the rows the reference would evaluate it on: the rows of `rows` (a BOOL expression, NULL =
// every row) narrowed by the IF / AND / OR / IFNULL / NULL-operand structure above it.
NodePtr GuardSignaling(const NodePtr& node, const vector<NodePtr>& below, const NodePtr& rows);
NodePtr MakeInputNode(const TupleSchema& schema, int position);
NodePtr MakeBinaryLogic(int op, const NodePtr& a, const NodePtr& b);

// ---- device session ------------------------------------------------------------------------
// One per process and device; created on first use. Fails loudly when no B200 is present.
class Session {
 public:
  static FailureOr<Session*> Get();
  ssb_ctx* ctx() const { return ctx_; }
  // Additional contexts (own stream each) on the same device, for copy/compute overlap.
  FailureOr<ssb_ctx*> lane(int i);
  // Waits for everything queued on the session's streams (the main context and the lanes that exist).
  void SyncAll();
  int device() const { return device_; }
  Exception* Error(int code, const char* what) const;
  static Exception* ErrorOn(ssb_ctx* ctx, int code, const char* what);
 private:
  Session(ssb_ctx* c, int device) : ctx_(c), device_(device) { lanes_[0] = lanes_[1] = NULL; }
  ssb_ctx* ctx_;
  int device_;
  ssb_ctx* lanes_[2];
};

// Lowers bound expression nodes to the C ABI's node array. `used` lists the referenced input
// columns in program order; `outs` / `pred` index into `nodes`.
void LowerProgram(const vector<NodePtr>& outputs, const NodePtr& predicate, vector<ssb_expr_node>* nodes,
                  vector<int>* used, vector<int32_t>* outs, int* pred);

// Caching allocator for device and pinned host memory: cudaMalloc / cudaMallocHost cost
// milliseconds, cursors are created per query. Blocks are rounded up to 1 MiB granules and
// returned to a free list instead of the driver (all users synchronise before releasing).
class MemoryPool {
 public:
  enum Kind { DEVICE = 0, PINNED = 1 };
  static FailureOr<void*> Acquire(Kind kind, size_t bytes, size_t* granted);
  static void Release(Kind kind, void* ptr, size_t granted);
};

// Device memory owned through the session.
class DeviceBuffer {
 public:
  DeviceBuffer() : ptr_(NULL), bytes_(0), granted_(0) {}
  ~DeviceBuffer() { Free(); }
  FailureOrVoid Allocate(size_t bytes);
  void Free();
  void* get() const { return ptr_; }
  size_t size() const { return bytes_; }
 private:
  DeviceBuffer(const DeviceBuffer&);
  void operator=(const DeviceBuffer&);
  void* ptr_;
  size_t bytes_;
  size_t granted_;
};

// STRING / BINARY columns. On the device a variable-length column is a column of INT64 codes into a
// dictionary of its distinct values in sorted order (code c = the c-th smallest value, so the codes compare like the
// strings: utils/strings/stringpiece.h:268-283): the relational kernels run on the codes, the bytes are touched by
// ssb_string_rank (building the codes) and when a result goes back to the host.
struct HostDict { vector<int64> offsets; vector<char> bytes; };
struct DeviceDict {
  DeviceBuffer offsets;   // INT64[n + 1]
  DeviceBuffer bytes;
  int64 n, total_bytes, max_len;
  std::shared_ptr<const HostDict> host;   // filled by the first download
  DeviceDict() : n(0), total_bytes(0), max_len(0) {}
};
inline bool IsVariableLength(DataType t) { return t == STRING || t == BINARY; }
inline DataType DeviceType(DataType t) { return IsVariableLength(t) ? INT64 : t; }   // what the kernels see
inline size_t DeviceWidth(DataType t) { return IsVariableLength(t) ? 8 : GetTypeInfo(t).size(); }
TupleSchema DeviceSchema(const TupleSchema& s);   // variable-length attributes as INT64 (codes)

// A set of equally long device columns: either borrowed (device pointers handed in through
// ScanView) or owned buffers.
struct DeviceColumnRef {
  ssb_column col;
  std::shared_ptr<DeviceBuffer> data, nulls;   // empty when borrowed
  std::shared_ptr<DeviceDict> dict;            // variable-length columns: col holds INT64 codes into it
};

// StringPiece cells of a host column -> codes + dictionary on the device.
FailureOrVoid UploadStringColumn(Session* s, const StringPiece* cells, const bool* is_null, rowcount_t rows, DeviceColumnRef* out);
// Codes + dictionary -> StringPiece cells pointing into storage the block keeps alive (NULL rows are left empty).
FailureOrVoid DownloadStringColumn(Session* s, const DeviceColumnRef& col, int64 rows, StringPiece* cells, Block* owner);
// Re-encodes the variable-length columns `cols` (rows[i] rows each) and the constants against ONE merged dictionary,
// so that their codes compare with each other; codes of the constants (dense ranks in the merged dictionary) are
// returned. Columns that already share one dictionary, with no constants, are left as they are.
FailureOrVoid UnifyDictionaries(Session* s, const vector<DeviceColumnRef*>& cols, const vector<int64>& rows,
                                const vector<string>& constants, vector<int64>* constant_codes);
struct DeviceTable {
  TupleSchema schema;
  vector<DeviceColumnRef> columns;
  int64 rows;
  DeviceTable() : rows(0) {}
  // Allocates storage for `capacity` rows of `s` (bitmaps for NULLABLE attributes).
  FailureOrVoid Allocate(const TupleSchema& s, int64 capacity, bool force_nulls = false);
  // Copies rows [0, rows) into a host block (converting bitmaps to bool per row).
  FailureOrVoid Download(Block* block, rowcount_t block_offset = 0) const;
};

// True when `p` points to device memory.
bool IsDevicePointer(const void* p);

// Uploads rows [offset, offset+rows) of a host view's columns `cols` into device columns.
// Device-resident source columns are referenced, not copied (no nulls conversion is possible
// for them: a device column's is_null must be NULL).
FailureOrVoid UploadColumns(const View& view, const vector<int>& cols, rowcount_t offset, rowcount_t rows,
                            DeviceTable* out);

// ---- compiled expression program ----------------------------------------------------------
class DeviceProgram {
 public:
  ~DeviceProgram();
  // outputs / predicate are expressions over `input_schema`; only referenced input columns
  // are bound (used_inputs() lists them in program order).
  static FailureOrOwned<DeviceProgram> Create(const TupleSchema& input_schema, const vector<NodePtr>& outputs,
                                              const NodePtr& predicate);
  const vector<int>& used_inputs() const { return used_; }
  bool has_predicate() const { return has_pred_; }
  int output_count() const { return n_out_; }
  ssb_program* handle() const { return prog_; }   // for the fused entry points of the C ABI
  // inputs: device columns in used_inputs() order. outputs must be allocated for `rows` rows.
  // Returns the number of rows written.
  FailureOr<int64> Run(const vector<ssb_column>& inputs, int64 rows, const vector<ssb_column>& outputs);
 private:
  DeviceProgram() : prog_(NULL), has_pred_(false), n_out_(0) {}
  ssb_program* prog_;
  vector<int> used_;
  bool has_pred_;
  int n_out_;
};

// ---- cursors -------------------------------------------------------------------------------
// Base of every GPU cursor: produces its complete result as a DeviceTable on first use, then
// serves Next() from a host copy. Parents that are GPU cursors take the DeviceTable directly.
class GpuCursor : public Cursor {
 public:
  virtual const TupleSchema& schema() const { return schema_; }
  virtual ResultView Next(rowcount_t max_row_count);
  virtual void Interrupt() { interrupted_.store(true, std::memory_order_relaxed); }
  virtual void AppendDebugDescription(string* target) const { target->append(name_); }
  // Runs the operator; the table stays valid while the cursor lives.
  FailureOr<const DeviceTable*> Produce();
 protected:
  GpuCursor(const TupleSchema& schema, BufferAllocator* allocator, const char* name)
      : schema_(schema), allocator_(allocator), name_(name), produced_(false), interrupted_(false),
        view_(schema), offset_(0) {}
  virtual FailureOrVoid Run(DeviceTable* result) = 0;
  virtual void DropFusion() {}   // a child was replaced (ApplyToChildren): plans fused with it no longer apply
  bool interrupted() const { return interrupted_.load(std::memory_order_relaxed); }
  BufferAllocator* allocator() const { return allocator_; }
 private:
  TupleSchema schema_;
  BufferAllocator* allocator_;
  const char* name_;
  bool produced_;
  std::atomic<bool> interrupted_;
  DeviceTable result_;
  std::unique_ptr<Block> host_;
  View view_;
  rowcount_t offset_;
};

// Obtains the complete output of `child` as device columns: directly when it is a GPU cursor,
// otherwise by draining it through Next() and uploading.
FailureOrVoid MaterializeOnDevice(Cursor* child, DeviceTable* out, std::unique_ptr<Block>* host_keepalive);

}  // namespace internal
}  // namespace supersonic
#endif  // SUPERSONIC_B200_HOST_INTERNAL_H_
