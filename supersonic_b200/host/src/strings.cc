// strings.cc -- STRING / BINARY columns between the host mirror and the device.
//
// The reference keeps a variable-length cell as a StringPiece into an arena (base/infrastructure/types.h:53-68,
// block.h:259-281) and every operator compares / hashes / copies the bytes row by row. Here the bytes cross the
// boundary once: a host column is packed (offsets + bytes), copied to HBM and ranked by ssb_string_rank into dense
// order-preserving INT64 codes plus the sorted dictionary of its distinct values; operators work on the codes
// (csrc/strings.cu explains why that is exact), and a result column goes back as codes + dictionary.
#include <string.h>

#include "internal.h"

namespace supersonic {
namespace internal {

#define SSB_CALL(session, call, what)                              \
  do {                                                             \
    const int rc_ = (call);                                        \
    if (rc_ != 0) THROW((session)->Error(rc_, what));              \
  } while (0)

TupleSchema DeviceSchema(const TupleSchema& s) {
  TupleSchema out;
  for (int i = 0; i < s.attribute_count(); ++i) {
    const Attribute& a = s.attribute(i);
    out.add_attribute(Attribute(a.name(), DeviceType(a.type()), a.nullability()));
  }
  return out;
}

namespace {

// Ranks `rows` strings given as (offsets, bytes) on the device: codes into d_codes, the sorted dictionary into *dict.
FailureOrVoid RankIntoDictionary(Session* s, const int64_t* d_offsets, const uint8_t* d_bytes, int64 rows, int64 max_len,
                                 int64_t* d_codes, std::shared_ptr<DeviceDict>* dict) {
  std::shared_ptr<DeviceDict> d(new DeviceDict);
  d->max_len = max_len;
  if (rows > 0) {
    DeviceBuffer first_rows;
    PROPAGATE_ON_FAILURE(first_rows.Allocate(static_cast<size_t>(rows) * 8 + 128));
    int64_t distinct = 0;
    SSB_CALL(s, ssb_string_rank(s->ctx(), d_offsets, d_bytes, rows, max_len, d_codes, static_cast<int64_t*>(first_rows.get()), &distinct),
             "string ranking");
    d->n = distinct;
    PROPAGATE_ON_FAILURE(d->offsets.Allocate(static_cast<size_t>(distinct + 1) * 8 + 128));
    int64_t total = 0;
    SSB_CALL(s, ssb_string_gather_offsets(s->ctx(), d_offsets, static_cast<const int64_t*>(first_rows.get()), distinct,
                                          static_cast<int64_t*>(d->offsets.get()), &total), "dictionary offsets");
    d->total_bytes = total;
    PROPAGATE_ON_FAILURE(d->bytes.Allocate(static_cast<size_t>(total) + 128));
    SSB_CALL(s, ssb_string_gather_bytes(s->ctx(), d_offsets, d_bytes, static_cast<const int64_t*>(first_rows.get()), distinct,
                                        static_cast<const int64_t*>(d->offsets.get()), static_cast<uint8_t*>(d->bytes.get())),
             "dictionary bytes");
    SSB_CALL(s, ssb_ctx_sync(s->ctx()), "sync");
  } else {
    PROPAGATE_ON_FAILURE(d->offsets.Allocate(8 + 128));
    PROPAGATE_ON_FAILURE(d->bytes.Allocate(128));
    SSB_CALL(s, ssb_memset(s->ctx(), d->offsets.get(), 0, 8), "memset");
    SSB_CALL(s, ssb_ctx_sync(s->ctx()), "sync");
  }
  *dict = d;
  return Success();
}

}  // namespace

FailureOrVoid UploadStringColumn(Session* s, const StringPiece* cells, const bool* is_null, rowcount_t rows, DeviceColumnRef* out) {
  // pack: offsets + bytes (NULL cells are empty and never read back)
  vector<int64> offsets(rows + 1);
  int64 total = 0, max_len = 0;
  for (rowcount_t i = 0; i < rows; ++i) {
    offsets[i] = total;
    if (is_null == NULL || !is_null[i]) {
      const int64 len = static_cast<int64>(cells[i].size());
      total += len;
      if (len > max_len) max_len = len;
    }
  }
  offsets[rows] = total;
  vector<char> bytes(static_cast<size_t>(total) + 1);
  for (rowcount_t i = 0; i < rows; ++i) {
    const int64 len = offsets[i + 1] - offsets[i];
    if (len > 0) memcpy(&bytes[offsets[i]], cells[i].data(), static_cast<size_t>(len));
  }
  DeviceBuffer d_offsets, d_bytes;
  PROPAGATE_ON_FAILURE(d_offsets.Allocate(static_cast<size_t>(rows + 1) * 8 + 128));
  PROPAGATE_ON_FAILURE(d_bytes.Allocate(static_cast<size_t>(total) + 128));
  SSB_CALL(s, ssb_memcpy_h2d(s->ctx(), d_offsets.get(), offsets.data(), static_cast<size_t>(rows + 1) * 8), "upload string offsets");
  if (total > 0) SSB_CALL(s, ssb_memcpy_h2d(s->ctx(), d_bytes.get(), bytes.data(), static_cast<size_t>(total)), "upload string bytes");
  SSB_CALL(s, ssb_ctx_sync(s->ctx()), "sync");
  out->data.reset(new DeviceBuffer);
  PROPAGATE_ON_FAILURE(out->data->Allocate(static_cast<size_t>(rows) * 8 + 128));
  out->col.data = out->data->get();
  out->col.dtype = INT64;
  out->col.reserved = 0;
  PROPAGATE_ON_FAILURE(RankIntoDictionary(s, static_cast<const int64_t*>(d_offsets.get()), static_cast<const uint8_t*>(d_bytes.get()),
                                          static_cast<int64>(rows), max_len, static_cast<int64_t*>(out->col.data), &out->dict));
  return Success();
}

FailureOrVoid DownloadStringColumn(Session* s, const DeviceColumnRef& col, int64 rows, StringPiece* cells, Block* owner) {
  if (rows <= 0) return Success();
  if (!col.dict) THROW(new Exception(ERROR_UNKNOWN_ERROR, "internal: a variable-length device column without a dictionary"));
  DeviceDict* d = col.dict.get();
  if (!d->host) {
    std::shared_ptr<HostDict> h(new HostDict);
    h->offsets.resize(static_cast<size_t>(d->n) + 1);
    h->bytes.resize(static_cast<size_t>(d->total_bytes) + 1);
    SSB_CALL(s, ssb_memcpy_d2h(s->ctx(), h->offsets.data(), d->offsets.get(), static_cast<size_t>(d->n + 1) * 8), "download dictionary");
    if (d->total_bytes > 0) {
      SSB_CALL(s, ssb_memcpy_d2h(s->ctx(), h->bytes.data(), d->bytes.get(), static_cast<size_t>(d->total_bytes)), "download dictionary");
    }
    SSB_CALL(s, ssb_ctx_sync(s->ctx()), "sync");
    d->host = h;
  }
  vector<int64> codes(static_cast<size_t>(rows));
  SSB_CALL(s, ssb_memcpy_d2h(s->ctx(), codes.data(), col.col.data, static_cast<size_t>(rows) * 8), "download codes");
  SSB_CALL(s, ssb_ctx_sync(s->ctx()), "sync");
  const HostDict& h = *d->host;
  for (int64 i = 0; i < rows; ++i) {
    const int64 c = codes[static_cast<size_t>(i)];
    if (c < 0 || c >= d->n) { cells[i] = StringPiece(); continue; }   // a NULL row of a column whose dictionary is empty
    cells[i] = StringPiece(h.bytes.data() + h.offsets[static_cast<size_t>(c)],
                           static_cast<size_t>(h.offsets[static_cast<size_t>(c) + 1] - h.offsets[static_cast<size_t>(c)]));
  }
  owner->KeepAlive(d->host);
  return Success();
}

FailureOrVoid UnifyDictionaries(Session* s, const vector<DeviceColumnRef*>& cols, const vector<int64>& rows,
                                const vector<string>& constants, vector<int64>* constant_codes) {
  if (constant_codes) constant_codes->clear();
  // the distinct dictionaries, in order of first use
  vector<std::shared_ptr<DeviceDict> > dicts;
  vector<int> dict_of(cols.size(), -1);
  for (size_t i = 0; i < cols.size(); ++i) {
    if (!cols[i]->dict) THROW(new Exception(ERROR_UNKNOWN_ERROR, "internal: a variable-length device column without a dictionary"));
    size_t k = 0;
    while (k < dicts.size() && dicts[k].get() != cols[i]->dict.get()) ++k;
    if (k == dicts.size()) dicts.push_back(cols[i]->dict);
    dict_of[i] = static_cast<int>(k);
  }
  if (dicts.size() <= 1 && constants.empty()) return Success();
  // the pool: every dictionary's values, then the constants
  vector<int64> first(dicts.size() + 1, 0), byte_base(dicts.size() + 1, 0);
  int64 max_len = 0;
  for (size_t k = 0; k < dicts.size(); ++k) {
    first[k + 1] = first[k] + dicts[k]->n;
    byte_base[k + 1] = byte_base[k] + dicts[k]->total_bytes;
    if (dicts[k]->max_len > max_len) max_len = dicts[k]->max_len;
  }
  vector<int64> c_off(constants.size() + 1, 0);
  string c_bytes;
  for (size_t i = 0; i < constants.size(); ++i) {
    c_off[i] = byte_base[dicts.size()] + static_cast<int64>(c_bytes.size());
    c_bytes += constants[i];
    if (static_cast<int64>(constants[i].size()) > max_len) max_len = static_cast<int64>(constants[i].size());
  }
  const int64 pool = first[dicts.size()] + static_cast<int64>(constants.size());
  const int64 pool_bytes = byte_base[dicts.size()] + static_cast<int64>(c_bytes.size());
  c_off[constants.size()] = pool_bytes;
  DeviceBuffer p_off, p_bytes, p_codes;
  PROPAGATE_ON_FAILURE(p_off.Allocate(static_cast<size_t>(pool + 1) * 8 + 128));
  PROPAGATE_ON_FAILURE(p_bytes.Allocate(static_cast<size_t>(pool_bytes) + 128));
  PROPAGATE_ON_FAILURE(p_codes.Allocate(static_cast<size_t>(pool + 1) * 8 + 128));
  int64_t* po = static_cast<int64_t*>(p_off.get());
  for (size_t k = 0; k < dicts.size(); ++k) {
    SSB_CALL(s, ssb_string_shift_offsets(s->ctx(), static_cast<const int64_t*>(dicts[k]->offsets.get()), dicts[k]->n, byte_base[k],
                                         po + first[k]), "dictionary merge");
    if (dicts[k]->total_bytes > 0) {
      SSB_CALL(s, ssb_memcpy_d2d(s->ctx(), static_cast<char*>(p_bytes.get()) + byte_base[k], dicts[k]->bytes.get(),
                                 static_cast<size_t>(dicts[k]->total_bytes)), "dictionary merge");
    }
  }
  // the constants' offsets and the terminating offset
  SSB_CALL(s, ssb_memcpy_h2d(s->ctx(), po + first[dicts.size()], c_off.data(), (constants.size() + 1) * 8), "dictionary merge");
  if (!c_bytes.empty()) {
    SSB_CALL(s, ssb_memcpy_h2d(s->ctx(), static_cast<char*>(p_bytes.get()) + byte_base[dicts.size()], c_bytes.data(), c_bytes.size()),
             "dictionary merge");
  }
  SSB_CALL(s, ssb_ctx_sync(s->ctx()), "sync");
  std::shared_ptr<DeviceDict> merged;
  PROPAGATE_ON_FAILURE(RankIntoDictionary(s, po, static_cast<const uint8_t*>(p_bytes.get()), pool, max_len,
                                          static_cast<int64_t*>(p_codes.get()), &merged));
  // columns: new code = pool code of (old dictionary's segment + old code)
  for (size_t i = 0; i < cols.size(); ++i) {
    DeviceColumnRef* c = cols[i];
    std::shared_ptr<DeviceBuffer> fresh(new DeviceBuffer);
    PROPAGATE_ON_FAILURE(fresh->Allocate(static_cast<size_t>(rows[i]) * 8 + 128));
    if (rows[i] > 0 && dicts[dict_of[i]]->n > 0) {
      ssb_column map, dst;
      map.data = static_cast<int64_t*>(p_codes.get()) + first[dict_of[i]]; map.nulls = NULL; map.dtype = INT64; map.reserved = 0;
      dst = map; dst.data = fresh->get();
      SSB_CALL(s, ssb_gather(s->ctx(), &map, static_cast<const int64_t*>(c->col.data), rows[i], &dst), "re-encoding");
    } else if (rows[i] > 0) {
      SSB_CALL(s, ssb_memset(s->ctx(), fresh->get(), 0, static_cast<size_t>(rows[i]) * 8), "memset");
    }
    c->data = fresh;
    c->col.data = fresh->get();
    c->dict = merged;
  }
  if (!constants.empty()) {
    vector<int64> codes(constants.size());
    SSB_CALL(s, ssb_memcpy_d2h(s->ctx(), codes.data(), static_cast<int64_t*>(p_codes.get()) + first[dicts.size()], constants.size() * 8),
             "constant codes");
    SSB_CALL(s, ssb_ctx_sync(s->ctx()), "sync");
    if (constant_codes) *constant_codes = codes;
  }
  SSB_CALL(s, ssb_ctx_sync(s->ctx()), "sync");
  return Success();
}

}  // namespace internal
}  // namespace supersonic
