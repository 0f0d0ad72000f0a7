// SPDX-License-Identifier: Apache-2.0
#include "device_scan.h"

#include <arrow/util/bit_util.h>
#include <arrow/util/bitmap_ops.h>

#include <chrono>
#include <cstring>
#include <mutex>
#include <stdexcept>

#include "decode_types.h"
#include "filter_eval.h"
#include "general_merge.h"
#include "zstd_decode.h"
#include "page_plan.h"

namespace lsb {

namespace {

[[noreturn]] void fail(const std::string& m) { throw std::runtime_error(m); }

#define SCAN_CUDA(x)                                                                      \
  do {                                                                                    \
    cudaError_t e_ = (x);                                                                 \
    if (e_ != cudaSuccess) fail(std::string(#x) + " failed: " + cudaGetErrorString(e_));  \
  } while (0)

double now_ms() {
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

// ---------------------------------------------------------------------------------------
// pinned host pool: cudaMallocHost is expensive (page pinning), readers come and go per
// bucket, so blocks are recycled by power-of-two size class.
// ---------------------------------------------------------------------------------------
class PinnedPool {
 public:
  static PinnedPool& get() {
    static PinnedPool* p = new PinnedPool();  // leaked on purpose: outlives late release callbacks
    return *p;
  }
  std::shared_ptr<uint8_t> alloc(size_t n) {
    size_t cls = 1 << 16;
    while (cls < n) cls <<= 1;
    uint8_t* p = nullptr;
    {
      std::lock_guard<std::mutex> g(mu_);
      auto& fl = free_[cls];
      if (!fl.empty()) {
        p = fl.back();
        fl.pop_back();
        cached_ -= cls;
      }
    }
    if (!p) {
      // host-logic tests on machines without a GPU (framing of files, tests/test_lazy_files.py) ask for pageable memory
      // -- honoured only when there is no CUDA device at all: with a GPU every buffer is pinned
      static const bool test_host_alloc = [] {
        if (getenv("LAKESOUL_B200_TEST_HOST_ALLOC") == nullptr) return false;
        int n = 0;
        return cudaGetDeviceCount(&n) != cudaSuccess || n == 0;
      }();
      if (test_host_alloc) {
        p = static_cast<uint8_t*>(aligned_alloc(4096, cls));
        if (!p) fail("out of memory");
        return std::shared_ptr<uint8_t>(p, [](uint8_t* q) { free(q); });
      }
      cudaError_t e = cudaMallocHost(reinterpret_cast<void**>(&p), cls);
      if (e != cudaSuccess) fail(std::string("cudaMallocHost failed: ") + cudaGetErrorString(e));
    }
    return std::shared_ptr<uint8_t>(p, [this, cls](uint8_t* q) { release(q, cls); });
  }

 private:
  void release(uint8_t* p, size_t cls) {
    {
      std::lock_guard<std::mutex> g(mu_);
      if (cached_ + cls <= kMaxCached) {
        free_[cls].push_back(p);
        cached_ += cls;
        return;
      }
    }
    cudaFreeHost(p);
  }
  static constexpr size_t kMaxCached = 8ull << 30;
  std::mutex mu_;
  std::map<size_t, std::vector<uint8_t*>> free_;
  size_t cached_ = 0;
};

}  // namespace

std::shared_ptr<uint8_t> pinned_alloc(size_t bytes) { return PinnedPool::get().alloc(bytes); }

// ---------------------------------------------------------------------------------------
// scan slots: a stream + a device scratch arena, checked out by one reader for the duration of
// its scan and recycled afterwards (readers come and go per bucket; their threads usually persist)
// ---------------------------------------------------------------------------------------
namespace {
struct ScanSlot {
  cudaStream_t stream = nullptr;
  DeviceArena arena;
  HostStage stage;  // pinned staging of the descriptor uploads (see mor_engine.h)
  int device = -1;
  // two side streams + events: the three ZSTD launches of a scan (large pages, two-block pages, one-block pages) are
  // independent and each is a ~10 ms chain of serial work -- side by side they cost the longest instead of the sum
  cudaStream_t side[2] = {nullptr, nullptr};
  cudaEvent_t fork = nullptr, join[2] = {nullptr, nullptr};
};
class SlotPool {
 public:
  static SlotPool& get() {
    static SlotPool* p = new SlotPool();  // leaked on purpose (no CUDA calls at process teardown)
    return *p;
  }
  ScanSlot* acquire() {
    int dev = 0;
    cudaGetDevice(&dev);
    {
      std::lock_guard<std::mutex> g(mu_);
      for (size_t i = 0; i < free_.size(); ++i) {
        if (free_[i]->device == dev) {
          ScanSlot* s = free_[i];
          free_.erase(free_.begin() + (long)i);
          return s;
        }
      }
    }
    auto* s = new ScanSlot();
    s->device = dev;
    cudaError_t e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
      delete s;
      fail(std::string("cudaStreamCreate failed: ") + cudaGetErrorString(e));
    }
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
      e = cudaStreamCreateWithFlags(&s->side[i], cudaStreamNonBlocking);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->join[i], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->fork, cudaEventDisableTiming);
    if (e != cudaSuccess) fail(std::string("side stream creation failed: ") + cudaGetErrorString(e));
    return s;
  }
  void release(ScanSlot* s) {
    s->arena.reset();
    s->stage.reset();
    std::lock_guard<std::mutex> g(mu_);
    free_.push_back(s);
  }
 private:
  std::mutex mu_;
  std::vector<ScanSlot*> free_;
};
}  // namespace

class ScanSession {
 public:
  ScanSlot* slot = nullptr;
  int64_t h2d_bytes = 0;
  ~ScanSession() {
    if (!slot) return;
    stream_wait(slot->stream);
    if (arena_current() == &slot->arena) arena_install(nullptr);
    if (host_stage_current() == &slot->stage) host_stage_install(nullptr);
    SlotPool::get().release(slot);
  }
};
std::shared_ptr<ScanSession> begin_scan_session() {
  auto s = std::make_shared<ScanSession>();
  s->slot = SlotPool::get().acquire();
  arena_install(&s->slot->arena);
  host_stage_install(&s->slot->stage);
  return s;
}
uint8_t* session_upload(ScanSession& s, const uint8_t* pinned, size_t bytes) {
  uint8_t* d = static_cast<uint8_t*>(s.slot->arena.alloc(bytes + 64));
  if (bytes > 0) {
    cudaError_t e = cudaMemcpyAsync(d, pinned, bytes, cudaMemcpyHostToDevice, s.slot->stream);
    if (e != cudaSuccess) fail(std::string("cudaMemcpyAsync failed: ") + cudaGetErrorString(e));
    s.h2d_bytes += (int64_t)bytes;
  }
  return d;
}
int64_t session_h2d_bytes(const ScanSession& s) { return s.h2d_bytes; }

namespace {
bool trace_on() {
  static const bool on = [] { const char* v = getenv("LAKESOUL_B200_TRACE"); return v && *v == '1'; }();
  return on;
}
struct TraceBuf { std::vector<std::pair<const char*, double>> marks; };
thread_local TraceBuf t_trace;
}  // namespace
void trace_mark(const char* phase) {
  if (trace_on()) t_trace.marks.emplace_back(phase, now_ms());
}
void trace_flush(const char* tag) {
  if (!trace_on() || t_trace.marks.empty()) return;
  std::string line = std::string("LSBTRACE ") + tag;
  char buf[96];
  snprintf(buf, sizeof(buf), " t0=%.3f", t_trace.marks.front().second);
  line += buf;
  for (size_t i = 1; i < t_trace.marks.size(); ++i) {
    snprintf(buf, sizeof(buf), " %s=%.3f", t_trace.marks[i].first, t_trace.marks[i].second - t_trace.marks[i - 1].second);
    line += buf;
  }
  line += "\n";
  fputs(line.c_str(), stderr);
  t_trace.marks.clear();
}

void launch_dict_index(const DecodeArgs& A, cudaStream_t s);
void launch_page_index(const DecodeArgs& A, cudaStream_t s);
void launch_block_count(const DecodeArgs& A, cudaStream_t s);
void launch_block_scan(const DecodeArgs& A, const int32_t* seg_bounds, const int64_t* seg_base, int nseg,
                       int64_t* seg_total, cudaStream_t s);
void launch_block_decode(const DecodeArgs& A, cudaStream_t s);
void launch_zstd_pages(const PqZstdPage* pages, int npages, const ZstdBlockDesc* blocks, ZstdBlockPlan* plans, ZstdTables* tables,
                       uint8_t* lit, uint64_t* seq, int32_t* err, cudaStream_t s);
void launch_snappy_pages(const PqZstdPage* pages, int npages, int32_t* err, cudaStream_t s);
void launch_zstd_batch(const PqZstdPage* pages, const int32_t* order_one, int none, const ZstdBlockDesc* blocks, ZstdBlockPlan* plans,
                       uint8_t* lit, uint64_t* seq, int32_t* err, cudaStream_t s);
void launch_zstd_pages_v2(const PqZstdPage* pages, const int32_t* order_small, int nsmall, const int32_t* order_big, int nbig,
                          const ZstdBlockDesc* blocks, ZstdBlockPlan* plans, ZstdTables* tables_v1, uint8_t* lit, uint64_t* seq, int32_t* err,
                          cudaStream_t s);

namespace {

class KeepAliveBuffer : public arrow::Buffer {
 public:
  KeepAliveBuffer(const uint8_t* p, int64_t n, std::shared_ptr<void> owner)
      : arrow::Buffer(p, n), owner_(std::move(owner)) {}

 private:
  std::shared_ptr<void> owner_;
};

struct DevMem {
  void* p = nullptr;
  size_t bytes = 0;
};

}  // namespace

ColumnKind classify_type(const arrow::DataType& t) {
  ColumnKind k;
  using T = arrow::Type;
  auto fixed = [&](int w, bool pk, uint8_t kind) {
    k.width = w; k.ok = true; k.pk_ok = pk; k.pk_kind = kind;
  };
  switch (t.id()) {
    case T::BOOL: k.width = 0; k.ok = true; break;
    case T::INT8: fixed(1, true, PK_SIGNED); break;
    case T::UINT8: fixed(1, true, PK_UNSIGNED); break;
    case T::INT16: fixed(2, true, PK_SIGNED); break;
    case T::UINT16: fixed(2, true, PK_UNSIGNED); break;
    case T::HALF_FLOAT: fixed(2, false, 0); break;
    case T::INT32: case T::DATE32: case T::TIME32: fixed(4, true, PK_SIGNED); break;
    case T::UINT32: fixed(4, true, PK_UNSIGNED); break;
    case T::FLOAT: fixed(4, true, PK_FLOAT); break;
    case T::INT64: case T::DATE64: case T::TIME64: case T::TIMESTAMP: case T::DURATION:
      fixed(8, true, PK_SIGNED); break;
    case T::UINT64: fixed(8, true, PK_UNSIGNED); break;
    case T::DOUBLE: fixed(8, true, PK_FLOAT); break;
    case T::INTERVAL_MONTHS: fixed(4, false, 0); break;
    case T::INTERVAL_DAY_TIME: fixed(8, false, 0); break;
    case T::INTERVAL_MONTH_DAY_NANO: fixed(16, false, 0); break;
    case T::DECIMAL128: fixed(16, false, 0); break;
    case T::DECIMAL256: fixed(32, false, 0); break;
    case T::FIXED_SIZE_BINARY: {
      const int w = static_cast<const arrow::FixedSizeBinaryType&>(t).byte_width();
      if (w == 1 || w == 2 || w == 4 || w == 8 || w == 16 || w == 32) fixed(w, false, 0);
      break;
    }
    case T::STRING: case T::BINARY: k.var = true; k.ok = true; break;
    default: break;
  }
  return k;
}

static int num_class_of(const arrow::DataType& t) {
  switch (t.id()) {
    case arrow::Type::INT8: return NC_I8;
    case arrow::Type::INT16: return NC_I16;
    case arrow::Type::INT32: return NC_I32;
    case arrow::Type::INT64: return NC_I64;
    case arrow::Type::UINT8: return NC_U8;
    case arrow::Type::UINT16: return NC_U16;
    case arrow::Type::UINT32: return NC_U32;
    case arrow::Type::UINT64: return NC_U64;
    case arrow::Type::FLOAT: return NC_F32;
    case arrow::Type::DOUBLE: return NC_F64;
    default: return NC_NONE;
  }
}

MergeOp merge_op_from_name(const std::string& name) {
  if (name == "UseLast") return MergeOp::UseLast;
  if (name == "UseLastNotNull") return MergeOp::UseLastNotNull;
  if (name == "SumAll") return MergeOp::SumAll;
  if (name == "SumLast") return MergeOp::SumLast;
  if (name == "JoinedLastByComma") return MergeOp::JoinedLastByComma;
  if (name == "JoinedLastBySemicolon") return MergeOp::JoinedLastBySemicolon;
  if (name == "JoinedAllByComma") return MergeOp::JoinedAllByComma;
  if (name == "JoinedAllBySemicolon") return MergeOp::JoinedAllBySemicolon;
  fail("Invalid MergeOperator name");
}

// ---------------------------------------------------------------------------------------
struct DeviceScan::Impl : std::enable_shared_from_this<DeviceScan::Impl> {
  ScanSpec spec;
  std::vector<BucketFiles> buckets;
  cudaStream_t stream = nullptr;
  ScanCounters ctr;

  // input layout
  int nb = 0, F = 0, C = 0;
  std::vector<int32_t> bucket_file_begin;
  std::vector<int64_t> file_rows, file_row_base;   // rows of every file in the column buffers (what was framed), their first row
  std::vector<int64_t> live_rows, live_row_base;   // the rows the merge reads: all of them, or the page-pruned range
  std::vector<const FileInput*> file_in;
  int64_t span = 0;

  struct DevColumn {
    ColumnKind kind;
    DevMem values, validity, data;
    std::vector<char> present;         // per file
    std::vector<int64_t> bucket_data;  // var: data base of each bucket
    DevMem d_bucket_data;
    int64_t data_bytes = 0;
    bool any_nulls = false;
    bool all_present = true;
    bool has_raw = false;
  };
  std::vector<DevColumn> cols;

  MergePlan* plan = nullptr;
  GeneralMerge* general = nullptr;

  // outputs (pinned host)
  struct OutColumn {
    ColumnKind kind;
    bool has_valid = false;
    std::shared_ptr<uint8_t> h_values, h_valid, h_offsets, h_data;
    std::shared_ptr<int64_t> h_bstart;  // per-batch start of the string bytes (+ closing entry)
  };
  std::vector<OutColumn> outs;
  std::vector<int64_t> out_base, batch_base;
  int64_t vstride = 0;
  bool ran = false;
  std::shared_ptr<uint8_t> h_sorted_buckets;  // sort mode with buckets: uint32 per output row (pinned)

  std::shared_ptr<ScanSession> session;
  ScanSlot* slot = nullptr;

  ~Impl() { end_device(); }

  // every device buffer of the scan lives in the slot's arena (no per-buffer driver calls)
  void* dalloc(size_t bytes) { return slot->arena.alloc(std::max<size_t>(bytes, 16)); }
  void begin_device() {
    if (!session) session = begin_scan_session();
    slot = session->slot;
    stream = slot->stream;
    ctr.h2d_bytes += session->h2d_bytes;
  }
  // must run on the thread that opened the session, with all results already on the host
  void end_device() {
    if (general) { general_destroy(general); general = nullptr; }
    if (plan) { plan_destroy(plan); plan = nullptr; }
    session.reset();  // waits for the stream, uninstalls the arena, recycles the slot
    slot = nullptr;
    stream = nullptr;
  }
  void h2d(void* dst, const void* src, size_t n) {
    if (n == 0) return;
    SCAN_CUDA(h2d_staged(dst, src, n, stream));  // descriptors and host-decoded columns: never from pageable memory
    ctr.h2d_bytes += (int64_t)n;
  }
  void d2h(void* dst, const void* src, size_t n) {
    if (n == 0) return;
    SCAN_CUDA(cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToHost, stream));
    ctr.d2h_bytes += (int64_t)n;
  }

  std::vector<int32_t> scratch_order;  // page order of the ZSTD launches (outlives the asynchronous upload)
  std::shared_ptr<uint8_t> h_small;  // pinned: [decode error flag | string segment totals]
  bool decode_err_pending = false;
  void check_decode_error();
  void layout();
  void upload();
  void upload_host_file(int c, int b, int f, int64_t dpos);
  void merge_and_gather();
};

void DeviceScan::Impl::layout() {
  nb = (int)buckets.size();
  C = spec.physical_schema->num_fields();
  bucket_file_begin.assign(nb + 1, 0);
  for (int b = 0; b < nb; ++b)
    bucket_file_begin[b + 1] = bucket_file_begin[b] + (int)buckets[b].files.size();
  F = bucket_file_begin[nb];
  file_rows.resize(F);
  file_row_base.resize(F);
  live_rows.resize(F);
  live_row_base.resize(F);
  file_in.resize(F);
  int64_t pos = 0;
  for (int b = 0, f = 0; b < nb; ++b) {
    for (auto& fi : buckets[b].files) {
      file_in[f] = &fi;
      file_rows[f] = fi.raw ? fi.raw->layout_rows() : fi.table->num_rows();
      // 64-row alignment keeps validity bitmaps byte-copyable and 16-byte alignment for
      // every fixed width; the spare row after each file holds the closing string offset.
      file_row_base[f] = pos;
      // page pruning by a pushed key predicate (RawParquetFile::PrunePages): only the pages that overlap the rows that
      // can match were framed; the merge reads that row range of the file, the boundary pages' other rows are decoded
      // but never referenced
      live_rows[f] = fi.raw ? fi.raw->live_rows() : file_rows[f];
      live_row_base[f] = pos + (fi.raw ? fi.raw->live_offset() : 0);
      pos = (pos + file_rows[f] + 1 + 63) & ~63ll;
      ctr.rows_in += live_rows[f];
      if (!fi.raw) ctr.host_decoded_files += 1;
      ++f;
    }
  }
  span = pos;
}

// Host-decoded file -> column buffers (used for files whose pages the device cannot decode).
void DeviceScan::Impl::upload_host_file(int c, int b, int f, int64_t dpos) {
  const auto& field = spec.physical_schema->field(c);
  DevColumn& dc = cols[c];
  auto col = file_in[f]->table->GetColumnByName(field->name());
  const int w = dc.kind.width;
  const int64_t base = file_row_base[f];
  const int64_t n = file_rows[f];
  std::vector<uint8_t> bits;
  std::vector<int32_t> offs;
  if (dc.validity.p) {  // zero-initialised bitmap: valid rows must be set explicitly
    bits.assign((size_t)((n + 7) / 8), 0xff);
    int64_t r0 = 0;
    for (const auto& ch : col->chunks()) {
      if (ch->null_count() > 0 && ch->null_bitmap_data())
        arrow::internal::CopyBitmap(ch->null_bitmap_data(), ch->offset(), ch->length(), bits.data(), r0);
      r0 += ch->length();
    }
    if (n & 7) bits.back() &= (uint8_t)((1u << (n & 7)) - 1);
    h2d(static_cast<uint8_t*>(dc.validity.p) + base / 8, bits.data(), bits.size());
  }
  int64_t r0 = 0;
  if (w == 0 && !dc.kind.var) {
    bits.assign((size_t)((n + 7) / 8), 0);
    for (const auto& ch : col->chunks()) {
      const auto& ba = static_cast<const arrow::BooleanArray&>(*ch);
      arrow::internal::CopyBitmap(ba.values()->data(), ba.offset(), ba.length(), bits.data(), r0);
      r0 += ch->length();
    }
    h2d(static_cast<uint8_t*>(dc.values.p) + base / 8, bits.data(), bits.size());
    return;
  }
  for (const auto& ch : col->chunks()) {
    const int64_t len = ch->length();
    if (len == 0) continue;
    if (dc.kind.var) {
      const auto& ba = static_cast<const arrow::BinaryArray&>(*ch);
      const int32_t* o = ba.raw_value_offsets();  // already offset-adjusted
      const int64_t first = o[0];
      const int64_t nbytes = (int64_t)o[len] - first;
      const int64_t delta = dpos - dc.bucket_data[b] - first;
      offs.resize((size_t)len + 1);
      for (int64_t i = 0; i <= len; ++i) offs[(size_t)i] = (int32_t)(o[i] + delta);
      h2d(static_cast<int32_t*>(dc.values.p) + base + r0, offs.data(), sizeof(int32_t) * (size_t)(len + 1));
      if (nbytes > 0)
        h2d(static_cast<uint8_t*>(dc.data.p) + dpos, ba.value_data()->data() + first, (size_t)nbytes);
      dpos += nbytes;
    } else {
      const uint8_t* src = ch->data()->buffers[1]->data() + (int64_t)ch->offset() * w;
      h2d(static_cast<uint8_t*>(dc.values.p) + (base + r0) * w, src, (size_t)len * w);
    }
    r0 += len;
  }
  if (dc.kind.var && n == 0) {
    const int32_t z = (int32_t)(dpos - dc.bucket_data[b]);
    h2d(static_cast<int32_t*>(dc.values.p) + base, &z, sizeof(int32_t));
  }
}

void DeviceScan::Impl::upload() {
  cols.resize(C);
  // ---- pass 1 (host): presence, nullability, host-decoded string sizes ---------------------
  std::vector<std::vector<int64_t>> host_bytes(C, std::vector<int64_t>(nb, 0));  // var: per bucket
  std::vector<std::vector<int64_t>> file_host_bytes(C);
  bool any_raw = false;
  for (int c = 0; c < C; ++c) {
    const auto& field = spec.physical_schema->field(c);
    DevColumn& dc = cols[c];
    dc.kind = classify_type(*field->type());
    if (!dc.kind.ok) fail("unsupported column type for the GPU reader: " + field->ToString());
    dc.present.assign(F, 0);
    file_host_bytes[c].assign(F, 0);
    for (int b = 0; b < nb; ++b) {
      for (int f = bucket_file_begin[b]; f < bucket_file_begin[b + 1]; ++f) {
        const FileInput& fi = *file_in[f];
        if (fi.raw) {
          any_raw = true;
          const RawColumn& rc = fi.raw->columns()[c];
          if (!rc.present) { dc.all_present = false; continue; }
          dc.present[f] = 1;
          dc.has_raw = true;
          if (rc.may_have_nulls) dc.any_nulls = true;
          continue;
        }
        auto col = fi.table->GetColumnByName(field->name());
        if (!col) { dc.all_present = false; continue; }
        if (!col->type()->Equals(*field->type()))
          fail("column '" + field->name() + "' has type " + col->type()->ToString() + " in a file but " +
               field->type()->ToString() + " is expected");
        dc.present[f] = 1;
        if (col->null_count() > 0) dc.any_nulls = true;
        if (dc.kind.var) {
          int64_t bytes = 0;
          for (const auto& ch : col->chunks()) {
            const auto& ba = static_cast<const arrow::BinaryArray&>(*ch);
            bytes += ba.value_offset(ba.length()) - ba.value_offset(0);
          }
          file_host_bytes[c][f] = bytes;
          host_bytes[c][b] += bytes;
        }
      }
    }
    // rows of a file that lacks the column must come out null on every path (the concat path and the
    // full merge never look at present[]): the zero-initialised validity bitmap marks them, and a null
    // row's values / offsets are never dereferenced (DefaultColumnStream null-fills, default_column.rs:62-108)
    if (!dc.all_present) dc.any_nulls = true;
  }

  // ---- raw files: image upload, page descriptors, string sizing -----------------------------
  std::vector<PqChunk> chunks;
  std::vector<PqPage> pages;
  std::vector<int32_t> seg_start;
  std::vector<int64_t> seg_base;
  std::vector<int> seg_col, seg_bucket;
  std::vector<int64_t> seg_total;
  PqChunk* d_chunks = nullptr;
  PqPage* d_pages = nullptr;
  int32_t* d_dict_pos = nullptr;
  int32_t* d_err = nullptr;
  DecodeArgs args{};
  if (any_raw) {
    // device copies of the file images (already sent by the reader, or sent now) and of the
    // decompressed pages; page offsets below are absolute device addresses (args.raw = nullptr)
    std::vector<int64_t> img_off(F, 0), aux_off(F, 0);
    for (int f = 0; f < F; ++f) {
      RawParquetFile* rf = file_in[f]->raw.get();
      if (!rf) continue;
      uint8_t* d_img = rf->device_image();
      if (!d_img) {
        d_img = static_cast<uint8_t*>(dalloc((size_t)rf->image_bytes() + 64));
        h2d(d_img, rf->image(), (size_t)rf->image_bytes());
      }
      img_off[f] = (int64_t)reinterpret_cast<uintptr_t>(d_img);
      if (rf->aux_bytes() > 0) {
        uint8_t* d_aux = static_cast<uint8_t*>(dalloc((size_t)rf->aux_bytes() + 64));
        h2d(d_aux, rf->aux(), (size_t)rf->aux_bytes());
        aux_off[f] = (int64_t)reinterpret_cast<uintptr_t>(d_aux);
      }
    }
    int64_t dict_entries = 0, nblocks = 0, vpos_entries = 0, zstd_bytes = 0;
    bool any_dict_var = false;
    std::vector<ZstdBlockDesc> zblocks;
    int64_t zlit_bytes = 0, zseq_recs = 0;
    std::vector<PqZstdPage> zpages, spages;     // ZSTD / SNAPPY payloads expanded on the device
    struct Fixup { bool snappy; size_t z; bool is_dict; size_t target; };  // whose address becomes the expansion's
    std::vector<Fixup> fixups;
    // registers one compressed payload of file image `fi`; the expansion lands at a 64-byte aligned slot of one buffer
    auto add_expansion = [&](const FileInput& fi, int f, uint8_t codec, int64_t src_off, int32_t comp_size, int32_t raw_size,
                             bool is_dict, size_t target, int c) {
      PqZstdPage zp{};
      zp.src_off = img_off[f] + src_off;
      zp.comp_size = comp_size;
      zp.raw_size = raw_size;
      zp.dst_off = zstd_bytes;  // relative for now; the buffer is allocated once the total is known
      zstd_bytes = (zstd_bytes + raw_size + 64 + 63) & ~63ll;
      if (codec == kCodecZstd) {
        // block headers of the frame (three bytes each), read from the pinned image
        zblocks.resize(zblocks.size() + 4096);
        int64_t frame_end = 0;
        const size_t b0 = zblocks.size() - 4096;
        const int64_t nblk = zstd_scan_frame(fi.raw->image() + src_off, comp_size, zblocks.data() + b0, 4096, &frame_end);
        if (nblk < 0) fail("malformed ZSTD page in " + spec.physical_schema->field(c)->name());
        zblocks.resize(b0 + (size_t)nblk);
        zp.first_block = (int32_t)b0;
        zp.nblocks = (int32_t)nblk;
        zp.lit_off = zlit_bytes;
        zp.seq_off = zseq_recs;
        zlit_bytes += ((int64_t)raw_size + 63) & ~63ll;
        zseq_recs += raw_size / 3 + 8;
        fixups.push_back({false, zpages.size(), is_dict, target});
        zpages.push_back(zp);
      } else {
        fixups.push_back({true, spages.size(), is_dict, target});
        spages.push_back(zp);
      }
    };
    for (int c = 0; c < C; ++c) {
      if (!cols[c].has_raw) continue;
      for (int b = 0; b < nb; ++b) {
        const size_t seg_first_page = pages.size();
        for (int f = bucket_file_begin[b]; f < bucket_file_begin[b + 1]; ++f) {
          const FileInput& fi = *file_in[f];
          if (!fi.raw) continue;
          const RawColumn& rc = fi.raw->columns()[c];
          if (!rc.present) continue;
          // one chunk record per dictionary (plus one for the dictionary-less pages)
          std::vector<int> chunk_of_dict(rc.dicts.size() + 1, -1);
          for (size_t pi = 0; pi < rc.pages.size(); ++pi) {
            const RawColumn::Page& rp = rc.pages[pi];
            const int di = rc.page_dict[pi];
            int& ci = chunk_of_dict[(size_t)(di + 1)];
            if (ci < 0) {
              PqChunk pc = rc.proto;
              pc.col = c;
              pc.bucket = b;
              pc.dict_nvals = 0;
              pc.dict_size = 0;
              pc.dict_off = 0;
              pc.dict_index = 0;
              if (di >= 0) {
                const RawColumn::Dict& d = rc.dicts[(size_t)di];
                pc.dict_nvals = d.num_values;
                pc.dict_size = d.size;
                pc.dict_off = (d.aux ? aux_off[f] : img_off[f]) + d.src_off;
                if (pc.phys == PQ_BYTE_ARRAY) {
                  pc.dict_index = dict_entries;
                  dict_entries += d.num_values + 1;
                }
                if (d.codec) add_expansion(fi, f, d.codec, d.src_off, d.comp_size, d.size, true, chunks.size(), c);
              }
              chunks.push_back(pc);
              ci = (int)chunks.size() - 1;
            }
            PqPage pg{};
            pg.src_off = (rp.aux ? aux_off[f] : img_off[f]) + rp.src_off;
            pg.size = rp.size;
            if (rp.codec)  // expanded on the device: the decoders read the expansion, not the file image
              add_expansion(fi, f, rp.codec, rp.src_off, rp.comp_size, rp.size, false, pages.size(), c);
            pg.num_values = rp.num_values;
            pg.encoding = rp.encoding;
            pg.def_encoding = rp.def_encoding;
            pg.v2_def_len = rp.v2_def_len;
            pg.chunk = ci;
            pg.closes_file = (pi + 1 == rc.pages.size()) ? 1 : 0;
            pg.row_start = file_row_base[f] + rp.first_row;
            pg.first_block = (int32_t)nblocks;
            nblocks += (rp.num_values + kDecodeBlockRows - 1) / kDecodeBlockRows;
            if (rc.proto.conv == CONV_VAR) {
              if (rp.encoding == ENC_PLAIN) {
                pg.vpos_base = vpos_entries;
                vpos_entries += (int64_t)rp.num_values + 1;
              } else {
                any_dict_var = true;
              }
            }
            pages.push_back(pg);
          }
          if (rc.pages.empty() && cols[c].kind.var) {
            // empty file: its single offset entry is written by the host below
          }
        }
        if (cols[c].kind.var && pages.size() > seg_first_page) {
          seg_start.push_back((int32_t)seg_first_page);
          seg_base.push_back(host_bytes[c][b]);
          seg_col.push_back(c);
          seg_bucket.push_back(b);
        }
      }
    }
    // explicit [start, end) page range of every (column, bucket) segment
    std::vector<int32_t> seg_bounds;
    for (size_t s = 0; s < seg_start.size(); ++s) {
      seg_bounds.push_back(seg_start[s]);
      // end = next page whose (col,bucket) differs
      int32_t e = seg_start[s];
      while (e < (int32_t)pages.size() && chunks[(size_t)pages[(size_t)e].chunk].col == seg_col[s] &&
             chunks[(size_t)pages[(size_t)e].chunk].bucket == seg_bucket[s])
        ++e;
      seg_bounds.push_back(e);
    }
    ctr.device_decoded_pages += (int64_t)pages.size();
    const int nseg = (int)seg_start.size();
    d_chunks = static_cast<PqChunk*>(dalloc(sizeof(PqChunk) * chunks.size()));
    d_pages = static_cast<PqPage*>(dalloc(sizeof(PqPage) * pages.size()));
    d_dict_pos = static_cast<int32_t*>(dalloc(sizeof(int32_t) * (size_t)(dict_entries + 1)));
    d_err = static_cast<int32_t*>(dalloc(sizeof(int32_t)));
    PqBlock* d_blocks = static_cast<PqBlock*>(dalloc(sizeof(PqBlock) * (size_t)(nblocks + 1)));
    int32_t* d_vpos = static_cast<int32_t*>(dalloc(sizeof(int32_t) * (size_t)(vpos_entries + 1)));
    SCAN_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int32_t), stream));
    if (!zpages.empty() || !spages.empty()) {
      // compressed payloads -> plain payloads, before anything reads the pages / dictionaries
      uint8_t* d_zout = static_cast<uint8_t*>(dalloc((size_t)zstd_bytes + 64));
      for (const Fixup& fx : fixups) {
        PqZstdPage& zp = fx.snappy ? spages[fx.z] : zpages[fx.z];
        zp.dst_off += (int64_t)reinterpret_cast<uintptr_t>(d_zout);
        if (fx.is_dict) chunks[fx.target].dict_off = zp.dst_off;
        else pages[fx.target].src_off = zp.dst_off;
      }
    }
    if (!spages.empty()) {
      PqZstdPage* d_spages = static_cast<PqZstdPage*>(dalloc(sizeof(PqZstdPage) * spages.size()));
      h2d(d_spages, spages.data(), sizeof(PqZstdPage) * spages.size());
      launch_snappy_pages(d_spages, (int)spages.size(), d_err, stream);
      ctr.kernel_launches += 1;
      ctr.device_decompressed_pages += (int64_t)spages.size();
    }
    if (!zpages.empty()) {
      PqZstdPage* d_zpages = static_cast<PqZstdPage*>(dalloc(sizeof(PqZstdPage) * zpages.size()));
      ZstdBlockDesc* d_zblocks = static_cast<ZstdBlockDesc*>(dalloc(sizeof(ZstdBlockDesc) * zblocks.size()));
      ZstdBlockPlan* d_zplans = static_cast<ZstdBlockPlan*>(dalloc(sizeof(ZstdBlockPlan) * zblocks.size()));
      uint8_t* d_zlit = static_cast<uint8_t*>(dalloc((size_t)zlit_bytes + 64));
      uint64_t* d_zseq = static_cast<uint64_t*>(dalloc(sizeof(uint64_t) * (size_t)(zseq_recs + 8)));
      h2d(d_zpages, zpages.data(), sizeof(PqZstdPage) * zpages.size());
      h2d(d_zblocks, zblocks.data(), sizeof(ZstdBlockDesc) * zblocks.size());
      static const bool zstd_v1 = [] { const char* v = getenv("LAKESOUL_B200_ZSTD_V1"); return v && *v == '1'; }();
      if (zstd_v1) {
        ZstdTables* d_ztables = static_cast<ZstdTables*>(dalloc(sizeof(ZstdTables) * zblocks.size()));
        launch_zstd_pages(d_zpages, (int)zpages.size(), d_zblocks, d_zplans, d_ztables, d_zlit, d_zseq, d_err, stream);
        ctr.kernel_launches += 1;
      } else {
        // one-block pages are batched eight to a thread block (zstd_batch_kernel), pages of two blocks run on two
        // warps, larger ones on eight (one warp per 128 KiB block)
        static const bool no_batch = [] { const char* v = getenv("LAKESOUL_B200_ZSTD_NO_BATCH"); return v && *v == '1'; }();
        std::vector<int32_t> order_one, order_small, order_big;
        for (size_t z = 0; z < zpages.size(); ++z) {
          const int nbk = zpages[z].nblocks;
          (nbk == 1 && !no_batch ? order_one : (nbk <= 2 ? order_small : order_big)).push_back((int32_t)z);
        }
        // largest pages first inside a batch class: pages of similar size share a thread block
        std::sort(order_one.begin(), order_one.end(), [&](int32_t a, int32_t b) { return zpages[(size_t)a].raw_size > zpages[(size_t)b].raw_size; });
        std::vector<int32_t>& order = scratch_order;
        order = order_big;
        order.insert(order.end(), order_small.begin(), order_small.end());
        order.insert(order.end(), order_one.begin(), order_one.end());
        int32_t* d_order = static_cast<int32_t*>(dalloc(sizeof(int32_t) * order.size()));
        h2d(d_order, order.data(), sizeof(int32_t) * order.size());
        static const bool small_v1 = [] { const char* v = getenv("LAKESOUL_B200_ZSTD_SMALL_V1"); return v && *v == '1'; }();
        ZstdTables* d_ztables = small_v1 ? static_cast<ZstdTables*>(dalloc(sizeof(ZstdTables) * zblocks.size())) : nullptr;
        static const bool side_streams = [] { const char* v = getenv("LAKESOUL_B200_ZSTD_STREAMS"); return !v || *v != '1'; }();
        const int classes = (order_big.empty() ? 0 : 1) + (order_small.empty() ? 0 : 1) + (order_one.empty() ? 0 : 1);
        if (side_streams && classes > 1 && !small_v1) {
          SCAN_CUDA(cudaEventRecord(slot->fork, stream));
          int used = 0;
          cudaStream_t where[3] = {stream, stream, stream};  // big, small, one
          bool first = true;
          const bool have[3] = {!order_big.empty(), !order_small.empty(), !order_one.empty()};
          for (int c = 0; c < 3; ++c) {
            if (!have[c]) continue;
            if (first) { first = false; continue; }  // the first class stays on the scan's stream
            where[c] = slot->side[used];
            SCAN_CUDA(cudaStreamWaitEvent(slot->side[used], slot->fork, 0));
            ++used;
          }
          if (have[0])
            launch_zstd_pages_v2(d_zpages, nullptr, 0, d_order, (int)order_big.size(), d_zblocks, d_zplans, nullptr, d_zlit, d_zseq, d_err, where[0]);
          if (have[1])
            launch_zstd_pages_v2(d_zpages, d_order + order_big.size(), (int)order_small.size(), nullptr, 0, d_zblocks, d_zplans, nullptr, d_zlit,
                                 d_zseq, d_err, where[1]);
          if (have[2])
            launch_zstd_batch(d_zpages, d_order + order_big.size() + order_small.size(), (int)order_one.size(), d_zblocks, d_zplans, d_zlit,
                              d_zseq, d_err, where[2]);
          for (int i = 0; i < used; ++i) {
            SCAN_CUDA(cudaEventRecord(slot->join[i], slot->side[i]));
            SCAN_CUDA(cudaStreamWaitEvent(stream, slot->join[i], 0));
          }
        } else {
          launch_zstd_pages_v2(d_zpages, d_order + order_big.size(), (int)order_small.size(), d_order, (int)order_big.size(), d_zblocks,
                               d_zplans, d_ztables, d_zlit, d_zseq, d_err, stream);
          launch_zstd_batch(d_zpages, d_order + order_big.size() + order_small.size(), (int)order_one.size(), d_zblocks, d_zplans, d_zlit,
                            d_zseq, d_err, stream);
        }
        ctr.kernel_launches += order_one.empty() ? 0 : 1;
        ctr.kernel_launches += (order_small.empty() ? 0 : 1) + (order_big.empty() ? 0 : 1);
      }
      ctr.device_decompressed_pages += (int64_t)zpages.size();
    }
    h2d(d_chunks, chunks.data(), sizeof(PqChunk) * chunks.size());
    h2d(d_pages, pages.data(), sizeof(PqPage) * pages.size());
    args.raw = nullptr;
    args.chunks = d_chunks;
    args.pages = d_pages;
    args.dict_pos = d_dict_pos;
    args.blocks = d_blocks;
    args.vpos = d_vpos;
    args.npages = (int32_t)pages.size();
    args.nchunks = (int32_t)chunks.size();
    args.nblocks = (int32_t)nblocks;
    args.err = d_err;
    args.cols = nullptr;
    seg_total.assign((size_t)nseg, 0);
    if (nseg > 0) {
      int32_t* d_bounds = static_cast<int32_t*>(dalloc(sizeof(int32_t) * seg_bounds.size()));
      int64_t* d_base = static_cast<int64_t*>(dalloc(sizeof(int64_t) * (size_t)nseg));
      int64_t* d_total = static_cast<int64_t*>(dalloc(sizeof(int64_t) * (size_t)nseg));
      h2d(d_bounds, seg_bounds.data(), sizeof(int32_t) * seg_bounds.size());
      h2d(d_base, seg_base.data(), sizeof(int64_t) * (size_t)nseg);
      trace_mark("up_enqueue");
      launch_dict_index(args, stream);
      launch_page_index(args, stream);
      if (any_dict_var) launch_block_count(args, stream);
      launch_block_scan(args, d_bounds, d_base, nseg, d_total, stream);
      ctr.kernel_launches += any_dict_var ? 4 : 3;
      bool host_var = false;
      for (size_t s2 = 0; s2 < seg_base.size(); ++s2) host_var |= seg_base[s2] != 0;
      if (nb == 1 && !any_dict_var && !host_var) {
        // one bucket of PLAIN strings: the page payload bounds the string bytes, so the column buffers
        // can be sized without waiting for the index pass (one GPU round trip less per reader)
        for (int s2 = 0; s2 < nseg; ++s2) {
          int64_t bound = 0;
          for (int32_t i = seg_bounds[2 * (size_t)s2]; i < seg_bounds[2 * (size_t)s2 + 1]; ++i) bound += pages[(size_t)i].size;
          seg_total[(size_t)s2] = bound;
        }
      } else {
        h_small = pinned_alloc(sizeof(int64_t) * (size_t)nseg + 64);
        int64_t* h_total = reinterpret_cast<int64_t*>(h_small.get() + 64);
        SCAN_CUDA(cudaMemcpyAsync(h_total, d_total, sizeof(int64_t) * (size_t)nseg, cudaMemcpyDeviceToHost, stream));
        SCAN_CUDA(stream_wait(stream));
        for (int s2 = 0; s2 < nseg; ++s2) seg_total[(size_t)s2] = h_total[s2];
      }
      trace_mark("wait_index");
    } else {
      launch_dict_index(args, stream);
      launch_page_index(args, stream);
      ctr.kernel_launches += 2;
    }
  }

  // ---- allocate the column buffers, upload host-decoded files ---------------------------------
  std::vector<PqColOut> colout((size_t)C);
  for (int c = 0; c < C; ++c) {
    DevColumn& dc = cols[c];
    const int w = dc.kind.width;
    dc.bucket_data.assign(nb, 0);
    std::vector<int64_t> dev_bytes(nb, 0);
    for (size_t s = 0; s < seg_col.size(); ++s)
      if (seg_col[s] == c) dev_bytes[(size_t)seg_bucket[s]] = seg_total[s];
    int64_t data_pos = 0;
    for (int b = 0; b < nb; ++b) {
      dc.bucket_data[b] = data_pos;
      const int64_t bytes = host_bytes[c][b] + dev_bytes[(size_t)b];
      if (bytes > 0x7fffff00ll)
        fail("variable-width column '" + spec.physical_schema->field(c)->name() + "' exceeds 2 GiB in one hash bucket");
      data_pos += bytes;
    }
    dc.data_bytes = data_pos;
    if (dc.kind.var) {
      dc.values.bytes = (size_t)(span + 1) * 4;
      dc.data.bytes = (size_t)data_pos;
      dc.data.p = dalloc(dc.data.bytes + 16);
      dc.d_bucket_data.p = dalloc(sizeof(int64_t) * nb);
      h2d(dc.d_bucket_data.p, dc.bucket_data.data(), sizeof(int64_t) * nb);
    } else if (w == 0) {
      dc.values.bytes = (size_t)(span / 8 + 8);
    } else {
      dc.values.bytes = (size_t)span * w;
    }
    dc.values.p = dalloc(dc.values.bytes);
    if (w == 0 && !dc.kind.var && dc.has_raw) SCAN_CUDA(cudaMemsetAsync(dc.values.p, 0, dc.values.bytes, stream));
    if (dc.any_nulls) {
      dc.validity.bytes = (size_t)(span / 8 + 8);
      dc.validity.p = dalloc(dc.validity.bytes);
      SCAN_CUDA(cudaMemsetAsync(dc.validity.p, 0, dc.validity.bytes, stream));
    }
    for (int b = 0; b < nb; ++b) {
      int64_t dpos = dc.bucket_data[b];
      for (int f = bucket_file_begin[b]; f < bucket_file_begin[b + 1]; ++f) {
        if (!dc.present[f]) continue;
        if (file_in[f]->raw) {
          if (dc.kind.var && file_rows[f] == 0) {
            const int32_t z = 0;
            h2d(static_cast<int32_t*>(dc.values.p) + file_row_base[f], &z, sizeof(int32_t));
          }
          continue;
        }
        upload_host_file(c, b, f, dpos);
        dpos += file_host_bytes[c][f];
      }
    }
    colout[(size_t)c].values = dc.values.p;
    colout[(size_t)c].validity = static_cast<uint8_t*>(dc.validity.p);
    colout[(size_t)c].data = static_cast<uint8_t*>(dc.data.p);
    colout[(size_t)c].bucket_data = static_cast<const int64_t*>(dc.d_bucket_data.p);
  }
  if (any_raw && !pages.empty()) {
    PqColOut* d_colout = static_cast<PqColOut*>(dalloc(sizeof(PqColOut) * (size_t)C));
    h2d(d_colout, colout.data(), sizeof(PqColOut) * (size_t)C);
    args.cols = d_colout;
    launch_block_decode(args, stream);
    ctr.kernel_launches += 1;
    // the error flag is read back with the merge counts (check_decode_error), not with a sync of its own
    if (!h_small) h_small = pinned_alloc(64);
    *reinterpret_cast<int32_t*>(h_small.get()) = 0;
    SCAN_CUDA(cudaMemcpyAsync(h_small.get(), d_err, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    decode_err_pending = true;
    trace_mark("dec_enqueue");
  }
}

void DeviceScan::Impl::check_decode_error() {
  if (!decode_err_pending) return;
  decode_err_pending = false;
  const int32_t err = *reinterpret_cast<const int32_t*>(h_small.get());
  if (err != 0) fail("device Parquet decode met a malformed or unsupported page (code " + std::to_string(err) + ")");
}

void DeviceScan::Impl::merge_and_gather() {
  const bool do_merge = spec.merge && !spec.primary_keys.empty();
  // ---- which path -----------------------------------------------------------------------
  bool all_use_last = true;
  for (auto op : spec.ops) all_use_last &= (op == MergeOp::UseLast);
  bool partial = false;
  for (int c = 0; c < C; ++c) partial |= !cols[c].all_present;
  const bool general_path = do_merge && (!all_use_last || partial);

  // ---- primary key ------------------------------------------------------------------------
  // Fast paths: up to 4 fixed-width columns packed into 64 / 128 bits, or one Utf8 / Binary column. Every other
  // shape (a string inside a composite key, decimal keys, more than 128 bits or 4 columns) is merged on row-encoded
  // keys (plan_encode_keys), the order the reference gets from arrow-row (stream/mod.rs:104-122).
  PkDesc pk{};
  int key_bits = 64;
  bool encoded_keys = false;
  std::vector<EncKeyCol> enc;
  // one key item = a physical column, or (c < 0) a device-computed uint32 column over global rows (the hash bucket of
  // the write side)
  struct KeyItem { int c; const void* syn; std::string name; };
  auto enc_col_of = [&](const KeyItem& it) -> EncKeyCol {
    EncKeyCol e{};
    if (it.c < 0) {
      e.values = it.syn;
      e.width = 4;
      e.kind = ENC_UNSIGNED;
      return e;
    }
    const DevColumn& dc = cols[it.c];
    const auto& type = *spec.physical_schema->field(it.c)->type();
    e.values = dc.values.p;
    e.validity = static_cast<const uint8_t*>(dc.validity.p);
    e.data = static_cast<const uint8_t*>(dc.data.p);
    e.bucket_data = static_cast<const int64_t*>(dc.d_bucket_data.p);
    e.width = dc.kind.width;
    if (dc.kind.var) e.kind = ENC_BYTES;
    else if (type.id() == arrow::Type::BOOL) e.kind = ENC_BOOL_BITS;
    else if (type.id() == arrow::Type::FLOAT || type.id() == arrow::Type::DOUBLE || type.id() == arrow::Type::HALF_FLOAT) e.kind = ENC_FLOAT;
    else if (arrow::is_unsigned_integer(type.id())) e.kind = ENC_UNSIGNED;
    else if (arrow::is_signed_integer(type.id()) || arrow::is_decimal(type.id()) || arrow::is_temporal(type.id())) e.kind = ENC_SIGNED;
    else
      fail("key column '" + it.name + "' of type " + type.ToString() + " is not supported by the GPU engine");
    return e;
  };
  auto key_items_of = [&](const std::vector<std::string>& names, std::vector<KeyItem>& items) {
    for (const auto& nme : names) {
      const int c = spec.physical_schema->GetFieldIndex(nme);
      if (c < 0) fail("primary key column '" + nme + "' is not in the schema");
      if (!cols[c].all_present) fail("primary key column '" + nme + "' is missing from a file");
      items.push_back(KeyItem{c, nullptr, nme});
    }
  };
  auto build_key = [&](const std::vector<KeyItem>& items) {
    const int nk = (int)items.size();
    int bits = 0;
    bool fixed_ok = nk <= kMaxPkCols;
    for (const auto& it : items) {
      if (it.c < 0) { bits += 32; continue; }
      const DevColumn& dc = cols[it.c];
      if (dc.kind.var || !dc.kind.pk_ok) fixed_ok = false;
      bits += dc.kind.width * 8 + (dc.validity.p ? 1 : 0);
    }
    if (fixed_ok && bits > 128) fixed_ok = false;
    const bool single_bytes = nk == 1 && items[0].c >= 0 && cols[items[0].c].kind.var;
    if (single_bytes) {
      const DevColumn& dc = cols[items[0].c];
      pk.ncols = 1;
      pk.col[0].values = dc.values.p;
      pk.col[0].validity = static_cast<const uint8_t*>(dc.validity.p);
      pk.col[0].width = 0;
      pk.col[0].kind = PK_BYTES;
      pk.col[0].data = static_cast<const uint8_t*>(dc.data.p);
      pk.col[0].bucket_data = static_cast<const int64_t*>(dc.d_bucket_data.p);
      key_bits = 128;
    } else if (fixed_ok) {
      pk.ncols = nk;
      for (int i = 0; i < nk; ++i) {
        if (items[i].c < 0) {
          pk.col[i].values = items[i].syn;
          pk.col[i].validity = nullptr;
          pk.col[i].width = 4;
          pk.col[i].kind = PK_UNSIGNED;
          continue;
        }
        const DevColumn& dc = cols[items[i].c];
        pk.col[i].values = dc.values.p;
        pk.col[i].validity = static_cast<const uint8_t*>(dc.validity.p);
        pk.col[i].width = (uint8_t)dc.kind.width;
        pk.col[i].kind = dc.kind.pk_kind;
      }
      key_bits = bits;
    } else {
      encoded_keys = true;
      key_bits = 128;
      for (const auto& it : items) enc.push_back(enc_col_of(it));
    }
    pk.key_bits = key_bits;
  };
  if (do_merge) {
    std::vector<KeyItem> items;
    key_items_of(spec.primary_keys, items);
    build_key(items);
  }

  const double t0 = now_ms();
  plan = plan_create(nb, bucket_file_begin.data(), live_rows.data(), live_row_base.data(), span,
                     do_merge ? key_bits : 64, spec.batch_size, stream);
  if (do_merge && encoded_keys) {
    // the encoder walks string offsets of the key columns: a malformed page must surface as its own error first
    SCAN_CUDA(stream_wait(stream));
    check_decode_error();
    plan_encode_keys(plan, enc.data(), (int)enc.size(), &pk, stream);
    ctr.kernel_launches += 5;
  }
  if (do_merge) plan_merge(plan, pk, stream);
  else plan_identity(plan, stream);
  plan_sync_counts(plan, stream);
  check_decode_error();  // same stream, enqueued before the merge
  // ---- write side: hash buckets + sort of the row list ----------------------------
  uint32_t* d_row_bucket = nullptr;
  if (spec.sort.enabled) {
    if (do_merge || nb != 1) fail("sort mode takes one group of unmerged inputs");
    std::vector<KeyItem> items;
    key_items_of(spec.sort.prefix_cols, items);
    if (spec.sort.with_buckets) {
      std::vector<KeyItem> hitems;
      key_items_of(spec.sort.hash_cols, hitems);
      std::vector<EncKeyCol> henc;
      for (const auto& it : hitems) henc.push_back(enc_col_of(it));
      d_row_bucket = static_cast<uint32_t*>(dalloc(sizeof(uint32_t) * (size_t)std::max<int64_t>(span, 1)));
      plan_hash_buckets(plan, henc.data(), (int)henc.size(), std::max<uint32_t>(spec.sort.bucket_num, 1), d_row_bucket, stream);
      ctr.kernel_launches += 1;
      items.push_back(KeyItem{-1, d_row_bucket, "<hash bucket>"});
    }
    key_items_of(spec.sort.hash_cols, items);
    key_items_of(spec.sort.suffix_cols, items);
    if (!items.empty()) {
      build_key(items);
      if (encoded_keys) {
        plan_encode_keys(plan, enc.data(), (int)enc.size(), &pk, stream);
        ctr.kernel_launches += 5;
      }
      plan_sort(plan, pk, stream);
    }
  }
  ctr.kernel_launches += plan_stats(plan).launches;

  // per-column winner lists / operator results (partial merge, non-UseLast operators)
  std::vector<GeneralColumnIn> gin;
  if (general_path) {
    ctr.path = "merge-general";
    gin.resize(C);
    for (int c = 0; c < C; ++c) {
      gin[c].op = spec.ops.empty() ? MergeOp::UseLast : spec.ops[c];
      gin[c].present = cols[c].present.data();
      gin[c].kind = cols[c].kind;
      gin[c].num_class = num_class_of(*spec.physical_schema->field(c)->type());
      gin[c].values = cols[c].values.p;
      gin[c].validity = static_cast<const uint8_t*>(cols[c].validity.p);
      gin[c].data = static_cast<const uint8_t*>(cols[c].data.p);
      gin[c].bucket_data = static_cast<const int64_t*>(cols[c].d_bucket_data.p);
      gin[c].data_bytes = cols[c].data_bytes;
    }
    general = general_create(plan, pk, gin, stream);
    ctr.kernel_launches += general_launches(general);
  } else {
    ctr.path = do_merge ? "merge-full" : "concat";
  }

  // ---- post-merge predicate ------------------------------------------------------------
  if (spec.filter) {
    std::vector<FilterColumnIn> fin(C);
    for (int c = 0; c < C; ++c) {
      fin[c].name = spec.physical_schema->field(c)->name();
      fin[c].type = spec.physical_schema->field(c)->type();
      fin[c].kind = cols[c].kind;
      fin[c].values = general ? general_values(general, c, cols[c].values.p) : cols[c].values.p;
      fin[c].validity = general ? general_validity(general, c, static_cast<const uint8_t*>(cols[c].validity.p))
                                : static_cast<const uint8_t*>(cols[c].validity.p);
      fin[c].data = general ? general_data(general, c, static_cast<const uint8_t*>(cols[c].data.p))
                            : static_cast<const uint8_t*>(cols[c].data.p);
      fin[c].bucket_data = general ? general_bucket_data(general, c, static_cast<const int64_t*>(cols[c].d_bucket_data.p))
                                   : static_cast<const int64_t*>(cols[c].d_bucket_data.p);
      fin[c].win = general ? general_winners(general, c) : nullptr;
    }
    std::vector<uint32_t*> extra;
    if (general) extra = general_winner_lists(general);
    ctr.kernel_launches += apply_filter(plan, *spec.filter, fin, extra, stream);
  }
  ctr.merge_ms += now_ms() - t0;
  trace_mark("merge");

  out_base = plan_out_base(plan);
  batch_base = plan_batch_base(plan);
  vstride = plan_valid_stride_bytes(plan);
  const int64_t U = out_base[nb];
  const int64_t TB = batch_base[nb];
  ctr.rows_out = U;

  // ---- gather ---------------------------------------------------------------------------
  // Every fixed-size output (validity, values, string offsets, per-batch data starts) of every column
  // lives in ONE device region mirrored by ONE pinned host block: a single D2H copy brings them back.
  const double t1 = now_ms();
  outs.resize(C);
  struct Seg { size_t valid = 0, main = 0, bstart = 0; size_t valid_n = 0, main_n = 0, bstart_n = 0; };
  std::vector<Seg> seg(C);
  size_t region = 0;
  auto take = [&](size_t n) { const size_t o = region; region = (region + n + 255) & ~(size_t)255; return o; };
  std::vector<const uint32_t*> wins(C, nullptr);
  std::vector<const void*> src_values(C, nullptr);
  std::vector<const uint8_t*> src_valid(C, nullptr);
  for (int c = 0; c < C; ++c) {
    DevColumn& dc = cols[c];
    OutColumn& oc = outs[c];
    oc.kind = dc.kind;
    wins[c] = general ? general_winners(general, c) : nullptr;
    src_values[c] = general ? general_values(general, c, dc.values.p) : dc.values.p;
    src_valid[c] = general ? general_validity(general, c, static_cast<const uint8_t*>(dc.validity.p))
                           : static_cast<const uint8_t*>(dc.validity.p);
    oc.has_valid = src_valid[c] != nullptr || (general && general_may_be_null(general, c));
    if (U == 0) continue;
    Seg& sg = seg[c];
    if (oc.has_valid) { sg.valid_n = (size_t)(TB * vstride); sg.valid = take(sg.valid_n); }
    if (dc.kind.var) {
      sg.main_n = (size_t)(U + TB) * 4;
      sg.main = take(sg.main_n);
      sg.bstart_n = (size_t)(TB + 1) * 8;
      sg.bstart = take(sg.bstart_n);
    } else {
      const int w = dc.kind.width;
      sg.main_n = w == 0 ? (size_t)(TB * vstride) : (size_t)U * w;
      sg.main = take(sg.main_n);
    }
  }
  uint8_t* d_region = U > 0 ? static_cast<uint8_t*>(dalloc(region + 256)) : nullptr;
  std::vector<void*> d_data(C, nullptr);
  std::vector<FixedCol> fixed;
  std::vector<VarCol> vars;
  for (int c = 0; c < C && U > 0; ++c) {
    DevColumn& dc = cols[c];
    const Seg& sg = seg[c];
    uint8_t* d_valid = outs[c].has_valid ? d_region + sg.valid : nullptr;
    if (dc.kind.var) {
      const uint8_t* src_data = general ? general_data(general, c, static_cast<const uint8_t*>(dc.data.p))
                                        : static_cast<const uint8_t*>(dc.data.p);
      const int64_t* bdata = general ? general_bucket_data(general, c, static_cast<const int64_t*>(dc.d_bucket_data.p))
                                     : static_cast<const int64_t*>(dc.d_bucket_data.p);
      const int64_t cap = general ? general_data_bytes(general, c, dc.data_bytes) : dc.data_bytes;
      d_data[c] = dalloc((size_t)cap);
      VarCol vc{};
      vc.src_offsets = static_cast<const int32_t*>(src_values[c]);
      vc.src_data = src_data;
      vc.src_data_base = bdata;
      vc.src_valid = src_valid[c];
      vc.win = wins[c];
      vc.dst_offsets = reinterpret_cast<int32_t*>(d_region + sg.main);
      vc.dst_data = static_cast<uint8_t*>(d_data[c]);
      vc.dst_valid = d_valid;
      vc.batch_data_start = reinterpret_cast<int64_t*>(d_region + sg.bstart);
      // Joined* operators gather from a computed scratch column addressed by output row, not by file row
      vc.direct = (general && general_data(general, c, nullptr) != nullptr) ? 1 : 0;
      vars.push_back(vc);
    } else {
      FixedCol fc{};
      fc.src = src_values[c];
      fc.src_valid = src_valid[c];
      fc.dst = d_region + sg.main;
      fc.dst_valid = d_valid;
      fc.win = wins[c];
      fc.width = dc.kind.width;
      fixed.push_back(fc);
    }
  }
  if (U > 0 && (!fixed.empty() || !vars.empty())) {
    gather_all(plan, fixed.data(), (int)fixed.size(), vars.data(), (int)vars.size(), stream);
    ctr.kernel_launches += (fixed.empty() ? 0 : 1) + (((int)vars.size() + 15) / 16);
  }

  // ---- download ---------------------------------------------------------------------------
  auto& pool = PinnedPool::get();
  std::shared_ptr<uint8_t> h_region;
  if (U > 0) {
    h_region = pool.alloc(region + 256);
    d2h(h_region.get(), d_region, region);
    for (int c = 0; c < C; ++c) {
      OutColumn& oc = outs[c];
      const Seg& sg = seg[c];
      if (oc.has_valid) oc.h_valid = std::shared_ptr<uint8_t>(h_region, h_region.get() + sg.valid);
      if (oc.kind.var) {
        oc.h_offsets = std::shared_ptr<uint8_t>(h_region, h_region.get() + sg.main);
        oc.h_bstart = std::shared_ptr<int64_t>(h_region, reinterpret_cast<int64_t*>(h_region.get() + sg.bstart));
      } else {
        oc.h_values = std::shared_ptr<uint8_t>(h_region, h_region.get() + sg.main);
      }
    }
  }
  if (d_row_bucket && U > 0) {  // bucket of every output row, in output order
    uint32_t* d_sorted = static_cast<uint32_t*>(dalloc(sizeof(uint32_t) * (size_t)U));
    plan_take_u32(plan, d_row_bucket, d_sorted, stream);
    h_sorted_buckets = pool.alloc(sizeof(uint32_t) * (size_t)U);
    d2h(h_sorted_buckets.get(), d_sorted, sizeof(uint32_t) * (size_t)U);
    ctr.kernel_launches += 1;
  }
  // error flag of the gather kernels (a look-back scan that never saw its predecessor publish), same wait
  std::shared_ptr<uint8_t> h_gerr = pinned_alloc(64);
  *reinterpret_cast<int32_t*>(h_gerr.get()) = 0;
  SCAN_CUDA(cudaMemcpyAsync(h_gerr.get(), plan_error_flag_dev(plan), sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
  trace_mark("gather_enqueue");
  SCAN_CUDA(stream_wait(stream));
  trace_mark("wait_gather");
  if (*reinterpret_cast<const int32_t*>(h_gerr.get()) != 0)
    fail("gather kernel reported error code " + std::to_string(*reinterpret_cast<const int32_t*>(h_gerr.get())));
  ctr.gather_ms += now_ms() - t1;
  const double t2 = now_ms();
  // string bytes: exact sizes are known now; one pinned block for all columns
  size_t str_total = 0;
  std::vector<size_t> str_off(C, 0);
  for (int c = 0; c < C && U > 0; ++c) {
    OutColumn& oc = outs[c];
    if (!oc.kind.var) continue;
    const int64_t* bs = oc.h_bstart.get();
    for (int64_t g = 0; g < TB; ++g)
      if (bs[g + 1] - bs[g] > 0x7fffffffll) fail("string data of one output batch exceeds 2 GiB");
    str_off[c] = str_total;
    str_total = (str_total + (size_t)bs[TB] + 255) & ~(size_t)255;
  }
  if (str_total > 0) {
    std::shared_ptr<uint8_t> h_str = pool.alloc(str_total);
    for (int c = 0; c < C; ++c) {
      OutColumn& oc = outs[c];
      if (!oc.kind.var) continue;
      oc.h_data = std::shared_ptr<uint8_t>(h_str, h_str.get() + str_off[c]);
      d2h(oc.h_data.get(), d_data[c], (size_t)oc.h_bstart.get()[TB]);
    }
    SCAN_CUDA(stream_wait(stream));
  } else {
    for (int c = 0; c < C && U > 0; ++c)
      if (outs[c].kind.var) outs[c].h_data = h_region;  // empty strings only: any valid pointer
  }
  trace_mark("wait_strings");
  ctr.download_ms += now_ms() - t2;
}

DeviceScan::DeviceScan(ScanSpec spec, std::vector<BucketFiles> buckets, std::shared_ptr<ScanSession> session)
    : impl_(std::make_shared<Impl>()) {
  impl_->session = std::move(session);
  impl_->spec = std::move(spec);
  impl_->buckets = std::move(buckets);
  if (impl_->spec.ops.empty())
    impl_->spec.ops.assign(impl_->spec.physical_schema->num_fields(), MergeOp::UseLast);
}
DeviceScan::~DeviceScan() = default;

void DeviceScan::Run() {
  Impl& I = *impl_;
  if (I.ran) return;
  I.begin_device();
  try {
    I.layout();
    const double t0 = now_ms();
    I.upload();
    I.ctr.upload_ms += now_ms() - t0;
    I.merge_and_gather();
  } catch (...) {
    I.end_device();
    throw;
  }
  // device memory is no longer needed: only the pinned host copy backs the exported batches
  I.end_device();
  I.buckets.clear();  // drop the decoded inputs
  I.file_in.clear();
  I.ran = true;
}

int DeviceScan::num_buckets() const { return impl_->nb; }
int64_t DeviceScan::num_rows(int b) const { return impl_->out_base[b + 1] - impl_->out_base[b]; }
int64_t DeviceScan::num_batches(int b) const { return impl_->batch_base[b + 1] - impl_->batch_base[b]; }
const ScanCounters& DeviceScan::counters() const { return impl_->ctr; }
const uint32_t* DeviceScan::sorted_buckets() const {
  return impl_->h_sorted_buckets ? reinterpret_cast<const uint32_t*>(impl_->h_sorted_buckets.get()) : nullptr;
}

std::shared_ptr<arrow::RecordBatch> DeviceScan::Batch(int b, int64_t i) const {
  const Impl& I = *impl_;
  const int64_t bs = I.spec.batch_size;
  const int64_t U = I.out_base[b + 1] - I.out_base[b];
  const int64_t rows = std::min<int64_t>(bs, U - i * bs);
  if (rows <= 0) fail("batch index out of range");
  const int64_t gb = I.batch_base[b] + i;
  const int64_t u0 = I.out_base[b] + i * bs;
  std::shared_ptr<void> owner = std::const_pointer_cast<Impl>(I.shared_from_this());
  std::vector<std::shared_ptr<arrow::Array>> arrays;
  for (int c = 0; c < I.C; ++c) {
    const auto& oc = I.outs[c];
    const auto& type = I.spec.physical_schema->field(c)->type();
    std::shared_ptr<arrow::Buffer> valid;
    int64_t null_count = 0;
    if (oc.has_valid) {
      const uint8_t* vp = oc.h_valid.get() + gb * I.vstride;
      null_count = rows - arrow::internal::CountSetBits(vp, 0, rows);
      if (null_count > 0) valid = std::make_shared<KeepAliveBuffer>(vp, (rows + 7) / 8, owner);
    }
    std::shared_ptr<arrow::ArrayData> data;
    if (oc.kind.var) {
      const uint8_t* op = oc.h_offsets.get() + (u0 + gb) * 4;
      const int64_t d0 = oc.h_bstart.get()[gb], d1 = oc.h_bstart.get()[gb + 1];
      auto offsets = std::make_shared<KeepAliveBuffer>(op, (rows + 1) * 4, owner);
      auto bytes = std::make_shared<KeepAliveBuffer>(oc.h_data.get() + d0, d1 - d0, owner);
      data = arrow::ArrayData::Make(type, rows, {valid, offsets, bytes}, null_count, 0);
    } else if (oc.kind.width == 0) {
      auto values = std::make_shared<KeepAliveBuffer>(oc.h_values.get() + gb * I.vstride, (rows + 7) / 8, owner);
      data = arrow::ArrayData::Make(type, rows, {valid, values}, null_count, 0);
    } else {
      const int w = oc.kind.width;
      auto values = std::make_shared<KeepAliveBuffer>(oc.h_values.get() + u0 * w, rows * w, owner);
      data = arrow::ArrayData::Make(type, rows, {valid, values}, null_count, 0);
    }
    arrays.push_back(arrow::MakeArray(data));
  }
  return arrow::RecordBatch::Make(I.spec.physical_schema, rows, std::move(arrays));
}

}  // namespace lsb
