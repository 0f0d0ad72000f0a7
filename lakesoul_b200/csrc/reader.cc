// SPDX-License-Identifier: Apache-2.0
//
// reader.cc -- per-hash-bucket reader behind the lakesoul-io-c C-ABI.
// Follows the observable behaviour of LakeSoulReader::start / next_rb
// (rust/lakesoul-io/src/reader.rs:138-232, 355-363), LakeSoulIOSession::build_physical_plan
// (session.rs:720-882), LakeSoulParquetFormat::create_physical_plan (file_format.rs:234-310),
// MergeParquetExec (physical_plan/merge/mod.rs:61-120, 283-374) and DefaultColumnStream
// (stream/default_column.rs:49-108) -- with the merge itself on the GPU (device_scan.cc).
#include "reader.h"

#include <arrow/compute/api.h>
#include <arrow/io/file.h>
#include <parquet/arrow/reader.h>
#include <parquet/arrow/schema.h>
#include <parquet/file_reader.h>
#include <sys/stat.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <future>
#include <set>
#include <stdexcept>

#include <arrow/io/memory.h>

#include "bucket_hash.h"
#include "device_scan.h"
#include "filter.h"
#include "page_plan.h"

namespace lsb {

namespace {

[[noreturn]] void fail(const std::string& m) { throw std::runtime_error(m); }

constexpr const char* kNullString = "__L@KE$OUL_NULL__";         // constant.rs:18
constexpr const char* kEmptyString = "__L@KE$OUL_EMPTY_STRING__";  // constant.rs:19

double now_ms() {
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

std::string upper(std::string s) {
  for (auto& c : s) c = (char)toupper((unsigned char)c);
  return s;
}

std::string strip_scheme(const std::string& path) {
  // local files only: "file:///a/b" or "file:/a/b" or plain paths
  if (path.rfind("file://", 0) == 0) return path.substr(7);
  if (path.rfind("file:", 0) == 0) return path.substr(5);
  if (path.find("://") != std::string::npos)
    fail("object store '" + path.substr(0, path.find("://")) + "' is not supported by the GPU reader (local files only)");
  return path;
}

// helpers/transform.rs:35-55 : timestamps with a zone are normalised to UTC
std::shared_ptr<arrow::Field> uniform_field(const std::shared_ptr<arrow::Field>& f) {
  if (f->type()->id() == arrow::Type::TIMESTAMP) {
    const auto& t = static_cast<const arrow::TimestampType&>(*f->type());
    if (!t.timezone().empty()) return f->WithType(arrow::timestamp(t.unit(), "UTC"));
  }
  return f;
}

// helpers/transform.rs:356-513 make_default_array
std::shared_ptr<arrow::Array> make_default_array(const std::shared_ptr<arrow::DataType>& type,
                                                 const std::string& value, int64_t rows) {
  auto nulls = [&]() { return arrow::MakeArrayOfNull(type, rows).ValueOrDie(); };
  if (value == kNullString) return nulls();
  std::shared_ptr<arrow::Scalar> scalar;
  auto parse_i64 = [&](int64_t* out) {
    if (value.empty()) return false;
    char* end = nullptr;
    errno = 0;
    long long v = strtoll(value.c_str(), &end, 10);
    if (*end != 0 || errno != 0) return false;
    *out = v;
    return true;
  };
  switch (type->id()) {
    case arrow::Type::STRING:
      scalar = std::make_shared<arrow::StringScalar>(value == kEmptyString ? std::string() : value);
      break;
    case arrow::Type::INT32: {
      int64_t v;
      if (!parse_i64(&v) || v < INT32_MIN || v > INT32_MAX) fail("invalid digit found in string: " + value);
      scalar = std::make_shared<arrow::Int32Scalar>((int32_t)v);
    } break;
    case arrow::Type::INT64: {
      int64_t v;
      if (!parse_i64(&v)) fail("invalid digit found in string: " + value);
      scalar = std::make_shared<arrow::Int64Scalar>(v);
    } break;
    case arrow::Type::DATE32: {
      int64_t v;
      if (parse_i64(&v)) {
        scalar = std::make_shared<arrow::Date32Scalar>((int32_t)v);
      } else {
        auto r = arrow::Scalar::Parse(type, value);  // YYYY-MM-DD
        if (!r.ok()) fail("cannot parse date partition value '" + value + "'");
        scalar = *r;
      }
    } break;
    case arrow::Type::TIMESTAMP: {
      int64_t v;
      if (parse_i64(&v)) {
        scalar = std::make_shared<arrow::TimestampScalar>(v, type);
      } else {
        std::string s = value;
        auto r = arrow::Scalar::Parse(arrow::timestamp(static_cast<const arrow::TimestampType&>(*type).unit()), s);
        if (!r.ok()) fail("cannot parse timestamp partition value '" + value + "'");
        scalar = std::make_shared<arrow::TimestampScalar>(static_cast<const arrow::TimestampScalar&>(**r).value, type);
      }
    } break;
    case arrow::Type::BOOL: {
      if (value == "true") scalar = std::make_shared<arrow::BooleanScalar>(true);
      else if (value == "false") scalar = std::make_shared<arrow::BooleanScalar>(false);
      else fail("provided string was not `true` or `false`");
    } break;
    default: {
      auto r = arrow::Scalar::Parse(type, value);
      if (!r.ok()) {
        fprintf(stderr, "[lakesoul_b200] make_default_array() datatype not match, datatype=%s, value=%s\n",
                type->ToString().c_str(), value.c_str());
        return nulls();
      }
      scalar = *r;
    }
  }
  return arrow::MakeArrayFromScalar(*scalar, rows).ValueOrDie();
}

struct FileInfo {
  std::string path;
  std::shared_ptr<RawParquetFile> raw;    // whole file in pinned memory + parsed footer
  std::shared_ptr<arrow::Schema> schema;  // partition columns removed
  int64_t rows = 0;
  int group = 0;                          // hash bucket of the reader's file list this file belongs to (b200_file_groups)
};

// eq(column, literal) terms at the top level of a conjunction
void collect_eq_terms(const FilterExpr& e, std::map<std::string, std::shared_ptr<arrow::Scalar>>* out) {
  if (e.kind == FilterExpr::kAnd) {
    collect_eq_terms(*e.left, out);
    collect_eq_terms(*e.right, out);
  } else if (e.kind == FilterExpr::kCmp && e.op == FilterExpr::kEq && e.literal && e.literal->is_valid) {
    out->emplace(e.column, e.literal);
  }
}

bool force_host_decode() {
  const char* v = getenv("LAKESOUL_B200_HOST_DECODE");
  return v && (v[0] == '1' || v[0] == 't');
}

}  // namespace

// ---------------------------------------------------------------------------------------
// Streaming of a bucket by primary-key range. The reference streams a bucket batch by batch
// (sorted_stream_merger.rs:454-524); the GPU reader merges a whole scan at once, so a bucket too large for that is read
// as a sequence of scans over ascending, disjoint ranges of the FIRST primary-key column: every version of a key falls
// into the same range, so merging the ranges one after the other and concatenating gives the bucket's merge result, in
// the same order (nulls first). A range is an ordinary pushed key predicate: row groups outside it are pruned by their
// statistics (page_plan.cc PruneRowGroups), only the byte span of the kept row groups is read (LoadKept), and rows of
// straddling row groups are dropped by the post-merge filter on the device.
// ---------------------------------------------------------------------------------------
namespace {
struct SplitKey {
  int64_t i = 0;
  std::string s;
  bool bytes = false;
  bool operator<(const SplitKey& o) const { return bytes ? s < o.s : i < o.i; }
  bool operator==(const SplitKey& o) const { return bytes ? s == o.s : i == o.i; }
};

// literal of the column's own filter-literal type (filter.cc parse_literal); nullptr = type not usable for splitting
std::shared_ptr<arrow::Scalar> split_literal(const arrow::DataType& type, const SplitKey& k) {
  switch (type.id()) {
    case arrow::Type::INT8: return std::make_shared<arrow::Int8Scalar>((int8_t)k.i);
    case arrow::Type::INT16: return std::make_shared<arrow::Int16Scalar>((int16_t)k.i);
    case arrow::Type::INT32: return std::make_shared<arrow::Int32Scalar>((int32_t)k.i);
    case arrow::Type::INT64: return std::make_shared<arrow::Int64Scalar>(k.i);
    case arrow::Type::DATE32: return std::make_shared<arrow::Date32Scalar>((int32_t)k.i);
    case arrow::Type::TIMESTAMP:
      if (static_cast<const arrow::TimestampType&>(type).unit() != arrow::TimeUnit::MICRO) return nullptr;
      return std::make_shared<arrow::TimestampScalar>(k.i, arrow::timestamp(arrow::TimeUnit::MICRO, "UTC"));
    case arrow::Type::STRING: return std::make_shared<arrow::StringScalar>(k.s);
    case arrow::Type::BINARY: return std::make_shared<arrow::BinaryScalar>(arrow::Buffer::FromString(k.s));
    default: return nullptr;
  }
}

std::shared_ptr<FilterExpr> cmp_expr(const std::string& col, FilterExpr::Op op, std::shared_ptr<arrow::Scalar> lit) {
  auto e = std::make_shared<FilterExpr>();
  e->kind = FilterExpr::kCmp;
  e->op = op;
  e->column = col;
  e->literal = std::move(lit);
  return e;
}
std::shared_ptr<FilterExpr> bin_expr(FilterExpr::Kind k, std::shared_ptr<FilterExpr> l, std::shared_ptr<FilterExpr> r) {
  auto e = std::make_shared<FilterExpr>();
  e->kind = k;
  e->left = std::move(l);
  e->right = std::move(r);
  return e;
}
}  // namespace

std::vector<std::shared_ptr<FilterExpr>> plan_stream_passes(const std::vector<std::string>& paths, const std::string& pk,
                                                            bool pk_nullable, int64_t pass_bytes) {
  std::vector<std::shared_ptr<FilterExpr>> none;
  struct Ev { SplitKey key, max; int64_t bytes; };
  std::vector<Ev> evs;
  std::shared_ptr<arrow::DataType> type;
  int64_t total = 0, head_bytes = 0;
  for (const auto& path : paths) {
    struct stat st;
    if (stat(path.c_str(), &st) != 0 || st.st_size < 8) continue;  // reported by the scan itself
    auto f = RawParquetFile::OpenLazy(path);
    auto fld = f->schema()->GetFieldByName(pk);
    if (!fld) return none;
    if (!type) type = fld->type();
    else if (!type->Equals(*fld->type())) return none;  // a cast below the merge: leave it to one scan
    pk_nullable = pk_nullable || fld->nullable();
    const auto ranges = f->KeyRanges(pk);
    if (ranges.size() != f->kept_row_groups().size()) return none;
    for (const auto& r : ranges) {
      total += r.bytes;
      if (r.rows == 0) continue;
      if (r.fam == RowGroupKeyRange::kNone) {
        if (r.nulls == r.rows) {  // only null keys: they sort first, the first pass takes them
          head_bytes += r.bytes;
          continue;
        }
        return none;  // no statistics: where the rows fall is unknown
      }
      Ev e;
      e.key.bytes = e.max.bytes = r.fam == RowGroupKeyRange::kBytes;
      e.key.i = r.imin; e.key.s = r.bmin;
      e.max.i = r.imax; e.max.s = r.bmax;
      e.bytes = r.bytes;
      evs.push_back(std::move(e));
    }
  }
  if (!type || evs.empty() || total <= pass_bytes) return none;
  if (!split_literal(*type, evs.front().key)) return none;
  std::stable_sort(evs.begin(), evs.end(), [](const Ev& a, const Ev& b) { return a.key < b.key; });
  // Sweep over the distinct row-group minima m_1 < m_2 < ...: the pass [lo, m_i) decodes every row group with
  // min < m_i and max >= lo ("active"). It is closed at m_i when taking the row groups that start at m_i would exceed
  // the budget AND it finishes at least half a budget of row groups (max < m_i) -- row groups that span the whole key
  // space (small delta files) are decoded by every pass whatever the cuts, they must not force tiny passes.
  std::vector<SplitKey> cuts;
  std::vector<const Ev*> active;
  int64_t acc = head_bytes;
  for (size_t i = 0; i < evs.size();) {
    size_t j = i;
    int64_t g = 0;
    while (j < evs.size() && evs[j].key == evs[i].key) g += evs[j++].bytes;
    if (i > 0 && acc + g > pass_bytes) {
      int64_t done = 0;
      for (const Ev* a : active) if (a->max < evs[i].key) done += a->bytes;
      if (done >= pass_bytes / 2) {
        cuts.push_back(evs[i].key);
        std::vector<const Ev*> keep;
        acc = 0;
        for (const Ev* a : active)
          if (!(a->max < evs[i].key)) { keep.push_back(a); acc += a->bytes; }
        active.swap(keep);
      }
    }
    for (size_t k = i; k < j; ++k) active.push_back(&evs[k]);
    acc += g;
    i = j;
  }
  if (cuts.empty()) return none;
  std::vector<std::shared_ptr<FilterExpr>> out;
  for (size_t k = 0; k <= cuts.size(); ++k) {
    std::shared_ptr<FilterExpr> e;
    if (k > 0) e = cmp_expr(pk, FilterExpr::kGe, split_literal(*type, cuts[k - 1]));
    if (k < cuts.size()) {
      auto lt = cmp_expr(pk, FilterExpr::kLt, split_literal(*type, cuts[k]));
      e = e ? bin_expr(FilterExpr::kAnd, e, lt) : lt;
    }
    if (k == 0 && pk_nullable) {
      auto isn = std::make_shared<FilterExpr>();
      isn->kind = FilterExpr::kIsNull;
      isn->column = pk;
      e = bin_expr(FilterExpr::kOr, e, isn);
    }
    out.push_back(e);
  }
  return out;
}

// test hook (tests/test_stream_passes.py): the pass filters for `files` (newline-separated), one per line
extern "C" __attribute__((visibility("default"))) int64_t lsb_test_stream_passes(const char* files, const char* pk, const char* opts,
                                                                                 int64_t pass_bytes, char* out, int64_t cap) {
  std::string text;
  try {
    std::vector<std::string> paths;
    std::string cur;
    for (const char* p = files; *p; ++p) {
      if (*p == '\n') { if (!cur.empty()) paths.push_back(cur); cur.clear(); }
      else cur.push_back(*p);
    }
    if (!cur.empty()) paths.push_back(cur);
    for (const auto& e : plan_stream_passes(paths, pk, opts && strstr(opts, "nullable"), pass_bytes)) text += filter_to_string(*e) + "\n";
  } catch (const std::exception& e) {
    snprintf(out, (size_t)cap, "%s", e.what());
    return -1;
  }
  if ((int64_t)text.size() + 1 > cap) return -1;
  memcpy(out, text.c_str(), text.size() + 1);
  return (int64_t)text.size();
}

// ---------------------------------------------------------------------------------------
bool IOConfig::option(const std::string& key, std::string* out) const {
  auto it = options.find(key);
  if (it != options.end()) { *out = it->second; return true; }
  const std::string env = "LAKESOUL_" + upper(key);
  if (const char* v = getenv(env.c_str())) { *out = v; return true; }
  return false;
}
std::string IOConfig::cdc_column() const {
  std::string v;
  return option("cdc_column", &v) ? v : std::string();
}
bool IOConfig::is_compacted() const {
  std::string v;
  return option("is_compacted", &v) && v == "true";
}
bool IOConfig::skip_merge_on_read() const {
  std::string v;
  return option("skip_merge_on_read", &v) && v == "true";
}

int ensure_device() {
  // JVM task threads start readers concurrently: the device is chosen exactly once
  struct Choice { int dev = -1; std::string err; };
  static const Choice choice = [] {
    Choice c;
    // One reader per task slot = one stream per concurrent bucket. With the default of 8 hardware work queues,
    // 16+ streams share queues and a long kernel (page decompression) of one reader blocks the short kernels other
    // readers queued behind it: 32 queues lifted the 16-slot rate from 17 to 51 M rows/s (r2, config 5, ZSTD).
    // Must be set before the CUDA context exists; an explicit setting of the embedding process wins.
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
      c.err = std::string("lakesoul_b200: no CUDA device available (") +
              (e != cudaSuccess ? cudaGetErrorString(e) : "0 devices") +
              "); the merge-on-read engine has no CPU fallback";
      return c;
    }
    int want = 0;
    if (const char* v = getenv("LAKESOUL_B200_DEVICE")) want = atoi(v);
    else if (const char* v2 = getenv("LOCAL_RANK")) want = atoi(v2) % n;
    c.dev = (want >= 0 && want < n) ? want : 0;
    // keep freed blocks in the stream-ordered pool: readers are created per bucket
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, c.dev) == cudaSuccess) {
      uint64_t thr = UINT64_MAX;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    return c;
  }();
  if (choice.dev < 0) fail(choice.err);
  cudaSetDevice(choice.dev);
  static const int cache_pref = [] {
    const char* v = getenv("LAKESOUL_B200_CACHE_PREF");  // experiment: 1 = prefer shared, 2 = prefer L1, 3 = equal
    const int p = v ? atoi(v) : 0;
    if (p == 1) cudaDeviceSetCacheConfig(cudaFuncCachePreferShared);
    else if (p == 2) cudaDeviceSetCacheConfig(cudaFuncCachePreferL1);
    else if (p == 3) cudaDeviceSetCacheConfig(cudaFuncCachePreferEqual);
    return p;
  }();
  (void)cache_pref;
  return choice.dev;
}

// ---------------------------------------------------------------------------------------
struct Reader::Impl {
  IOConfig cfg;
  ReaderStats stats;
  std::shared_ptr<arrow::Schema> out_schema;
  std::unique_ptr<DeviceScan> scan;
  int64_t next_batch = 0, n_batches = 0;  // inside the current bucket of the scan
  int cur_bucket = 0;
  bool started = false;

  // count-only reads (empty projection, nothing to merge)
  bool count_only = false;
  int64_t count_rows_left = 0;

  struct OutField {
    enum Kind { kPhysical, kDefault, kNull } kind = kNull;
    int phys = -1;
    std::shared_ptr<arrow::Array> constant;  // batch_size rows, sliced for the last batch
    std::shared_ptr<arrow::DataType> type;
  };
  std::vector<OutField> out_fields;

  // a bucket streamed by key range: one inner reader per pass, the next one prepared while this one is consumed
  std::vector<std::shared_ptr<FilterExpr>> pass_filters;
  size_t pass = 0;
  std::unique_ptr<Reader> inner;
  std::future<std::unique_ptr<Reader>> next_inner;

  std::unique_ptr<Reader> make_pass(size_t k) const {
    IOConfig c = cfg;
    c.pass_filter = pass_filters[k];
    c.lazy_files = true;
    auto r = std::make_unique<Reader>(std::move(c));
    r->Start();
    return r;
  }
  void prefetch() {
    if (pass + 1 >= pass_filters.size()) return;
    static const bool off = [] { const char* v = getenv("LAKESOUL_B200_STREAM_PREFETCH"); return v && *v == '0'; }();
    const size_t k = pass + 1;
    next_inner = std::async(off ? std::launch::deferred : std::launch::async, [this, k] { return make_pass(k); });
  }
  void absorb(const ReaderStats& s) {
    stats.rows_in += s.rows_in; stats.rows_out += s.rows_out;
    stats.h2d_bytes += s.h2d_bytes; stats.d2h_bytes += s.d2h_bytes;
    stats.kernel_launches += s.kernel_launches;
    stats.device_decoded_pages += s.device_decoded_pages; stats.host_decoded_files += s.host_decoded_files;
    stats.device_decompressed_pages += s.device_decompressed_pages;
    stats.pruned_pages += s.pruned_pages;
    stats.decode_ms += s.decode_ms; stats.upload_ms += s.upload_ms; stats.merge_ms += s.merge_ms;
    stats.gather_ms += s.gather_ms; stats.download_ms += s.download_ms;
    stats.path = s.path;
  }
};

Reader::Reader(IOConfig cfg) : impl_(new Impl()) {
  if (cfg.files.empty()) fail("LakeSoulReader has the wrong number of files");  // reader.rs:108-110
  impl_->cfg = std::move(cfg);
}
Reader::~Reader() {
  if (impl_ && impl_->next_inner.valid()) {
    try { impl_->next_inner.get(); } catch (...) {}
  }
}
const ReaderStats& Reader::stats() const { return impl_->stats; }
const IOConfig& Reader::config() const { return impl_->cfg; }
std::shared_ptr<arrow::Schema> Reader::schema() const { return impl_->out_schema; }

void Reader::Start() {
  Impl& I = *impl_;
  if (I.started) return;
  const IOConfig& cfg = I.cfg;
  if (cfg.filter_protos > 0)
    fail("substrait filters (add_filter_proto) are not supported by the GPU reader; "
         "refusing to read without applying a pushed filter");
  // ---- hash-bucket elimination (N2), before a single byte is read: an equality filter on the WHOLE primary key
  // selects one hash bucket; when every file of this reader belongs to another one, no row can match
  // (helpers/mod.rs:929-1065 drops such files while planning). Needs the table's bucket count
  // (set_hash_bucket_num), LakeSoul file names and the key columns in the requested schema.
  if (!cfg.filter_strs.empty() && !cfg.primary_keys.empty() && !cfg.inferring_schema && cfg.target_schema &&
      getenv("LAKESOUL_B200_NO_PRUNING") == nullptr) {
    const long nbuckets = atol(cfg.hash_bucket_num.c_str());
    std::map<std::string, std::shared_ptr<arrow::Scalar>> eqs;
    bool ok = nbuckets > 1;
    if (ok) {
      try {
        for (const auto& fs : cfg.filter_strs) collect_eq_terms(*parse_filter(fs, *cfg.target_schema), &eqs);
      } catch (const std::exception&) {
        ok = false;  // the full parse below reports the error
      }
    }
    uint32_t h = kHashSeed;
    for (const auto& pkc : cfg.primary_keys) {
      if (!ok) break;
      auto it = eqs.find(pkc);
      auto fld = cfg.target_schema->GetFieldByName(pkc);
      if (it == eqs.end() || !fld || !it->second->type->Equals(*fld->type())) { ok = false; break; }
      ok = hash_scalar(*it->second, h, &h);
    }
    if (ok) {
      const int want = (int)(h % (uint32_t)nbuckets);
      bool all_other = true;
      for (const auto& raw : cfg.files) {
        const int id = bucket_id_of_file(raw);
        if (id < 0 || id == want) { all_other = false; break; }
      }
      if (all_other) {
        std::vector<std::shared_ptr<arrow::Field>> out_f;
        for (const auto& f : cfg.target_schema->fields()) out_f.push_back(uniform_field(f));
        I.out_schema = arrow::schema(out_f);
        I.stats.pruned_buckets = 1;
        I.stats.path = "bucket-pruned";
        I.started = true;  // empty stream
        return;
      }
    }
  }
  ensure_device();

  // ---- open files, read footers (session.rs:546-593) -----------------------------------------
  const double t_dec0 = now_ms();
  trace_mark("start");
  // the scan slot (stream + device arena) is taken now: every file image is sent to the GPU as soon as
  // it is in pinned memory, overlapping the read of the next file
  std::shared_ptr<ScanSession> session;  // taken below, once it is known that this reader scans by itself
  const bool early_upload = !force_host_decode();
  // Extension (option b200_file_groups = "n0,n1,..."): the file list is several hash buckets back to back, n_i files
  // each; every bucket is merged on its own, in ONE device scan (one framing pass, one launch set -- SURVEY.md 7.2
  // "tiny units at scale"), and the stream returns bucket 0's batches, then bucket 1's, ... (no batch spans two buckets).
  std::vector<int> group_of(cfg.files.size(), 0);
  int ngroups = 1;
  {
    std::string gs;
    if (cfg.option("b200_file_groups", &gs) && !gs.empty()) {
      size_t pos = 0, f = 0;
      int g = 0;
      while (pos < gs.size()) {
        const size_t comma = gs.find(',', pos);
        const long n = atol(gs.substr(pos, comma == std::string::npos ? std::string::npos : comma - pos).c_str());
        if (n <= 0 || f + (size_t)n > cfg.files.size()) fail("b200_file_groups does not match the file list");
        for (long i = 0; i < n; ++i) group_of[f++] = g;
        ++g;
        if (comma == std::string::npos) break;
        pos = comma + 1;
      }
      if (f != cfg.files.size()) fail("b200_file_groups does not match the file list");
      ngroups = g;
    }
  }
  // The whole scan group is resident at once (file images in pinned memory, decoded columns and outputs in HBM). A
  // merged bucket whose files hold more than the pass budget is streamed instead: a sequence of scans over ascending
  // primary-key ranges (plan_stream_passes above), each an inner reader with the range pushed as a key predicate;
  // pass k + 1 is prepared while the caller consumes pass k. What cannot be split (no key, no statistics, file
  // groups) stays one scan and is refused above LAKESOUL_B200_MAX_SCAN_BYTES with a clear message instead of failing
  // in an allocator.
  if (!cfg.lazy_files) {
    static const int64_t max_bytes = [] {
      const char* v = getenv("LAKESOUL_B200_MAX_SCAN_BYTES");
      return v ? atoll(v) : ((int64_t)24 << 30);
    }();
    int64_t pass_bytes = (int64_t)4 << 30;  // uncompressed bytes per pass
    {
      std::string v;
      if (cfg.option("b200_stream_pass_bytes", &v) && !v.empty()) pass_bytes = atoll(v.c_str());
    }
    int64_t total = 0;
    std::vector<std::string> paths;
    for (const auto& raw : cfg.files) {
      struct stat st0;
      if (raw.find("://") == std::string::npos || raw.rfind("file:", 0) == 0) {
        paths.push_back(strip_scheme(raw));
        if (stat(paths.back().c_str(), &st0) == 0) total += (int64_t)st0.st_size;
      }
    }
    const bool merged = !cfg.primary_keys.empty() && !cfg.skip_merge_on_read();
    // (file bytes are a lower bound of the uncompressed bytes: footers are read only when streaming is plausible)
    if (pass_bytes > 0 && merged && ngroups == 1 && paths.size() == cfg.files.size() && total > pass_bytes / 8) {
      bool nullable = true;
      if (!cfg.inferring_schema && cfg.target_schema)
        if (auto f = cfg.target_schema->GetFieldByName(cfg.primary_keys[0])) nullable = f->nullable();
      I.pass_filters = plan_stream_passes(paths, cfg.primary_keys[0], nullable, pass_bytes);
      if (I.pass_filters.size() > 1) {
        I.pass = 0;
        I.inner = I.make_pass(0);
        I.out_schema = I.inner->schema();
        I.absorb(I.inner->stats());
        I.stats.pruned_row_groups = 0;  // (row groups outside a pass are not "pruned by the filter")
        I.stats.stream_passes = (int32_t)I.pass_filters.size();
        I.prefetch();
        I.started = true;
        return;
      }
      I.pass_filters.clear();
    }
    if (total > max_bytes)
      fail("the files of this reader hold " + std::to_string(total >> 20) + " MiB, more than LAKESOUL_B200_MAX_SCAN_BYTES (" +
           std::to_string(max_bytes >> 20) + " MiB), and cannot be streamed by key range (no primary-key statistics, a single "
           "key value, or several file groups): the GPU reader keeps a whole scan resident; compact the bucket or raise the limit");
  }
  session = begin_scan_session();
  // Lazy IO -- footers first, then only the byte segments of the row groups / pages / columns that are read: always for
  // the passes of a streamed bucket, and for readers whose pushed filters name a primary-key column (a key lookup or a
  // key range reads a few pages of every file instead of the files). Everything else reads whole files and sends each
  // to the GPU while the next one is read.
  bool lazy_io = cfg.lazy_files;
  if (!lazy_io && !cfg.primary_keys.empty() && getenv("LAKESOUL_B200_NO_PRUNING") == nullptr) {
    static const bool on = [] { const char* v = getenv("LAKESOUL_B200_LAZY_FILTERED"); return !(v && *v == '0'); }();
    for (const auto& fs : cfg.filter_strs)
      for (const auto& pk : cfg.primary_keys) lazy_io = lazy_io || (on && fs.find(pk) != std::string::npos);
  }
  std::vector<FileInfo> files;
  size_t file_no = 0;
  for (const auto& raw : cfg.files) {
    const int this_group = group_of[file_no++];
    const std::string path = strip_scheme(raw);
    struct stat st;
    if (stat(path.c_str(), &st) != 0) fail("Object at location " + path + " not found");
    if (st.st_size < 8) {  // session.rs:423-434
      fprintf(stderr, "[lakesoul_b200] File %s, size %lld, is invalid\n", path.c_str(), (long long)st.st_size);
      continue;
    }
    FileInfo fi;
    fi.path = path;
    fi.group = this_group;
    fi.raw = lazy_io ? RawParquetFile::OpenLazy(path) : RawParquetFile::Open(path);
    if (early_upload && !lazy_io)
      fi.raw->set_device_image(session_upload(*session, fi.raw->image(), (size_t)fi.raw->image_bytes()));
    std::vector<std::shared_ptr<arrow::Field>> kept;
    for (const auto& f : fi.raw->schema()->fields())
      if (!cfg.partition_schema->GetFieldByName(f->name())) kept.push_back(f);
    fi.schema = arrow::schema(kept);
    fi.rows = fi.raw->num_rows();
    files.push_back(std::move(fi));
  }
  trace_mark("open_read");

  // ---- table schema = target (uniformed) + file-only fields (session.rs:596-626) ---------------
  std::vector<std::shared_ptr<arrow::Field>> target_fields;
  if (!cfg.inferring_schema)
    for (const auto& f : cfg.target_schema->fields()) target_fields.push_back(uniform_field(f));
  std::vector<std::shared_ptr<arrow::Field>> table_fields = target_fields;
  auto in_fields = [](const std::vector<std::shared_ptr<arrow::Field>>& v, const std::string& n) {
    for (const auto& f : v) if (f->name() == n) return true;
    return false;
  };
  for (const auto& fi : files)
    for (const auto& f : fi.schema->fields())
      if (!in_fields(table_fields, f->name())) table_fields.push_back(f);
  auto table_schema = arrow::schema(table_fields);
  // requested columns: the target schema, or everything when inferring
  std::vector<std::shared_ptr<arrow::Field>> requested = cfg.inferring_schema ? table_fields : target_fields;

  // ---- filters (config/mod.rs:702-731) -------------------------------------------------------
  std::vector<std::shared_ptr<FilterExpr>> fexprs;
  for (const auto& s : cfg.filter_strs) fexprs.push_back(parse_filter(s, *table_schema));
  if (cfg.pass_filter) fexprs.push_back(cfg.pass_filter);
  std::shared_ptr<FilterExpr> filter = conjunction(fexprs);
  if (filter) {
    // predicates on default-valued / partition columns and on columns no file carries are constants: the reference
    // evaluates pushed filters after DefaultColumnStream filled those columns in (session.rs:839-877), the device
    // never carries them
    filter = fold_constant_columns(filter, [&](const std::string& name, std::shared_ptr<arrow::Scalar>* out) {
      auto fld = table_schema->GetFieldByName(name);
      if (!fld) return false;
      auto dv = cfg.default_column_value.find(name);
      if (dv != cfg.default_column_value.end()) {
        auto arr = make_default_array(fld->type(), dv->second, 1);
        auto sc = arr->GetScalar(0);
        *out = sc.ok() ? *sc : nullptr;
        return true;
      }
      for (const auto& fi : files)
        if (fi.schema->GetFieldByName(name)) return false;
      *out = nullptr;  // in the table schema, in no file: NULL for every row
      return true;
    });
  }
  std::set<std::string> fcols;
  if (filter) filter_columns(*filter, &fcols);
  // key predicates prune whole row groups by their min/max statistics before anything is decoded
  if (filter && !cfg.primary_keys.empty() && getenv("LAKESOUL_B200_NO_PRUNING") == nullptr) {
    const std::set<std::string> pk_set(cfg.primary_keys.begin(), cfg.primary_keys.end());
    bool touches_pk = false;
    for (const auto& c : fcols) touches_pk = touches_pk || pk_set.count(c) > 0;
    if (touches_pk) {
      static const bool page_pruning = [] { const char* v = getenv("LAKESOUL_B200_PAGE_PRUNING"); return !(v && *v == '0'); }();
      for (auto& fi : files) {
        I.stats.pruned_row_groups += fi.raw->PruneRowGroups(*filter, pk_set);
        if (page_pruning) fi.raw->PrunePages(*filter, pk_set);  // session.rs:685-717: the reference prunes by page index too
        fi.rows = fi.raw->num_rows();
      }
    }
  }

  // ---- merged schema: requested + filter + pk + cdc columns, table order (file_format.rs:254-265)
  const std::string cdc = cfg.cdc_column();
  auto is_pk = [&](const std::string& n) {
    for (const auto& p : cfg.primary_keys) if (p == n) return true;
    return false;
  };
  std::vector<std::shared_ptr<arrow::Field>> merged_fields;
  for (const auto& f : table_fields) {
    const std::string& n = f->name();
    if (in_fields(requested, n) || fcols.count(n) || is_pk(n) || (!cdc.empty() && n == cdc))
      merged_fields.push_back(f);
  }
  auto in_any_file = [&](const std::string& n) {
    for (const auto& fi : files) if (fi.schema->GetFieldByName(n)) return true;
    return false;
  };

  // ---- physical schema: merged minus default-valued and never-present columns ---------------
  std::vector<std::shared_ptr<arrow::Field>> phys_fields;
  for (const auto& f : merged_fields) {
    if (cfg.default_column_value.count(f->name())) continue;  // merge/mod.rs:314-327
    if (!in_any_file(f->name())) continue;                      // stays null
    phys_fields.push_back(f);
  }
  auto physical_schema = arrow::schema(phys_fields);

  // ---- output fields + nullability (merge/mod.rs:75-97) --------------------------------------
  std::vector<std::shared_ptr<arrow::Field>> out_f;
  for (const auto& f : requested) {
    Impl::OutField of;
    of.type = f->type();
    bool nullable = f->nullable();
    for (const auto& fi : files) {
      auto ff = fi.schema->GetFieldByName(f->name());
      nullable = nullable || !ff || ff->nullable();
    }
    auto dv = cfg.default_column_value.find(f->name());
    const int pi = physical_schema->GetFieldIndex(f->name());
    if (dv != cfg.default_column_value.end()) {
      of.kind = Impl::OutField::kDefault;
      of.constant = make_default_array(f->type(), dv->second, (int64_t)cfg.batch_size);
    } else if (pi >= 0) {
      of.kind = Impl::OutField::kPhysical;
      of.phys = pi;
    } else {
      of.kind = Impl::OutField::kNull;
      of.constant = arrow::MakeArrayOfNull(f->type(), (int64_t)cfg.batch_size).ValueOrDie();
    }
    I.out_fields.push_back(of);
    out_f.push_back(f->WithNullable(nullable));
  }
  I.out_schema = arrow::schema(out_f);

  // ---- merge decision (merge/mod.rs:297-304) --------------------------------------------------
  bool merge_on_read;
  if (cfg.skip_merge_on_read() || cfg.primary_keys.empty()) merge_on_read = false;
  else merge_on_read = cfg.lazy_files || !(cfg.files.size() == 1 && cfg.merge_operators.empty() && cfg.is_compacted());
  // (with file groups, and in the passes of a streamed bucket, a one-file bucket is merged like the others: a sorted,
  // compacted file comes out unchanged)

  if (files.empty()) {  // no valid file: empty stream with the requested schema
    I.started = true;
    return;
  }

  // ---- count-only read with nothing to merge or filter ----------------------------------------
  if (phys_fields.empty()) {
    I.count_only = true;
    for (const auto& fi : files) I.count_rows_left += fi.rows;
    I.stats.rows_in = I.stats.rows_out = I.count_rows_left;
    I.stats.path = "count";
    I.started = true;
    return;
  }

  if (lazy_io) {
    // only now is it known which row groups, pages and columns this pass reads: the byte segments of the kept row
    // groups -- with a page index, of the live pages -- of the physical columns
    std::vector<std::string> names;
    for (const auto& pf : phys_fields) names.push_back(pf->name());
    for (auto& fi : files) {
      fi.raw->LoadKept(&names);
      if (early_upload && fi.raw->image_bytes() > 0)
        fi.raw->set_device_image(session_upload(*session, fi.raw->image(), (size_t)fi.raw->image_bytes()));
    }
    trace_mark("load_kept");
  }

  // ---- frame every file for the device decoders; host Arrow decode only where they do not apply
  std::vector<BucketFiles> groups((size_t)ngroups);
  const bool host_only = force_host_decode();
  for (auto& fi : files) {
    bool any = false;
    for (const auto& pf : phys_fields) any = any || fi.schema->GetFieldByName(pf->name());
    if (!any) fail("file " + fi.path + " carries none of the requested columns");
    std::string why;
    if (!host_only && fi.raw->Plan(phys_fields, &why)) {
      FileInput in;
      in.raw = fi.raw;
      I.stats.pruned_pages += fi.raw->pruned_pages();
      groups[(size_t)fi.group].files.push_back(std::move(in));
      continue;
    }
    if (!host_only && getenv("LAKESOUL_B200_VERBOSE"))
      fprintf(stderr, "[lakesoul_b200] %s: host decode (%s)\n", fi.path.c_str(), why.c_str());
    // host fallback: Arrow C++ decode from the pinned image already in memory (a lazily opened file: image + pread)
    auto rd = parquet::arrow::OpenFile(fi.raw->source(), arrow::default_memory_pool());
    if (!rd.ok()) fail(rd.status().ToString());
    std::unique_ptr<parquet::arrow::FileReader> reader = std::move(*rd);
    reader->set_use_threads(false);
    const auto& manifest = reader->manifest();
    std::vector<int> leaf;
    std::vector<std::shared_ptr<arrow::Field>> want;
    for (const auto& pf : phys_fields) {
      if (fi.schema->GetFieldIndex(pf->name()) < 0) continue;
      int full_idx = -1;
      for (size_t s2 = 0; s2 < manifest.schema_fields.size(); ++s2)
        if (manifest.schema_fields[s2].field->name() == pf->name()) { full_idx = (int)s2; break; }
      if (full_idx < 0 || manifest.schema_fields[full_idx].column_index < 0)
        fail("nested column '" + pf->name() + "' is not supported by the GPU reader");
      leaf.push_back(manifest.schema_fields[full_idx].column_index);
      want.push_back(pf);
    }
    auto tr = reader->ReadRowGroups(fi.raw->kept_row_groups(), leaf);
    if (!tr.ok()) fail(tr.status().ToString());
    std::shared_ptr<arrow::Table> table = *tr;
    if (fi.raw->live_rows() != table->num_rows()) table = table->Slice(fi.raw->live_begin(), fi.raw->live_rows());  // page-pruned
    // a2: reorder to target order + cast to the target type (helpers/transform.rs:107-354)
    std::vector<std::shared_ptr<arrow::ChunkedArray>> cols;
    std::vector<std::shared_ptr<arrow::Field>> flds;
    for (const auto& pf : want) {
      auto col = table->GetColumnByName(pf->name());
      if (!col) fail("column '" + pf->name() + "' missing after decode of " + fi.path);
      if (!col->type()->Equals(*pf->type())) {
        if (col->type()->id() == arrow::Type::TIMESTAMP && pf->type()->id() == arrow::Type::TIMESTAMP &&
            static_cast<const arrow::TimestampType&>(*col->type()).unit() ==
                static_cast<const arrow::TimestampType&>(*pf->type()).unit()) {
          auto v = col->View(pf->type());
          if (!v.ok()) fail(v.status().ToString());
          col = *v;
        } else {
          auto r = arrow::compute::Cast(arrow::Datum(col), pf->type(), arrow::compute::CastOptions::Safe());
          if (!r.ok())
            fail("Failed to cast type from " + col->type()->ToString() + " to " + pf->type()->ToString() +
                 ": " + r.status().ToString());
          col = r->chunked_array();
        }
      }
      cols.push_back(col);
      flds.push_back(pf);
    }
    FileInput in;
    in.table = arrow::Table::Make(arrow::schema(flds), cols, table->num_rows());
    groups[(size_t)fi.group].files.push_back(std::move(in));
  }
  I.stats.decode_ms = now_ms() - t_dec0;
  trace_mark("frame");

  ScanSpec spec;
  spec.physical_schema = physical_schema;
  spec.primary_keys = cfg.primary_keys;
  spec.merge = merge_on_read;
  spec.batch_size = (int)cfg.batch_size;
  spec.filter = filter;
  for (const auto& f : phys_fields) {
    auto it = cfg.merge_operators.find(f->name());
    spec.ops.push_back(it == cfg.merge_operators.end() ? MergeOp::UseLast : merge_op_from_name(it->second));
  }
  std::vector<BucketFiles> buckets;
  for (auto& g : groups)
    if (!g.files.empty()) buckets.push_back(std::move(g));  // a bucket whose files were all skipped has no rows
  I.scan.reset(new DeviceScan(std::move(spec), std::move(buckets), std::move(session)));
  I.scan->Run();
  trace_mark("cleanup");
  trace_flush("reader");
  I.cur_bucket = 0;
  I.n_batches = I.scan->num_batches(0);
  const ScanCounters& c = I.scan->counters();
  I.stats.rows_in = c.rows_in;
  I.stats.rows_out = c.rows_out;
  I.stats.h2d_bytes = c.h2d_bytes;
  I.stats.d2h_bytes = c.d2h_bytes;
  I.stats.kernel_launches = c.kernel_launches;
  I.stats.device_decoded_pages = c.device_decoded_pages;
  I.stats.device_decompressed_pages = c.device_decompressed_pages;
  I.stats.host_decoded_files = c.host_decoded_files;
  I.stats.upload_ms = c.upload_ms;
  I.stats.merge_ms = c.merge_ms;
  I.stats.gather_ms = c.gather_ms;
  I.stats.download_ms = c.download_ms;
  I.stats.path = c.path;
  I.started = true;
}

std::shared_ptr<arrow::RecordBatch> Reader::Next() {
  Impl& I = *impl_;
  if (!I.started) fail("reader not started");
  const int64_t bs = (int64_t)I.cfg.batch_size;
  int64_t rows = 0;
  std::shared_ptr<arrow::RecordBatch> phys;
  if (I.count_only) {
    if (I.count_rows_left <= 0) return nullptr;
    rows = std::min(bs, I.count_rows_left);
    I.count_rows_left -= rows;
  } else if (!I.pass_filters.empty()) {  // streamed bucket: the passes one after the other
    while (true) {
      if (!I.inner) return nullptr;
      if (auto b = I.inner->Next()) return b;  // (the inner reader already produced the output fields)
      I.inner.reset();
      if (I.pass + 1 >= I.pass_filters.size()) return nullptr;
      I.inner = I.next_inner.get();  // rethrows what the pass threw
      I.pass += 1;
      I.absorb(I.inner->stats());
      I.prefetch();
    }
  } else {
    if (!I.scan) return nullptr;
    while (I.next_batch >= I.n_batches) {  // next bucket of the scan
      if (I.cur_bucket + 1 >= I.scan->num_buckets()) return nullptr;
      I.cur_bucket += 1;
      I.next_batch = 0;
      I.n_batches = I.scan->num_batches(I.cur_bucket);
    }
    phys = I.scan->Batch(I.cur_bucket, I.next_batch++);
    rows = phys->num_rows();
  }
  std::vector<std::shared_ptr<arrow::Array>> arrays;
  for (const auto& of : I.out_fields) {
    if (of.kind == Impl::OutField::kPhysical) arrays.push_back(phys->column(of.phys));
    else arrays.push_back(rows == of.constant->length() ? of.constant : of.constant->Slice(0, rows));
  }
  return arrow::RecordBatch::Make(I.out_schema, rows, std::move(arrays));
}

}  // namespace lsb
