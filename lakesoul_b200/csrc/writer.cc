// SPDX-License-Identifier: Apache-2.0
//
// writer.cc -- see writer.h. Observable behaviour follows
//   create_writer                         rust/lakesoul-io/src/writer/mod.rs:61-128 (which writer, file names)
//   SyncSendableMutableLakeSoulWriter     writer/mod.rs:145-400 (write_batch, flush_and_close summary, abort)
//   MultiPartAsyncWriter::try_new         writer/async_writer/multipart_writer.rs:63-140 (file schema, writer properties)
//   PartitioningAsyncWriter::pull_and_sink writer/async_writer/partitioning_writer.rs:245-375 (paths, partition_desc)
// The row order inside every file is the reference's: SortExec over (range partitions, primary keys, aux sort
// columns) followed by the hash repartitioning keeps rows of one (partition, bucket) in sorted order.
#include "writer.h"

#include <arrow/io/file.h>
#include <parquet/arrow/writer.h>
#include <parquet/properties.h>
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <filesystem>
#include <random>
#include <stdexcept>

#include "device_scan.h"

namespace lsb {

namespace {

[[noreturn]] void fail(const std::string& m) { throw std::runtime_error(m); }

constexpr const char* kNullString = "__L@KE$OUL_NULL__";          // constant.rs:18
constexpr const char* kEmptyString = "__L@KE$OUL_EMPTY_STRING__";   // constant.rs:19
constexpr const char* kEqString = "__L@KE$OUL_EQ__";               // constant.rs:20
constexpr const char* kTbdPartitionDesc = "-4";                    // constant.rs:24
constexpr const char* kNoPartitionDesc = "-5";                     // helpers/mod.rs:479-480

double now_ms() {
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

std::string strip_scheme(const std::string& path) {
  if (path.rfind("file://", 0) == 0) return path.substr(7);
  if (path.rfind("file:", 0) == 0) return path.substr(5);
  if (path.find("://") != std::string::npos)
    fail("object store '" + path.substr(0, path.find("://")) + "' is not supported by the GPU writer (local files only)");
  return path;
}

std::string random_str(size_t n) {  // utils/mod.rs:17-23
  static const char kAlnum[] = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789";
  static thread_local std::mt19937_64 rng(std::random_device{}() ^ (uint64_t)getpid() ^
                                          (uint64_t)std::chrono::steady_clock::now().time_since_epoch().count());
  std::string s(n, 'a');
  for (auto& c : s) c = kAlnum[rng() % 62];
  return s;
}

std::shared_ptr<arrow::Field> uniform_field(const std::shared_ptr<arrow::Field>& f) {  // helpers/transform.rs:35-55
  if (f->type()->id() == arrow::Type::TIMESTAMP) {
    const auto& t = static_cast<const arrow::TimestampType&>(*f->type());
    if (!t.timezone().empty()) return f->WithType(arrow::timestamp(t.unit(), "UTC"));
  }
  return f;
}

std::string civil_from_days(int64_t z, const char* sep_time, int64_t secs_of_day, int frac_digits, int64_t frac) {
  // days since 1970-01-01 -> proleptic Gregorian date (the algorithm chrono uses: era / day-of-era arithmetic)
  z += 719468;
  const int64_t era = (z >= 0 ? z : z - 146096) / 146097;
  const int64_t doe = z - era * 146097;
  const int64_t yoe = (doe - doe / 1460 + doe / 36524 - doe / 146096) / 365;
  int64_t y = yoe + era * 400;
  const int64_t doy = doe - (365 * yoe + yoe / 4 - yoe / 100);
  const int64_t mp = (5 * doy + 2) / 153;
  const int64_t d = doy - (153 * mp + 2) / 5 + 1;
  const int64_t m = mp + (mp < 10 ? 3 : -9);
  if (m <= 2) ++y;
  char buf[96];
  if (!sep_time) {
    snprintf(buf, sizeof buf, "%04lld-%02lld-%02lld", (long long)y, (long long)m, (long long)d);
    return buf;
  }
  int n = snprintf(buf, sizeof buf, "%04lld-%02lld-%02lld%s%02lld:%02lld:%02lld", (long long)y, (long long)m, (long long)d, sep_time,
                   (long long)(secs_of_day / 3600), (long long)(secs_of_day / 60 % 60), (long long)(secs_of_day % 60));
  if (frac_digits > 0) snprintf(buf + n, sizeof buf - n, ".%0*lld", frac_digits, (long long)frac);
  return buf;
}

std::string format_timestamp(int64_t v, int64_t per_sec, int frac_digits) {
  int64_t secs = v / per_sec, frac = v % per_sec;
  if (frac < 0) { frac += per_sec; --secs; }
  int64_t days = secs / 86400, sod = secs % 86400;
  if (sod < 0) { sod += 86400; --days; }
  return civil_from_days(days, "T", sod, frac_digits, frac);
}

}  // namespace

std::string format_partition_value(const arrow::Array& col, int64_t row) {
  using T = arrow::Type;
  if (col.IsNull(row)) {
    switch (col.type_id()) {  // helpers/mod.rs:213,263-268: these nulls print the null marker; the others fall through to Display
      case T::DATE32: case T::NA: case T::BINARY: case T::LARGE_BINARY: case T::FIXED_SIZE_BINARY: return kNullString;
      default: return "NULL";  // ScalarValue's Display of a null
    }
  }
  switch (col.type_id()) {
    case T::DATE32: return civil_from_days(static_cast<const arrow::Date32Array&>(col).Value(row), nullptr, 0, 0, 0);
    case T::STRING: {
      const std::string s = static_cast<const arrow::StringArray&>(col).GetString(row);
      if (s.empty()) return kEmptyString;
      std::string out;  // helpers/mod.rs:218: ',' -> the EQ marker (the reference's second replace never matches)
      for (char c : s) {
        if (c == ',') out += kEqString;
        else out += c;
      }
      return out;
    }
    case T::TIMESTAMP: {
      const auto unit = static_cast<const arrow::TimestampType&>(*col.type()).unit();
      const int64_t v = static_cast<const arrow::TimestampArray&>(col).Value(row);
      switch (unit) {
        case arrow::TimeUnit::SECOND: return format_timestamp(v, 1, 0);
        case arrow::TimeUnit::MILLI: return format_timestamp(v, 1000, 3);
        case arrow::TimeUnit::MICRO: return format_timestamp(v, 1000000, 6);
        default: return format_timestamp(v, 1000000000, 9);
      }
    }
    case T::BINARY: case T::LARGE_BINARY: case T::FIXED_SIZE_BINARY: {
      auto sc = col.GetScalar(row).ValueOrDie();
      const auto& b = static_cast<const arrow::BaseBinaryScalar&>(*sc).value;
      static const char* hex = "0123456789abcdef";
      std::string out;
      for (int64_t i = 0; i < b->size(); ++i) { out += hex[b->data()[i] >> 4]; out += hex[b->data()[i] & 15]; }
      return out;
    }
    case T::DECIMAL128: {  // the unscaled integer (helpers/mod.rs:261)
      const auto& a = static_cast<const arrow::Decimal128Array&>(col);
      return arrow::Decimal128(a.GetValue(row)).ToIntegerString();
    }
    case T::DECIMAL256: {
      const auto& a = static_cast<const arrow::Decimal256Array&>(col);
      return arrow::Decimal256(a.GetValue(row)).ToIntegerString();
    }
    case T::BOOL: return static_cast<const arrow::BooleanArray&>(col).Value(row) ? "true" : "false";
    default: {
      auto sc = col.GetScalar(row);
      if (!sc.ok()) fail("cannot format a range partition value of type " + col.type()->ToString());
      return (*sc)->ToString();
    }
  }
}

struct Writer::Impl {
  enum class Mode { Plain, Sort, Partition };
  IOConfig cfg;
  Mode mode = Mode::Plain;
  std::shared_ptr<arrow::Schema> in_schema;    // what Write() takes
  std::shared_ptr<arrow::Schema> file_schema;  // in_schema minus range partitions minus aux sort columns
  std::vector<int> file_cols, range_cols;
  std::string prefix;
  std::string write_id;
  uint32_t bucket_num = 1;
  int64_t budget = (int64_t)1 << 30;

  std::vector<std::shared_ptr<arrow::RecordBatch>> buffered;
  int64_t buffered_bytes = 0;
  std::vector<FlushOutput> outs;
  std::vector<std::string> created;
  WriterStats st;
  bool closed = false;

  struct Open {
    std::unique_ptr<parquet::arrow::FileWriter> w;
    std::shared_ptr<arrow::io::FileOutputStream> out;
    std::string path, desc;
    int64_t rows = 0;
  } open;

  std::shared_ptr<parquet::WriterProperties> props;
  std::shared_ptr<parquet::ArrowWriterProperties> arrow_props;

  void make_props() {
    // multipart_writer.rs:111-133: row group length, write batch size, ZSTD (default level), dictionaries off
    const size_t ncol = std::max<size_t>(1, (size_t)in_schema->num_fields());
    size_t max_rg = cfg.max_row_group_size;
    if (cfg.max_row_group_size * ncol > cfg.max_row_group_num_values) max_rg = std::max(cfg.batch_size, cfg.max_row_group_num_values / ncol);
    parquet::WriterProperties::Builder b;
    b.max_row_group_length((int64_t)std::max<size_t>(max_rg, 1));
    b.write_batch_size((int64_t)std::max<size_t>(cfg.batch_size, 1));
    b.compression(parquet::Compression::ZSTD);
    b.compression_level(3);  // zstd crate's default level
    b.disable_dictionary();
    b.created_by("lakesoul_b200 (sm_100a) via parquet-cpp-arrow");
    props = b.build();
    parquet::ArrowWriterProperties::Builder ab;
    ab.store_schema();  // parquet-rs ArrowWriter embeds ARROW:schema as well
    arrow_props = ab.build();
  }

  void open_file(const std::string& path, const std::string& desc) {
    close_file();
    std::error_code ec;
    const auto parent = std::filesystem::path(path).parent_path();
    if (!parent.empty()) std::filesystem::create_directories(parent, ec);
    auto o = arrow::io::FileOutputStream::Open(path);
    if (!o.ok()) fail("cannot create '" + path + "': " + o.status().ToString());
    open.out = *o;
    created.push_back(path);
    auto w = parquet::arrow::FileWriter::Open(*file_schema, arrow::default_memory_pool(), open.out, props, arrow_props);
    if (!w.ok()) fail("parquet writer for '" + path + "': " + w.status().ToString());
    open.w = std::move(*w);
    open.path = path;
    open.desc = desc;
    open.rows = 0;
  }

  void close_file() {
    if (!open.w) return;
    auto s = open.w->Close();
    if (!s.ok()) fail("closing '" + open.path + "': " + s.ToString());
    s = open.out->Close();
    if (!s.ok()) fail("closing '" + open.path + "': " + s.ToString());
    open.w.reset();
    open.out.reset();
    struct stat sb {};
    if (stat(open.path.c_str(), &sb) != 0) fail("cannot stat '" + open.path + "'");
    FlushOutput fo;
    fo.partition_desc = open.desc;
    fo.file_path = open.path;
    fo.file_size = (int64_t)sb.st_size;
    for (int i = 0; i < file_schema->num_fields(); ++i) {
      if (i) fo.file_exist_cols += ",";
      fo.file_exist_cols += file_schema->field(i)->name();
    }
    outs.push_back(std::move(fo));
    st.files += 1;
  }

  void write_slice(const std::shared_ptr<arrow::RecordBatch>& batch, int64_t start, int64_t len) {
    if (len <= 0) return;
    std::vector<std::shared_ptr<arrow::Array>> arrays;
    for (int c : file_cols) arrays.push_back(batch->column(c)->Slice(start, len));
    auto rb = arrow::RecordBatch::Make(file_schema, len, std::move(arrays));
    const double t0 = now_ms();
    auto s = open.w->WriteRecordBatch(*rb);
    if (!s.ok()) fail("writing '" + open.path + "': " + s.ToString());
    st.encode_ms += now_ms() - t0;
    open.rows += len;
    st.rows_out += len;
  }

  std::string single_file_path() const {
    if (!cfg.files.empty()) return strip_scheme(cfg.files[0]);
    if (prefix.empty()) fail("writer needs a file name or a prefix");
    std::string id;
    int bucket = 0;
    if (cfg.option("hash_bucket_id", &id) && !id.empty()) bucket = atoi(id.c_str());
    char buf[32];
    snprintf(buf, sizeof buf, "%04d", bucket);
    return prefix + "/part-" + random_str(16) + "_" + buf + ".parquet";  // writer/mod.rs:95-100
  }

  // ---- the sorting modes: one device pass over everything buffered --------------------------------------------
  bool same_group(const arrow::RecordBatch& a, int64_t ra, const arrow::RecordBatch& b, int64_t rb, const uint32_t* buckets, int64_t ua,
                  int64_t ub) const {
    if (buckets && buckets[ua] != buckets[ub]) return false;
    for (int c : range_cols)
      if (!a.column(c)->RangeEquals(ra, ra + 1, rb, *b.column(c))) return false;
    return true;
  }

  void flush_cycle() {
    if (buffered.empty()) return;
    st.flush_cycles += 1;
    const double t0 = now_ms();
    auto table = arrow::Table::FromRecordBatches(in_schema, buffered);
    if (!table.ok()) fail("buffered batches: " + table.status().ToString());
    buffered.clear();
    buffered_bytes = 0;
    ScanSpec spec;
    spec.physical_schema = in_schema;
    spec.merge = false;
    spec.batch_size = (int)std::max<size_t>(cfg.batch_size, 1);
    spec.sort.enabled = true;
    if (mode == Mode::Partition) {
      spec.sort.prefix_cols = cfg.range_partitions;
      spec.sort.with_buckets = true;
      spec.sort.bucket_num = bucket_num;
    }
    spec.sort.hash_cols = cfg.primary_keys;
    spec.sort.suffix_cols = cfg.aux_sort_cols;
    std::vector<BucketFiles> groups(1);
    FileInput fi;
    fi.table = *table;
    groups[0].files.push_back(std::move(fi));
    DeviceScan scan(std::move(spec), std::move(groups));
    scan.Run();
    st.sort_ms += now_ms() - t0;
    const auto& ctr = scan.counters();
    st.h2d_bytes += ctr.h2d_bytes;
    st.d2h_bytes += ctr.d2h_bytes;
    st.kernel_launches += ctr.kernel_launches;
    const uint32_t* buckets = scan.sorted_buckets();
    const int64_t nbatches = scan.num_batches(0);
    const int64_t bs = std::max<size_t>(cfg.batch_size, 1);
    if (mode == Mode::Sort) {
      if (!open.w) open_file(single_file_path(), kTbdPartitionDesc);
      for (int64_t i = 0; i < nbatches; ++i) {
        auto rb = scan.Batch(0, i);
        write_slice(rb, 0, rb->num_rows());
      }
      return;
    }
    // Partition mode: the rows arrive grouped by (range values, bucket); one file is open at a time
    std::shared_ptr<arrow::RecordBatch> group_batch;  // batch and row of the open group's first row
    int64_t group_row = 0, group_u = 0;
    for (int64_t i = 0; i < nbatches; ++i) {
      auto rb = scan.Batch(0, i);
      const int64_t n = rb->num_rows(), u0 = i * bs;
      int64_t r = 0;
      while (r < n) {
        const bool cont = group_batch && same_group(*group_batch, group_row, *rb, r, buckets, group_u, u0 + r);
        if (!cont) {
          // new (partition, bucket): partitioning_writer.rs:276-291
          std::string desc, sub = "/";
          for (size_t k = 0; k < range_cols.size(); ++k) {
            const std::string kv = in_schema->field(range_cols[k])->name() + "=" + format_partition_value(*rb->column(range_cols[k]), r);
            desc += (k ? "," : "") + kv;
            sub += kv + "/";
          }
          if (range_cols.empty()) desc = kNoPartitionDesc;
          char buf[32];
          snprintf(buf, sizeof buf, "%04u", buckets ? buckets[u0 + r] : 0u);
          open_file(prefix + sub + "part-" + write_id + "_" + buf + ".parquet", desc);
          group_batch = rb;
          group_row = r;
          group_u = u0 + r;
        }
        // rows of this batch that still belong to the group: the predicate is monotone (sorted input) -> bisection
        int64_t lo = r, hi = n;  // rows [r, lo] belong; find the first that does not
        if (same_group(*group_batch, group_row, *rb, n - 1, buckets, group_u, u0 + n - 1)) {
          lo = n;
        } else {
          while (hi - lo > 1) {
            const int64_t mid = (lo + hi) / 2;
            if (same_group(*group_batch, group_row, *rb, mid, buckets, group_u, u0 + mid)) lo = mid; else hi = mid;
          }
          lo = hi;
        }
        write_slice(rb, r, lo - r);
        r = lo;
      }
    }
    close_file();  // a later cycle starts new files (new write id)
    write_id = random_str(16);
  }
};

Writer::Writer(IOConfig cfg) : impl_(new Impl) {
  Impl& I = *impl_;
  I.cfg = std::move(cfg);
  std::vector<std::shared_ptr<arrow::Field>> f;
  for (const auto& x : I.cfg.target_schema->fields()) f.push_back(uniform_field(x));
  I.in_schema = arrow::schema(f);
  if (I.in_schema->num_fields() == 0) fail("writer needs a schema");
  I.prefix = strip_scheme(I.cfg.prefix);
  while (I.prefix.size() > 1 && I.prefix.back() == '/') I.prefix.pop_back();
  std::string v;
  const bool keep_order = I.cfg.option("keep_orders", &v) && v == "true";
  // writer/mod.rs:84-125
  if (I.cfg.use_dynamic_partition) I.mode = Impl::Mode::Partition;
  else if (!I.cfg.primary_keys.empty() && !keep_order) I.mode = Impl::Mode::Sort;
  else I.mode = Impl::Mode::Plain;
  auto require = [&](const std::vector<std::string>& names, const char* what) {
    for (const auto& n : names)
      if (I.in_schema->GetFieldIndex(n) < 0) fail(std::string(what) + " column '" + n + "' is not in the schema");
  };
  require(I.cfg.primary_keys, "primary key");
  require(I.cfg.range_partitions, "range partition");
  require(I.cfg.aux_sort_cols, "aux sort");
  for (int i = 0; i < I.in_schema->num_fields(); ++i) {
    const std::string& n = I.in_schema->field(i)->name();
    const bool is_range = std::find(I.cfg.range_partitions.begin(), I.cfg.range_partitions.end(), n) != I.cfg.range_partitions.end();
    const bool is_aux = std::find(I.cfg.aux_sort_cols.begin(), I.cfg.aux_sort_cols.end(), n) != I.cfg.aux_sort_cols.end();
    if (!is_range && !is_aux) I.file_cols.push_back(i);
  }
  for (const auto& n : I.cfg.range_partitions) I.range_cols.push_back(I.in_schema->GetFieldIndex(n));
  std::vector<std::shared_ptr<arrow::Field>> ff;
  for (int c : I.file_cols) ff.push_back(I.in_schema->field(c));
  I.file_schema = arrow::schema(ff);
  if (I.mode == Impl::Mode::Partition && I.prefix.empty()) fail("dynamic partition writer needs a prefix");
  const long long bn = atoll(I.cfg.hash_bucket_num.c_str());  // config/mod.rs:217-221
  I.bucket_num = (uint32_t)std::max<long long>(bn, 1);
  if (I.cfg.primary_keys.empty()) I.bucket_num = std::max<uint32_t>(I.bucket_num, 1);
  // flush budget: everything buffered goes through the GPU in one pass (32-bit string offsets: < 2 GiB per column)
  if (I.cfg.option("mem_limit", &v) && !v.empty()) I.budget = std::max<int64_t>((int64_t)1 << 20, atoll(v.c_str()));
  if (I.cfg.option("max_file_size", &v) && !v.empty()) I.budget = std::max<int64_t>((int64_t)1 << 20, atoll(v.c_str()));
  if (const char* e = getenv("LAKESOUL_B200_WRITER_BUFFER_BYTES")) I.budget = std::max<int64_t>((int64_t)1 << 16, atoll(e));
  I.budget = std::min<int64_t>(I.budget, (int64_t)3 << 29);
  I.write_id = random_str(16);
  I.make_props();
  if (I.mode != Impl::Mode::Plain) ensure_device();  // no CPU fallback for the sort
  if (I.mode == Impl::Mode::Plain) I.open_file(I.single_file_path(), kTbdPartitionDesc);
}

Writer::~Writer() {
  if (impl_ && !impl_->closed) {
    try { AbortAndClose(); } catch (...) {}
  }
}

std::shared_ptr<arrow::Schema> Writer::schema() const { return impl_->in_schema; }
const std::vector<FlushOutput>& Writer::outputs() const { return impl_->outs; }
const WriterStats& Writer::stats() const { return impl_->st; }

void Writer::Write(const std::shared_ptr<arrow::RecordBatch>& batch_in) {
  Impl& I = *impl_;
  if (I.closed) fail("Invalid state of inner writer");  // writer/mod.rs:328
  if (!batch_in) fail("null record batch");
  auto batch = batch_in;
  if (batch->num_columns() != I.in_schema->num_fields()) fail("record batch does not match the writer schema");
  // columns by position, like the reference (the batch schema must be the target schema); time zones are normalised
  std::vector<std::shared_ptr<arrow::Array>> cols;
  for (int i = 0; i < batch->num_columns(); ++i) {
    auto col = batch->column(i);
    const auto& want = I.in_schema->field(i)->type();
    if (!col->type()->Equals(*want)) {
      if (col->type()->id() == want->id() && want->id() == arrow::Type::TIMESTAMP &&
          static_cast<const arrow::TimestampType&>(*col->type()).unit() == static_cast<const arrow::TimestampType&>(*want).unit()) {
        auto d = col->data()->Copy();
        d->type = want;
        col = arrow::MakeArray(d);
      } else {
        fail("column '" + I.in_schema->field(i)->name() + "' has type " + col->type()->ToString() + ", the writer schema says " +
             want->ToString());
      }
    }
    cols.push_back(col);
  }
  batch = arrow::RecordBatch::Make(I.in_schema, batch->num_rows(), std::move(cols));
  I.st.rows_in += batch->num_rows();
  if (batch->num_rows() == 0) return;
  if (I.mode == Impl::Mode::Plain) {
    I.write_slice(batch, 0, batch->num_rows());
    return;
  }
  int64_t bytes = 0;
  for (const auto& c : batch->columns())
    for (const auto& b : c->data()->buffers)
      if (b) bytes += b->size();
  I.buffered.push_back(batch);
  I.buffered_bytes += bytes;
  if (I.buffered_bytes >= I.budget) I.flush_cycle();
}

std::string Writer::FlushAndClose() {
  Impl& I = *impl_;
  if (I.closed) return std::string();
  try {
    if (I.mode != Impl::Mode::Plain) I.flush_cycle();
    if (I.mode == Impl::Mode::Sort && !I.open.w && I.outs.empty()) I.open_file(I.single_file_path(), kTbdPartitionDesc);  // an empty file, like the reference
    I.close_file();
  } catch (...) {
    I.closed = true;
    for (const auto& p : I.created) unlink(p.c_str());
    throw;
  }
  I.closed = true;
  if (const char* tr = getenv("LAKESOUL_B200_TRACE"); tr && *tr == '1')
    fprintf(stderr, "LSBTRACE writer rows_in=%lld rows_out=%lld files=%lld flush_cycles=%lld device_pass_ms=%.2f encode_ms=%.2f h2d=%lld d2h=%lld launches=%d\n",
            (long long)I.st.rows_in, (long long)I.st.rows_out, (long long)I.st.files, (long long)I.st.flush_cycles, I.st.sort_ms, I.st.encode_ms,
            (long long)I.st.h2d_bytes, (long long)I.st.d2h_bytes, I.st.kernel_launches);
  // writer/mod.rs:346-378 (insertion order of the partitions instead of a hash map's)
  std::vector<std::string> order;
  std::vector<std::vector<std::string>> files;
  for (const auto& o : I.outs) {
    const std::string enc = o.file_path + "\x03" + std::to_string(o.file_size) + "\x03" + o.file_exist_cols;
    size_t k = 0;
    for (; k < order.size(); ++k)
      if (order[k] == o.partition_desc) break;
    if (k == order.size()) { order.push_back(o.partition_desc); files.emplace_back(); }
    files[k].push_back(enc);
  }
  std::string summary = std::to_string(order.size());
  for (size_t k = 0; k < order.size(); ++k) {
    summary += "\x01";
    summary += order[k];
    for (const auto& f : files[k]) { summary += "\x02"; summary += f; }
  }
  return summary;
}

void Writer::AbortAndClose() {
  Impl& I = *impl_;
  if (I.closed) return;
  I.closed = true;
  I.buffered.clear();
  if (I.open.w) {
    (void)I.open.w->Close();
    (void)I.open.out->Close();
    I.open.w.reset();
    I.open.out.reset();
  }
  for (const auto& p : I.created) unlink(p.c_str());
  I.outs.clear();
}

}  // namespace lsb
