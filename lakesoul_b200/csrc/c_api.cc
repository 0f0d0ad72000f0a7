// SPDX-License-Identifier: Apache-2.0
//
// c_api.cc -- extern "C" surface of liblakesoul_io_c.so: the reference's lakesoul-io-c ABI
// (include/lakesoul_c_bindings.h, rust/lakesoul-io-c/src/lib.rs:164-1508) plus the B200
// extensions (include/lakesoul_b200_ext.h).
#include <arrow/c/bridge.h>
#include <arrow/io/memory.h>
#include <arrow/ipc/reader.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>

#pragma GCC visibility push(default)
#include "lakesoul_b200_ext.h"
#pragma GCC visibility pop
#include "mor_engine.h"
#include "reader.h"
#include "writer.h"

namespace {

char* dup_cstr(const std::string& s) {
  char* p = static_cast<char*>(malloc(s.size() + 1));
  memcpy(p, s.c_str(), s.size() + 1);
  return p;
}

CStatus* new_status(int32_t status, const std::string* err = nullptr) {
  auto* st = static_cast<CStatus*>(malloc(sizeof(CStatus)));
  st->status = status;
  st->err = err ? dup_cstr(*err) : nullptr;
  return st;
}

struct BuilderImpl { lsb::IOConfig cfg; };
struct ConfigImpl { lsb::IOConfig cfg; };
struct RuntimeBuilderImpl { size_t threads = 2; };
struct RuntimeImpl { size_t threads = 2; };
struct ReaderImpl {
  std::unique_ptr<lsb::Reader> reader;
  std::mutex mu;  // one reader is used by one thread at a time (reader.rs:278-282)
};

BuilderImpl* B(IOConfigBuilder* b) { return reinterpret_cast<BuilderImpl*>(b); }
IOConfigBuilder* O(BuilderImpl* b) { return reinterpret_cast<IOConfigBuilder*>(b); }
ReaderImpl* R(CResult_Reader* r) { return r ? reinterpret_cast<ReaderImpl*>(r->ptr) : nullptr; }
struct WriterImpl {
  std::unique_ptr<lsb::Writer> writer;
  std::mutex mu;
};
WriterImpl* W(CResult_Writer* w) { return w ? reinterpret_cast<WriterImpl*>(w->ptr) : nullptr; }

std::shared_ptr<arrow::Schema> import_schema(c_ptrdiff_t addr) {
  // takes ownership like FFI_ArrowSchema::from_raw (lib.rs:295): ImportSchema releases it
  auto r = arrow::ImportSchema(reinterpret_cast<struct ArrowSchema*>(addr));
  if (!r.ok()) {
    // the reference panics on a bad schema (unwrap, lib.rs:296)
    fprintf(stderr, "lakesoul_b200: invalid Arrow schema: %s\n", r.status().ToString().c_str());
    abort();
  }
  return *r;
}

thread_local std::string g_ext_error;

template <typename F>
int32_t guard(F&& f) {
  try {
    f();
    return 0;
  } catch (const std::exception& e) {
    g_ext_error = e.what();
    return -1;
  }
}

// returns rows (>0), 0 at end, or throws
int32_t next_into(ReaderImpl* ri, c_ptrdiff_t array_addr, c_ptrdiff_t schema_addr) {
  std::lock_guard<std::mutex> g(ri->mu);
  auto batch = ri->reader->Next();
  if (!batch) return 0;
  auto* out_array = reinterpret_cast<struct ArrowArray*>(array_addr);
  auto* out_schema = schema_addr ? reinterpret_cast<struct ArrowSchema*>(schema_addr) : nullptr;
  auto st = arrow::ExportRecordBatch(*batch, out_array, out_schema);
  if (!st.ok()) throw std::runtime_error(st.ToString());
  return (int32_t)batch->num_rows();
}

}  // namespace

extern "C" {

// ---- builder -------------------------------------------------------------------------------
IOConfigBuilder* new_lakesoul_io_config_builder(void) { return O(new BuilderImpl()); }

#define BUILDER_STR(fn, stmt)                                    \
  IOConfigBuilder* fn(IOConfigBuilder* builder, const char* v) { \
    auto* b = B(builder);                                        \
    std::string s(v);                                            \
    stmt;                                                        \
    return builder;                                              \
  }

BUILDER_STR(lakesoul_config_builder_with_prefix, b->cfg.prefix = s)
BUILDER_STR(lakesoul_config_builder_add_single_file, b->cfg.files.push_back(s))
BUILDER_STR(lakesoul_config_builder_add_single_column, (void)b; (void)s)
BUILDER_STR(lakesoul_config_builder_add_single_aux_sort_column, b->cfg.aux_sort_cols.push_back(s))
BUILDER_STR(lakesoul_config_builder_add_filter, b->cfg.filter_strs.push_back(s))
BUILDER_STR(lakesoul_config_builder_add_single_primary_key, b->cfg.primary_keys.push_back(s))
BUILDER_STR(lakesoul_config_builder_add_single_range_partition, b->cfg.range_partitions.push_back(s))

IOConfigBuilder* lakesoul_config_builder_add_filter_proto(IOConfigBuilder* builder, c_ptrdiff_t, int32_t) {
  // Substrait plans cannot be decoded here; the reader refuses to start rather than ignore a
  // pushed filter.
  B(builder)->cfg.filter_protos += 1;
  return builder;
}
IOConfigBuilder* lakesoul_config_builder_set_schema(IOConfigBuilder* builder, c_ptrdiff_t schema_addr) {
  B(builder)->cfg.target_schema = import_schema(schema_addr);
  return builder;
}
IOConfigBuilder* lakesoul_config_builder_set_partition_schema(IOConfigBuilder* builder, c_ptrdiff_t schema_addr) {
  B(builder)->cfg.partition_schema = import_schema(schema_addr);
  return builder;
}
IOConfigBuilder* lakesoul_config_builder_set_thread_num(IOConfigBuilder* builder, c_size_t n) {
  B(builder)->cfg.thread_num = n;
  return builder;
}
IOConfigBuilder* lakesoul_config_builder_set_dynamic_partition(IOConfigBuilder* builder, bool enable) {
  B(builder)->cfg.use_dynamic_partition = enable;
  return builder;
}
IOConfigBuilder* lakesoul_config_builder_set_inferring_schema(IOConfigBuilder* builder, bool enable) {
  B(builder)->cfg.inferring_schema = enable;
  return builder;
}
IOConfigBuilder* lakesoul_config_builder_set_batch_size(IOConfigBuilder* builder, c_size_t n) {
  B(builder)->cfg.batch_size = n;
  return builder;
}
IOConfigBuilder* lakesoul_config_builder_set_max_row_group_size(IOConfigBuilder* builder, c_size_t n) {
  B(builder)->cfg.max_row_group_size = n;
  return builder;
}
IOConfigBuilder* lakesoul_config_builder_set_max_row_group_num_values(IOConfigBuilder* builder, c_size_t n) {
  B(builder)->cfg.max_row_group_num_values = n;
  return builder;
}
IOConfigBuilder* lakesoul_config_builder_set_buffer_size(IOConfigBuilder* builder, c_size_t n) {
  B(builder)->cfg.prefetch_size = n;
  return builder;
}
IOConfigBuilder* lakesoul_config_builder_set_hash_bucket_num(IOConfigBuilder* builder, c_size_t n) {
  B(builder)->cfg.hash_bucket_num = std::to_string(n);
  return builder;
}
IOConfigBuilder* lakesoul_config_builder_set_object_store_option(IOConfigBuilder* builder, const char* key,
                                                                 const char* value) {
  B(builder)->cfg.object_store_options[key] = value;
  return builder;
}
IOConfigBuilder* lakesoul_config_builder_set_option(IOConfigBuilder* builder, const char* key, const char* value) {
  B(builder)->cfg.options[key] = value;
  return builder;
}
IOConfigBuilder* lakesoul_config_builder_add_files(IOConfigBuilder* builder, const char* const* files,
                                                   c_size_t file_num) {
  for (c_size_t i = 0; i < file_num; ++i) B(builder)->cfg.files.emplace_back(files[i]);
  return builder;
}
IOConfigBuilder* lakesoul_config_builder_add_merge_op(IOConfigBuilder* builder, const char* field,
                                                      const char* merge_op) {
  B(builder)->cfg.merge_operators[field] = merge_op;
  return builder;
}
IOConfigBuilder* lakesoul_config_builder_add_primary_keys(IOConfigBuilder* builder, const char* const* pks,
                                                          c_size_t pk_num) {
  for (c_size_t i = 0; i < pk_num; ++i) B(builder)->cfg.primary_keys.emplace_back(pks[i]);
  return builder;
}
IOConfigBuilder* lakesoul_config_builder_set_default_column_value(IOConfigBuilder* builder, const char* field,
                                                                  const char* value) {
  B(builder)->cfg.default_column_value[field] = value;
  return builder;
}
IOConfig* create_lakesoul_io_config_from_builder(IOConfigBuilder* builder) {
  auto* b = B(builder);
  auto* c = new ConfigImpl{std::move(b->cfg)};
  delete b;
  return reinterpret_cast<IOConfig*>(c);
}

// ---- runtime shims ---------------------------------------------------------------------------
TokioRuntimeBuilder* new_tokio_runtime_builder(void) {
  return reinterpret_cast<TokioRuntimeBuilder*>(new RuntimeBuilderImpl());
}
TokioRuntimeBuilder* tokio_runtime_builder_set_thread_num(TokioRuntimeBuilder* builder, c_size_t n) {
  reinterpret_cast<RuntimeBuilderImpl*>(builder)->threads = n;
  return builder;
}
TokioRuntime* create_tokio_runtime_from_builder(TokioRuntimeBuilder* builder) {
  auto* rb = reinterpret_cast<RuntimeBuilderImpl*>(builder);
  auto* rt = new RuntimeImpl{rb->threads};
  delete rb;
  return reinterpret_cast<TokioRuntime*>(rt);
}

// ---- reader ----------------------------------------------------------------------------------
static CResult_Reader* make_reader(IOConfig* config) {
  auto* c = reinterpret_cast<ConfigImpl*>(config);
  auto* res = static_cast<CResult_Reader*>(malloc(sizeof(CResult_Reader)));
  res->ptr = nullptr;
  res->err = nullptr;
  try {
    auto* ri = new ReaderImpl();
    ri->reader.reset(new lsb::Reader(std::move(c->cfg)));
    res->ptr = reinterpret_cast<Reader*>(ri);
  } catch (const std::exception& e) {
    res->err = dup_cstr(e.what());
  }
  delete c;
  return res;
}
CResult_Reader* create_lakesoul_reader_from_config(IOConfig* config, TokioRuntime* runtime) {
  delete reinterpret_cast<RuntimeImpl*>(runtime);
  return make_reader(config);
}
CResult_Reader* create_lakesoul_reader_from_config_with_global_runtime(IOConfig* config) {
  return make_reader(config);
}
const char* check_reader_created(CResult_Reader* reader) { return reader->err; }

CStatus* start_reader(CResult_Reader* reader) {
  try {
    auto* ri = R(reader);
    if (!ri) throw std::runtime_error(reader->err ? reader->err : "reader was not created");
    std::lock_guard<std::mutex> g(ri->mu);
    ri->reader->Start();
    return new_status(0);
  } catch (const std::exception& e) {
    std::string m = e.what();
    return new_status(-1, &m);
  }
}
void start_reader_with_data(CResult_Reader* reader, const void* data, DataResultCallback callback) {
  CStatus* st = start_reader(reader);
  callback(st->status == 0, st->err, data);
  free_c_status(st);
}
CStatus* next_record_batch_blocked(CResult_Reader* reader, c_ptrdiff_t array_addr) {
  try {
    auto* ri = R(reader);
    if (!ri) throw std::runtime_error("reader was not created");
    return new_status(next_into(ri, array_addr, 0));
  } catch (const std::exception& e) {
    std::string m = e.what();
    return new_status(-1, &m);
  }
}
void next_record_batch(CResult_Reader* reader, c_ptrdiff_t schema_addr, c_ptrdiff_t array_addr,
                       I32ResultCallback callback) {
  try {
    auto* ri = R(reader);
    if (!ri) throw std::runtime_error("reader was not created");
    callback(next_into(ri, array_addr, schema_addr), nullptr);
  } catch (const std::exception& e) {
    callback(-1, e.what());  // the library owns `err`: valid only during the callback (lib.rs:731-743)
  }
}
void next_record_batch_with_data(CResult_Reader* reader, c_ptrdiff_t schema_addr, c_ptrdiff_t array_addr,
                                 const void* data, I32DataResultCallback callback) {
  try {
    auto* ri = R(reader);
    if (!ri) throw std::runtime_error("reader was not created");
    callback(next_into(ri, array_addr, schema_addr), nullptr, data);
  } catch (const std::exception& e) {
    callback(-1, e.what(), data);
  }
}
void lakesoul_reader_get_schema(CResult_Reader* reader, c_ptrdiff_t schema_addr) {
  auto* ri = R(reader);
  std::shared_ptr<arrow::Schema> s = ri ? ri->reader->schema() : nullptr;
  if (!s) s = arrow::schema({});
  (void)arrow::ExportSchema(*s, reinterpret_cast<struct ArrowSchema*>(schema_addr));
}

// ---- free ------------------------------------------------------------------------------------
void free_lakesoul_reader(CResult_Reader* reader) {
  if (!reader) return;
  delete reinterpret_cast<ReaderImpl*>(reader->ptr);
  free(const_cast<char*>(reader->err));
  free(reader);
}
void free_lakesoul_writer(CResult_Writer* writer) {
  if (!writer) return;
  delete reinterpret_cast<WriterImpl*>(writer->ptr);
  free(const_cast<char*>(writer->err));
  free(writer);
}
void free_lakesoul_io_config_builder(IOConfigBuilder* builder) { delete B(builder); }
void free_lakesoul_io_config(IOConfig* io_config) { delete reinterpret_cast<ConfigImpl*>(io_config); }
void free_tokio_runtime_builder(TokioRuntimeBuilder* builder) {
  delete reinterpret_cast<RuntimeBuilderImpl*>(builder);
}
void free_tokio_runtime(TokioRuntime* runtime) { delete reinterpret_cast<RuntimeImpl*>(runtime); }
void free_c_status(CStatus* status) {
  if (!status) return;
  free(const_cast<char*>(status->err));
  free(status);
}

// ---- writer side (lib.rs:1081-1377) -----------------------------------------------
CResult_Writer* create_lakesoul_writer_from_config(IOConfig* io_config, TokioRuntime* runtime) {
  auto* c = reinterpret_cast<ConfigImpl*>(io_config);
  delete reinterpret_cast<RuntimeImpl*>(runtime);
  auto* res = static_cast<CResult_Writer*>(malloc(sizeof(CResult_Writer)));
  res->ptr = nullptr;
  res->err = nullptr;
  try {
    auto* wi = new WriterImpl();
    wi->writer.reset(new lsb::Writer(std::move(c->cfg)));
    res->ptr = reinterpret_cast<Writer*>(wi);
  } catch (const std::exception& e) {
    res->err = dup_cstr(e.what());
  }
  delete c;
  return res;
}
CStatus* check_writer_created(CResult_Writer* writer) {
  if (!writer || writer->err || !writer->ptr) {
    std::string m = writer && writer->err ? writer->err : "writer was not created";
    return new_status(-1, &m);
  }
  return new_status(0);
}
// the C structs are moved from, like ptr::replace(.., FFI_ArrowArray::empty()) (lib.rs:1136-1141)
static void write_imported(WriterImpl* wi, c_ptrdiff_t schema_addr, c_ptrdiff_t array_addr) {
  auto batch = arrow::ImportRecordBatch(reinterpret_cast<struct ArrowArray*>(array_addr), reinterpret_cast<struct ArrowSchema*>(schema_addr));
  if (!batch.ok()) throw std::runtime_error(batch.status().ToString());
  std::lock_guard<std::mutex> g(wi->mu);
  wi->writer->Write(*batch);
}
void write_record_batch(CResult_Writer* writer, c_ptrdiff_t schema_addr, c_ptrdiff_t array_addr, ResultCallback callback) {
  try {
    auto* wi = W(writer);
    if (!wi) throw std::runtime_error("writer was not created");
    write_imported(wi, schema_addr, array_addr);
    callback(true, nullptr);
  } catch (const std::exception& e) {
    callback(false, e.what());
  }
}
CStatus* write_record_batch_blocked(CResult_Writer* writer, c_ptrdiff_t schema_addr, c_ptrdiff_t array_addr) {
  try {
    auto* wi = W(writer);
    if (!wi) throw std::runtime_error("writer was not created");
    write_imported(wi, schema_addr, array_addr);
    return new_status(0);
  } catch (const std::exception& e) {
    std::string m = e.what();
    return new_status(-1, &m);
  }
}
// an Arrow IPC stream; status = rows written (lib.rs:1211-1250)
CStatus* write_record_batch_ipc_blocked(CResult_Writer* writer, c_ptrdiff_t ipc_addr, int64_t len) {
  try {
    auto* wi = W(writer);
    if (!wi) throw std::runtime_error("writer was not created");
    auto buf = std::make_shared<arrow::Buffer>(reinterpret_cast<const uint8_t*>(ipc_addr), len);
    auto in = std::make_shared<arrow::io::BufferReader>(buf);
    auto rd = arrow::ipc::RecordBatchStreamReader::Open(in);
    if (!rd.ok()) throw std::runtime_error(rd.status().ToString());
    int64_t rows = 0;
    while (true) {
      std::shared_ptr<arrow::RecordBatch> rb;
      auto st = (*rd)->ReadNext(&rb);
      if (!st.ok()) throw std::runtime_error(st.ToString());
      if (!rb) break;
      // the batch borrows the caller's buffer: copy before buffering
      std::vector<std::shared_ptr<arrow::Array>> cols;
      for (const auto& c : rb->columns()) {
        auto cc = arrow::Concatenate({c});
        if (!cc.ok()) throw std::runtime_error(cc.status().ToString());
        cols.push_back(*cc);
      }
      auto copy = arrow::RecordBatch::Make(rb->schema(), rb->num_rows(), std::move(cols));
      std::lock_guard<std::mutex> g(wi->mu);
      wi->writer->Write(copy);
      rows += rb->num_rows();
    }
    return new_status((int32_t)rows);
  } catch (const std::exception& e) {
    std::string m = e.what();
    return new_status(-1, &m);
  }
}
struct BytesImpl { std::string bytes; };
static CResult_BytesResult* make_bytes(std::string v) {
  auto* res = static_cast<CResult_BytesResult*>(malloc(sizeof(CResult_BytesResult)));
  auto* bi = new BytesImpl();
  bi->bytes = std::move(v);
  res->ptr = reinterpret_cast<BytesResult*>(bi);
  res->err = nullptr;
  return res;
}
void free_bytes_result(CResult_BytesResult* bytes) {
  if (!bytes) return;
  delete reinterpret_cast<BytesImpl*>(bytes->ptr);
  free(const_cast<char*>(bytes->err));
  free(bytes);
}
// copies the bytes (+ a closing 0) to `addr` and frees the result (lib.rs:1262-1300)
void export_bytes_result(void (*callback)(bool, const char*), CResult_BytesResult* bytes, int32_t len, c_ptrdiff_t addr) {
  auto* bi = bytes ? reinterpret_cast<BytesImpl*>(bytes->ptr) : nullptr;
  if (!bi || (int64_t)bi->bytes.size() != (int64_t)len) {
    free_bytes_result(bytes);
    callback(false, "Size of buffer and result mismatch at export_bytes_result.");
    return;
  }
  memcpy(reinterpret_cast<void*>(addr), bi->bytes.c_str(), (size_t)len + 1);
  free_bytes_result(bytes);
  callback(true, nullptr);
}
static CResult_BytesResult* empty_bytes() { return make_bytes(std::string()); }
CResult_BytesResult* flush_and_close_writer(CResult_Writer* writer, I32ResultCallback callback) {
  auto* wi = W(writer);
  std::string out, err;
  bool ok = false;
  if (wi) {
    try {
      std::lock_guard<std::mutex> g(wi->mu);
      out = wi->writer->FlushAndClose();
      ok = true;
    } catch (const std::exception& e) {
      err = e.what();
    }
  } else {
    err = writer && writer->err ? writer->err : "writer was not created";
  }
  free_lakesoul_writer(writer);
  if (!ok) {
    callback(-1, err.c_str());
    return empty_bytes();
  }
  callback((int32_t)out.size(), nullptr);
  return make_bytes(std::move(out));
}
void abort_and_close_writer(CResult_Writer* writer, ResultCallback callback) {
  auto* wi = W(writer);
  std::string err;
  bool ok = true;
  if (wi) {
    try {
      std::lock_guard<std::mutex> g(wi->mu);
      wi->writer->AbortAndClose();
    } catch (const std::exception& e) {
      ok = false;
      err = e.what();
    }
  }
  free_lakesoul_writer(writer);
  callback(ok, ok ? nullptr : err.c_str());
}
CResult_BytesResult* apply_partition_filter(void (*callback)(int32_t, const char*), int32_t, c_ptrdiff_t,
                                            c_ptrdiff_t, int32_t, c_ptrdiff_t) {
  callback(-1, "lakesoul_b200: apply_partition_filter (metadata plane) is not part of the merge-on-read engine");
  return empty_bytes();
}
void rust_logger_init(void) {}

// ---- extensions ------------------------------------------------------------------------------
const char* lsb_last_error(void) { return g_ext_error.c_str(); }
const char* lsb_version(void) { return "lakesoul_b200 0.1 (sm_100a)"; }
int32_t lsb_init_device(void) {
  int dev = -1;
  if (guard([&] { dev = lsb::ensure_device(); }) != 0) return -1;
  return dev;
}
int32_t lsb_reader_get_stats(CResult_Reader* reader, LsbReaderStats* out) {
  auto* ri = R(reader);
  if (!ri) { g_ext_error = "reader was not created"; return -1; }
  const lsb::ReaderStats& s = ri->reader->stats();
  out->rows_in = s.rows_in; out->rows_out = s.rows_out;
  out->h2d_bytes = s.h2d_bytes; out->d2h_bytes = s.d2h_bytes;
  out->kernel_launches = s.kernel_launches; out->stream_passes = s.stream_passes;
  out->decode_ms = s.decode_ms; out->upload_ms = s.upload_ms; out->merge_ms = s.merge_ms;
  out->gather_ms = s.gather_ms; out->download_ms = s.download_ms;
  out->device_decoded_pages = s.device_decoded_pages; out->host_decoded_files = s.host_decoded_files;
  out->pruned_row_groups = s.pruned_row_groups;
  out->pruned_buckets = s.pruned_buckets;
  out->pruned_pages = s.pruned_pages;
  out->device_decompressed_pages = s.device_decompressed_pages;
  return 0;
}

int32_t lsb_writer_get_stats(CResult_Writer* writer, LsbWriterStats* out) {
  auto* wi = W(writer);
  if (!wi) { g_ext_error = "writer was not created"; return -1; }
  const lsb::WriterStats& s = wi->writer->stats();
  out->rows_in = s.rows_in; out->rows_out = s.rows_out; out->files = s.files; out->flush_cycles = s.flush_cycles;
  out->h2d_bytes = s.h2d_bytes; out->d2h_bytes = s.d2h_bytes;
  out->kernel_launches = s.kernel_launches; out->pad = 0;
  out->sort_ms = s.sort_ms; out->encode_ms = s.encode_ms;
  return 0;
}

struct LsbPlan { lsb::MergePlan* p = nullptr; };

LsbPlan* lsb_plan_create(int32_t nb, const int32_t* bucket_file_begin, const int64_t* file_rows,
                         const int64_t* file_row_base, int64_t row_span, int32_t key_bits, int32_t batch_size,
                         void* stream) {
  LsbPlan* h = nullptr;
  if (guard([&] {
        lsb::ensure_device();
        auto* p = lsb::plan_create(nb, bucket_file_begin, file_rows, file_row_base, row_span, key_bits,
                                   batch_size, static_cast<cudaStream_t>(stream));
        h = new LsbPlan{p};
      }) != 0)
    return nullptr;
  return h;
}
void lsb_plan_destroy(LsbPlan* plan) {
  if (!plan) return;
  lsb::plan_destroy(plan->p);
  delete plan;
}
int32_t lsb_plan_merge(LsbPlan* plan, int32_t n_pk, const void* const* pk_values, const void* const* pk_validity,
                       const int32_t* pk_width, const int32_t* pk_kind, void* stream) {
  return guard([&] {
    if (n_pk < 1 || n_pk > lsb::kMaxPkCols) throw std::runtime_error("1..4 primary key columns expected");
    lsb::PkDesc pk{};
    pk.ncols = n_pk;
    int bits = 0;
    for (int i = 0; i < n_pk; ++i) {
      pk.col[i].values = pk_values[i];
      pk.col[i].validity = static_cast<const uint8_t*>(pk_validity ? pk_validity[i] : nullptr);
      pk.col[i].width = (uint8_t)pk_width[i];
      pk.col[i].kind = (uint8_t)pk_kind[i];
      pk.col[i].data = nullptr;
      pk.col[i].bucket_data = nullptr;
      bits += pk_width[i] * 8 + (pk.col[i].validity ? 1 : 0);
    }
    pk.key_bits = bits;
    lsb::plan_merge(plan->p, pk, static_cast<cudaStream_t>(stream));
  });
}
int32_t lsb_plan_identity(LsbPlan* plan, void* stream) {
  return guard([&] { lsb::plan_identity(plan->p, static_cast<cudaStream_t>(stream)); });
}
int32_t lsb_plan_sync_counts(LsbPlan* plan, int64_t* out_base, int64_t* batch_base, void* stream) {
  return guard([&] {
    lsb::plan_sync_counts(plan->p, static_cast<cudaStream_t>(stream));
    const auto& ob = lsb::plan_out_base(plan->p);
    const auto& bb = lsb::plan_batch_base(plan->p);
    if (out_base) memcpy(out_base, ob.data(), ob.size() * sizeof(int64_t));
    if (batch_base) memcpy(batch_base, bb.data(), bb.size() * sizeof(int64_t));
  });
}
const void* lsb_plan_winners(LsbPlan* plan) { return lsb::plan_winners(plan->p); }
int64_t lsb_plan_valid_stride(LsbPlan* plan) { return lsb::plan_valid_stride_bytes(plan->p); }
int32_t lsb_plan_stats(LsbPlan* plan, int64_t* stats) {
  const lsb::MergeStats& s = lsb::plan_stats(plan->p);
  stats[0] = s.n_rows; stats[1] = s.n_out; stats[2] = s.tiles_l0; stats[3] = s.levels;
  stats[4] = s.log_s; stats[5] = s.tile_cap; stats[6] = s.launches; stats[7] = 0;
  return 0;
}
int32_t lsb_plan_gather_fixed(LsbPlan* plan, int32_t ncols, const void* const* src, const void* const* src_valid,
                              void* const* dst, void* const* dst_valid, const int32_t* width, void* stream) {
  return guard([&] {
    std::vector<lsb::FixedCol> cols((size_t)ncols);
    for (int i = 0; i < ncols; ++i) {
      cols[i].src = src[i];
      cols[i].src_valid = static_cast<const uint8_t*>(src_valid ? src_valid[i] : nullptr);
      cols[i].dst = dst[i];
      cols[i].dst_valid = static_cast<uint8_t*>(dst_valid ? dst_valid[i] : nullptr);
      cols[i].win = nullptr;
      cols[i].width = width[i];
    }
    lsb::gather_fixed(plan->p, cols.data(), ncols, static_cast<cudaStream_t>(stream));
  });
}
int32_t lsb_plan_gather_var(LsbPlan* plan, const void* src_offsets, const void* src_data, const void* src_data_base,
                            const void* src_valid, void* dst_offsets, void* dst_data, void* dst_valid,
                            void* batch_data_start, void* stream) {
  return guard([&] {
    lsb::VarCol vc{};
    vc.src_offsets = static_cast<const int32_t*>(src_offsets);
    vc.src_data = static_cast<const uint8_t*>(src_data);
    vc.src_data_base = static_cast<const int64_t*>(src_data_base);
    vc.src_valid = static_cast<const uint8_t*>(src_valid);
    vc.win = nullptr;
    vc.dst_offsets = static_cast<int32_t*>(dst_offsets);
    vc.dst_data = static_cast<uint8_t*>(dst_data);
    vc.dst_valid = static_cast<uint8_t*>(dst_valid);
    vc.batch_data_start = static_cast<int64_t*>(batch_data_start);
    lsb::gather_var(plan->p, vc, static_cast<cudaStream_t>(stream));
  });
}

int32_t lsb_plan_gather_var_multi(LsbPlan* plan, int32_t ncols, const void* const* src_offsets,
                                  const void* const* src_data, const void* const* src_data_base,
                                  const void* const* src_valid, void* const* dst_offsets, void* const* dst_data,
                                  void* const* dst_valid, void* const* batch_data_start, void* stream) {
  return guard([&] {
    std::vector<lsb::VarCol> cols((size_t)ncols);
    for (int i = 0; i < ncols; ++i) {
      lsb::VarCol& vc = cols[i];
      vc.src_offsets = static_cast<const int32_t*>(src_offsets[i]);
      vc.src_data = static_cast<const uint8_t*>(src_data[i]);
      vc.src_data_base = static_cast<const int64_t*>(src_data_base[i]);
      vc.src_valid = static_cast<const uint8_t*>(src_valid ? src_valid[i] : nullptr);
      vc.win = nullptr;
      vc.dst_offsets = static_cast<int32_t*>(dst_offsets[i]);
      vc.dst_data = static_cast<uint8_t*>(dst_data[i]);
      vc.dst_valid = static_cast<uint8_t*>(dst_valid ? dst_valid[i] : nullptr);
      vc.batch_data_start = static_cast<int64_t*>(batch_data_start[i]);
    }
    lsb::gather_var_multi(plan->p, cols.data(), ncols, static_cast<cudaStream_t>(stream));
  });
}

int32_t lsb_plan_gather_all(LsbPlan* plan, int32_t nfixed, const void* const* src, const void* const* src_valid,
                            void* const* dst, void* const* dst_valid, const int32_t* width, int32_t nvars,
                            const void* const* v_src_offsets, const void* const* v_src_data,
                            const void* const* v_src_data_base, const void* const* v_src_valid,
                            void* const* v_dst_offsets, void* const* v_dst_data, void* const* v_dst_valid,
                            void* const* v_batch_data_start, void* stream) {
  return guard([&] {
    std::vector<lsb::FixedCol> fc((size_t)nfixed);
    for (int i = 0; i < nfixed; ++i) {
      fc[i].src = src[i];
      fc[i].src_valid = static_cast<const uint8_t*>(src_valid ? src_valid[i] : nullptr);
      fc[i].dst = dst[i];
      fc[i].dst_valid = static_cast<uint8_t*>(dst_valid ? dst_valid[i] : nullptr);
      fc[i].win = nullptr;
      fc[i].width = width[i];
    }
    std::vector<lsb::VarCol> vc((size_t)nvars);
    for (int i = 0; i < nvars; ++i) {
      vc[i].src_offsets = static_cast<const int32_t*>(v_src_offsets[i]);
      vc[i].src_data = static_cast<const uint8_t*>(v_src_data[i]);
      vc[i].src_data_base = static_cast<const int64_t*>(v_src_data_base[i]);
      vc[i].src_valid = static_cast<const uint8_t*>(v_src_valid ? v_src_valid[i] : nullptr);
      vc[i].win = nullptr;
      vc[i].dst_offsets = static_cast<int32_t*>(v_dst_offsets[i]);
      vc[i].dst_data = static_cast<uint8_t*>(v_dst_data[i]);
      vc[i].dst_valid = static_cast<uint8_t*>(v_dst_valid ? v_dst_valid[i] : nullptr);
      vc[i].batch_data_start = static_cast<int64_t*>(v_batch_data_start[i]);
    }
    lsb::gather_all(plan->p, fc.data(), nfixed, vc.data(), nvars, static_cast<cudaStream_t>(stream));
  });
}

}  // extern "C"
