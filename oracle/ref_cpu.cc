// SPDX-License-Identifier: Apache-2.0
//
// ref_cpu.cc -- CPU port of the reference's merge-on-read reader, structured like the reference.
// TEST / BENCH INFRASTRUCTURE ONLY: it is the "cpu_baseline" (kind "port") and the
// `bench.py --impl reference` arm; nothing under lakesoul_b200/ links or calls it.
//
// The reference's Rust reader cannot be built in this image (no cargo/rustc, ~400 crates), so this
// restates its structure in C++17/20 on Arrow C++ (pyarrow's libarrow/libparquet):
//   * one reader per hash bucket, one merge thread per reader (lakesoul-io-c/src/lib.rs:1385-1391),
//     readers spread over the host cores;
//   * per file: a streaming Parquet decode in batches of `batch_size` rows
//     (merge/mod.rs:66-72 -> DataSourceExec(ParquetSource));
//   * cursor = the PK column buffer for a single primitive PK (stream/mod.rs:176-188) or
//     arrow-row style byte-comparable rows for composite PKs (stream/mod.rs:104-122);
//   * SortKeyBatchRange = run of equal keys inside one batch (sort_key_range.rs:56-113);
//   * loser tree over the k streams ordered by (key, stream_idx) (combiner.rs:680-785,
//     sort_key_range.rs:182-187); the pending output row keeps the last (batch,row) seen
//     (sort_key_range.rs:402-412); at batch_size rows one index vector drives a per-column
//     interleave (combiner.rs:529-605).
// Input (stdin), written by oracle/ref_cpu_driver.py:
//   line 1: <threads> <batch_size> <verify 0|1>
//   line 2: primary keys, tab separated      line 3: columns, tab separated
//   then one line per bucket: file paths, tab separated
// Output: one JSON line {"seconds":..,"threads":..,"rows_in":..,"rows_out":..,"checksum":..}
// verify = 1: every bucket's merged batches are also written as an Arrow IPC file
//   $REF_CPU_VERIFY_DIR/bucket_<index>.arrow (tests/test_ref_cpu.py compares them with mor_oracle.py)
#include <arrow/api.h>
#include <arrow/io/file.h>
#include <arrow/ipc/writer.h>
#include <parquet/arrow/reader.h>
#include <parquet/file_reader.h>

#include <atomic>
#include <chrono>
#include <cstring>
#include <iostream>
#include <mutex>
#include <sstream>
#include <thread>
#include <vector>

namespace {

struct Cursor {  // comparable view of one decoded batch
  const int64_t* i64 = nullptr;       // single int64 PK fast path
  std::vector<uint8_t> rows;          // composite: fixed-width byte-comparable rows
  int row_width = 0;
  int64_t n = 0;
  bool eq(int64_t a, const Cursor& o, int64_t b) const {
    if (i64) return i64[a] == o.i64[b];
    return memcmp(rows.data() + a * row_width, o.rows.data() + b * o.row_width, row_width) == 0;
  }
  int cmp(int64_t a, const Cursor& o, int64_t b) const {
    if (i64) return i64[a] < o.i64[b] ? -1 : (i64[a] > o.i64[b] ? 1 : 0);
    return memcmp(rows.data() + a * row_width, o.rows.data() + b * o.row_width, row_width);
  }
};

struct Stream {
  std::unique_ptr<parquet::arrow::FileReader> fr;
  std::shared_ptr<arrow::RecordBatchReader> rbr;
  std::shared_ptr<arrow::RecordBatch> batch;
  std::shared_ptr<Cursor> cur;
  int64_t begin = 0, end = 0;  // current range
  int batch_slot = -1;         // index into the live-batch table
  bool exhausted = false;
};

void put_be(uint8_t* p, uint64_t v, int bytes) {
  for (int i = 0; i < bytes; ++i) p[i] = (uint8_t)(v >> (8 * (bytes - 1 - i)));
}

std::shared_ptr<Cursor> make_cursor(const arrow::RecordBatch& b, const std::vector<int>& pk_idx) {
  auto c = std::make_shared<Cursor>();
  c->n = b.num_rows();
  if (pk_idx.size() == 1 && b.column(pk_idx[0])->type_id() == arrow::Type::INT64 && b.column(pk_idx[0])->null_count() == 0) {
    c->i64 = static_cast<const arrow::Int64Array&>(*b.column(pk_idx[0])).raw_values();
    return c;
  }
  int w = 0;
  for (int i : pk_idx) w += 1 + b.column(i)->type()->byte_width();
  c->row_width = w;
  c->rows.resize((size_t)w * c->n);
  int off = 0;
  for (int i : pk_idx) {
    const auto& arr = *b.column(i);
    const int bw = arr.type()->byte_width();
    const uint8_t* vals = arr.data()->buffers[1]->data() + arr.offset() * bw;
    for (int64_t r = 0; r < c->n; ++r) {
      uint8_t* p = c->rows.data() + r * w + off;
      if (arr.IsNull(r)) { memset(p, 0, 1 + bw); continue; }
      p[0] = 1;
      uint64_t v = 0;
      memcpy(&v, vals + r * bw, bw);
      v ^= 1ull << (bw * 8 - 1);  // signed integers only (int32/int64/date/timestamp keys)
      put_be(p + 1, v, bw);
    }
    off += 1 + bw;
  }
  return c;
}

struct OutColumn {
  std::vector<uint8_t> values, validity, data;
  std::vector<int32_t> offsets;
};

struct Reader {
  std::vector<Stream> streams;
  std::vector<int> pk_idx;
  int batch_size = 8192;
  int ncols = 0;
  std::vector<std::shared_ptr<arrow::RecordBatch>> live;  // batches referenced by pending rows
  std::vector<std::pair<int, int64_t>> in_progress;        // (live slot, row)
  std::vector<int> tree;
  int64_t rows_in = 0, rows_out = 0;
  uint64_t checksum = 0;
  bool keep_output = false;  // verify mode: the merged batches are kept as Arrow arrays
  std::shared_ptr<arrow::Schema> out_schema;
  std::vector<std::shared_ptr<arrow::RecordBatch>> out_batches;
  bool have_pending = false;
  int pend_stream = -1;
  int64_t pend_row = 0;
  std::shared_ptr<Cursor> pend_cur;
  int pend_slot = -1;
  int64_t pend_last_row = 0;

  bool advance_range(Stream& s) {  // SortKeyBatchRange::advance
    s.begin = s.end;
    if (s.begin >= s.cur->n) return false;
    s.end = s.begin + 1;
    while (s.end < s.cur->n && s.cur->eq(s.end, *s.cur, s.begin)) ++s.end;
    return true;
  }
  bool poll(Stream& s) {  // maybe_poll_stream: next non-empty batch
    while (true) {
      std::shared_ptr<arrow::RecordBatch> b;
      auto st = s.rbr->ReadNext(&b);
      if (!st.ok()) throw std::runtime_error(st.ToString());
      if (!b) { s.exhausted = true; s.batch.reset(); s.cur.reset(); return false; }
      if (b->num_rows() == 0) continue;
      rows_in += b->num_rows();
      s.batch = b;
      s.cur = make_cursor(*b, pk_idx);
      live.push_back(b);
      s.batch_slot = (int)live.size() - 1;
      s.begin = s.end = 0;
      advance_range(s);
      return true;
    }
  }
  bool less(int a, int b) const {  // (key, stream_idx); exhausted streams always lose
    const Stream &x = streams[a], &y = streams[b];
    if (x.exhausted) return false;
    if (y.exhausted) return true;
    int c = x.cur->cmp(x.begin, *y.cur, y.begin);
    return c < 0 || (c == 0 && a < b);
  }
  void init_tree() {
    const int k = (int)streams.size();
    tree.assign(k, -1);
    for (int i = 0; i < k; ++i) {
      int winner = i;
      int node = (k + i) / 2;
      while (node != 0 && tree[node] != -1) {
        int challenger = tree[node];
        if (less(challenger, winner)) { tree[node] = winner; winner = challenger; }
        node /= 2;
      }
      tree[node] = winner;
    }
  }
  void update_tree() {
    const int k = (int)streams.size();
    int winner = tree[0];
    int node = (k + winner) / 2;
    while (node != 0) {
      int challenger = tree[node];
      if (less(challenger, winner)) { tree[node] = winner; winner = challenger; }
      node /= 2;
    }
    tree[0] = winner;
  }

  void flush_pending() {
    if (!have_pending) return;
    in_progress.emplace_back(pend_slot, pend_last_row);
    have_pending = false;
    if ((int)in_progress.size() == batch_size) build_batch();
  }

  void build_batch() {  // build_record_batch_full_merge + interleave
    const int64_t n = (int64_t)in_progress.size();
    if (n == 0) return;
    std::vector<std::shared_ptr<arrow::Array>> kept;
    for (int c = 0; c < ncols; ++c) {
      OutColumn oc;
      const auto& type = live[in_progress[0].first]->column(c)->type();
      bool any_null = false;
      for (auto& pr : in_progress) any_null |= live[pr.first]->column(c)->null_count() > 0;
      if (any_null) oc.validity.assign((size_t)(n + 7) / 8, 0);
      if (type->id() == arrow::Type::STRING || type->id() == arrow::Type::BINARY) {
        oc.offsets.resize((size_t)n + 1);
        int64_t total = 0;
        for (int64_t i = 0; i < n; ++i) {
          const auto& a = static_cast<const arrow::BinaryArray&>(*live[in_progress[i].first]->column(c));
          oc.offsets[(size_t)i] = (int32_t)total;
          if (!a.IsNull(in_progress[i].second)) total += a.value_length(in_progress[i].second);
        }
        oc.offsets[(size_t)n] = (int32_t)total;
        oc.data.resize((size_t)total);
        for (int64_t i = 0; i < n; ++i) {
          const auto& a = static_cast<const arrow::BinaryArray&>(*live[in_progress[i].first]->column(c));
          const int64_t r = in_progress[i].second;
          if (a.IsNull(r)) continue;
          if (any_null) oc.validity[(size_t)i >> 3] |= (uint8_t)(1u << (i & 7));
          int32_t len;
          const uint8_t* p = a.GetValue(r, &len);
          memcpy(oc.data.data() + oc.offsets[(size_t)i], p, (size_t)len);
        }
        if (!oc.data.empty()) checksum += oc.data[oc.data.size() / 2];
      } else {
        const int bw = type->byte_width();
        oc.values.resize((size_t)n * bw);
        for (int64_t i = 0; i < n; ++i) {
          const auto& a = *live[in_progress[i].first]->column(c);
          const int64_t r = in_progress[i].second;
          memcpy(oc.values.data() + i * bw, a.data()->buffers[1]->data() + (a.offset() + r) * bw, (size_t)bw);
          if (any_null && a.IsValid(r)) oc.validity[(size_t)i >> 3] |= (uint8_t)(1u << (i & 7));
        }
        checksum += oc.values[oc.values.size() / 2];
      }
      if (keep_output) {
        auto buf = [](const void* p, size_t bytes) -> std::shared_ptr<arrow::Buffer> {
          auto b = arrow::AllocateBuffer((int64_t)bytes).ValueOrDie();
          if (bytes) memcpy(b->mutable_data(), p, bytes);
          return std::shared_ptr<arrow::Buffer>(std::move(b));
        };
        std::shared_ptr<arrow::Buffer> valid = any_null ? buf(oc.validity.data(), oc.validity.size()) : nullptr;
        std::shared_ptr<arrow::ArrayData> ad;
        if (!oc.offsets.empty())
          ad = arrow::ArrayData::Make(type, n, {valid, buf(oc.offsets.data(), oc.offsets.size() * 4), buf(oc.data.data(), oc.data.size())});
        else
          ad = arrow::ArrayData::Make(type, n, {valid, buf(oc.values.data(), oc.values.size())});
        kept.push_back(arrow::MakeArray(ad));
      }
    }
    if (keep_output) out_batches.push_back(arrow::RecordBatch::Make(out_schema, n, kept));
    rows_out += n;
    in_progress.clear();
    // drop batches nobody references any more (streams keep their current one)
    std::vector<std::shared_ptr<arrow::RecordBatch>> keep;
    for (auto& s : streams) {
      if (!s.exhausted && s.batch) { keep.push_back(s.batch); s.batch_slot = (int)keep.size() - 1; }
    }
    if (have_pending) {
      // the pending row's batch may belong to a stream that moved on
      bool found = false;
      for (size_t i = 0; i < keep.size(); ++i)
        if (keep[i] == live[pend_slot]) { pend_slot = (int)i; found = true; break; }
      if (!found) { keep.push_back(live[pend_slot]); pend_slot = (int)keep.size() - 1; }
    }
    live.swap(keep);
  }

  void run() {
    for (auto& s : streams) poll(s);
    init_tree();
    while (true) {
      const int w = tree[0];
      Stream& s = streams[w];
      if (s.exhausted) break;
      // match_row: same key as the pending output row?
      if (have_pending && !(pend_cur->eq(pend_row, *s.cur, s.begin))) flush_pending();
      have_pending = true;
      pend_cur = s.cur;
      pend_row = s.begin;
      pend_slot = s.batch_slot;
      pend_last_row = s.end - 1;
      if (!advance_range(s)) poll(s);
      update_tree();
    }
    flush_pending();
    build_batch();
  }
};

std::vector<std::string> split_tabs(const std::string& line) {
  std::vector<std::string> out;
  std::stringstream ss(line);
  std::string tok;
  while (std::getline(ss, tok, '\t'))
    if (!tok.empty()) out.push_back(tok);
  return out;
}

}  // namespace

int main() {
  std::string line;
  std::getline(std::cin, line);
  int threads = 1, batch_size = 8192, verify = 0;
  { std::stringstream ss(line); ss >> threads >> batch_size >> verify; }
  std::getline(std::cin, line);
  const auto pks = split_tabs(line);
  std::getline(std::cin, line);
  const auto columns = split_tabs(line);
  std::vector<std::vector<std::string>> buckets;
  while (std::getline(std::cin, line)) {
    auto f = split_tabs(line);
    if (!f.empty()) buckets.push_back(f);
  }
  std::atomic<size_t> next{0};
  std::atomic<int64_t> rows_in{0}, rows_out{0};
  std::atomic<uint64_t> checksum{0};
  std::string error;
  std::mutex err_mu;
  auto t0 = std::chrono::steady_clock::now();
  auto worker = [&]() {
    try {
      while (true) {
        const size_t b = next.fetch_add(1);
        if (b >= buckets.size()) break;
        Reader rd;
        rd.batch_size = batch_size;
        for (const auto& path : buckets[b]) {
          Stream s;
          auto in = arrow::io::ReadableFile::Open(path).ValueOrDie();
          s.fr = parquet::arrow::OpenFile(in, arrow::default_memory_pool()).ValueOrDie();
          s.fr->set_use_threads(false);
          s.fr->set_batch_size(batch_size);
          std::shared_ptr<arrow::Schema> schema;
          auto st = s.fr->GetSchema(&schema);
          if (!st.ok()) throw std::runtime_error(st.ToString());
          std::vector<int> idx;
          for (const auto& c : columns) {
            int i = schema->GetFieldIndex(c);
            if (i < 0) throw std::runtime_error("column missing: " + c);
            idx.push_back(i);
          }
          std::vector<int> rgs(s.fr->num_row_groups());
          for (size_t i = 0; i < rgs.size(); ++i) rgs[i] = (int)i;
          auto rr = s.fr->GetRecordBatchReader(rgs, idx);
          if (!rr.ok()) throw std::runtime_error(rr.status().ToString());
          s.rbr = std::move(*rr);
          rd.streams.push_back(std::move(s));
        }
        auto out_schema = rd.streams[0].rbr->schema();
        rd.ncols = out_schema->num_fields();
        for (const auto& p : pks) {
          const int i = out_schema->GetFieldIndex(p);
          if (i < 0) throw std::runtime_error("primary key not in the column list: " + p);
          rd.pk_idx.push_back(i);
        }
        rd.keep_output = verify != 0;
        rd.out_schema = out_schema;
        rd.run();
        if (verify) {
          const char* dir = getenv("REF_CPU_VERIFY_DIR");
          if (!dir) throw std::runtime_error("verify mode needs REF_CPU_VERIFY_DIR");
          auto sink = arrow::io::FileOutputStream::Open(std::string(dir) + "/bucket_" + std::to_string(b) + ".arrow").ValueOrDie();
          auto wr = arrow::ipc::MakeFileWriter(sink, out_schema).ValueOrDie();
          for (const auto& rb : rd.out_batches) {
            auto st2 = wr->WriteRecordBatch(*rb);
            if (!st2.ok()) throw std::runtime_error(st2.ToString());
          }
          if (!wr->Close().ok() || !sink->Close().ok()) throw std::runtime_error("cannot write the verify output");
        }
        rows_in += rd.rows_in;
        rows_out += rd.rows_out;
        checksum += rd.checksum;
      }
    } catch (const std::exception& e) {
      std::lock_guard<std::mutex> g(err_mu);
      error = e.what();
    }
  };
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; ++t) pool.emplace_back(worker);
  for (auto& t : pool) t.join();
  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (!error.empty()) {
    std::cerr << "ref_cpu: " << error << std::endl;
    return 1;
  }
  std::cout << "{\"seconds\": " << secs << ", \"threads\": " << threads << ", \"rows_in\": " << rows_in.load()
            << ", \"rows_out\": " << rows_out.load() << ", \"checksum\": " << checksum.load() << "}" << std::endl;
  return 0;
}
