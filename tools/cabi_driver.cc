// SPDX-License-Identifier: Apache-2.0
//
// cabi_driver.cc -- a native consumer of the lakesoul-io-c C-ABI, shaped like an engine's task
// executor: T worker threads ("task slots"), each one running the exact call sequence of
// NativeIOReader.initializeReader / nextBatchBlocked (native-io/lakesoul-io-java/.../NativeIOReader.java:60-129)
// for one hash bucket at a time and releasing every Arrow batch it receives.
// Used by bench.py for the end-to-end leg (no Python/GIL in the timed path) -- it calls nothing but
// the symbols declared in include/lakesoul_c_bindings.h (+ lsb_reader_get_stats for the byte counters).
//
// stdin:  line 1: <threads> <batch_size> <repeat>
//         line 2: primary keys (tab separated)   line 3: options key=value (tab separated, may be empty)
//         line 4: filters (tab separated, may be empty)
//         then one line per bucket: file paths (tab separated)
// The requested schema is inferred from the files (set_inferring_schema), i.e. all columns.
// stdout: one JSON line.
#include <dlfcn.h>
#include <sys/resource.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "lakesoul_b200_ext.h"

// Arrow C Data Interface (only what a consumer needs)
struct ArrowArray {
  int64_t length, null_count, offset, n_buffers, n_children;
  const void** buffers;
  struct ArrowArray** children;
  struct ArrowArray* dictionary;
  void (*release)(struct ArrowArray*);
  void* private_data;
};

static std::vector<std::string> split_tabs(const std::string& line) {
  std::vector<std::string> out;
  std::stringstream ss(line);
  std::string tok;
  while (std::getline(ss, tok, '\t'))
    if (!tok.empty()) out.push_back(tok);
  return out;
}

int main() {
  std::string line;
  std::getline(std::cin, line);
  int threads = 1, batch_size = 8192, repeat = 1;
  { std::stringstream ss(line); ss >> threads >> batch_size >> repeat; }
  std::getline(std::cin, line);
  const auto pks = split_tabs(line);
  std::getline(std::cin, line);
  const auto opts = split_tabs(line);
  std::getline(std::cin, line);
  const auto filters = split_tabs(line);
  std::vector<std::vector<std::string>> buckets;
  while (std::getline(std::cin, line)) {
    auto f = split_tabs(line);
    if (!f.empty()) buckets.push_back(f);
  }
  if (lsb_init_device() < 0) {
    fprintf(stderr, "cabi_driver: %s\n", lsb_last_error());
    return 2;
  }
  std::string error;
  std::mutex mu;
  double best = 1e30, best_cpu = 0;
  auto cpu_seconds = [] {
    struct rusage ru;
    getrusage(RUSAGE_SELF, &ru);
    return ru.ru_utime.tv_sec + ru.ru_stime.tv_sec + 1e-6 * (ru.ru_utime.tv_usec + ru.ru_stime.tv_usec);
  };
  int64_t rows_out = 0, rows_in = 0, h2d = 0, d2h = 0, launches = 0, dev_pages = 0, host_files = 0, batches = 0;
  uint64_t checksum = 0;
  for (int rep = 0; rep < repeat; ++rep) {
    std::atomic<size_t> next{0};
    std::atomic<int64_t> a_out{0}, a_in{0}, a_h2d{0}, a_d2h{0}, a_l{0}, a_dp{0}, a_hf{0}, a_b{0};
    std::atomic<uint64_t> a_ck{0};
    auto t0 = std::chrono::steady_clock::now();
    const double cpu0 = cpu_seconds();
    auto worker = [&]() {
      while (true) {
        const size_t b = next.fetch_add(1);
        if (b >= buckets.size()) break;
        IOConfigBuilder* cb = new_lakesoul_io_config_builder();
        for (const auto& f : buckets[b]) cb = lakesoul_config_builder_add_single_file(cb, f.c_str());
        for (const auto& p : pks) cb = lakesoul_config_builder_add_single_primary_key(cb, p.c_str());
        cb = lakesoul_config_builder_set_inferring_schema(cb, true);
        cb = lakesoul_config_builder_set_batch_size(cb, (c_size_t)batch_size);
        cb = lakesoul_config_builder_set_thread_num(cb, 2);
        for (const auto& kv : opts) {
          const size_t eq = kv.find('=');
          if (eq != std::string::npos) cb = lakesoul_config_builder_set_option(cb, kv.substr(0, eq).c_str(), kv.substr(eq + 1).c_str());
        }
        for (const auto& fl : filters) cb = lakesoul_config_builder_add_filter(cb, fl.c_str());
        TokioRuntime* rt = create_tokio_runtime_from_builder(new_tokio_runtime_builder());
        IOConfig* cfg = create_lakesoul_io_config_from_builder(cb);
        CResult_Reader* rd = create_lakesoul_reader_from_config(cfg, rt);
        const char* err = check_reader_created(rd);
        if (err) {
          std::lock_guard<std::mutex> g(mu);
          error = err;
          free_lakesoul_reader(rd);
          return;
        }
        CStatus* st = start_reader(rd);
        if (st->status != 0) {
          std::lock_guard<std::mutex> g(mu);
          error = st->err ? st->err : "start_reader failed";
          free_c_status(st);
          free_lakesoul_reader(rd);
          return;
        }
        free_c_status(st);
        while (true) {
          ArrowArray arr;
          memset(&arr, 0, sizeof(arr));
          st = next_record_batch_blocked(rd, (c_ptrdiff_t)&arr);
          const int n = st->status;
          if (n < 0) {
            std::lock_guard<std::mutex> g(mu);
            error = st->err ? st->err : "next_record_batch_blocked failed";
          }
          free_c_status(st);
          if (n <= 0) break;
          a_out += n;
          a_b += 1;
          // touch the data like a consumer would: first value buffer of the first and last column
          uint64_t ck = (uint64_t)arr.length;
          if (arr.n_children > 0) {
            const ArrowArray* c0 = arr.children[0];
            const ArrowArray* cl = arr.children[arr.n_children - 1];
            if (c0->n_buffers > 1 && c0->buffers[1]) ck += static_cast<const uint8_t*>(c0->buffers[1])[0];
            if (cl->n_buffers > 1 && cl->buffers[1]) ck += static_cast<const uint8_t*>(cl->buffers[1])[0];
          }
          a_ck += ck;
          if (arr.release) arr.release(&arr);
        }
        LsbReaderStats rs;
        if (lsb_reader_get_stats(rd, &rs) == 0) {
          a_in += rs.rows_in; a_h2d += rs.h2d_bytes; a_d2h += rs.d2h_bytes; a_l += rs.kernel_launches;
          a_dp += rs.device_decoded_pages; a_hf += rs.host_decoded_files;
        }
        free_lakesoul_reader(rd);
      }
    };
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) pool.emplace_back(worker);
    for (auto& t : pool) t.join();
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (!error.empty()) break;
    if (secs < best) { best = secs; best_cpu = cpu_seconds() - cpu0; }
    rows_out = a_out; rows_in = a_in; h2d = a_h2d; d2h = a_d2h; launches = a_l; dev_pages = a_dp; host_files = a_hf;
    batches = a_b; checksum = a_ck;
  }
  if (!error.empty()) {
    fprintf(stderr, "cabi_driver: %s\n", error.c_str());
    return 1;
  }
  printf("{\"seconds\": %.6f, \"cpu_seconds\": %.6f, \"threads\": %d, \"repeat\": %d, \"rows_in\": %lld, \"rows_out\": %lld, \"batches\": %lld, "
         "\"h2d_bytes\": %lld, \"d2h_bytes\": %lld, \"kernel_launches\": %lld, \"device_decoded_pages\": %lld, "
         "\"host_decoded_files\": %lld, \"checksum\": %llu}\n",
         best, best_cpu, threads, repeat, (long long)rows_in, (long long)rows_out, (long long)batches, (long long)h2d,
         (long long)d2h, (long long)launches, (long long)dev_pages, (long long)host_files, (unsigned long long)checksum);
  return 0;
}
