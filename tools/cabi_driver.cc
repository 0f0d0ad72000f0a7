// SPDX-License-Identifier: Apache-2.0
//
// cabi_driver.cc -- a native consumer of the lakesoul-io-c C-ABI, shaped like an engine's task
// executor: T worker threads ("task slots"), each one running the exact call sequence of
// NativeIOReader.initializeReader / nextBatchBlocked (native-io/lakesoul-io-java/.../NativeIOReader.java:60-129)
// for one hash bucket at a time and releasing every Arrow batch it receives.
// Used by bench.py for the end-to-end leg (no Python/GIL in the timed path) -- it calls nothing but
// the symbols declared in include/lakesoul_c_bindings.h (+ lsb_reader_get_stats for the byte counters).
//
// usage: cabi_driver <job-file> [--serve]
// job file:  line 1: <threads> <batch_size> <passes>
//            line 2: primary keys (tab separated)   line 3: options key=value (tab separated, may be empty)
//            line 4: filters (tab separated, may be empty)
//            then one line per bucket: file paths (tab separated)
// The requested schema is inferred from the files (set_inferring_schema), i.e. all columns.
// default mode: runs <passes> passes over all buckets, prints ONE JSON line (seconds = median pass,
//            "pass_seconds" = every pass).
// --serve:   prints "READY" after device initialisation, then for every line "GO" read from stdin runs one
//            pass ("GO <n>": over the first n buckets only) and prints its JSON line; "QUIT" (or EOF) ends. bench.py puts a barrier across the ranks
//            in front of every "GO", so the passes of all ranks start together.
#include <sys/resource.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
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

struct Job {
  int threads = 1, batch_size = 8192, passes = 1;
  int group = 1;  // hash buckets per reader (option b200_file_groups of the library): one device scan for `group` buckets
  std::vector<std::string> pks, opts, filters;
  std::vector<std::vector<std::string>> buckets;
};

struct PassResult {
  double seconds = 0, cpu_seconds = 0;
  int64_t rows_out = 0, rows_in = 0, h2d = 0, d2h = 0, launches = 0, dev_pages = 0, host_files = 0, batches = 0,
          dev_decomp = 0;
  uint64_t checksum = 0;
  std::string error;
};

static double cpu_seconds() {
  struct rusage ru;
  getrusage(RUSAGE_SELF, &ru);
  return ru.ru_utime.tv_sec + ru.ru_stime.tv_sec + 1e-6 * (ru.ru_utime.tv_usec + ru.ru_stime.tv_usec);
}

static PassResult run_pass(const Job& job, size_t nbuckets) {
  PassResult R;
  nbuckets = std::min(nbuckets, job.buckets.size());
  std::mutex mu;
  std::atomic<size_t> next{0};
  std::atomic<int64_t> a_out{0}, a_in{0}, a_h2d{0}, a_d2h{0}, a_l{0}, a_dp{0}, a_hf{0}, a_b{0}, a_dz{0};
  std::atomic<uint64_t> a_ck{0};
  std::atomic<bool> failed{false};
  const auto t0 = std::chrono::steady_clock::now();
  const double cpu0 = cpu_seconds();
  auto fail = [&](const std::string& m) {
    std::lock_guard<std::mutex> g(mu);
    if (R.error.empty()) R.error = m;
    failed = true;
  };
  auto worker = [&]() {
    while (!failed) {
      const size_t b = next.fetch_add((size_t)job.group);
      if (b >= nbuckets) break;
      const size_t b_end = std::min(nbuckets, b + (size_t)job.group);
      IOConfigBuilder* cb = new_lakesoul_io_config_builder();
      std::string groups;
      for (size_t bb = b; bb < b_end; ++bb) {
        for (const auto& f : job.buckets[bb]) cb = lakesoul_config_builder_add_single_file(cb, f.c_str());
        groups += (groups.empty() ? "" : ",") + std::to_string(job.buckets[bb].size());
      }
      if (b_end - b > 1) cb = lakesoul_config_builder_set_option(cb, "b200_file_groups", groups.c_str());
      for (const auto& p : job.pks) cb = lakesoul_config_builder_add_single_primary_key(cb, p.c_str());
      cb = lakesoul_config_builder_set_inferring_schema(cb, true);
      cb = lakesoul_config_builder_set_batch_size(cb, (c_size_t)job.batch_size);
      cb = lakesoul_config_builder_set_thread_num(cb, 2);
      for (const auto& kv : job.opts) {
        const size_t eq = kv.find('=');
        if (eq != std::string::npos) cb = lakesoul_config_builder_set_option(cb, kv.substr(0, eq).c_str(), kv.substr(eq + 1).c_str());
      }
      for (const auto& fl : job.filters) cb = lakesoul_config_builder_add_filter(cb, fl.c_str());
      TokioRuntime* rt = create_tokio_runtime_from_builder(new_tokio_runtime_builder());
      IOConfig* cfg = create_lakesoul_io_config_from_builder(cb);
      CResult_Reader* rd = create_lakesoul_reader_from_config(cfg, rt);
      const char* err = check_reader_created(rd);
      if (err) {
        fail(err);
        free_lakesoul_reader(rd);
        return;
      }
      CStatus* st = start_reader(rd);
      if (st->status != 0) {
        fail(st->err ? st->err : "start_reader failed");
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
        if (n < 0) fail(st->err ? st->err : "next_record_batch_blocked failed");
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
        a_dp += rs.device_decoded_pages; a_hf += rs.host_decoded_files; a_dz += rs.device_decompressed_pages;
      }
      free_lakesoul_reader(rd);
    }
  };
  std::vector<std::thread> pool;
  for (int t = 0; t < job.threads; ++t) pool.emplace_back(worker);
  for (auto& t : pool) t.join();
  R.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  R.cpu_seconds = cpu_seconds() - cpu0;
  R.rows_out = a_out; R.rows_in = a_in; R.h2d = a_h2d; R.d2h = a_d2h; R.launches = a_l; R.dev_pages = a_dp;
  R.host_files = a_hf; R.batches = a_b; R.checksum = a_ck; R.dev_decomp = a_dz;
  return R;
}

static void print_result(const Job& job, const PassResult& r, const std::vector<double>& all) {
  std::string passes = "[";
  for (size_t i = 0; i < all.size(); ++i) {
    char b[32];
    snprintf(b, sizeof(b), "%s%.6f", i ? ", " : "", all[i]);
    passes += b;
  }
  passes += "]";
  printf("{\"seconds\": %.6f, \"cpu_seconds\": %.6f, \"threads\": %d, \"buckets\": %zu, \"rows_in\": %lld, \"rows_out\": %lld, "
         "\"batches\": %lld, \"h2d_bytes\": %lld, \"d2h_bytes\": %lld, \"kernel_launches\": %lld, \"device_decoded_pages\": %lld, "
         "\"device_decompressed_pages\": %lld, \"host_decoded_files\": %lld, \"checksum\": %llu, \"pass_seconds\": %s}\n",
         r.seconds, r.cpu_seconds, job.threads, job.buckets.size(), (long long)r.rows_in, (long long)r.rows_out,
         (long long)r.batches, (long long)r.h2d, (long long)r.d2h, (long long)r.launches, (long long)r.dev_pages,
         (long long)r.dev_decomp, (long long)r.host_files, (unsigned long long)r.checksum, passes.c_str());
  fflush(stdout);
}

int main(int argc, char** argv) {
  if (argc < 2) {
    fprintf(stderr, "usage: cabi_driver <job-file> [--serve]\n");
    return 2;
  }
  const bool serve = argc > 2 && std::string(argv[2]) == "--serve";
  std::ifstream in(argv[1]);
  if (!in) {
    fprintf(stderr, "cabi_driver: cannot open %s\n", argv[1]);
    return 2;
  }
  Job job;
  std::string line;
  std::getline(in, line);
  { std::stringstream ss(line); ss >> job.threads >> job.batch_size >> job.passes; if (!(ss >> job.group) || job.group < 1) job.group = 1; }
  std::getline(in, line);
  job.pks = split_tabs(line);
  std::getline(in, line);
  job.opts = split_tabs(line);
  std::getline(in, line);
  job.filters = split_tabs(line);
  while (std::getline(in, line)) {
    auto f = split_tabs(line);
    if (!f.empty()) job.buckets.push_back(f);
  }
  if (lsb_init_device() < 0) {
    fprintf(stderr, "cabi_driver: %s\n", lsb_last_error());
    return 2;
  }
  if (serve) {
    printf("READY\n");
    fflush(stdout);
    while (std::getline(std::cin, line)) {
      if (line == "QUIT") break;
      if (line.rfind("GO", 0) != 0) continue;
      size_t n = job.buckets.size();  // "GO <n>": only the first n buckets (warm-up)
      if (line.size() > 3) n = (size_t)std::max(1L, atol(line.c_str() + 3));
      PassResult r = run_pass(job, n);
      if (!r.error.empty()) {
        fprintf(stderr, "cabi_driver: %s\n", r.error.c_str());
        printf("{\"error\": \"pass failed\"}\n");
        fflush(stdout);
        return 1;
      }
      print_result(job, r, {r.seconds});
    }
    return 0;
  }
  std::vector<PassResult> results;
  std::vector<double> secs;
  for (int p = 0; p < std::max(job.passes, 1); ++p) {
    PassResult r = run_pass(job, job.buckets.size());
    if (!r.error.empty()) {
      fprintf(stderr, "cabi_driver: %s\n", r.error.c_str());
      return 1;
    }
    secs.push_back(r.seconds);
    results.push_back(r);
  }
  // the pass whose time is the median
  std::vector<size_t> order(secs.size());
  for (size_t i = 0; i < order.size(); ++i) order[i] = i;
  std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return secs[a] < secs[b]; });
  print_result(job, results[order[order.size() / 2]], secs);
  return 0;
}
