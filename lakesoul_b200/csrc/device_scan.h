// SPDX-License-Identifier: Apache-2.0
//
// device_scan.h -- one scan group (a set of hash buckets) on the GPU:
// decoded file columns -> HBM column buffers -> PK merge -> winner gather -> pinned host
// Arrow buffers, exported as zero-copy RecordBatches of <= batch_size rows.
//
// This is the B200 replacement for MergeParquetExec::execute + merge_stream
// (reference rust/lakesoul-io/src/physical_plan/merge/mod.rs:233-374).
#pragma once
#include <arrow/api.h>
#include <cuda_runtime.h>

#include <map>
#include <memory>
#include <string>
#include <vector>

#include "filter.h"
#include "mor_engine.h"
#include "scan_types.h"

namespace lsb {

ColumnKind classify_type(const arrow::DataType& t);

// Development trace (LAKESOUL_B200_TRACE=1): phase marks of the calling thread's current reader,
// printed to stderr as one line per reader by trace_flush(). No-ops when the variable is unset.
void trace_mark(const char* phase);
void trace_flush(const char* tag);

// merge_operator.rs:41-53 -- unknown names are an error (the reference panics)
MergeOp merge_op_from_name(const std::string& name);

// Write side: the rows of ONE group of unmerged inputs come out sorted by
// (prefix_cols, [hash bucket of hash_cols], hash_cols, suffix_cols), ascending, nulls first, ties in input order
// -- the order SortExec + RepartitionByRangeAndHashExec give the reference's writers (sort_writer.rs:59-76,
// partitioning_writer.rs:139-241).
struct SortSpec {
  bool enabled = false;
  std::vector<std::string> prefix_cols;   // range partition columns
  std::vector<std::string> hash_cols;     // primary keys
  std::vector<std::string> suffix_cols;   // aux sort columns
  bool with_buckets = false;              // hash the rows into buckets and sort by bucket before the primary key
  uint32_t bucket_num = 1;
};

struct ScanSpec {
  std::shared_ptr<arrow::Schema> physical_schema;  // columns carried through the merge, target order
  std::vector<std::string> primary_keys;
  std::vector<MergeOp> ops;     // one per physical column (UseLast by default)
  bool merge = true;            // false = plain concatenation in file order (merge/mod.rs:297-312)
  int batch_size = 8192;
  std::shared_ptr<FilterExpr> filter;  // evaluated on the merged rows (session.rs:839-877), may be null
  SortSpec sort;                // merge must be false
};

class RawParquetFile;

// One delta file: either already decoded on the host (Arrow table; used for shapes the device
// decoders do not cover) or a raw Parquet image whose pages are decoded on the GPU (page_plan.h).
struct FileInput {
  std::shared_ptr<arrow::Table> table;
  std::shared_ptr<RawParquetFile> raw;
};

struct BucketFiles {
  // commit order; a file may lack some physical columns
  std::vector<FileInput> files;
};

struct ScanCounters {
  int64_t rows_in = 0, rows_out = 0, h2d_bytes = 0, d2h_bytes = 0;
  int64_t device_decoded_pages = 0, host_decoded_files = 0, device_decompressed_pages = 0;
  int32_t kernel_launches = 0;
  double upload_ms = 0, merge_ms = 0, gather_ms = 0, download_ms = 0;
  std::string path;
};

// A checked-out scan slot (stream + device scratch arena), installed on the calling thread until the
// session is destroyed. A reader opens one before it reads its files so that every file image
// travels to the GPU while the next file is still being read.
class ScanSession;
std::shared_ptr<ScanSession> begin_scan_session();
// copies a pinned host block to fresh device memory of the session (asynchronous on its stream)
uint8_t* session_upload(ScanSession& s, const uint8_t* pinned, size_t bytes);
int64_t session_h2d_bytes(const ScanSession& s);

class DeviceScan {
 public:
  DeviceScan(ScanSpec spec, std::vector<BucketFiles> buckets, std::shared_ptr<ScanSession> session = nullptr);
  ~DeviceScan();
  // runs upload -> merge -> (filter) -> gather -> download; synchronous on return
  void Run();
  int num_buckets() const;
  int64_t num_batches(int bucket) const;
  int64_t num_rows(int bucket) const;
  // zero-copy over pinned host memory; columns == spec.physical_schema
  std::shared_ptr<arrow::RecordBatch> Batch(int bucket, int64_t i) const;
  const ScanCounters& counters() const;
  // sort mode with buckets: hash bucket of every output row of bucket 0, in output order (valid while the scan lives)
  const uint32_t* sorted_buckets() const;

 private:
  struct Impl;
  std::shared_ptr<Impl> impl_;
};

}  // namespace lsb
