// SPDX-License-Identifier: Apache-2.0
//
// writer.h -- write side of the native-IO boundary: the C++ mirror of
// SyncSendableMutableLakeSoulWriter (reference rust/lakesoul-io/src/writer/mod.rs:61-400) and the three writers
// behind it:
//   * MultiPartAsyncWriter   (writer/async_writer/multipart_writer.rs:63-226): rows as they come, one file;
//   * SortAsyncWriter        (sort_writer.rs:48-120): rows sorted by (primary keys, aux sort columns), one file;
//   * PartitioningAsyncWriter(partitioning_writer.rs:68-447): rows sorted by (range partitions, primary keys, aux
//     sort columns) and split by range partition value and spark-murmur3 hash bucket of the primary keys
//     (physical_plan/repartition/mod.rs:222-300), one file per (partition, bucket), range columns dropped.
// The hash, the sort and the reordering of every column run on the GPU (device_scan.h sort mode: plan_hash_buckets,
// plan_sort, the gather kernels); the host encodes Parquet (ZSTD, dictionaries off -- multipart_writer.rs:121-133)
// with Arrow C++. A GPU is mandatory for the sorting modes.
#pragma once
#include <arrow/api.h>

#include <memory>
#include <string>
#include <vector>

#include "reader.h"

namespace lsb {

struct FlushOutput {          // writer/async_writer/mod.rs:46-55
  std::string partition_desc; // "-5" when the table has no range partitions (helpers/mod.rs:476-488)
  std::string file_path;
  int64_t file_size = 0;
  std::string file_exist_cols;  // comma-joined leaf column names of the file (helpers/mod.rs:897-906)
};

struct WriterStats {
  int64_t rows_in = 0, rows_out = 0, files = 0, flush_cycles = 0;
  int64_t h2d_bytes = 0, d2h_bytes = 0;
  int32_t kernel_launches = 0;
  double sort_ms = 0, encode_ms = 0;
};

class Writer {
 public:
  explicit Writer(IOConfig cfg);  // throws std::runtime_error on an unusable configuration
  ~Writer();
  std::shared_ptr<arrow::Schema> schema() const;            // what write() expects (target schema, uniformed)
  void Write(const std::shared_ptr<arrow::RecordBatch>& batch);
  // writer/mod.rs:335-383: "<groups>" then per partition "\x01<desc>\x02<path>\x03<size>\x03<cols>[\x02...]"
  std::string FlushAndClose();
  void AbortAndClose();                                     // removes every file this writer created
  const std::vector<FlushOutput>& outputs() const;          // valid after FlushAndClose
  const WriterStats& stats() const;

 private:
  struct Impl;
  std::unique_ptr<Impl> impl_;
};

// helpers/mod.rs:205-274 format_scalar_value for row `row` of a range-partition column
std::string format_partition_value(const arrow::Array& column, int64_t row);

}  // namespace lsb
