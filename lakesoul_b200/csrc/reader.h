// SPDX-License-Identifier: Apache-2.0
//
// reader.h -- host side of the drop-in reader: the C++ mirror of the reference's
// LakeSoulIOConfig(Builder) (rust/lakesoul-io/src/config/mod.rs:39-118, 330-700) and
// LakeSoulReader / SyncSendableMutableLakeSoulReader (rust/lakesoul-io/src/reader.rs:95-363).
// Everything compute-heavy goes to the CUDA engine in mor_engine.h; a GPU is mandatory --
// there is no CPU fallback for the merge.
#pragma once
#include <arrow/api.h>

#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace lsb {

struct FilterExpr;

struct IOConfig {
  std::string prefix;
  std::vector<std::string> files;             // commit order: later file wins
  std::vector<std::string> primary_keys;
  std::vector<std::string> range_partitions;
  std::vector<std::string> aux_sort_cols;
  std::vector<std::string> filter_strs;       // parquet-mr FilterPredicate.toString form
  int filter_protos = 0;                      // substrait plans handed in (unsupported -> error at start)
  std::shared_ptr<arrow::Schema> target_schema = arrow::schema({});
  std::shared_ptr<arrow::Schema> partition_schema = arrow::schema({});
  size_t batch_size = 8192;                   // config/mod.rs:67
  size_t thread_num = 2;
  size_t prefetch_size = 1;
  size_t max_row_group_size = 250000;
  size_t max_row_group_num_values = 2147483647;
  std::string hash_bucket_num = "1";
  bool use_dynamic_partition = false;
  bool inferring_schema = false;
  std::map<std::string, std::string> object_store_options;
  std::map<std::string, std::string> merge_operators;       // column -> operator name
  std::map<std::string, std::string> default_column_value;  // column -> string value
  std::map<std::string, std::string> options;

  // set by the reader itself for the passes of a bucket streamed by key range (never through the C-ABI)
  std::shared_ptr<FilterExpr> pass_filter;    // key range of this pass, ANDed to the pushed filters
  bool lazy_files = false;                    // footers first, then only the byte span of the row groups the pass keeps

  // config/mod.rs:157-162: option map first, then env LAKESOUL_<KEY>
  bool option(const std::string& key, std::string* out) const;
  std::string cdc_column() const;
  bool is_compacted() const;
  bool skip_merge_on_read() const;
};

struct ReaderStats {
  int64_t rows_in = 0, rows_out = 0;
  int64_t h2d_bytes = 0, d2h_bytes = 0;
  int32_t kernel_launches = 0;
  int64_t device_decoded_pages = 0, host_decoded_files = 0, pruned_row_groups = 0, device_decompressed_pages = 0;
  int32_t stream_passes = 0;   // > 1: the bucket was read in that many passes over ascending primary-key ranges
  int64_t pruned_pages = 0;    // data pages outside the rows in which a pushed key predicate can hold (column index)
  int64_t pruned_buckets = 0;  // 1: an equality filter on the whole key selects another hash bucket than this reader's
  double decode_ms = 0, upload_ms = 0, merge_ms = 0, gather_ms = 0, download_ms = 0;
  std::string path;  // "merge-full" | "merge-general" | "concat"
};

class Reader {
 public:
  explicit Reader(IOConfig cfg);  // throws std::runtime_error("LakeSoulReader has the wrong number of files")
  ~Reader();
  void Start();                                   // throws on error
  std::shared_ptr<arrow::Schema> schema() const;  // valid after Start()
  std::shared_ptr<arrow::RecordBatch> Next();     // nullptr == end of stream; never an empty batch
  const ReaderStats& stats() const;
  const IOConfig& config() const;

 private:
  struct Impl;
  std::unique_ptr<Impl> impl_;
};

// Key ranges for streaming a bucket (reader.cc): filters of the passes over ascending ranges of the first primary-key
// column so that every pass decodes about `pass_bytes` (uncompressed) -- cut at row-group minima, from footers only.
// Empty when the bucket cannot be split (no statistics, unsupported key type, one key value) or fits one pass.
std::vector<std::shared_ptr<FilterExpr>> plan_stream_passes(const std::vector<std::string>& paths, const std::string& pk,
                                                            bool pk_nullable, int64_t pass_bytes);

// device selection for this process (LOCAL_RANK / LAKESOUL_B200_DEVICE aware); throws when no GPU.
int ensure_device();

}  // namespace lsb
