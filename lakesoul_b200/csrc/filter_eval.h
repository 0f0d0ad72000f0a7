// SPDX-License-Identifier: Apache-2.0
//
// filter_eval.h -- device evaluation of the post-merge predicate + stream compaction of the
// surviving winners (reference: FilterExec above the merge, session.rs:839-877; CDC form
// rust/lakesoul-datafusion/src/datasource/file_format/metadata_format.rs:310-320).
#pragma once
#include <arrow/api.h>
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "device_scan.h"
#include "filter.h"
#include "mor_engine.h"

namespace lsb {

struct FilterColumnIn {
  std::string name;
  std::shared_ptr<arrow::DataType> type;
  ColumnKind kind;
  const void* values = nullptr;
  const uint8_t* validity = nullptr;
  const uint8_t* data = nullptr;
  const int64_t* bucket_data = nullptr;
  const uint32_t* win = nullptr;  // per-column winners or nullptr
};

// Evaluates `expr` on every merged row, keeps rows where it is TRUE, compacts the plan's winner
// list (and `extra_lists`) and refreshes the per-bucket counts (plan is re-synced on return).
// Returns the number of kernels launched. Throws on unsupported column types.
int apply_filter(MergePlan* plan, const FilterExpr& expr, const std::vector<FilterColumnIn>& cols,
                 const std::vector<uint32_t*>& extra_lists, cudaStream_t stream);

}  // namespace lsb
