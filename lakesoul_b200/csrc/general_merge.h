// SPDX-License-Identifier: Apache-2.0
//
// general_merge.h -- the non-default merge paths:
//   * partial merge: delta files with different column sets (reference UseLastRangeCombiner<_, true>,
//     combiner.rs:612-667 + sort_key_range.rs:387-401: the winner is kept per target column);
//   * merge operators (reference MinHeapSortKeyBatchRangeCombiner + MergeOperator::merge,
//     combiner.rs:136-336, merge_operator.rs:55-579).
// Both are resolved from one "presence matrix" R[u][j] = last row of output key u inside file j
// (or kNoRow), produced by a second pass of the tile-merge kernel.
#pragma once
#include <cuda_runtime.h>

#include <vector>

#include "mor_engine.h"
#include "scan_types.h"

namespace lsb {

struct GeneralColumnIn {
  MergeOp op = MergeOp::UseLast;
  const char* present = nullptr;  // host, F entries: file carries the column
  ColumnKind kind;
  int num_class = 0;              // NumClass (SumAll / SumLast)
  const void* values = nullptr;   // device: values / int32 offsets / boolean bits
  const uint8_t* validity = nullptr;
  const uint8_t* data = nullptr;
  const int64_t* bucket_data = nullptr;
  int64_t data_bytes = 0;
};

struct GeneralMerge;
// requires plan_merge + plan_sync_counts on `plan`
GeneralMerge* general_create(MergePlan* plan, const PkDesc& pk, const std::vector<GeneralColumnIn>& cols,
                             cudaStream_t stream);
void general_destroy(GeneralMerge*);
int general_launches(const GeneralMerge*);
// per-column gather inputs: winner list (nullptr = the plan's winners) and source buffers
const uint32_t* general_winners(const GeneralMerge*, int c);
const void* general_values(const GeneralMerge*, int c, const void* dflt);
const uint8_t* general_validity(const GeneralMerge*, int c, const uint8_t* dflt);
const uint8_t* general_data(const GeneralMerge*, int c, const uint8_t* dflt);
const int64_t* general_bucket_data(const GeneralMerge*, int c, const int64_t* dflt);
int64_t general_data_bytes(const GeneralMerge*, int c, int64_t dflt);
bool general_may_be_null(const GeneralMerge*, int c);
std::vector<uint32_t*> general_winner_lists(GeneralMerge*);  // distinct per-column lists (for compaction)

}  // namespace lsb
