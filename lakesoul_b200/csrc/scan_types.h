// SPDX-License-Identifier: Apache-2.0
// Small POD types shared by host (g++) and device (nvcc) translation units.
#pragma once
#include <cstdint>

namespace lsb {

struct ColumnKind {
  int width = -1;       // bytes; 0 = boolean bits; -1 = not fixed
  bool var = false;     // Utf8 / Binary (int32 offsets)
  bool ok = false;      // supported as a payload column
  bool pk_ok = false;   // supported as a primary-key column
  uint8_t pk_kind = 0;  // PkKind
};

// reference rust/lakesoul-io/src/physical_plan/merge/sorted/merge_operator.rs:22-32
enum class MergeOp : uint8_t {
  UseLast, UseLastNotNull, SumAll, SumLast,
  JoinedLastByComma, JoinedLastBySemicolon, JoinedAllByComma, JoinedAllBySemicolon
};

// numeric class of a column for SumAll / SumLast (merge_operator.rs:197-387)
enum NumClass : int32_t { NC_NONE = 0, NC_I8, NC_I16, NC_I32, NC_I64, NC_U8, NC_U16, NC_U32, NC_U64, NC_F32, NC_F64 };

}  // namespace lsb
