// SPDX-License-Identifier: Apache-2.0
// POD description of a predicate for the device interpreter (no arrow types: shared by g++ and nvcc).
#pragma once
#include <cstdint>

namespace lsb {

enum FilterOpcode : int32_t { FO_CMP = 0, FO_ISNULL, FO_ISNOTNULL, FO_AND, FO_OR, FO_NOT, FO_TRUE, FO_FALSE, FO_NULL };
enum FilterClass : int32_t { FC_INT = 0, FC_UINT, FC_F32, FC_F64, FC_BOOL, FC_DEC128, FC_BYTES };

struct FilterColDev {
  const void* values;
  const uint8_t* validity;
  const uint8_t* data;
  const int64_t* bucket_data;
  const uint32_t* win;
  int32_t width;
  int32_t cls;
  int64_t mul;  // integer columns are scaled by mul/div before the comparison (timestamp units)
  int64_t div;
};

struct FilterInstr {
  int32_t opcode;
  int32_t col;
  int32_t cmp;      // FilterExpr::Op
  int32_t lit_len;  // bytes literal length
  int64_t lit_i;
  uint64_t lit_u;
  double lit_f;
  uint64_t lit_hi, lit_lo;  // decimal128 two's complement
  int64_t lit_ofs;          // offset of a bytes literal inside the literal pool
};

constexpr int kMaxFilterInstr = 64;
constexpr int kMaxFilterCols = 32;

struct FilterProgram {
  int32_t n_instr;
  int32_t n_cols;
  FilterInstr instr[kMaxFilterInstr];
  FilterColDev cols[kMaxFilterCols];
  const uint8_t* lit_pool;  // device
};

}  // namespace lsb
