// SPDX-License-Identifier: Apache-2.0
#include "filter_eval.h"

#include <arrow/util/decimal.h>

#include <cstring>
#include <stdexcept>

#include "filter_prog.h"

namespace lsb {

void launch_filter_kernel(const FilterProgram* d_prog, const uint32_t* plan_win, const int64_t* out_base,
                          int nb, int64_t U, uint8_t* keep, cudaStream_t stream);

namespace {

[[noreturn]] void fail(const std::string& m) { throw std::runtime_error(m); }

struct Builder {
  FilterProgram prog{};
  std::string pool;
  const std::vector<FilterColumnIn>* cols = nullptr;
  std::vector<int> col_slot;  // physical column -> slot in prog.cols or -1

  int slot_for(const std::string& name) {
    int c = -1;
    for (size_t i = 0; i < cols->size(); ++i)
      if ((*cols)[i].name == name) { c = (int)i; break; }
    if (c < 0) fail("filter column '" + name + "' is not carried through the merge");
    if (col_slot[c] >= 0) return col_slot[c];
    if (prog.n_cols >= kMaxFilterCols) fail("filter references too many columns");
    const FilterColumnIn& in = (*cols)[c];
    FilterColDev d{};
    d.values = in.values;
    d.validity = in.validity;
    d.data = in.data;
    d.bucket_data = in.bucket_data;
    d.win = in.win;
    d.width = in.kind.width;
    d.mul = 1;
    d.div = 1;
    using T = arrow::Type;
    switch (in.type->id()) {
      case T::INT8: case T::INT16: case T::INT32: case T::INT64: case T::DATE32: d.cls = FC_INT; break;
      case T::TIMESTAMP: {
        d.cls = FC_INT;
        switch (static_cast<const arrow::TimestampType&>(*in.type).unit()) {
          case arrow::TimeUnit::SECOND: d.mul = 1000000; break;
          case arrow::TimeUnit::MILLI: d.mul = 1000; break;
          case arrow::TimeUnit::MICRO: d.mul = 1; break;
          default: d.mul = 1; d.div = -1000; break;  // ns: the literal is scaled instead (marker)
        }
      } break;
      case T::FLOAT: d.cls = FC_F32; break;
      case T::DOUBLE: d.cls = FC_F64; break;
      case T::BOOL: d.cls = FC_BOOL; break;
      case T::DECIMAL128: d.cls = FC_DEC128; break;
      case T::STRING: case T::BINARY: d.cls = FC_BYTES; break;
      default:
        fail("filter on column '" + name + "' of type " + in.type->ToString() +
             " is not supported by the GPU reader");
    }
    prog.cols[prog.n_cols] = d;
    col_slot[c] = prog.n_cols;
    return prog.n_cols++;
  }

  void emit(const FilterInstr& in) {
    if (prog.n_instr >= kMaxFilterInstr) fail("filter expression too large");
    prog.instr[prog.n_instr++] = in;
  }

  void build(const FilterExpr& e) {
    FilterInstr in{};
    switch (e.kind) {
      case FilterExpr::kAnd: build(*e.left); build(*e.right); in.opcode = FO_AND; emit(in); return;
      case FilterExpr::kOr: build(*e.left); build(*e.right); in.opcode = FO_OR; emit(in); return;
      case FilterExpr::kNot: build(*e.right); in.opcode = FO_NOT; emit(in); return;
      case FilterExpr::kTrue: in.opcode = FO_TRUE; emit(in); return;
      case FilterExpr::kFalse: in.opcode = FO_FALSE; emit(in); return;
      case FilterExpr::kNull: in.opcode = FO_NULL; emit(in); return;
      case FilterExpr::kIsNull: in.opcode = FO_ISNULL; in.col = slot_for(e.column); emit(in); return;
      case FilterExpr::kIsNotNull: in.opcode = FO_ISNOTNULL; in.col = slot_for(e.column); emit(in); return;
      case FilterExpr::kCmp: break;
    }
    in.opcode = FO_CMP;
    in.col = slot_for(e.column);
    in.cmp = (int)e.op;
    const FilterColDev& d = prog.cols[in.col];
    const arrow::Scalar& lit = *e.literal;
    if (!lit.is_valid) {  // comparison with a NULL literal is NULL for every row
      FilterInstr f{};
      f.opcode = FO_NULL;
      emit(f);
      return;
    }
    using T = arrow::Type;
    switch (lit.type->id()) {
      case T::INT8: in.lit_i = static_cast<const arrow::Int8Scalar&>(lit).value; break;
      case T::INT16: in.lit_i = static_cast<const arrow::Int16Scalar&>(lit).value; break;
      case T::INT32: in.lit_i = static_cast<const arrow::Int32Scalar&>(lit).value; break;
      case T::INT64: in.lit_i = static_cast<const arrow::Int64Scalar&>(lit).value; break;
      case T::DATE32: in.lit_i = static_cast<const arrow::Date32Scalar&>(lit).value; break;
      case T::TIMESTAMP:
        in.lit_i = static_cast<const arrow::TimestampScalar&>(lit).value;
        if (d.div == -1000) in.lit_i *= 1000;
        break;
      case T::BOOL: in.lit_i = static_cast<const arrow::BooleanScalar&>(lit).value ? 1 : 0; break;
      case T::FLOAT: in.lit_f = static_cast<const arrow::FloatScalar&>(lit).value; break;
      case T::DOUBLE: in.lit_f = static_cast<const arrow::DoubleScalar&>(lit).value; break;
      case T::DECIMAL128: {
        const auto& v = static_cast<const arrow::Decimal128Scalar&>(lit).value;
        in.lit_hi = (uint64_t)v.high_bits();
        in.lit_lo = v.low_bits();
      } break;
      case T::STRING: case T::BINARY: {
        const auto& buf = static_cast<const arrow::BaseBinaryScalar&>(lit).value;
        if (d.cls != FC_BYTES)
          fail("filter literal for column '" + e.column + "' could not be parsed for its type");
        in.lit_ofs = (int64_t)pool.size();
        in.lit_len = (int32_t)buf->size();
        pool.append(reinterpret_cast<const char*>(buf->data()), (size_t)buf->size());
      } break;
      default: fail("unsupported filter literal type " + lit.type->ToString());
    }
    emit(in);
  }
};

}  // namespace

int apply_filter(MergePlan* plan, const FilterExpr& expr, const std::vector<FilterColumnIn>& cols,
                 const std::vector<uint32_t*>& extra_lists, cudaStream_t stream) {
  Builder b;
  b.cols = &cols;
  b.col_slot.assign(cols.size(), -1);
  b.build(expr);
  const int nb = plan_num_buckets(plan);
  const int64_t U = plan_out_base(plan)[nb];
  if (U == 0) return 0;
  uint8_t* d_pool = nullptr;
  FilterProgram* d_prog = nullptr;
  uint8_t* d_keep = nullptr;
  auto check = [](cudaError_t e, const char* what) {
    if (e != cudaSuccess) fail(std::string(what) + ": " + cudaGetErrorString(e));
  };
  d_pool = static_cast<uint8_t*>(device_alloc_bytes(b.pool.size() + 16, stream));
  d_prog = static_cast<FilterProgram*>(device_alloc_bytes(sizeof(FilterProgram), stream));
  d_keep = static_cast<uint8_t*>(device_alloc_bytes((size_t)U, stream));
  if (!b.pool.empty())
    check(h2d_staged(d_pool, b.pool.data(), b.pool.size(), stream), "memcpy pool");
  b.prog.lit_pool = d_pool;
  check(h2d_staged(d_prog, &b.prog, sizeof(FilterProgram), stream), "memcpy prog");
  launch_filter_kernel(d_prog, plan_winners(plan), plan_out_base_dev(plan), nb, U, d_keep, stream);
  plan_compact(plan, d_keep, extra_lists, stream);
  plan_sync_counts(plan, stream);
  device_free_bytes(d_pool, stream);
  device_free_bytes(d_prog, stream);
  device_free_bytes(d_keep, stream);
  return 1 + 8 + 2 * (int)extra_lists.size();
}

}  // namespace lsb
