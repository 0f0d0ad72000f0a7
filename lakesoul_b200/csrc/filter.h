// SPDX-License-Identifier: Apache-2.0
//
// filter.h -- the parquet-mr FilterPredicate string grammar the JVM readers hand to the
// native reader (reference rust/lakesoul-io/src/filter/parser.rs:50-265), parsed into a
// small expression tree that the device evaluates on the merged rows (post-merge FilterExec,
// reference session.rs:839-877) with SQL three-valued logic: a row survives only when the
// predicate is TRUE.
#pragma once
#include <arrow/api.h>

#include <functional>
#include <memory>
#include <set>
#include <string>

namespace lsb {

struct FilterExpr {
  enum Kind : int { kAnd, kOr, kNot, kCmp, kIsNull, kIsNotNull, kTrue, kFalse, kNull };  // kNull: the constant NULL
  enum Op : int { kEq, kNe, kGt, kGe, kLt, kLe };
  Kind kind = kTrue;
  Op op = kEq;
  std::shared_ptr<FilterExpr> left, right;  // kNot uses `right` like the reference parser
  std::string column;                       // kCmp / kIsNull / kIsNotNull
  std::shared_ptr<arrow::Scalar> literal;   // kCmp
};

// Throws std::runtime_error on malformed strings ("wrong filter str") or unparsable literals.
std::shared_ptr<FilterExpr> parse_filter(const std::string& filter, const arrow::Schema& schema);
std::shared_ptr<FilterExpr> conjunction(const std::vector<std::shared_ptr<FilterExpr>>& exprs);
void filter_columns(const FilterExpr& e, std::set<std::string>* out);
std::string filter_to_string(const FilterExpr& e);

// Replaces every predicate on a column whose value is the same for all rows of the scan -- a default-valued /
// partition column (stream/default_column.rs:62-108, helpers/transform.rs:356-513) or a column no file carries
// (all NULL) -- by the constant it evaluates to. The reference evaluates pushed filters over the merged schema
// AFTER those columns were filled in, so such filters are legal; the device never sees the constant columns.
// `constant_of(name, &scalar)` returns true when `name` is such a column (scalar: its value, null scalar = NULL).
std::shared_ptr<FilterExpr> fold_constant_columns(
    const std::shared_ptr<FilterExpr>& e,
    const std::function<bool(const std::string&, std::shared_ptr<arrow::Scalar>*)>& constant_of);

}  // namespace lsb
