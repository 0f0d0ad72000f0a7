// SPDX-License-Identifier: Apache-2.0
//
// filter.cc -- parser for the filter-string grammar `op(lhs, rhs)`
// (behavioural mirror of rust/lakesoul-io/src/filter/parser.rs:50-265; written from the
// grammar, not translated).
#include "filter.h"

#include <arrow/util/decimal.h>

#include <cstdlib>
#include <stdexcept>

namespace lsb {

namespace {

[[noreturn]] void bad(const std::string& msg) { throw std::runtime_error(msg); }

std::string trim(const std::string& s) {
  size_t a = 0, b = s.size();
  while (a < b && isspace((unsigned char)s[a])) ++a;
  while (b > a && isspace((unsigned char)s[b - 1])) --b;
  return s.substr(a, b - a);
}

// split "op(left, right)" at the first top-level comma (parser.rs:92-135)
void split_call(const std::string& filter, std::string* op, std::string* left, std::string* right) {
  const size_t lp = filter.find('(');
  if (lp == std::string::npos) bad("wrong filter str");
  if (filter.empty() || filter.back() != ')') bad("wrong filter str");
  *op = filter.substr(0, lp);
  const std::string inner = filter.substr(lp + 1, filter.size() - lp - 2);
  int depth = 0;
  size_t comma = 0;
  for (size_t i = 0; i < inner.size(); ++i) {
    const char ch = inner[i];
    if (ch == '(') ++depth;
    else if (ch == ')') --depth;
    else if (ch == ',' && depth == 0 && comma == 0) comma = i;
  }
  if (depth != 0) bad("Invalid filter string");
  if (*op == "not") {
    *left = trim(inner.substr(0, comma));
    *right = trim(inner.substr(comma));
  } else {
    *left = trim(inner.substr(0, comma));
    *right = comma + 1 <= inner.size() ? trim(inner.substr(comma + 1)) : std::string();
  }
}

// "Binary{[1, -2, 3]}"-style byte lists (parser.rs:232-265): none when there is no '['
bool parse_byte_list(const std::string& v, std::string* out) {
  const size_t l = v.find('[');
  const size_t r = v.find(']');
  const size_t lpos = l == std::string::npos ? 0 : l;
  const size_t rpos = r == std::string::npos ? 0 : r;
  if (lpos == 0) return false;
  out->clear();
  if (lpos + 1 == rpos) return true;
  std::string body = v.substr(lpos + 1, rpos - lpos - 1);
  std::string tok;
  auto flush = [&]() {
    if (tok.empty()) bad("cannot parse byte list in filter literal");
    char* end = nullptr;
    long x = strtol(tok.c_str(), &end, 10);
    if (*end != 0 || x < -32768 || x > 32767) bad("cannot parse byte list in filter literal");
    if (x < 0) x += 256;
    out->push_back((char)(uint8_t)x);
    tok.clear();
  };
  for (char ch : body) {
    if (ch == ' ') continue;
    if (ch == ',') flush(); else tok.push_back(ch);
  }
  flush();
  return true;
}

template <typename T>
T parse_int(const std::string& s) {
  if (s.empty()) bad("cannot parse integer literal ''");
  errno = 0;
  char* end = nullptr;
  long long v = strtoll(s.c_str(), &end, 10);
  if (*end != 0 || errno != 0) bad("cannot parse integer literal '" + s + "'");
  if (v < (long long)std::numeric_limits<T>::min() || v > (long long)std::numeric_limits<T>::max())
    bad("integer literal out of range '" + s + "'");
  return (T)v;
}

double parse_f64(const std::string& s) {
  char* end = nullptr;
  double v = strtod(s.c_str(), &end);
  if (s.empty() || *end != 0) bad("cannot parse float literal '" + s + "'");
  return v;
}

std::shared_ptr<arrow::Scalar> parse_literal(const arrow::Field& field, const std::string& value) {
  const auto& t = field.type();
  switch (t->id()) {
    case arrow::Type::DECIMAL128: {
      const auto& dt = static_cast<const arrow::Decimal128Type&>(*t);
      if (dt.precision() <= 18) {
        // unscaled integer
        auto r = arrow::Decimal128::FromString(value);
        if (!r.ok()) bad("cannot parse decimal literal '" + value + "'");
        return std::make_shared<arrow::Decimal128Scalar>(*r, t);
      }
      std::string bytes;
      if (!parse_byte_list(value, &bytes)) bad("parse binary array failed");
      if (bytes.size() > 16) bad("decimal byte literal longer than 16 bytes");
      uint8_t be[16] = {0};
      for (size_t i = 0; i < bytes.size(); ++i) be[i + 16 - bytes.size()] = (uint8_t)bytes[i];
      uint64_t hi = 0, lo = 0;
      for (int i = 0; i < 8; ++i) hi = (hi << 8) | be[i];
      for (int i = 8; i < 16; ++i) lo = (lo << 8) | be[i];
      return std::make_shared<arrow::Decimal128Scalar>(arrow::Decimal128((int64_t)hi, lo), t);
    }
    case arrow::Type::BOOL: {
      if (value == "true") return std::make_shared<arrow::BooleanScalar>(true);
      if (value == "false") return std::make_shared<arrow::BooleanScalar>(false);
      bad("cannot parse bool literal '" + value + "'");
    }
    case arrow::Type::BINARY: {
      std::string bytes;
      if (!parse_byte_list(value, &bytes)) return arrow::MakeNullScalar(arrow::binary());
      return std::make_shared<arrow::BinaryScalar>(arrow::Buffer::FromString(bytes));
    }
    case arrow::Type::FLOAT: return std::make_shared<arrow::FloatScalar>((float)parse_f64(value));
    case arrow::Type::DOUBLE: return std::make_shared<arrow::DoubleScalar>(parse_f64(value));
    case arrow::Type::INT8: return std::make_shared<arrow::Int8Scalar>(parse_int<int8_t>(value));
    case arrow::Type::INT16: return std::make_shared<arrow::Int16Scalar>(parse_int<int16_t>(value));
    case arrow::Type::INT32: return std::make_shared<arrow::Int32Scalar>(parse_int<int32_t>(value));
    case arrow::Type::INT64: return std::make_shared<arrow::Int64Scalar>(parse_int<int64_t>(value));
    case arrow::Type::DATE32: return std::make_shared<arrow::Date32Scalar>(parse_int<int32_t>(value));
    case arrow::Type::TIMESTAMP:
      // the reference always builds a microsecond/UTC literal (parser.rs:212-218); the column is
      // compared after coercion to that unit, done at evaluation time.
      return std::make_shared<arrow::TimestampScalar>(parse_int<int64_t>(value),
                                                      arrow::timestamp(arrow::TimeUnit::MICRO, "UTC"));
    case arrow::Type::STRING: {
      // Binary{"text"}  ->  text   (parser.rs:220-223: strips 8 leading and 2 trailing chars)
      if (value.size() < 10) bad("malformed string literal '" + value + "'");
      return std::make_shared<arrow::StringScalar>(value.substr(8, value.size() - 10));
    }
    default:
      return std::make_shared<arrow::StringScalar>(value);
  }
}

}  // namespace

std::shared_ptr<FilterExpr> parse_filter(const std::string& filter, const arrow::Schema& schema) {
  std::string op, left, right;
  split_call(filter, &op, &left, &right);
  auto e = std::make_shared<FilterExpr>();
  if (op == "or" || op == "and") {
    e->kind = op == "or" ? FilterExpr::kOr : FilterExpr::kAnd;
    e->left = parse_filter(left, schema);
    e->right = parse_filter(right, schema);
    return e;
  }
  if (op == "not") {
    e->kind = FilterExpr::kNot;
    e->right = parse_filter(right, schema);
    return e;
  }
  auto field = schema.GetFieldByName(left);
  if (!field) {
    // nested (a.b.c) references resolve only for struct columns, which the device path does not
    // carry; an unknown column is the constant FALSE in the reference (parser.rs:86-88).
    e->kind = FilterExpr::kFalse;
    return e;
  }
  e->column = left;
  if (right == "null") {
    if (op == "eq") e->kind = FilterExpr::kIsNull;
    else if (op == "noteq") e->kind = FilterExpr::kIsNotNull;
    else e->kind = FilterExpr::kTrue;
    return e;
  }
  e->literal = parse_literal(*field, right);
  e->kind = FilterExpr::kCmp;
  if (op == "eq") e->op = FilterExpr::kEq;
  else if (op == "noteq") e->op = FilterExpr::kNe;
  else if (op == "gt") e->op = FilterExpr::kGt;
  else if (op == "gteq") e->op = FilterExpr::kGe;
  else if (op == "lt") e->op = FilterExpr::kLt;
  else if (op == "lteq") e->op = FilterExpr::kLe;
  else e->kind = FilterExpr::kTrue;
  return e;
}

std::shared_ptr<FilterExpr> conjunction(const std::vector<std::shared_ptr<FilterExpr>>& exprs) {
  std::shared_ptr<FilterExpr> acc;
  for (const auto& x : exprs) {
    if (!acc) { acc = x; continue; }
    auto e = std::make_shared<FilterExpr>();
    e->kind = FilterExpr::kAnd;
    e->left = acc;
    e->right = x;
    acc = e;
  }
  return acc;
}

void filter_columns(const FilterExpr& e, std::set<std::string>* out) {
  switch (e.kind) {
    case FilterExpr::kAnd:
    case FilterExpr::kOr:
      filter_columns(*e.left, out);
      filter_columns(*e.right, out);
      break;
    case FilterExpr::kNot: filter_columns(*e.right, out); break;
    case FilterExpr::kCmp:
    case FilterExpr::kIsNull:
    case FilterExpr::kIsNotNull: out->insert(e.column); break;
    default: break;
  }
}

std::string filter_to_string(const FilterExpr& e) {
  static const char* ops[] = {"eq", "noteq", "gt", "gteq", "lt", "lteq"};
  switch (e.kind) {
    case FilterExpr::kAnd: return "and(" + filter_to_string(*e.left) + ", " + filter_to_string(*e.right) + ")";
    case FilterExpr::kOr: return "or(" + filter_to_string(*e.left) + ", " + filter_to_string(*e.right) + ")";
    case FilterExpr::kNot: return "not(" + filter_to_string(*e.right) + ")";
    case FilterExpr::kCmp: return std::string(ops[e.op]) + "(" + e.column + ", " + e.literal->ToString() + ")";
    case FilterExpr::kIsNull: return "isnull(" + e.column + ")";
    case FilterExpr::kIsNotNull: return "isnotnull(" + e.column + ")";
    case FilterExpr::kTrue: return "true";
    case FilterExpr::kNull: return "null";
    default: return "false";
  }
}

namespace {

// three-way compare of a constant column value with a literal of the column's family; false = not comparable
bool compare_scalars(const arrow::Scalar& v, const arrow::Scalar& lit, int* out) {
  using T = arrow::Type;
  auto sgn = [](auto a, auto b) { return a < b ? -1 : (a > b ? 1 : 0); };
  auto as_i64 = [](const arrow::Scalar& s, int64_t* x) {
    switch (s.type->id()) {
      case T::INT8: *x = static_cast<const arrow::Int8Scalar&>(s).value; return true;
      case T::INT16: *x = static_cast<const arrow::Int16Scalar&>(s).value; return true;
      case T::INT32: *x = static_cast<const arrow::Int32Scalar&>(s).value; return true;
      case T::INT64: *x = static_cast<const arrow::Int64Scalar&>(s).value; return true;
      case T::UINT8: *x = static_cast<const arrow::UInt8Scalar&>(s).value; return true;
      case T::UINT16: *x = static_cast<const arrow::UInt16Scalar&>(s).value; return true;
      case T::UINT32: *x = static_cast<const arrow::UInt32Scalar&>(s).value; return true;
      case T::DATE32: *x = static_cast<const arrow::Date32Scalar&>(s).value; return true;
      case T::BOOL: *x = static_cast<const arrow::BooleanScalar&>(s).value ? 1 : 0; return true;
      case T::TIMESTAMP: {
        int64_t t = static_cast<const arrow::TimestampScalar&>(s).value;
        switch (static_cast<const arrow::TimestampType&>(*s.type).unit()) {  // literals are microseconds
          case arrow::TimeUnit::SECOND: t *= 1000000; break;
          case arrow::TimeUnit::MILLI: t *= 1000; break;
          case arrow::TimeUnit::NANO: t = t / 1000 - ((t % 1000 != 0 && t < 0) ? 1 : 0); break;
          default: break;
        }
        *x = t;
        return true;
      }
      default: return false;
    }
  };
  int64_t a, b;
  if (as_i64(v, &a) && as_i64(lit, &b)) { *out = sgn(a, b); return true; }
  auto as_f64 = [](const arrow::Scalar& s, double* x) {
    if (s.type->id() == T::FLOAT) { *x = static_cast<const arrow::FloatScalar&>(s).value; return true; }
    if (s.type->id() == T::DOUBLE) { *x = static_cast<const arrow::DoubleScalar&>(s).value; return true; }
    return false;
  };
  double fa, fb;
  if (as_f64(v, &fa) && as_f64(lit, &fb)) { *out = sgn(fa, fb); return fa == fa && fb == fb; }
  auto is_bytes = [](const arrow::Scalar& s) { return s.type->id() == T::STRING || s.type->id() == T::BINARY; };
  if (is_bytes(v) && is_bytes(lit)) {
    const auto& x = static_cast<const arrow::BaseBinaryScalar&>(v).value;
    const auto& y = static_cast<const arrow::BaseBinaryScalar&>(lit).value;
    if (!x || !y) return false;
    *out = sgn(x->ToString().compare(y->ToString()), 0);
    return true;
  }
  if (v.type->id() == T::DECIMAL128 && lit.type->id() == T::DECIMAL128) {
    const auto& x = static_cast<const arrow::Decimal128Scalar&>(v).value;
    const auto& y = static_cast<const arrow::Decimal128Scalar&>(lit).value;
    *out = x < y ? -1 : (x > y ? 1 : 0);
    return true;
  }
  return false;
}

std::shared_ptr<FilterExpr> constant(FilterExpr::Kind k) {
  auto e = std::make_shared<FilterExpr>();
  e->kind = k;
  return e;
}

}  // namespace

std::shared_ptr<FilterExpr> fold_constant_columns(
    const std::shared_ptr<FilterExpr>& e,
    const std::function<bool(const std::string&, std::shared_ptr<arrow::Scalar>*)>& constant_of) {
  if (!e) return e;
  switch (e->kind) {
    case FilterExpr::kAnd:
    case FilterExpr::kOr: {
      auto n = std::make_shared<FilterExpr>(*e);
      n->left = fold_constant_columns(e->left, constant_of);
      n->right = fold_constant_columns(e->right, constant_of);
      return n;
    }
    case FilterExpr::kNot: {
      auto n = std::make_shared<FilterExpr>(*e);
      n->right = fold_constant_columns(e->right, constant_of);
      return n;
    }
    case FilterExpr::kIsNull:
    case FilterExpr::kIsNotNull: {
      std::shared_ptr<arrow::Scalar> v;
      if (!constant_of(e->column, &v)) return e;
      const bool is_null = !v || !v->is_valid;
      return constant((e->kind == FilterExpr::kIsNull) == is_null ? FilterExpr::kTrue : FilterExpr::kFalse);
    }
    case FilterExpr::kCmp: {
      std::shared_ptr<arrow::Scalar> v;
      if (!constant_of(e->column, &v)) return e;
      if (!v || !v->is_valid || !e->literal || !e->literal->is_valid) return constant(FilterExpr::kNull);
      int c = 0;
      if (!compare_scalars(*v, *e->literal, &c))
        bad("filter on the constant column '" + e->column + "' of type " + v->type->ToString() + " cannot be evaluated");
      bool r = false;
      switch (e->op) {
        case FilterExpr::kEq: r = c == 0; break;
        case FilterExpr::kNe: r = c != 0; break;
        case FilterExpr::kGt: r = c > 0; break;
        case FilterExpr::kGe: r = c >= 0; break;
        case FilterExpr::kLt: r = c < 0; break;
        case FilterExpr::kLe: r = c <= 0; break;
      }
      return constant(r ? FilterExpr::kTrue : FilterExpr::kFalse);
    }
    default: return e;
  }
}

}  // namespace lsb
