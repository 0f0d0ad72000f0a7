// SPDX-License-Identifier: Apache-2.0
//
// bucket_hash.h -- Spark-compatible murmur3_32 of primary-key literals, for hash-bucket elimination
// (SURVEY.md 8f N2): a row lives in bucket murmur3(pk columns, chained, seed 42) mod bucket_num
// (reference rust/lakesoul-io/src/utils/hash/mod.rs:22-178, spark_murmur3.rs:42-87, repartition/mod.rs:260), so a
// reader whose files all belong to another bucket than the one an equality filter on the whole key selects has
// nothing to return (the reference drops such files while planning, helpers/mod.rs:929-1065).
#pragma once
#include <arrow/api.h>

#include <cstdint>
#include <cstring>
#include <string>

namespace lsb {

constexpr uint32_t kHashSeed = 42;  // utils/hash/mod.rs:22

inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

// spark_murmur3.rs:42-87: 4-byte little-endian words, then every trailing byte mixed as its own word
inline uint32_t spark_murmur3_32(const uint8_t* p, size_t n, uint32_t seed) {
  auto mix = [](uint32_t state, uint32_t k) {
    k *= 0xcc9e2d51u;
    k = rotl32(k, 15);
    k *= 0x1b873593u;
    state ^= k;
    state = rotl32(state, 13);
    return state * 5u + 0xe6546b64u;
  };
  uint32_t state = seed;
  size_t i = 0;
  for (; i + 4 <= n; i += 4) {
    uint32_t k;
    memcpy(&k, p + i, 4);
    state = mix(state, k);
  }
  for (; i < n; ++i) state = mix(state, (uint32_t)p[i]);
  uint32_t h = state ^ (uint32_t)n;
  h ^= h >> 16;
  h *= 0x85ebca6bu;
  h ^= h >> 13;
  h *= 0xc2b2ae35u;
  h ^= h >> 16;
  return h;
}

// HashValue::hash_one for a literal of the key column's type (utils/hash/mod.rs:46-120). Returns false for types the
// elimination does not cover (the reader then simply reads the bucket).
inline bool hash_scalar(const arrow::Scalar& s, uint32_t seed, uint32_t* out) {
  if (!s.is_valid) return false;
  using T = arrow::Type;
  auto h32 = [&](uint32_t v) { *out = spark_murmur3_32(reinterpret_cast<const uint8_t*>(&v), 4, seed); return true; };
  auto h64 = [&](uint64_t v) { *out = spark_murmur3_32(reinterpret_cast<const uint8_t*>(&v), 8, seed); return true; };
  switch (s.type->id()) {
    case T::BOOL: return h32(static_cast<const arrow::BooleanScalar&>(s).value ? 1u : 0u);
    case T::INT8: return h32((uint32_t)(int32_t) static_cast<const arrow::Int8Scalar&>(s).value);
    case T::INT16: return h32((uint32_t)(int32_t) static_cast<const arrow::Int16Scalar&>(s).value);
    case T::INT32: return h32((uint32_t) static_cast<const arrow::Int32Scalar&>(s).value);
    case T::UINT8: return h32(static_cast<const arrow::UInt8Scalar&>(s).value);
    case T::UINT16: return h32(static_cast<const arrow::UInt16Scalar&>(s).value);
    case T::UINT32: return h32(static_cast<const arrow::UInt32Scalar&>(s).value);
    case T::DATE32: return h32((uint32_t) static_cast<const arrow::Date32Scalar&>(s).value);
    case T::INT64: return h64((uint64_t) static_cast<const arrow::Int64Scalar&>(s).value);
    case T::UINT64: return h64(static_cast<const arrow::UInt64Scalar&>(s).value);
    case T::FLOAT: {
      const float f = static_cast<const arrow::FloatScalar&>(s).value;
      uint32_t bits = 0;
      if (!(f == 0.0f)) memcpy(&bits, &f, 4);  // -0.0 hashes like 0.0 (mod.rs:76-88)
      return h32(bits);
    }
    case T::DOUBLE: {
      const double d = static_cast<const arrow::DoubleScalar&>(s).value;
      uint64_t bits = 0;
      if (!(d == 0.0)) memcpy(&bits, &d, 8);
      return h64(bits);
    }
    case T::STRING:
    case T::BINARY: {
      const auto& b = static_cast<const arrow::BaseBinaryScalar&>(s).value;
      if (!b) return false;
      *out = spark_murmur3_32(b->data(), (size_t)b->size(), seed);
      return true;
    }
    default: return false;  // timestamps / decimals: literal units and scales differ from the stored values
  }
}

// bucket id carried by a LakeSoul data file name: part-<16 alnum>_<bucket:04>.parquet (utils/mod.rs:30-32); -1 = none
inline int bucket_id_of_file(const std::string& path) {
  const size_t dot = path.rfind(".parquet");
  if (dot == std::string::npos || dot + 8 != path.size()) return -1;
  size_t us = path.rfind('_', dot);
  if (us == std::string::npos || us + 1 >= dot) return -1;
  int v = 0;
  for (size_t i = us + 1; i < dot; ++i) {
    if (path[i] < '0' || path[i] > '9') return -1;
    v = v * 10 + (path[i] - '0');
    if (v > 1000000) return -1;
  }
  return v;
}

}  // namespace lsb
