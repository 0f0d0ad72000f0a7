# SPDX-License-Identifier: Apache-2.0
"""TEST INFRASTRUCTURE (oracle): scalar restatement of the reference's write-side row placement.

  * spark murmur3_32 of one value            rust/lakesoul-io/src/utils/hash/spark_murmur3.rs:42-87
  * per-type bytes that get hashed           utils/hash/mod.rs:46-136 (ints <= 32 bits as 4 LE bytes of the value cast to u32,
                                             64-bit as 8, i128 as 16, floats by bits with -0.0 -> 0, strings / binary bytes)
  * chaining over key columns, nulls skipped utils/hash/mod.rs:143-178, 304-361 (first column seeded with 42, the running hash
                                             seeds the next; the buffer starts at 0: physical_plan/repartition/mod.rs:247)
  * bucket = hash % n as u32                 physical_plan/repartition/mod.rs:260
  * order inside a (partition, bucket) file  SortExec over (range partitions, primary keys, aux sort columns), ascending,
                                             nulls first (writer/async_writer/partitioning_writer.rs:139-171, sort_writer.rs:59-76)
Pinned by the reference's hash known-answer tests G11 (tests/test_oracle_golden.py checks lakesoul_b200.hashing against
rust/lakesoul-datafusion/src/tests/hash_tests.rs:48-90; this module is checked against hashing.py in tests/test_write_oracle.py).
Only tests/ may import this module."""
from __future__ import annotations

import struct

import pyarrow as pa

M32 = 0xFFFFFFFF


def _rotl(x, r):
    return ((x << r) | (x >> (32 - r))) & M32


def _mix(state, k):
    k = (k * 0xCC9E2D51) & M32
    k = _rotl(k, 15)
    k = (k * 0x1B873593) & M32
    state ^= k
    state = _rotl(state, 13)
    return (state * 5 + 0xE6546B64) & M32


def murmur3_32(data: bytes, seed: int) -> int:
    state = seed & M32
    n = len(data)
    i = 0
    while i + 4 <= n:
        state = _mix(state, int.from_bytes(data[i:i + 4], "little"))
        i += 4
    while i < n:
        state = _mix(state, data[i])
        i += 1
    h = state ^ n
    h ^= h >> 16
    h = (h * 0x85EBCA6B) & M32
    h ^= h >> 13
    h = (h * 0xC2B2AE35) & M32
    h ^= h >> 16
    return h


def value_bytes(v, typ: pa.DataType) -> bytes:
    """bytes the reference hashes for one non-null value of an Arrow type"""
    t = typ
    if pa.types.is_boolean(t):
        return struct.pack("<I", 1 if v else 0)
    if pa.types.is_int8(t) or pa.types.is_int16(t) or pa.types.is_int32(t):
        return struct.pack("<I", v & M32)  # sign-extended to 32 bits, `as u32`
    if pa.types.is_uint8(t) or pa.types.is_uint16(t) or pa.types.is_uint32(t):
        return struct.pack("<I", v)
    if pa.types.is_date32(t) or pa.types.is_time32(t):
        return struct.pack("<i", v)
    if pa.types.is_int64(t) or pa.types.is_timestamp(t) or pa.types.is_date64(t) or pa.types.is_time64(t) or pa.types.is_duration(t):
        return struct.pack("<q", v)
    if pa.types.is_uint64(t):
        return struct.pack("<Q", v)
    if pa.types.is_float32(t):
        return struct.pack("<I", 0) if v == 0.0 else struct.pack("<f", v)
    if pa.types.is_float64(t):
        return struct.pack("<Q", 0) if v == 0.0 else struct.pack("<d", v)
    if pa.types.is_decimal128(t):
        return int(v).to_bytes(16, "little", signed=True)
    if pa.types.is_string(t):
        return v.encode() if isinstance(v, str) else bytes(v)
    if pa.types.is_binary(t):
        return bytes(v)
    raise NotImplementedError(str(t))


def _raw(col: pa.ChunkedArray | pa.Array):
    """python values with integers for the temporal / decimal types (what the hash sees)"""
    arr = col.combine_chunks() if isinstance(col, pa.ChunkedArray) else col
    t = arr.type
    if pa.types.is_timestamp(t) or pa.types.is_date64(t) or pa.types.is_time64(t) or pa.types.is_duration(t):
        return arr.cast(pa.int64()).to_pylist()
    if pa.types.is_date32(t) or pa.types.is_time32(t):
        return arr.cast(pa.int32()).to_pylist()
    if pa.types.is_decimal128(t):
        scale = t.scale
        return [None if v is None else int(v.scaleb(scale)) for v in arr.to_pylist()]
    return arr.to_pylist()


def row_hashes(table: pa.Table, key_cols) -> list:
    n = table.num_rows
    h = [0] * n
    for ci, name in enumerate(key_cols):
        typ = table.schema.field(name).type
        vals = _raw(table.column(name))
        for i, v in enumerate(vals):
            if v is None:
                continue
            h[i] = murmur3_32(value_bytes(v, typ), 42 if ci == 0 else h[i])
    return h


def buckets(table: pa.Table, key_cols, bucket_num: int) -> list:
    return [x % max(bucket_num, 1) for x in row_hashes(table, key_cols)]


def _sort_key(table: pa.Table, cols):
    """per row: tuple usable by sorted(); ascending, nulls first; floats in IEEE total order; strings bytewise"""
    import math
    keys = [[] for _ in range(table.num_rows)]
    for name in cols:
        typ = table.schema.field(name).type
        vals = _raw(table.column(name))
        for i, v in enumerate(vals):
            if v is None:
                keys[i].append((0, 0))
            elif pa.types.is_floating(typ):
                bits = struct.unpack("<q", struct.pack("<d", float(v)))[0]
                if pa.types.is_float32(typ):
                    b32 = struct.unpack("<i", struct.pack("<f", v))[0]
                    bits = b32 ^ ((b32 >> 31) & 0x7FFFFFFF)
                else:
                    bits = bits ^ ((bits >> 63) & 0x7FFFFFFFFFFFFFFF)
                keys[i].append((1, bits))
            elif isinstance(v, str):
                keys[i].append((1, v.encode()))
            else:
                keys[i].append((1, v))
    return [tuple(k) for k in keys]


def placed(table: pa.Table, range_cols, primary_keys, aux_cols, bucket_num: int, dynamic: bool):
    """-> list of (range values tuple, bucket, [row indices in file order]) in output order"""
    n = table.num_rows
    sk = _sort_key(table, list(range_cols) + list(primary_keys) + list(aux_cols))
    if not dynamic:
        order = sorted(range(n), key=lambda i: (sk[i], i))
        return [((), 0, order)]
    bk = buckets(table, primary_keys, bucket_num)
    rk = _sort_key(table, range_cols)
    order = sorted(range(n), key=lambda i: (rk[i], bk[i], sk[i], i))
    groups = []
    for i in order:
        g = (rk[i], bk[i])
        if not groups or groups[-1][0] != g:
            groups.append((g, []))
        groups[-1][1].append(i)
    return [(g[0], g[1], rows) for g, rows in groups]
