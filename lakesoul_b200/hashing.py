# SPDX-License-Identifier: Apache-2.0
"""Spark-compatible murmur3_32 (seed 42) bucket assignment, vectorised with numpy.

Host-side helper of the synthetic data generator and the bucket sharder: restates
rust/lakesoul-io/src/utils/hash/mod.rs:22,64-74,138-178 (hash_bytes_by_int for 4-byte aligned
payloads; int32 hashed as 4 LE bytes, int64 as 8; multi-column keys chain the running hash as the
next seed; bucket = hash as u32 mod bucket_num, repartition/mod.rs:260).
"""
from __future__ import annotations

import numpy as np

SEED = 42
_C1, _C2 = np.uint32(0xCC9E2D51), np.uint32(0x1B873593)


def _rotl(x, r):
    return (x << np.uint32(r)) | (x >> np.uint32(32 - r))


def _mix_k1(k1):
    k1 = k1 * _C1
    k1 = _rotl(k1, 15)
    return k1 * _C2


def _mix_h1(h1, k1):
    h1 = h1 ^ k1
    h1 = _rotl(h1, 13)
    return h1 * np.uint32(5) + np.uint32(0xE6546B64)


def _fmix(h1, length):
    h1 = h1 ^ np.uint32(length)
    h1 = h1 ^ (h1 >> np.uint32(16))
    h1 = h1 * np.uint32(0x85EBCA6B)
    h1 = h1 ^ (h1 >> np.uint32(13))
    h1 = h1 * np.uint32(0xC2B2AE35)
    return h1 ^ (h1 >> np.uint32(16))


def hash_int32(values: np.ndarray, seed) -> np.ndarray:
    with np.errstate(over="ignore"):
        v = np.asarray(values).astype(np.int32).view(np.uint32)
        h1 = np.broadcast_to(np.asarray(seed, dtype=np.uint32), v.shape).copy()
        h1 = _mix_h1(h1, _mix_k1(v))
        return _fmix(h1, 4)


def hash_int64(values: np.ndarray, seed) -> np.ndarray:
    with np.errstate(over="ignore"):
        v = np.asarray(values).astype(np.int64).view(np.uint64)
        lo = (v & np.uint64(0xFFFFFFFF)).astype(np.uint32)
        hi = (v >> np.uint64(32)).astype(np.uint32)
        h1 = np.broadcast_to(np.asarray(seed, dtype=np.uint32), v.shape).copy()
        h1 = _mix_h1(h1, _mix_k1(lo))
        h1 = _mix_h1(h1, _mix_k1(hi))
        return _fmix(h1, 8)


def hash_columns(columns) -> np.ndarray:
    """Chained row hash over int32/int64 key columns (first column seeded with 42)."""
    h = np.uint32(SEED)
    for c in columns:
        c = np.asarray(c)
        h = hash_int64(c, h) if c.dtype.itemsize == 8 else hash_int32(c, h)
    return h


def bucket_of(columns, bucket_num: int) -> np.ndarray:
    return (hash_columns(columns).astype(np.uint64) % np.uint64(bucket_num)).astype(np.int64)
