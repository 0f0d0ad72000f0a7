# SPDX-License-Identifier: Apache-2.0
"""The write-side oracle (oracle/write_oracle.py) against the KAT-pinned vectorised hash of lakesoul_b200.hashing (G11) and
the reference's own known answers (rust/lakesoul-datafusion/src/tests/hash_tests.rs:48-90 through tests/golden)."""
import numpy as np
import pyarrow as pa

import write_oracle as wo
from lakesoul_b200 import hashing


def test_scalar_hash_matches_vectorised_hash():
    rng = np.random.default_rng(3)
    a = rng.integers(-2**62, 2**62, 500, dtype=np.int64)
    b = rng.integers(-2**31, 2**31 - 1, 500, dtype=np.int64).astype(np.int32)
    t = pa.table({"a": a, "b": b})
    assert wo.row_hashes(t, ["a"]) == [int(x) for x in hashing.hash_columns([a])]
    assert wo.row_hashes(t, ["a", "b"]) == [int(x) for x in hashing.hash_columns([a, b])]
    assert wo.buckets(t, ["b", "a"], 7) == [int(x) for x in hashing.bucket_of([b, a], 7)]


def test_reference_known_answers():
    """every assertion of hash_value_test, rust/lakesoul-datafusion/src/tests/hash_tests.rs:47-93, through value_bytes"""
    i32 = lambda x: int(np.uint32(np.int32(x)))  # noqa: E731
    H = lambda v, t, seed=42: wo.murmur3_32(wo.value_bytes(v, t), seed)  # noqa: E731
    for v, want in zip([1, 2, 3, 4], [-559580957, 1765031574, -1823081949, -397064898]):
        assert H(v, pa.int32()) == i32(want)
    for v, want in zip([1, 2, 3, 4], [-1712319331, -797927272, 519220707, 1344313940]):
        assert H(v, pa.uint64()) == i32(want)
    for v, want in zip([1.0, 2.0, 3.0, 4.0], [-466301895, 1199227445, 1710391653, -1959694433]):
        assert H(v, pa.float32()) == i32(want)
    for v, want in zip([1.0, 2.0, 3.0, 4.0], [-460888942, -2030303457, 1075969934, 1290556682]):
        assert H(v, pa.float64()) == i32(want)
    for v, want in zip(["1", "2", "3", "4"], [1625004744, 870267989, -1756013582, -2142269034]):
        assert H(v, pa.string()) == i32(want)
    assert H("321", pa.string(), H("321", pa.string())) == i32(-218318595)
    assert H("12", pa.string(), H("1", pa.string())) == i32(891492135)
    assert H("22", pa.string(), H("2", pa.string())) == i32(1475972200)
    assert H(0.0, pa.float32()) == H(-0.0, pa.float32()) == 933211791
    assert H(0.0, pa.float64()) == H(-0.0, pa.float64()) == i32(-1670924195)
    assert H(49, pa.uint8()) == H(49, pa.int32()) == 766678906
    assert H(False, pa.bool_()) == 933211791
    assert H(1065353216, pa.int32()) == i32(-466301895)


def test_nulls_leave_the_running_hash():
    t = pa.table({"a": pa.array([None, 5], pa.int32()), "s": pa.array(["x", None])})
    h = wo.row_hashes(t, ["a", "s"])
    assert h[0] == wo.murmur3_32(b"x", 0)          # null first column: the buffer's initial 0 seeds the next column
    assert h[1] == wo.murmur3_32((5).to_bytes(4, "little"), 42)
