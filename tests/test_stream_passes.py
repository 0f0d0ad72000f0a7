# SPDX-License-Identifier: Apache-2.0
"""Pass planning for buckets streamed by primary-key range (reader.cc plan_stream_passes): from footers only, the key
space of the first primary-key column is cut at row-group minima so that every pass decodes about the pass budget.
The passes must be disjoint, ascending and cover every key (nulls in the first pass)."""
import ctypes
import os
import re

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq
import pytest

from lakesoul_b200 import load_library  # noqa: E402


def _passes(files, pk, pass_bytes, opts=""):
    lib = load_library()
    buf = ctypes.create_string_buffer(1 << 20)
    n = lib.lsb_test_stream_passes("\n".join(files).encode(), pk.encode(), opts.encode(), pass_bytes, buf, len(buf))
    assert n >= 0, buf.value.decode()
    return [ln for ln in buf.value.decode().split("\n") if ln]


def _bucket(tmp_path, keys_per_file, make_key=lambda a: pa.array(a), rg=2000):
    files = []
    for i, keys in enumerate(keys_per_file):
        k = make_key(np.sort(keys))
        t = pa.table([k, pa.array(np.arange(len(keys), dtype=np.int64))],
                     schema=pa.schema([pa.field("k", k.type, nullable=False), pa.field("v", pa.int64())]))
        p = str(tmp_path / f"part-{i:04d}_0000.parquet")
        pq.write_table(t, p, row_group_size=rg, compression="ZSTD", use_dictionary=False)
        files.append(p)
    return files


def _bounds(lines, parse=int):
    """[(lo, hi)] with None for an open side; checks the shape of every line"""
    out = []
    for i, ln in enumerate(lines):
        m_ge = re.search(r"gteq\(k, (.*?)\)", ln)
        m_lt = re.search(r"lt\(k, (.*?)\)", ln)
        lo = parse(m_ge.group(1)) if m_ge else None
        hi = parse(m_lt.group(1)) if m_lt else None
        assert (lo is None) == (i == 0), ln
        assert (hi is None) == (i == len(lines) - 1), ln
        out.append((lo, hi))
    return out


def test_int_key_passes_partition_the_key_space(tmp_path):
    rng = np.random.default_rng(1)
    base = np.arange(0, 40000, dtype=np.int64)
    deltas = [rng.choice(60000, 3000, replace=False).astype(np.int64) for _ in range(6)]
    files = _bucket(tmp_path, [base] + deltas)
    total = sum(pq.ParquetFile(f).metadata.row_group(g).total_byte_size for f in files
                for g in range(pq.ParquetFile(f).metadata.num_row_groups))
    lines = _passes(files, "k", total // 5)
    assert 3 <= len(lines) <= 10, lines
    b = _bounds(lines)
    for (lo0, hi0), (lo1, hi1) in zip(b, b[1:]):
        assert hi0 == lo1 and (lo0 is None or lo0 < hi0)
    # cuts are row-group minima of some file
    mins = set()
    for f in files:
        md = pq.ParquetFile(f).metadata
        for g in range(md.num_row_groups):
            mins.add(md.row_group(g).column(0).statistics.min)
    assert all(hi in mins for _, hi in b[:-1])
    assert "isnull" not in lines[0]
    # one pass when the bucket fits the budget; no split of a single key value
    assert _passes(files, "k", total * 2) == []
    same = _bucket(tmp_path / "same", [np.full(5000, 7, dtype=np.int64)] * 3) if (tmp_path / "same").mkdir() is None else None
    assert _passes(same, "k", 1000) == []


def test_string_key_and_nullable_first_pass(tmp_path):
    rng = np.random.default_rng(2)
    keys = [np.array(["key%07d" % i for i in rng.choice(100000, 8000, replace=False)]) for _ in range(4)]
    files = _bucket(tmp_path, keys, make_key=lambda a: pa.array(a.tolist(), pa.string()))
    lines = _passes(files, "k", 60000, "nullable")
    assert len(lines) >= 3
    assert lines[0].startswith("or(lt(k, ") and lines[0].endswith("isnull(k))"), lines[0]
    b = _bounds(lines, parse=str)
    assert all(x[1] == y[0] for x, y in zip(b, b[1:]))
    assert [x[1] for x in b[:-1]] == sorted(x[1] for x in b[:-1])


def test_unsplittable_buckets_stay_one_scan(tmp_path):
    # float keys: not a literal type the planner cuts on; a file without the key column; no statistics
    f1 = _bucket(tmp_path, [np.arange(9000, dtype=np.int64)] * 2, make_key=lambda a: pa.array(a.astype(np.float64)))
    assert _passes(f1, "k", 1000) == []
    (tmp_path / "b").mkdir()
    p = str(tmp_path / "b" / "nostat.parquet")
    pq.write_table(pa.table({"k": pa.array(np.arange(9000, dtype=np.int64))}), p, row_group_size=1000, write_statistics=False)
    assert _passes([p], "k", 1000) == []
    assert _passes([p], "missing", 1000) == []


@pytest.mark.parametrize("seed", range(8))
def test_random_buckets_give_ascending_disjoint_passes(tmp_path, seed):
    """random file counts / sizes / overlaps and budgets: the cuts ascend strictly, every cut is a row-group minimum, no
    pass is closed before it finished half a budget of row groups (except the last), one scan when nothing can be cut"""
    rng = np.random.default_rng(500 + seed)
    nfiles = int(rng.integers(1, 9))
    keys = []
    for i in range(nfiles):
        span = int(rng.integers(2000, 90000))
        off = int(rng.integers(0, 50000))
        keys.append(off + np.sort(rng.choice(span, int(rng.integers(1000, min(span, 20000))), replace=False)).astype(np.int64))
    files = _bucket(tmp_path, keys, rg=int(rng.integers(700, 5000)))
    mds = [pq.ParquetFile(f).metadata for f in files]
    rgs = [(md.row_group(g).column(0).statistics.min, md.row_group(g).column(0).statistics.max, md.row_group(g).total_byte_size)
           for md in mds for g in range(md.num_row_groups)]
    total = sum(b for _, _, b in rgs)
    budget = int(total / rng.uniform(1.5, 12))
    lines = _passes(files, "k", budget)
    if not lines:
        return
    b = _bounds(lines)
    cuts = [hi for _, hi in b[:-1]]
    assert cuts == sorted(set(cuts))
    mins = {mn for mn, _, _ in rgs}
    assert all(c in mins for c in cuts) and min(mins) < cuts[0]
    los = [None] + cuts
    for lo, hi in zip(los, cuts):  # every closed pass finishes at least half a budget of row groups
        done = sum(by for mn, mx, by in rgs if mx < hi and (lo is None or mx >= lo))
        assert done >= budget // 2, (lo, hi, done, budget)
