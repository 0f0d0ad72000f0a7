# SPDX-License-Identifier: Apache-2.0
"""Page-level pruning by the key columns' page index (page_plan.cc PrunePages; the reference enables the page index for
pushed filters, session.rs:685-717): the rows in which a pushed key predicate can hold are narrowed to one contiguous
range, only the pages overlapping it are framed -- for EVERY column, whose pages do not align with the key's -- and the
file's row space shrinks to the framed pages. Host logic only: which rows, which pages."""
import ctypes
import os

import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
import pyarrow.parquet as pq
import pytest

from lakesoul_b200 import load_library  # noqa: E402

N = 30000


def _prune_once(path, flt, pks, lazy):
    lib = load_library()
    buf = ctypes.create_string_buffer(1 << 22)
    n = lib.lsb_test_page_prune(path.encode(), flt.encode(), pks.encode(), lazy, buf, len(buf))
    assert n >= 0, buf.value.decode()
    lines = buf.value.decode().splitlines()
    head = lines[0].split()
    info = {head[i]: int(head[i + 1]) for i in range(0, len(head), 2)}
    cols = {}
    for ln in lines[1:]:
        parts = ln.split()
        if parts[0] == "K":
            info["kept_rgs"] = [int(x) for x in parts[1:]]
            continue
        cols[parts[1]] = [tuple(int(x) for x in p.split(":")) for p in parts[2:]]
    return info, cols


def _prune(path, flt, pks):
    """eager framing, checked against the lazy one: same live range, same pages, same payload bytes, fewer bytes read"""
    info, cols = _prune_once(path, flt, pks, 0)
    linfo, lcols = _prune_once(path, flt, pks, 1)
    assert lcols == cols, (flt, "lazy framing differs")
    for k in ("dropped_rg", "kept_rows", "live_begin", "live_rows", "live_offset", "layout_rows", "pruned_pages", "kept_rgs"):
        assert info[k] == linfo[k], (flt, k, info, linfo)
    assert info["image_bytes"] == info["file_bytes"] and linfo["image_bytes"] <= info["file_bytes"]
    info["lazy_image_bytes"] = linfo["image_bytes"]
    return info, {name: [(a, n) for a, n, _ in pages] for name, pages in cols.items()}


def _file(tmp_path, name="f.parquet", page_index=True, **kw):
    rng = np.random.default_rng(3)
    t = pa.table({
        "k": pa.array(np.arange(N, dtype=np.int64) * 3),
        "k2": pa.array(["id%07d" % i for i in range(N)], pa.string()),
        "v": pa.array(rng.integers(0, 1 << 40, N).astype(np.int64)),
        "s": pa.array([("x" * (i % 40)) if i % 7 else None for i in range(N)], pa.string()),
        "b": pa.array(rng.integers(0, 2, N).astype(bool)),
    })
    path = str(tmp_path / name)
    pq.write_table(t, path, row_group_size=10000, data_page_size=4 << 10, write_page_index=page_index,
                   compression="ZSTD", use_dictionary=False, **kw)
    return t, path


def _check(t, path, flt, pks, mask, slack_pages=2):
    info, cols = _prune(path, flt, pks)
    md = pq.ParquetFile(path).metadata
    rows_before = sum(md.row_group(g).num_rows for g in range(info["first_kept_rg"])) if info["first_kept_rg"] >= 0 else 0
    idx = np.flatnonzero(mask)
    lo = rows_before + info["live_begin"]
    hi = lo + info["live_rows"]
    if len(idx):
        assert lo <= idx[0] and idx[-1] < hi, (flt, lo, hi, idx[0], idx[-1])
    # tight: at most a couple of key pages of slack around the matching rows
    kpage = max(n for _, n in cols["k"]) if cols["k"] else 0
    if len(idx):
        assert info["live_rows"] <= (idx[-1] - idx[0] + 1) + slack_pages * 2 * max(kpage, 1), (flt, info)
    # every column: contiguous pages from >= 0, covering the live range, inside the layout
    for name, pages in cols.items():
        if info["live_rows"] == 0:
            assert pages == []
            continue
        assert pages[0][0] >= 0
        for (a, n), (b, _) in zip(pages, pages[1:]):
            assert a + n == b, name
        assert pages[0][0] <= info["live_offset"], name
        assert pages[-1][0] + pages[-1][1] >= info["live_offset"] + info["live_rows"], name
        assert pages[-1][0] + pages[-1][1] <= info["layout_rows"], name
    if info["live_rows"]:
        assert 0 <= min(p[0][0] for p in cols.values()) < 64 and info["live_offset"] % 64 == 0  # merge range on a 64-row boundary
        assert max(p[-1][0] + p[-1][1] for p in cols.values()) == info["layout_rows"]
    return info


def test_key_range_keeps_only_the_overlapping_pages(tmp_path):
    t, path = _file(tmp_path)
    k = t["k"].to_numpy()
    info = _check(t, path, "and(gteq(k, 30300), lt(k, 33000))", "k,k2", (k >= 30300) & (k < 33000))
    assert info["dropped_rg"] == 2 and info["pruned_pages"] > 20 and info["live_rows"] < 3000
    assert info["lazy_image_bytes"] < info["file_bytes"] * 0.08  # the pages of ~1000 of 30000 rows (+ the boolean chunk)
    # (the boolean column has one small page per row group: it alone keeps the layout at the row group's size)
    assert 10000 <= info["layout_rows"] < 10064
    # across a row-group boundary
    info = _check(t, path, "and(gteq(k, 29000), lteq(k, 31000))", "k,k2", (k >= 29000) & (k <= 31000))
    assert info["dropped_rg"] == 1 and info["live_rows"] < 3000
    # equality, open ranges, nothing
    _check(t, path, "eq(k, 45000)", "k,k2", k == 45000)
    _check(t, path, "lt(k, 600)", "k,k2", k < 600)
    _check(t, path, "gt(k, 89000)", "k,k2", k > 89000)
    info = _check(t, path, "and(gt(k, 30001), lt(k, 30002))", "k,k2", np.zeros(N, bool))
    assert info["live_rows"] <= 2 * 600  # between two keys: whatever page holds the gap, no more


def test_string_key_and_two_key_columns(tmp_path):
    t, path = _file(tmp_path)
    k2 = np.array(t["k2"].to_pylist())
    k = t["k"].to_numpy()
    info = _check(t, path, 'and(gteq(k2, Binary{"id0012000"}), lt(k2, Binary{"id0012100"}))', "k,k2", (k2 >= "id0012000") & (k2 < "id0012100"))
    assert info["pruned_pages"] > 0
    # predicates on both key columns: the intersection of the two page hulls
    m = (k >= 36000) & (k2 < "id0012500")
    info = _check(t, path, 'and(gteq(k, 36000), lt(k2, Binary{"id0012500"}))', "k,k2", m)
    assert 0 < info["live_rows"] <= 2048
    # a disjunction is not a range: the hull of its parts
    info, _ = _prune(path, "or(eq(k, 300), eq(k, 89700))", "k,k2")  # rows 100 (row group 0) and 29900 (row group 2)
    assert info["dropped_rg"] == 1 and info["live_begin"] <= 100 and info["live_begin"] + info["live_rows"] >= 19901


def test_non_key_predicates_and_files_without_index_prune_nothing(tmp_path):
    t, path = _file(tmp_path)
    info, cols = _prune(path, "lt(v, 5)", "k,k2")
    assert info["pruned_pages"] == 0 and info["live_rows"] == N and info["layout_rows"] == N and info["live_offset"] == 0
    t2, p2 = _file(tmp_path, "noindex.parquet", page_index=False)
    info, cols = _prune(p2, "and(gteq(k, 30300), lt(k, 33000))", "k,k2")
    assert info["dropped_rg"] == 2 and info["pruned_pages"] == 0 and info["live_rows"] == 10000 == info["layout_rows"]


def test_null_key_pages(tmp_path):
    # null keys sort first: a page of nulls only cannot satisfy a comparison, but satisfies IS NULL
    k = pa.array([None] * 4000 + list(range(6000)), pa.int64())
    t = pa.table({"k": k, "v": pa.array(np.arange(10000, dtype=np.int64))})
    path = str(tmp_path / "nulls.parquet")
    pq.write_table(t, path, row_group_size=10000, max_rows_per_page=500, write_page_index=True, use_dictionary=False)
    info, cols = _prune(path, "lt(k, 100)", "k")
    assert info["live_begin"] == 4000 and info["live_rows"] == 500, info  # the null pages are out, one page holds 0..99
    info, cols = _prune(path, "or(lt(k, 100), eq(k, null))", "k")
    assert info["live_begin"] == 0 and info["live_rows"] == 4500, info
    info, cols = _prune(path, "noteq(k, null)", "k")
    assert info["live_begin"] == 4000 and info["live_rows"] == 6000, info
    info, cols = _prune(path, "eq(k, null)", "k")
    assert info["live_begin"] == 0 and info["live_rows"] == 4000, info


@pytest.mark.parametrize("layout", ["plain", "snappy_dict", "zstd_dict", "v2"])
def test_lazy_segments_for_every_page_layout(tmp_path, layout):
    """dictionary pages travel with the live data pages (two byte segments per chunk); chunks libparquet has to read
    (v2 pages, host-decompressed ones) get what was not loaded by pread"""
    kw = dict(plain=dict(compression="NONE"), snappy_dict=dict(compression="SNAPPY", use_dictionary=True),
              zstd_dict=dict(compression="ZSTD", use_dictionary=["v", "s"]), v2=dict(data_page_version="2.0"))[layout]
    rng = np.random.default_rng(4)
    t = pa.table({"k": pa.array(np.arange(N, dtype=np.int64)), "v": pa.array(rng.integers(0, 50, N).astype(np.int64)),
                  "s": pa.array(["tag%02d" % (i % 37) if i % 5 else None for i in range(N)], pa.string())})
    path = str(tmp_path / "f.parquet")
    args = dict(row_group_size=10000, data_page_size=4 << 10, write_page_index=True, compression="ZSTD", use_dictionary=False)
    args.update(kw)
    pq.write_table(t, path, **args)
    k = t["k"].to_numpy()
    for lo, hi in [(12000, 13500), (9500, 10500), (0, 800), (29000, 10**9)]:
        info = _check(t, path, f"and(gteq(k, {lo}), lt(k, {hi}))", "k", (k >= lo) & (k < hi))
        assert info["lazy_image_bytes"] < info["file_bytes"] * 0.5
