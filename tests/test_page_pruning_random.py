# SPDX-License-Identifier: Apache-2.0
"""Randomised check of row-group + page pruning (page_plan.cc): for random sorted files (key type, nulls first, row-group
and page sizes on odd boundaries) and random pushed predicates, every row the predicate accepts -- evaluated by the
oracle's filter -- lies inside the live range, in kept-row coordinates; the framed pages of every column cover it; the
lazy (segment-wise) framing equals the whole-file framing. Host logic only."""
import os

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq
import pytest

import mor_oracle as mo  # noqa: E402
from test_page_pruning import _prune  # noqa: E402


def _random_file(rng, path, string_key):
    n = int(rng.integers(3000, 15000))
    nulls = int(rng.integers(0, 3)) * int(rng.integers(1, 400))
    ids = np.sort(rng.choice(n * 3, n - nulls, replace=False))
    if string_key:
        k = pa.array([None] * nulls + ["k%07d" % i for i in ids], pa.string())
    else:
        k = pa.array([None] * nulls + [int(i) - n for i in ids], pa.int64())  # negative and positive keys
    t = pa.table({"k": k,
                  "g": pa.array(rng.integers(0, 5, n).astype(np.int32)),
                  "s": pa.array([None if i % 5 == 0 else "v" * (i % 23) for i in range(n)], pa.string()),
                  "f": pa.array(rng.random(n))})
    pq.write_table(t, path, row_group_size=int(rng.integers(900, 6000)), max_rows_per_page=int(rng.integers(17, 500)),
                   write_batch_size=int(rng.integers(5, 64)), write_page_index=True,
                   compression=str(rng.choice(["ZSTD", "NONE", "SNAPPY"])), use_dictionary=bool(rng.integers(0, 2)))
    return t, ids, n


def _lit(rng, ids, n, string_key):
    x = int(rng.choice(ids)) + int(rng.integers(-2, 3))
    return 'Binary{"k%07d"}' % max(x, 0) if string_key else str(x - n)


def _random_filter(rng, ids, n, string_key):
    ops = ["eq", "noteq", "lt", "lteq", "gt", "gteq"]

    def term():
        r = rng.integers(0, 10)
        if r < 6:
            return f"{rng.choice(ops)}(k, {_lit(rng, ids, n, string_key)})"
        if r == 6:
            return "eq(k, null)"
        if r == 7:
            return "noteq(k, null)"
        if r == 8:
            return f"lt(g, {int(rng.integers(0, 6))})"  # not a key column
        a, b = sorted([_lit(rng, ids, n, string_key), _lit(rng, ids, n, string_key)], key=lambda v: (len(v), v) if not string_key else v)
        return f"and(gteq(k, {a}), lt(k, {b}))"

    def expr(depth):
        r = rng.integers(0, 10)
        if depth == 0 or r < 4:
            return term()
        if r < 7:
            return f"and({expr(depth - 1)}, {expr(depth - 1)})"
        if r < 9:
            return f"or({expr(depth - 1)}, {expr(depth - 1)})"
        return f"not({expr(depth - 1)})"

    return expr(2)


@pytest.mark.parametrize("seed", range(10))
def test_accepted_rows_lie_in_the_live_range(tmp_path, seed):
    rng = np.random.default_rng(1000 + seed)
    string_key = bool(seed % 2)
    path = str(tmp_path / "f.parquet")
    t, ids, n = _random_file(rng, path, string_key)
    md = pq.ParquetFile(path).metadata
    rg_rows = [md.row_group(g).num_rows for g in range(md.num_row_groups)]
    rg_start = np.concatenate([[0], np.cumsum(rg_rows)])
    indexed = t.append_column("__row", pa.array(np.arange(n, dtype=np.int64)))
    pruned_some = 0
    for _ in range(40):
        flt = _random_filter(rng, ids, n, string_key)
        info, cols = _prune(path, flt, "k")
        rows = indexed.filter(mo.parse_filter(flt, t.schema))["__row"].to_numpy()
        # file row -> kept-row coordinate
        kept = info["kept_rgs"]
        base = {g: sum(rg_rows[h] for h in kept if h < g) for g in kept}
        for r in rows:
            g = int(np.searchsorted(rg_start, r, side="right") - 1)
            assert g in base, (flt, "a row group with an accepted row was dropped", int(r), g, kept)
            kr = base[g] + int(r - rg_start[g])
            assert info["live_begin"] <= kr < info["live_begin"] + info["live_rows"], (flt, int(r), kr, info)
        if info["live_rows"]:
            for name, pages in cols.items():
                assert pages and pages[0][0] <= info["live_offset"], (flt, name)
                assert pages[-1][0] + pages[-1][1] >= info["live_offset"] + info["live_rows"], (flt, name)
                assert all(a + c == b for (a, c), (b, _) in zip(pages, pages[1:])), (flt, name)
        pruned_some += info["pruned_pages"] > 0 or info["dropped_rg"] > 0
    assert pruned_some >= 5
