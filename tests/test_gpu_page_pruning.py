# SPDX-License-Identifier: Apache-2.0
"""Page-level pruning end to end (reference session.rs:685-717 enables the page index): with a pushed
key predicate only the pages that overlap the rows that can match are decompressed and decoded -- the columns' pages do
not align, so a file's row space in the scan is the union of the framed pages and the merge reads a sub-range of it.
Results must equal the oracle's for every page layout, key shape and merge path; the statistics must show pages skipped."""
import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq
import pytest

import mor_oracle as mo
from conftest import assert_tables_equal

pytestmark = pytest.mark.gpu

N = 40000


def _bucket(tmp_path, key_of, *, deltas=4, seed=0, extra=None, **write_kw):
    rng = np.random.default_rng(seed)
    kw = dict(row_group_size=16000, data_page_size=4 << 10, write_page_index=True, compression="ZSTD", use_dictionary=False)
    kw.update(write_kw)
    tables, paths = [], []
    for i in range(deltas + 1):
        ids = np.arange(N) if i == 0 else np.sort(rng.choice(int(N * 1.2), N // 5, replace=False))
        n = len(ids)
        cols = dict(key_of(ids))
        cols["v"] = pa.array(rng.integers(-1000, 1000, n).astype(np.int64))
        cols["s"] = pa.array([None if j % 9 == 0 else "s%05d" % x * (1 + x % 4) for j, x in enumerate(rng.integers(0, 10**5, n))], pa.string())
        cols["b"] = pa.array((rng.integers(0, 3, n) > 0).tolist(), pa.bool_())
        cols["d"] = pa.array(rng.random(n))
        if extra:
            cols.update(extra(rng, n, i))
        t = pa.table(cols)
        p = str(tmp_path / f"part-{i:05d}_0000.parquet")
        pq.write_table(t, p, **kw)
        tables.append(t)
        paths.append(p)
    return tables[0].schema, paths


def _read(paths, schema, pks, filters, **kw):
    from lakesoul_b200 import read_bucket
    got, rd = read_bucket(paths, schema, pks, batch_size=4096, filters=filters, return_reader=True, **kw)
    st = rd.stats()
    rd.close()
    want = mo.read_bucket(paths, schema, pks, filters=list(filters), **kw)
    assert_tables_equal(got, want, f"filters {filters}")
    return got, st


INT_KEY = lambda ids: {"k": pa.array(ids.astype(np.int64))}  # noqa: E731


@pytest.mark.parametrize("layout", ["zstd", "plain", "snappy_dict", "v2"])
def test_key_range_over_every_page_layout(tmp_path, layout):
    kw = dict(zstd={}, plain=dict(compression="NONE"), snappy_dict=dict(compression="SNAPPY", use_dictionary=True),
              v2=dict(data_page_version="2.0"))[layout]
    schema, paths = _bucket(tmp_path, INT_KEY, **kw)
    total = sum(pq.read_metadata(p).num_rows for p in paths)
    got, st = _read(paths, schema, ["k"], ["and(gteq(k, 17000), lt(k, 19500))"])
    assert 2000 < got.num_rows <= 2500
    assert st["pruned_pages"] > 0 and st["rows_in"] < total * 0.4, st
    # a range that crosses a row-group boundary of the base file, an equality, an empty result
    got, st = _read(paths, schema, ["k"], ["and(gt(k, 15500), lteq(k, 16200))"])
    assert st["pruned_pages"] > 0
    got, st = _read(paths, schema, ["k"], ["eq(k, 31234)"])
    assert got.num_rows == 1 and st["rows_in"] < total * 0.3
    # a key lookup reads footers and the byte segments of a few pages per file, not the files
    import os
    file_bytes = sum(os.path.getsize(p) for p in paths)
    assert st["h2d_bytes"] < file_bytes * (0.3 if layout in ("zstd", "plain") else 0.9), (st["h2d_bytes"], file_bytes)
    got, st = _read(paths, schema, ["k"], ["gt(k, 99999999)"])
    assert got.num_rows == 0


def test_string_and_composite_keys(tmp_path):
    schema, paths = _bucket(tmp_path / "a", lambda ids: {"k": pa.array(["key-%07d" % i for i in ids], pa.string())}) \
        if (tmp_path / "a").mkdir() is None else None
    got, st = _read(paths, schema, ["k"], ['and(gteq(k, Binary{"key-0021000"}), lt(k, Binary{"key-0021700"}))'])
    assert got.num_rows >= 700 and st["pruned_pages"] > 0
    (tmp_path / "b").mkdir()
    schema, paths = _bucket(tmp_path / "b", lambda ids: {"a": pa.array((ids // 5).astype(np.int32)),
                                                          "z": pa.array(["z%d" % (i % 5) for i in ids], pa.string())})
    got, st = _read(paths, schema, ["a", "z"], ["and(gteq(a, 3000), lt(a, 3300))"])
    assert got.num_rows >= 1500 and st["pruned_pages"] > 0
    got, st = _read(paths, schema, ["a", "z"], ['and(eq(a, 4000), eq(z, Binary{"z3"}))'])
    assert got.num_rows == 1


def test_with_other_predicates_cdc_and_merge_operators(tmp_path):
    def extra(rng, n, file_no):
        kinds = np.array(["insert", "update", "delete"])
        return {"rowKinds": pa.array(kinds[rng.choice(3, n, p=[0.5, 0.35, 0.15] if file_no else [1, 0, 0])].tolist(), pa.string())}

    schema, paths = _bucket(tmp_path, INT_KEY, extra=extra)
    _, st = _read(paths, schema, ["k"], ['not(eq(rowKinds, Binary{"delete"}))', "and(gteq(k, 8000), lt(k, 12000))", "gt(v, -300)"],
                  options={"cdc_column": "rowKinds"})
    assert st["pruned_pages"] > 0
    _, st = _read(paths, schema, ["k"], ["and(gteq(k, 8000), lt(k, 12000))"], merge_ops={"v": "SumAll", "s": "JoinedAllByComma"})
    assert st["pruned_pages"] > 0
    # a disjunction of two far keys: the hull is decoded, the result is two rows
    got, st = _read(paths, schema, ["k"], ["or(eq(k, 100), eq(k, 39000))"])
    assert got.num_rows == 2
    # a predicate on a non-key column alone prunes nothing
    _, st = _read(paths, schema, ["k"], ["lt(v, -990)"])
    assert st["pruned_pages"] == 0


def test_odd_page_boundaries(tmp_path):
    """pages of 77 rows (parquet-rs pages end wherever the size limit says): live ranges and page starts on odd rows"""
    schema, paths = _bucket(tmp_path, INT_KEY, write_batch_size=11, max_rows_per_page=77)
    for flt in ["and(gteq(k, 17011), lt(k, 19513))", "eq(k, 12345)", "and(gt(k, 15990), lteq(k, 16013))", "lt(k, 131)"]:
        got, st = _read(paths, schema, ["k"], [flt])
        assert st["pruned_pages"] > 0 and got.num_rows > 0
    _read(paths, schema, ["k"], ["and(gteq(k, 3001), lt(k, 9007))"], merge_ops={"v": "SumAll", "s": "JoinedAllByComma"})


def test_streamed_bucket_decodes_each_page_once(tmp_path):
    """with a page index the passes of a streamed bucket decode the delta files' pages of their key range only: the rows
    decoded over all passes stay close to the bucket's rows (without it every pass decodes every straddling row group)"""
    from lakesoul_b200 import read_bucket
    schema, paths = _bucket(tmp_path, INT_KEY, deltas=5, row_group_size=8000, write_batch_size=128, max_rows_per_page=256)
    total_rows = sum(pq.read_metadata(p).num_rows for p in paths)
    total_bytes = sum(pq.read_metadata(p).row_group(g).total_byte_size for p in paths for g in range(pq.read_metadata(p).num_row_groups))
    got, rd = read_bucket(paths, schema, ["k"], batch_size=4096, options={"b200_stream_pass_bytes": str(total_bytes // 4)}, return_reader=True)
    st = rd.stats()
    rd.close()
    assert_tables_equal(got, mo.read_bucket(paths, schema, ["k"]), "streamed with page index")
    assert st["stream_passes"] >= 3 and st["pruned_pages"] > 0
    assert st["rows_in"] < total_rows * 1.5, (st, total_rows)
    # ... and read once: with the offset index a pass reads the byte segments of its pages only (every pass would read
    # every delta file whole otherwise: about 3 x the files here)
    import os
    file_bytes = sum(os.path.getsize(p) for p in paths)
    assert st["h2d_bytes"] < file_bytes * 2.0, (st["h2d_bytes"], file_bytes)
