# SPDX-License-Identifier: Apache-2.0
"""Buckets streamed by primary-key range (reader.cc: plan_stream_passes, lazy file reads): a bucket above the pass budget
(option b200_stream_pass_bytes) is read as several scans over ascending, disjoint key ranges. The stream must be exactly
what one scan returns -- same rows, same order -- for every key shape the planner cuts on, with filters, CDC deletes
and merge operators, and it must read only the byte span of the row groups each pass keeps. The reference streams a
bucket batch by batch (sorted_stream_merger.rs:454-524)."""
import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq
import pytest

import mor_oracle as mo
from conftest import assert_tables_equal

pytestmark = pytest.mark.gpu

N = 48000


def _write(tmp_path, tables, rg=4000, compression="ZSTD", dictionary=False):
    paths = []
    for i, t in enumerate(tables):
        p = str(tmp_path / f"part-{i:05d}_0000.parquet")
        pq.write_table(t, p, row_group_size=rg, compression=compression, use_dictionary=dictionary, data_page_size=16 << 10)
        paths.append(p)
    return paths


def _payload(rng, n, with_nulls=True):
    s = ["s%06d" % x * (1 + x % 3) for x in rng.integers(0, 10**6, n)]
    if with_nulls:
        s = [None if i % 13 == 0 else x for i, x in enumerate(s)]
    return {"v": pa.array(rng.integers(-1000, 1000, n).astype(np.int64)), "s": pa.array(s, pa.string())}


def _files(rng, key_of, nullable=False, extra=None, deltas=5):
    """a base file with every key 0..N-1 and `deltas` files with random subsets (also of new keys), all sorted by key"""
    tables = []
    for i in range(deltas + 1):
        ids = np.arange(N) if i == 0 else np.sort(rng.choice(int(N * 1.25), N // 6, replace=False))
        cols = dict(key_of(ids))
        cols.update(_payload(rng, len(ids)))
        if extra:
            cols.update(extra(rng, len(ids), i))
        tables.append(pa.table(cols))
    return tables


def _total_bytes(paths):
    return sum(pq.ParquetFile(p).metadata.row_group(g).total_byte_size for p in paths
               for g in range(pq.ParquetFile(p).metadata.num_row_groups))


def _check(paths, schema, pks, *, passes_min=3, filters=(), options=None, merge_ops=None, batch_size=2048):
    from lakesoul_b200 import read_bucket
    options = dict(options or {})
    one, r1 = read_bucket(paths, schema, pks, batch_size=batch_size, options={**options, "b200_stream_pass_bytes": "0"},
                          filters=filters, merge_ops=merge_ops, return_reader=True)
    s1 = r1.stats()
    r1.close()
    assert s1["stream_passes"] == 0
    want = mo.read_bucket(paths, schema, pks, options=options, filters=list(filters), merge_ops=merge_ops)
    assert_tables_equal(one, want, "one scan vs oracle")
    budget = _total_bytes(paths) // 4
    got, r2 = read_bucket(paths, schema, pks, batch_size=batch_size, options={**options, "b200_stream_pass_bytes": str(budget)},
                          filters=filters, merge_ops=merge_ops, return_reader=True)
    s2 = r2.stats()
    r2.close()
    assert s2["stream_passes"] >= passes_min, s2
    assert got.schema.equals(one.schema)
    assert_tables_equal(got, one, "streamed vs one scan")
    assert s2["rows_out"] == one.num_rows
    return s1, s2


def test_int_key(tmp_path):
    rng = np.random.default_rng(5)
    tables = _files(rng, lambda ids: {"k": pa.array(ids.astype(np.int64))})
    paths = _write(tmp_path, tables)
    s1, s2 = _check(paths, tables[0].schema, ["k"])
    # every pass decodes the base file's row groups of its range only (the base about once in total) plus the delta row
    # groups that span its range -- far less than the whole bucket per pass
    assert s2["rows_in"] < s1["rows_in"] * s2["stream_passes"] * 0.6, (s1["rows_in"], s2)


def test_string_key_uncompressed(tmp_path):
    rng = np.random.default_rng(6)
    tables = _files(rng, lambda ids: {"k": pa.array(["user-%08d" % i for i in ids], pa.string())})
    paths = _write(tmp_path, tables, compression="NONE")
    _check(paths, tables[0].schema, ["k"])


def test_nullable_key_nulls_in_first_pass(tmp_path):
    rng = np.random.default_rng(7)

    def key_of(ids):
        k = pa.array(ids.astype(np.int32))
        mask = np.zeros(len(ids), bool)
        mask[:3] = True  # null keys sort first and are equal to each other
        return {"k": pa.array(np.where(mask, 0, ids).astype(np.int32), mask=mask)}

    tables = _files(rng, key_of)
    paths = _write(tmp_path, tables, compression="SNAPPY", dictionary=True)
    _check(paths, tables[0].schema, ["k"])


def test_composite_key_split_on_first_column(tmp_path):
    rng = np.random.default_rng(8)
    tables = _files(rng, lambda ids: {"a": pa.array((ids // 7).astype(np.int32)), "b": pa.array(["b%02d" % (i % 7) for i in ids], pa.string())})
    paths = _write(tmp_path, tables)
    _check(paths, tables[0].schema, ["a", "b"])


def test_filters_cdc_and_merge_operators(tmp_path):
    rng = np.random.default_rng(9)

    def extra(rng, n, file_no):
        kinds = np.array(["insert", "update", "delete"])
        return {"rowKinds": pa.array(kinds[rng.choice(3, n, p=[0.5, 0.35, 0.15] if file_no else [1, 0, 0])].tolist(), pa.string())}

    tables = _files(rng, lambda ids: {"k": pa.array(ids.astype(np.int64))}, extra=extra)
    paths = _write(tmp_path, tables)
    schema = tables[0].schema
    _check(paths, schema, ["k"], options={"cdc_column": "rowKinds"},
           filters=['not(eq(rowKinds, Binary{"delete"}))', "gt(v, -500)"])
    # a pushed key range on top of the pass ranges
    _check(paths, schema, ["k"], filters=["and(gteq(k, 9000), lt(k, 41000))"], passes_min=2)
    _check(paths, schema, ["k"], merge_ops={"v": "SumAll", "s": "JoinedLastByComma"})


def test_single_compacted_file(tmp_path):
    rng = np.random.default_rng(10)
    tables = _files(rng, lambda ids: {"k": pa.array(ids.astype(np.int64))}, deltas=0)
    paths = _write(tmp_path, tables)
    _check(paths, tables[0].schema, ["k"], options={"is_compacted": "true"})


def test_unsplittable_bucket_is_one_scan(tmp_path):
    from lakesoul_b200 import read_bucket
    rng = np.random.default_rng(11)
    tables = _files(rng, lambda ids: {"k": pa.array(ids.astype(np.float64))}, deltas=2)
    paths = _write(tmp_path, tables)
    got, rd = read_bucket(paths, tables[0].schema, ["k"], options={"b200_stream_pass_bytes": "100000"}, return_reader=True)
    st = rd.stats()
    rd.close()
    assert st["stream_passes"] == 0
    assert_tables_equal(got, mo.read_bucket(paths, tables[0].schema, ["k"]), "float key, one scan")
