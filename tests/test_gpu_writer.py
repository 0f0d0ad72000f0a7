# SPDX-License-Identifier: Apache-2.0
"""Write side behind the C-ABI: create_lakesoul_writer_from_config / write_record_batch_blocked /
flush_and_close_writer. Files, their names, the flush summary and the row order inside every file are compared with
oracle/write_oracle.py (the reference's SortExec + RepartitionByRangeAndHashExec placement, KAT-pinned hash), and the
files are read back through the merge-on-read reader (compaction = merge + write closes the loop)."""
import decimal
import os

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq
import pytest

import mor_oracle as mo
import write_oracle as wo
from conftest import assert_tables_equal
from lakesoul_b200.native_io import NativeIOWriter, read_bucket

pytestmark = pytest.mark.gpu


def _writer(schema, prefix, pks=(), ranges=(), aux=(), buckets=1, dynamic=False, batch_size=1024, options=None):
    w = NativeIOWriter(schema)
    w.withPrefix(str(prefix))
    w.setPrimaryKeys(pks)
    w.setRangePartitions(ranges)
    w.setAuxSortColumns(aux)
    w.setHashBucketNum(buckets)
    w.useDynamicPartition(dynamic)
    w.setBatchSize(batch_size)
    for k, v in (options or {}).items():
        w.setOption(k, v)
    w.initializeWriter()
    return w


def _write_all(w, table, chunk=700):
    for b in table.to_batches(max_chunksize=chunk):
        w.write(b)


def _lineitem_like(n, seed, with_nulls=True):
    rng = np.random.default_rng(seed)
    k1 = rng.integers(0, n // 3, n, dtype=np.int64) * 7 - 1000
    k2 = rng.integers(0, 5, n).astype(np.int32)
    words = np.array(["alpha", "be,ta", "gamma delta", "", "epsilon-zeta-eta-theta-iota-kappa", "z"])
    s = words[rng.integers(0, len(words), n)]
    smask = rng.random(n) < 0.1 if with_nulls else np.zeros(n, bool)
    v = rng.standard_normal(n)
    vmask = rng.random(n) < 0.05 if with_nulls else np.zeros(n, bool)
    d = rng.integers(19000, 19003, n).astype(np.int32)
    flag = rng.random(n) < 0.5
    dec = [decimal.Decimal(int(x)).scaleb(-2) for x in rng.integers(-10**6, 10**6, n)]
    return pa.table({
        "k1": pa.array(k1), "k2": pa.array(k2), "s": pa.array(s.tolist(), mask=smask), "v": pa.array(v, mask=vmask),
        "d": pa.array(d, pa.date32()), "flag": pa.array(flag), "dec": pa.array(dec, pa.decimal128(15, 2)),
        "seq": pa.array(np.arange(n, dtype=np.int64)),
    })


def test_sort_writer_orders_by_primary_key_and_keeps_duplicates(tmp_path):
    """sort_writer.rs:59-76: one file, rows sorted by (k1, k2), ties in input order, every row kept"""
    t = _lineitem_like(20000, 1)
    w = _writer(t.schema, tmp_path, pks=["k1", "k2"], options={"hash_bucket_id": "3"})
    _write_all(w, t)
    st = w.writer_stats()
    res = w.flush()
    w.close()
    assert list(res.keys()) == ["-4"] and len(res["-4"]) == 1
    fr = res["-4"][0]
    assert os.path.basename(fr.filePath).startswith("part-") and fr.filePath.endswith("_0003.parquet")
    assert fr.fileSize == os.path.getsize(fr.filePath)
    assert fr.fileExistCols == ",".join(t.schema.names)
    got = pq.ParquetFile(fr.filePath).read()
    (_, _, order), = wo.placed(t, [], ["k1", "k2"], [], 1, False)
    assert_tables_equal(got, t.take(pa.array(order)), "sort writer")
    md = pq.ParquetFile(fr.filePath).metadata
    col = md.row_group(0).column(0)
    assert col.compression == "ZSTD" and "PLAIN" in col.encodings and "RLE_DICTIONARY" not in col.encodings  # multipart_writer.rs:128-129
    assert st["rows_in"] == 20000


@pytest.mark.parametrize("pks", [["k1"], ["k1", "k2"], ["s", "k2"], ["dec"], ["flag", "d", "k2", "k1", "s"]])
def test_partition_writer_files_buckets_and_order(tmp_path, pks):
    """partitioning_writer.rs:139-375 + repartition/mod.rs:222-300: one file per (range partition, hash bucket), named
    <prefix>/<col>=<value>/part-<write id>_<bucket:04>.parquet, range column dropped, rows sorted by the primary key"""
    t = _lineitem_like(12000, 2)
    nb = 4
    w = _writer(t.schema, tmp_path, pks=pks, ranges=["d"], buckets=nb, dynamic=True)
    _write_all(w, t)
    res = w.flush()
    w.close()
    want = wo.placed(t, ["d"], pks, [], nb, True)
    file_cols = [n for n in t.schema.names if n != "d"]
    seen = 0
    write_ids = set()
    for (rk, bucket, rows) in want:
        day = t.column("d")[rows[0]].as_py().isoformat()
        desc = f"d={day}"
        assert desc in res, (desc, list(res))
        match = [f for f in res[desc] if f.filePath.endswith(f"_{bucket:04d}.parquet")]
        assert len(match) == 1, (desc, bucket, res[desc])
        fr = match[0]
        assert os.path.dirname(fr.filePath) == os.path.join(str(tmp_path), desc)
        write_ids.add(os.path.basename(fr.filePath)[5:21])
        assert fr.fileExistCols == ",".join(file_cols)
        got = pq.ParquetFile(fr.filePath).read()
        assert_tables_equal(got, t.take(pa.array(rows)).select(file_cols), f"{desc} bucket {bucket}")
        seen += len(rows)
    assert seen == t.num_rows
    assert sum(len(v) for v in res.values()) == len(want)
    assert len(write_ids) == 1  # one write id per flush (partitioning_writer.rs:110)


def test_partition_writer_without_range_partitions_and_aux_sort(tmp_path):
    t = _lineitem_like(6000, 3)
    w = _writer(t.schema, tmp_path, pks=["k2"], aux=["seq"], buckets=3, dynamic=True)
    _write_all(w, t.take(pa.array(np.random.default_rng(0).permutation(t.num_rows))))  # shuffled input: aux column restores the order
    res = w.flush()
    w.close()
    assert list(res.keys()) == ["-5"]
    shuffled = t.take(pa.array(np.random.default_rng(0).permutation(t.num_rows)))
    want = wo.placed(shuffled, [], ["k2"], ["seq"], 3, True)
    cols = [n for n in t.schema.names if n != "seq"]  # aux sort columns are not written (writer/mod.rs:66-81)
    for (_, bucket, rows) in want:
        fr, = [f for f in res["-5"] if f.filePath.endswith(f"_{bucket:04d}.parquet")]
        assert os.path.dirname(fr.filePath) == str(tmp_path)
        assert_tables_equal(pq.ParquetFile(fr.filePath).read(), shuffled.take(pa.array(rows)).select(cols), f"bucket {bucket}")


def test_null_and_float_keys_hash_like_the_reference(tmp_path):
    """nulls leave the running hash unchanged and sort first; -0.0 hashes like 0.0 but sorts before it"""
    f = pa.array([0.0, -0.0, None, 1.5, -1.5, float("inf"), None, 0.0, -0.0], pa.float64())
    s = pa.array(["a", None, "a", "", None, "b", None, "a", "a"])
    t = pa.table({"f": f, "s": s, "i": pa.array(range(9), pa.int32())})
    w = _writer(t.schema, tmp_path, pks=["f", "s"], buckets=5, dynamic=True)
    w.write(t.to_batches()[0])
    res = w.flush()
    w.close()
    want = wo.placed(t, [], ["f", "s"], [], 5, True)
    assert sum(len(r) for _, _, r in want) == 9
    for (_, bucket, rows) in want:
        fr, = [x for x in res["-5"] if x.filePath.endswith(f"_{bucket:04d}.parquet")]
        got = pq.ParquetFile(fr.filePath).read()
        assert got.column("i").to_pylist() == [t.column("i")[r].as_py() for r in rows], bucket


def test_flush_cycles_bound_the_buffer(tmp_path, monkeypatch):
    """a small buffer budget: several device passes, several sorted files per bucket, nothing lost"""
    monkeypatch.setenv("LAKESOUL_B200_WRITER_BUFFER_BYTES", str(200_000))
    t = _lineitem_like(30000, 4, with_nulls=False)
    w = _writer(t.schema, tmp_path, pks=["k1"], buckets=2, dynamic=True)
    _write_all(w, t, chunk=1000)
    st = w.writer_stats()
    res = w.flush()
    w.close()
    assert st["flush_cycles"] >= 2
    rows = 0
    seqs = []
    for fr in res["-5"]:
        got = pq.ParquetFile(fr.filePath).read()
        k = got.column("k1").to_numpy()
        assert np.all(k[1:] >= k[:-1])
        rows += got.num_rows
        seqs += got.column("seq").to_pylist()
    assert rows == t.num_rows and sorted(seqs) == list(range(t.num_rows))


def test_write_then_merge_on_read_round_trip(tmp_path):
    """two commits written by the partitioning writer, every bucket read back by the merge-on-read reader: the result is
    the oracle's upsert of commit 2 over commit 1 (the compaction loop of CompactBucketIO: merge + write)"""
    rng = np.random.default_rng(9)
    schema = pa.schema([("id", pa.int64()), ("name", pa.string()), ("val", pa.float64())])

    def commit(n, lo, hi, tag):
        ids = rng.choice(np.arange(lo, hi), n, replace=False).astype(np.int64)
        return pa.table({"id": ids, "name": [f"{tag}-{i}" for i in ids], "val": rng.standard_normal(n)}, schema=schema)

    commits = [commit(5000, 0, 8000, "a"), commit(3000, 4000, 12000, "b")]
    nb = 4
    files = {b: [] for b in range(nb)}
    for c in commits:
        w = _writer(schema, tmp_path, pks=["id"], buckets=nb, dynamic=True)
        _write_all(w, c)
        res = w.flush()
        w.close()
        for fr in res["-5"]:
            files[int(fr.filePath[-12:-8])].append(fr.filePath)
    total = 0
    for b in range(nb):
        got = read_bucket(files[b], schema, ["id"])
        want = mo.read_bucket(files[b], schema, ["id"])
        assert_tables_equal(got, want, f"bucket {b}")
        assert all(x == b for x in wo.buckets(got, ["id"], nb))
        total += got.num_rows
    all_ids = set(commits[0].column("id").to_pylist()) | set(commits[1].column("id").to_pylist())
    assert total == len(all_ids)


def test_ipc_entry_point_and_abort(tmp_path):
    t = _lineitem_like(3000, 5)
    w = _writer(t.schema, tmp_path, pks=["k1"])
    sink = pa.BufferOutputStream()
    with pa.ipc.new_stream(sink, t.schema) as wr:
        for b in t.to_batches(max_chunksize=1000):
            wr.write_batch(b)
    assert w.writeIpc(sink.getvalue().to_pybytes()) == 3000
    res = w.flush()
    w.close()
    got = pq.ParquetFile(res["-4"][0].filePath).read()
    (_, _, order), = wo.placed(t, [], ["k1"], [], 1, False)
    assert_tables_equal(got, t.take(pa.array(order)), "ipc")
    # abort: nothing is left behind
    d2 = tmp_path / "aborted"
    w2 = _writer(t.schema, d2, pks=["k1"], buckets=2, dynamic=True)
    _write_all(w2, t)
    w2.abort()
    w2.close()
    assert not d2.exists() or not any(p.suffix == ".parquet" for p in d2.rglob("*"))


def test_partition_descriptions_follow_format_scalar_value(tmp_path):
    """helpers/mod.rs:205-274: strings (',' -> the EQ marker, '' -> the empty marker), integers, booleans, microsecond
    timestamps and nulls in the partition description and in the directory names (helpers/mod.rs:452-488)"""
    import datetime as dt
    ts = [dt.datetime(2024, 1, 2, 3, 4, 5, 123456), dt.datetime(1969, 12, 31, 23, 59, 59, 500000)]
    rows = [("a,b", 5, True, ts[0]), ("", -7, False, ts[1]), (None, 5, True, ts[0]), ("plain", 0, False, ts[1])]
    t = pa.table({
        "id": pa.array(range(len(rows) * 50), pa.int64()),
        "s": pa.array([rows[i % 4][0] for i in range(len(rows) * 50)], pa.string()),
        "i": pa.array([rows[i % 4][1] for i in range(len(rows) * 50)], pa.int32()),
        "b": pa.array([rows[i % 4][2] for i in range(len(rows) * 50)]),
        "ts": pa.array([rows[i % 4][3] for i in range(len(rows) * 50)], pa.timestamp("us")),
    })
    w = _writer(t.schema, tmp_path, pks=["id"], ranges=["s", "i", "b", "ts"], buckets=1, dynamic=True)
    _write_all(w, t)
    res = w.flush()
    w.close()
    want = {
        "s=a__L@KE$OUL_EQ__b,i=5,b=true,ts=2024-01-02T03:04:05.123456",
        "s=__L@KE$OUL_EMPTY_STRING__,i=-7,b=false,ts=1969-12-31T23:59:59.500000",
        "s=NULL,i=5,b=true,ts=2024-01-02T03:04:05.123456",
        "s=plain,i=0,b=false,ts=1969-12-31T23:59:59.500000",
    }
    assert set(res) == want
    for desc, files in res.items():
        assert len(files) == 1
        sub = "/".join(desc.split(","))
        assert files[0].filePath == os.path.join(str(tmp_path), sub, os.path.basename(files[0].filePath))
        assert files[0].fileExistCols == "id"
        assert pq.ParquetFile(files[0].filePath).read().num_rows == 50
