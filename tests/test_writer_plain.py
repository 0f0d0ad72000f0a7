# SPDX-License-Identifier: Apache-2.0
"""Write side without a GPU: the plain writer (no primary key, or keep_orders=true; writer/mod.rs:101-125 ->
MultiPartAsyncWriter) writes rows as they come -- no sort, so no device -- with the reference's file naming, range
columns dropped, ZSTD pages, and the flush summary of writer/mod.rs:335-383; the sorting writers refuse to start
without a GPU instead of falling back to the CPU."""
import os

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq
import pytest

from lakesoul_b200.native_io import NativeIOWriter


def _table(n=5000):
    rng = np.random.default_rng(0)
    return pa.table({"id": rng.integers(0, 1000, n), "name": [f"n{i % 17}" for i in range(n)],
                     "day": pa.array(["2024-01-01"] * n), "x": rng.standard_normal(n)})


def test_plain_writer_keeps_input_order_and_reports_the_file(tmp_path):
    t = _table()
    w = NativeIOWriter(t.schema)
    w.withPrefix(str(tmp_path))
    w.setRangePartitions(["day"])
    w.setOption("hash_bucket_id", "7")
    w.initializeWriter()
    for b in t.to_batches(max_chunksize=999):
        w.write(b)
    res = w.flush()
    w.close()
    assert list(res) == ["-4"]           # TBD_PARTITION_DESC (multipart_writer.rs:221, constant.rs:24)
    fr, = res["-4"]
    name = os.path.basename(fr.filePath)
    assert name.startswith("part-") and name.endswith("_0007.parquet") and len(name) == len("part-") + 16 + len("_0007.parquet")
    assert fr.fileSize == os.path.getsize(fr.filePath)
    assert fr.fileExistCols == "id,name,x"   # range partition columns are not written (multipart_writer.rs:96-109)
    got = pq.read_table(fr.filePath)
    assert got.equals(t.drop_columns(["day"]))
    md = pq.ParquetFile(fr.filePath).metadata
    assert md.row_group(0).column(0).compression == "ZSTD"


def test_keep_orders_with_primary_keys_is_plain(tmp_path):
    t = _table(100)
    w = NativeIOWriter(t.schema)
    w.addFile(str(tmp_path / "explicit.parquet"))
    w.setPrimaryKeys(["id"])
    w.setOption("keep_orders", "true")
    w.initializeWriter()
    w.write(t.to_batches()[0])
    res = w.flush()
    w.close()
    assert res["-4"][0].filePath == str(tmp_path / "explicit.parquet")
    assert pq.read_table(res["-4"][0].filePath).equals(t)


def test_abort_removes_the_file_and_schema_mismatch_is_an_error(tmp_path):
    t = _table(10)
    w = NativeIOWriter(t.schema)
    w.withPrefix(str(tmp_path))
    w.initializeWriter()
    with pytest.raises(IOError, match="type|schema"):
        w.write(pa.record_batch({"id": pa.array([1.5]), "name": ["a"], "day": ["d"], "x": [1.0]}))
    w.write(t.to_batches()[0])
    w.abort()
    w.close()
    assert [p for p in os.listdir(tmp_path) if p.endswith(".parquet")] == []


def test_sorting_writer_needs_a_gpu(tmp_path):
    import ctypes
    from lakesoul_b200 import load_library
    try:
        ctypes.CDLL("libcuda.so.1")
        have_driver = load_library().lsb_init_device() >= 0
    except OSError:
        have_driver = False
    if have_driver:
        pytest.skip("a GPU is present")
    t = _table(10)
    w = NativeIOWriter(t.schema)
    w.withPrefix(str(tmp_path))
    w.setPrimaryKeys(["id"])
    with pytest.raises(IOError, match="Init native writer failed"):
        w.initializeWriter()
    w.close()
