# SPDX-License-Identifier: Apache-2.0
"""Write-side timing probe (development): rows through the dynamic-partition writer, device pass vs host Parquet encode.
usage: dev_writer_probe.py <rows> <buckets>"""
import os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import pyarrow as pa
from lakesoul_b200.native_io import NativeIOWriter

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 16
rng = np.random.default_rng(0)
words = np.array(["DELIVER IN PERSON", "COLLECT COD", "NONE", "TAKE BACK RETURN"])
t = pa.table({
    "l_orderkey": rng.integers(0, n // 4, n, dtype=np.int64), "l_linenumber": rng.integers(1, 8, n).astype(np.int32),
    "l_partkey": rng.integers(0, 200000, n, dtype=np.int64), "l_quantity": rng.random(n),
    "l_extendedprice": rng.random(n) * 1e5, "l_shipdate": pa.array(rng.integers(8000, 10500, n).astype(np.int32), pa.date32()),
    "l_shipinstruct": words[rng.integers(0, 4, n)].tolist(),
    "l_comment": ["%016x" % x for x in rng.integers(0, 2**62, n)],
})
for rep in range(2):
    d = tempfile.mkdtemp(prefix="lsb_wr_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    w = NativeIOWriter(t.schema)
    w.withPrefix(d)
    w.setPrimaryKeys(["l_orderkey", "l_linenumber"])
    w.setHashBucketNum(nb)
    w.useDynamicPartition(True)
    w.setBatchSize(8192)
    w.initializeWriter()
    t0 = time.perf_counter()
    for b in t.to_batches(max_chunksize=65536):
        w.write(b)
    st0 = w.writer_stats()
    t1 = time.perf_counter()
    res = w.flush()
    t2 = time.perf_counter()
    w.close()
    files = sum(len(v) for v in res.values())
    size = sum(f.fileSize for v in res.values() for f in v)
    print(f"rep {rep}: {n} rows, {nb} buckets -> {files} files, {size/1e6:.1f} MB; write calls {t1-t0:.2f} s, flush {t2-t1:.2f} s "
          f"({n/(t2-t0)/1e6:.2f} M rows/s); before flush: {st0}")
