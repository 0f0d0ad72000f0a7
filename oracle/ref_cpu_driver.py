# SPDX-License-Identifier: Apache-2.0
"""Driver of the CPU baseline (TEST/BENCH INFRASTRUCTURE ONLY -- never imported by lakesoul_b200/).

Runs the reference-structured CPU reader over a list of buckets, one reader per host thread (the
reference's deployment shape: one merge thread per reader, rust/lakesoul-io-c/src/lib.rs:1385-1391).
Uses oracle/_build/ref_cpu (C++ port: Arrow C++ Parquet decode + loser-tree merge + interleave) when
it has been built, else the vectorised pyarrow oracle."""
from __future__ import annotations

import json
import os
import subprocess
import time
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REF_CPU = os.path.join(HERE, "_build", "ref_cpu")


def run(paths, schema, primary_keys, batch_size=8192, threads=None, verify_dir=None):
    """verify_dir: the C++ port also writes every bucket's merged output as <verify_dir>/bucket_<i>.arrow (IPC file)."""
    threads = threads or (os.cpu_count() or 1)
    if os.path.exists(REF_CPU):
        spec = (f"{threads} {batch_size} {1 if verify_dir else 0}\n" + "\t".join(primary_keys) + "\n" + "\t".join(schema.names) + "\n" +
                "\n".join("\t".join(p) for p in paths) + "\n")
        t0 = time.perf_counter()
        env = dict(os.environ)
        if verify_dir:
            env["REF_CPU_VERIFY_DIR"] = verify_dir
        out = subprocess.run([REF_CPU], input=spec, capture_output=True, text=True, check=True, env=env)
        res = json.loads(out.stdout.strip().splitlines()[-1])
        res.setdefault("seconds", time.perf_counter() - t0)
        res.update(kind="port", impl="C++ port of rust/lakesoul-io SortedStreamMerger/UseLastRangeCombiner (oracle/ref_cpu.cc)")
        return res
    import mor_oracle as mo
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        rows = list(ex.map(lambda fs: mo.read_bucket(fs, schema, primary_keys).num_rows, paths))
    return {"seconds": time.perf_counter() - t0, "threads": threads, "rows_out": sum(rows), "kind": "port",
            "impl": "vectorised pyarrow restatement (oracle/mor_oracle.py)"}
