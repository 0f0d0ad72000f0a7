# SPDX-License-Identifier: Apache-2.0
"""Builds (part of) a bench-scale table: full-size hash buckets of one BASELINE config, each generated
independently by a pool of processes, plus hard-linked replicas of them under further bucket ids.

    python tools/make_table.py --root DIR --config 5 --buckets 4096 --distinct 256 --encodings E3 \
        [--part R --parts N] [--procs P]

Buckets 0..distinct-1 are generated (this invocation builds those with index % parts == part); bucket b >= distinct
is a hard link to bucket b % distinct (same bytes, its own path: the reader opens, reads, decodes and merges it like
any other bucket). Prints one JSON line: {"rows": {bucket: input rows}, "seconds": ...}.
"""
import argparse
import json
import os
import sys
import time
from concurrent.futures import ProcessPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _one(args):
    root, config, bucket, encodings = args
    from lakesoul_b200 import datagen
    return bucket, datagen.write_full_bucket(root, config, bucket, encodings)


def _link(args):
    root, config, encs, distinct, buckets = args
    from lakesoul_b200 import datagen
    spec = datagen.config_spec(config, 1.0)
    for enc in encs:
        for b in buckets:
            src = datagen.bucket_file_paths(root, enc, b % distinct, spec.commits, spec.seed)
            dst = datagen.bucket_file_paths(root, enc, b, spec.commits, spec.seed)
            for s, d in zip(src, dst):
                if not os.path.exists(d):
                    os.link(s, d)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--root", required=True)
    ap.add_argument("--config", type=int, default=5)
    ap.add_argument("--buckets", type=int, default=4096)
    ap.add_argument("--distinct", type=int, default=256)
    ap.add_argument("--encodings", default="E3")
    ap.add_argument("--part", type=int, default=0)
    ap.add_argument("--parts", type=int, default=1)
    ap.add_argument("--procs", type=int, default=0)
    ap.add_argument("--links-only", action="store_true")
    a = ap.parse_args()
    from lakesoul_b200 import datagen
    encs = a.encodings.split(",")
    t0 = time.perf_counter()
    rows = {}
    distinct = min(a.distinct, a.buckets)
    if not a.links_only:
        mine = [d for d in range(distinct) if d % a.parts == a.part]
        procs = a.procs or max(1, len(os.sched_getaffinity(0)))
        with ProcessPoolExecutor(max_workers=min(procs, max(len(mine), 1))) as ex:
            for b, n in ex.map(_one, [(a.root, a.config, d, encs) for d in mine]):
                rows[b] = n
    else:
        todo = [b for b in range(distinct, a.buckets) if b % a.parts == a.part]
        procs = a.procs or max(1, len(os.sched_getaffinity(0)))
        chunks = [todo[i::procs] for i in range(procs) if todo[i::procs]]
        with ProcessPoolExecutor(max_workers=max(len(chunks), 1)) as ex:
            list(ex.map(_link, [(a.root, a.config, encs, distinct, c) for c in chunks]))
    print(json.dumps({"rows": rows, "seconds": time.perf_counter() - t0}))


if __name__ == "__main__":
    main()
