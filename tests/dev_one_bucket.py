# SPDX-License-Identifier: Apache-2.0
"""Read a few config-2 buckets through the C-ABI once or twice (development probe, meant to run under ncu)."""
import os, sys, time, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 1
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
enc = sys.argv[3] if len(sys.argv) > 3 else "E1"
tmp = tempfile.mkdtemp(prefix="lsb_one_")
spec, paths, rows = bench.make_e2e_files(nb, tmp, encoding=enc)
from lakesoul_b200 import load_library, read_bucket
lib = load_library(); assert lib.lsb_init_device() >= 0
for rep in range(reps):
    for files in paths:
        t = time.perf_counter()
        tb, rd = read_bucket(files, spec.schema, spec.primary_keys, batch_size=8192, return_reader=True)
        dt = time.perf_counter() - t
        print("rep", rep, "rows", tb.num_rows, "ms", round(dt * 1e3, 2), {k: (round(v, 2) if isinstance(v, float) else v) for k, v in rd.stats().items()})
        rd.close()
