# SPDX-License-Identifier: Apache-2.0
"""Where does the end-to-end time of the per-bucket C-ABI readers go? (development probe)"""
import os, sys, time, tempfile, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from concurrent.futures import ThreadPoolExecutor
import numpy as np
import bench

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 32
tmp = tempfile.mkdtemp(prefix="lsb_probe_")
t0 = time.time()
spec, paths, rows = bench.make_e2e_files(nb, tmp)
print("datagen+write s", round(time.time() - t0, 1), "rows", sum(rows))
from lakesoul_b200 import load_library, NativeIOReader, LakeSoulArrowReader
lib = load_library(); assert lib.lsb_init_device() >= 0

def read_one(files):
    r = NativeIOReader()
    for f in files: r.addFile(f)
    r.setPrimaryKeys(spec.primary_keys); r.setSchema(spec.schema); r.setBatchSize(8192)
    t = time.perf_counter(); r.initializeReader(); t1 = time.perf_counter()
    n = 0
    for b in LakeSoulArrowReader(r): n += b.num_rows
    t2 = time.perf_counter()
    st = r.stats(); st["init_s"] = t1 - t; st["iter_s"] = t2 - t1; r.close(); return st

for th in (1, 4, 16, 32):
    for rep in range(2):
        t = time.perf_counter()
        with ThreadPoolExecutor(max_workers=th) as ex: res = list(ex.map(read_one, paths))
        dt = time.perf_counter() - t
    agg = {k: float(np.mean([s[k] for s in res])) for k in ("decode_ms", "upload_ms", "merge_ms", "gather_ms", "download_ms", "init_s", "iter_s")}
    print(th, "threads:", round(sum(rows) / dt / 1e6, 2), "Mrows/s", {k: round(v, 2) for k, v in agg.items()})
if os.environ.get('PROBE_CPU') != '1': sys.exit(0)
import ref_cpu_driver
for th in (8, 32, 128):
    reps = max(1, th // len(paths))
    res = ref_cpu_driver.run(paths * reps, spec.schema, spec.primary_keys, threads=th)
    print("cpu port", th, "threads", round(sum(rows) * reps / res["seconds"] / 1e6, 2), "Mrows/s", res)
