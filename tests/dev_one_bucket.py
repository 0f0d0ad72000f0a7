# SPDX-License-Identifier: Apache-2.0
"""Read a few full-size buckets of one BASELINE config through the C-ABI (development probe, meant to run under ncu).
usage: dev_one_bucket.py <config> <buckets> <reps> <encoding>"""
import os, sys, time, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lakesoul_b200 import datagen, read_bucket
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 5
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 1
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
enc = sys.argv[4] if len(sys.argv) > 4 else "E3"
tmp = tempfile.mkdtemp(prefix="lsb_one_")
spec = datagen.config_spec(cfg, 1.0)
for b in range(nb):
    datagen.write_full_bucket(tmp, cfg, b, [enc])
for rep in range(reps):
    for b in range(nb):
        paths = datagen.bucket_file_paths(tmp, enc, b, spec.commits, spec.seed)
        t0 = time.perf_counter()
        got, rd = read_bucket(paths, spec.schema, spec.primary_keys, batch_size=8192, options=spec.options, return_reader=True)
        dt = time.perf_counter() - t0
        st = rd.stats(); rd.close()
        print(f"rep {rep} bucket {b}: {dt*1e3:.2f} ms rows_in {st['rows_in']} out {st['rows_out']} launches {st['kernel_launches']} "
              f"decode {st['decode_ms']:.2f} upload {st['upload_ms']:.2f} merge {st['merge_ms']:.2f} gather {st['gather_ms']:.2f} "
              f"download {st['download_ms']:.2f} devdecomp {st['device_decompressed_pages']}", flush=True)
