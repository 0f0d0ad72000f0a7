# SPDX-License-Identifier: Apache-2.0
"""One full-size bucket of configs 2..5 (all three encodings) through the C-ABI against the CPU oracle
(development check; the pytest suite runs the same configs scaled down)."""
import os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # repo root (this script lives in tests/: it uses the oracle)
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mor_oracle as mo
from conftest import assert_tables_equal
from lakesoul_b200 import datagen, read_bucket
for cfg in (2, 3, 4, 5):
    for enc in ("E1", "E2", "E3"):
        spec = datagen.config_spec(cfg, 1.0)
        spec.buckets = 1
        tmp = tempfile.mkdtemp(prefix="lsb_full_")
        paths = datagen.write_table(tmp, spec, encoding=enc)[0]
        t0 = time.perf_counter()
        got, rd = read_bucket(paths, spec.schema, spec.primary_keys, batch_size=8192, options=spec.options, return_reader=True)
        dt = time.perf_counter() - t0
        st = rd.stats(); rd.close()
        want = mo.read_bucket(paths, spec.schema, spec.primary_keys, options=spec.options)
        assert_tables_equal(got, want, f"config {cfg} {enc}")
        print(f"config {cfg} {enc}: files {len(paths)} rows_in {st['rows_in']} rows_out {st['rows_out']} ok  ({dt*1e3:.1f} ms, "
              f"{st['device_decoded_pages']} device pages, {st['host_decoded_files']} host files)", flush=True)
