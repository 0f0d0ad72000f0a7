# SPDX-License-Identifier: Apache-2.0
"""End-to-end rows/s of the native C-ABI task executor (tools/cabi_driver.cc) over task-slot counts."""
import os, sys, tempfile, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 64
threads = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [8, 16, 24, 32, 48]
enc = sys.argv[3] if len(sys.argv) > 3 else "E1"
tmp = tempfile.mkdtemp(prefix="lsb_sweep_")
spec, paths, rows = bench.make_e2e_files(nb, tmp, encoding=enc)
print("usable cores", bench.usable_cores(), "buckets", nb, "rows", sum(rows), "encoding", enc, flush=True)
for th in threads:
    r = bench.run_e2e_native(spec, paths, th, repeat=3)
    print(th, "slots:", round(sum(rows) / r["seconds"] / 1e6, 1), "Mrows/s", round(r["seconds"] * 1e3, 1), "ms",
          "h2d GB/s", round(r["h2d_bytes"] / r["seconds"] / 1e9, 1), "cpu cores busy", round(r.get("cpu_seconds", 0) / r["seconds"], 1), flush=True)
