# SPDX-License-Identifier: Apache-2.0
"""End-to-end rows/s of the native C-ABI task executor (tools/cabi_driver.cc) over task-slot counts (development probe).
usage: dev_slots_sweep.py <config> <buckets> <encoding> <slots,comma separated>"""
import json, os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lakesoul_b200 import datagen
cfg = int(sys.argv[1]); nb = int(sys.argv[2]); enc = sys.argv[3]; slots = [int(x) for x in sys.argv[4].split(",")]
tmp = tempfile.mkdtemp(prefix="lsb_sweep_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
subprocess.run([sys.executable, os.path.join(ROOT, "tools", "make_table.py"), "--root", tmp, "--config", str(cfg), "--buckets", str(nb),
                "--distinct", str(nb), "--encodings", enc], check=True, capture_output=True)
spec = datagen.config_spec(cfg, 1.0)
paths = [datagen.bucket_file_paths(tmp, enc, b, spec.commits, spec.seed) for b in range(nb)]
drv = os.path.join(ROOT, "lakesoul_b200", "lib", "cabi_driver")
for t in slots:
    job = os.path.join(tmp, "job.txt")
    with open(job, "w") as f:
        f.write(f"{t} 8192 {os.environ.get('SWEEP_PASSES', '3')} {os.environ.get('SWEEP_GROUP', '1')}\n" + "\t".join(spec.primary_keys) + "\n\n\n" + "\n".join("\t".join(p) for p in paths) + "\n")
    out = subprocess.run([drv, job], capture_output=True, text=True)
    if os.environ.get("LAKESOUL_B200_TRACE") == "1":
        sys.stderr.write(out.stderr)
    if out.returncode != 0:
        print("FAILED", out.stderr[-400:]); continue
    r = json.loads(out.stdout.strip().splitlines()[-1])
    print(f"slots {t:3d}: {r['rows_in'] / r['seconds'] / 1e6:7.1f} M rows/s  ({r['seconds']*1e3/nb:6.2f} ms/bucket, host cores busy "
          f"{r['cpu_seconds'] / r['seconds']:.1f}, device-decompressed pages {r['device_decompressed_pages']}) passes "
          + " ".join(f"{x*1e3/nb:.2f}" for x in r["pass_seconds"]), flush=True)
