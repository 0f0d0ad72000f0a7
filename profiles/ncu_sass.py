# SPDX-License-Identifier: Apache-2.0
"""SASS listing of one launch of an .ncu-rep in address order: instructions executed (warp level), samples, text.
usage: python profiles/ncu_sass.py <file.ncu-rep> [launch-index]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
sel = ["--launch-skip", sys.argv[2], "--launch-count", "1"] if len(sys.argv) > 2 else ["--launch-count", "1"]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + sel, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1] if rows[0] and rows[0][0] == "Kernel Name" else rows[0]
start = rows.index(h) + 1
idx = {n: i for i, n in enumerate(h)}
tot = 0
for r in rows[start:]:
    if len(r) != len(h):
        continue
    n = int(r[idx["Instructions Executed"]] or 0)
    tot += n
    print(f"{n:>11d} {r[idx['# Samples']]:>7s} {tot/1e6:9.1f}M  {r[idx['Source']].strip()}")
