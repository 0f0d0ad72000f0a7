# SPDX-License-Identifier: Apache-2.0
"""Per source line: executed warp instructions and stall samples of an .ncu-rep (needs -lineinfo + --import-source on).
usage: python profiles/ncu_lines.py <file.ncu-rep> [top-n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = next(r for r in rows if r and r[0] == "Line No")
ix = {n: i for i, n in enumerate(hdr)}  # first occurrence wins for duplicate names
ix["Source"] = hdr.index("Source")
lines = [r for r in rows if len(r) == len(hdr) and r[0] not in ("", "Line No")]
tot_i = sum(int(r[ix["Instructions Executed"]] or 0) for r in lines) or 1
tot_s = sum(int(r[ix["# Samples"]] or 0) for r in lines) or 1
print(f"total warp instructions {tot_i}  samples {tot_s}")
for r in sorted(lines, key=lambda r: -int(r[ix["Instructions Executed"]] or 0))[:top]:
    i = int(r[ix["Instructions Executed"]] or 0); s = int(r[ix["# Samples"]] or 0)
    print(f"{r[0]:>5s} inst {100*i/tot_i:5.1f}%  samples {100*s/tot_s:5.1f}%  {r[ix['Source']].strip()[:110]}")
