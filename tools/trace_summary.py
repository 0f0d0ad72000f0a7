# SPDX-License-Identifier: Apache-2.0
"""Summarise LSBTRACE lines (LAKESOUL_B200_TRACE=1): mean ms per phase, per group of N readers."""
import sys, collections
rows = []
for line in open(sys.argv[1]):
    if not line.startswith("LSBTRACE"): continue
    parts = line.split()[2:]
    d = collections.OrderedDict()
    for p in parts:
        k, v = p.split("="); d[k] = d.get(k, 0.0) + float(v) if k != "t0" else float(v)
    rows.append(d)
group = int(sys.argv[2]) if len(sys.argv) > 2 else len(rows)
for g0 in range(0, len(rows), group):
    rs = rows[g0:g0 + group]
    keys = [k for k in rs[0] if k != "t0"]
    span = max(r["t0"] + sum(v for k, v in r.items() if k != "t0") for r in rs) - min(r["t0"] for r in rs)
    print("readers %d span %.1f ms | " % (len(rs), span) + "  ".join("%s %.2f" % (k, sum(r.get(k, 0) for r in rs) / len(rs)) for k in keys)
          + "  | total %.2f" % (sum(sum(v for k, v in r.items() if k != "t0") for r in rs) / len(rs)))
