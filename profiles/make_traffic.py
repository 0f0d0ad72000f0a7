# SPDX-License-Identifier: Apache-2.0
"""profiles/r0N_traffic.json from an ncu launch list of ONE bench step
(ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none ... --steps 1 --warmup 1):
DRAM bytes and serialised time per kernel of the LAST step in the list.
usage: python profiles/make_traffic.py profiles/r01_launches_step_kernels.csv > profiles/r0N_traffic.json"""
import csv, collections, json, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
d = collections.OrderedDict()
for x in rows:
    e = d.setdefault(int(x["ID"]), {"name": x["Kernel Name"].split("(")[0]})
    v = float(x["Metric Value"].replace(",", "")); u = x["Metric Unit"]
    v *= {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
    e[x["Metric Name"]] = v
ids = sorted(d)
# one step = the launches after the last var_write of the previous step
ends = [i for i in ids if "var_write_kernel" in d[i]["name"]]
lo = ends[-2] + 1 if len(ends) > 1 else ids[0]
step = [d[i] for i in ids if lo <= i <= ends[-1]]
byt, ms = collections.OrderedDict(), collections.OrderedDict()
for e in step:
    byt[e["name"]] = byt.get(e["name"], 0.0) + e.get("dram__bytes_read.sum", 0) + e.get("dram__bytes_write.sum", 0)
    ms[e["name"]] = ms.get(e["name"], 0.0) + e.get("gpu__time_duration.sum", 0) / 1e6
print(json.dumps({"source": sys.argv[1] + " (one bench step: 256 buckets x 8 deltas, 60 M input rows)",
                  "dram_bytes_per_step": sum(byt.values()), "dram_bytes_per_kernel": byt, "ms_per_kernel_serialised": ms}, indent=1))
