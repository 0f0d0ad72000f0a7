import csv, collections, sys
lines=[l for l in open(sys.argv[1]) if not l.startswith("==")]
rows=list(csv.DictReader(lines))
d=collections.OrderedDict()
for x in rows:
    e=d.setdefault(x["ID"],{"name":x["Kernel Name"][:36],"grid":x["Grid Size"]})
    v=float(x["Metric Value"].replace(",","")); u=x["Metric Unit"]
    if u in ("Gbyte","GB"): v*=1e3
    if u in ("Kbyte","KB"): v/=1e3
    if u in ("byte",): v/=1e6
    e[x["Metric Name"]]=v
tot=0
for k,v in d.items():
    t=v.get("gpu__time_duration.sum",0)/1e6; tot+=t
    print(f'{v["name"]:38s} {v["grid"]:16s} {t:8.3f} ms  rd {v.get("dram__bytes_read.sum",0):9.1f} MB  wr {v.get("dram__bytes_write.sum",0):9.1f} MB')
print("total ms", tot)
