# SPDX-License-Identifier: Apache-2.0
"""Digest of an .ncu-rep: headline metrics, stall-reason totals and the hottest SASS instructions.
usage: python profiles/ncu_digest.py <file.ncu-rep>"""
import csv, subprocess, sys, io
rep = sys.argv[1]
# optional 3rd argument: index of the launch inside the report (default 0)
sel = ["--launch-skip", sys.argv[3], "--launch-count", "1"] if len(sys.argv) > 3 else ["--launch-count", "1"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"] + sel, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__shared_mem_per_block_allocated", "smsp__inst_executed.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct"]
print("==", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "")
for h, u, v in zip(hdr, units, vals):
    if h in want:
        print(f"{h:80s} {v} {u}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + sel, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
# one section per kernel: a "Kernel Name" row, a header row, then one row per SASS instruction
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] or [0]
k = min(int(sys.argv[3]) if len(sys.argv) > 3 else 0, len(starts) - 1) if len(starts) > 1 else 0
lo = starts[k]; hi = starts[k + 1] if k + 1 < len(starts) else len(rows)
h = rows[lo + 1]; data = [r for r in rows[lo + 2:hi] if len(r) == len(h)]
idx = {n: i for i, n in enumerate(h)}
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
tot = {s: sum(int(r[idx[s]] or 0) for r in data) for s in stalls}
T = sum(tot.values()) or 1
print("-- stall reasons")
for s, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]:
    print(f"   {s:26s} {100*v/T:5.1f}%")
print("-- hottest SASS")
for r in sorted(data, key=lambda r: -int(r[idx["# Samples"]] or 0))[:int(sys.argv[2]) if len(sys.argv) > 2 else 14]:
    print("  ", r[idx["# Samples"]].rjust(7), r[idx["Source"]].strip()[:64].ljust(64), "thr", r[idx["Avg. Threads Executed"]])
