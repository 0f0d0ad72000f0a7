# SPDX-License-Identifier: Apache-2.0
"""Times the fixed-width and the string gather of the bench shard separately (CUDA events, no profiler)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, bench
from lakesoul_b200.engine import PK_SIGNED, ScanPlan
sh = bench.Shard(bench.BUCKETS, 1)
plan = ScanPlan(sh.bucket_file_begin, sh.file_rows, key_bits=96, batch_size=bench.BATCH_SIZE)
plan.merge([sh.l_orderkey, sh.l_linenumber], [PK_SIGNED, PK_SIGNED])
ob, bb = plan.sync_counts()
U, TB = int(ob[-1]), int(bb[-1])
o = sh.alloc_outputs(U, TB, plan.valid_stride)
srcs = [sh.l_orderkey, sh.l_linenumber] + [sh.fixed[n] for n, _ in bench.FIXED_COLS]
dsts = [o["l_orderkey"], o["l_linenumber"]] + [o[n] for n, _ in bench.FIXED_COLS]
widths = [8, 4] + [w for _, w in bench.FIXED_COLS]
vcols = []
for name, _, _ in bench.STRING_COLS:
    offs, data, _ = sh.strings[name]
    do, dd, db = o[name]
    vcols.append(dict(src_offsets=offs, src_data=data, src_data_base=sh.zero_base, dst_offsets=do, dst_data=dd, batch_data_start=db))
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
print("rows out", U, "fixed ms", round(t(lambda: plan.gather_fixed(srcs, dsts, widths)), 3),
      "var(all 5) ms", round(t(lambda: plan.gather_var_multi(vcols)), 3))
for i, (name, lo, hi) in enumerate(bench.STRING_COLS):
    print(name, lo, hi, "ms", round(t(lambda: plan.gather_var_multi([vcols[i]])), 3))
print("both ms", round(t(lambda: plan.gather_all(srcs, dsts, widths, vcols)), 3))
