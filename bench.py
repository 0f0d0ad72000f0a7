# SPDX-License-Identifier: Apache-2.0
"""bench.py -- merge-on-read rows/s on B200 (BASELINE.json metric), one JSON line on rank 0.

A "step" = one full pass of the hot path over one scan batch of the N=1 workload (BASELINE.json
configs[1]: 256 hash buckets x 8 delta files, TPC-H lineitem schema, last-write-wins):
    plan (host) -> PK k-way merge kernels -> winner counts -> gather of all 16 columns into
    batch-segmented Arrow buffers.
`value`   : inputs (decoded columns) already resident in HBM when the timed region starts.
`e2e`     : the same metric through the reference-facing C-ABI (NativeIOReader per bucket, host
            Parquet files in, host Arrow batches out; H2D/D2H inside the timed region).
`roofline`: PK merge + gather kernels, algorithmic bytes / CUDA-event time.
`cpu_baseline` / `--impl reference`: the C++ port of the reference's loser-tree reader
            (oracle/ref_cpu, Arrow C++ Parquet decode) on the box's host cores.
Multi-GPU: one process per GPU (torchrun), hash buckets sharded by contiguous ranges balanced by
bytes (NCCL all-gather of per-bucket sizes), row counts / checksums all-reduced; weak scaling
(256 config-2 buckets per GPU).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

# config 2
BUCKETS = 256
K = 8
BASE_ROWS = 117_188
DELTA_ROWS = 16_741
BATCH_SIZE = 8192


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--buckets", type=int, default=BUCKETS, help="hash buckets per GPU")
    ap.add_argument("--e2e-buckets", type=int, default=64, help="buckets written as Parquet for the e2e leg")
    ap.add_argument("--e2e-threads", type=int, default=0,
                    help="concurrent per-bucket readers (task slots); 0 = usable host cores")
    ap.add_argument("--cpu-buckets", type=int, default=32, help="distinct buckets in the bounded CPU baseline sample")
    ap.add_argument("--cpu-threads", type=int, default=0, help="CPU baseline reader threads (0 = usable host cores)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--data-dir", default=None)
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.idx = device_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# synthetic config-2 shard resident in HBM
# ------------------------------------------------------------------------------------------------
FIXED_COLS = [("l_partkey", 8), ("l_suppkey", 8), ("l_quantity", 16), ("l_extendedprice", 16), ("l_discount", 16),
              ("l_tax", 16), ("l_shipdate", 4), ("l_commitdate", 4), ("l_receiptdate", 4)]
STRING_COLS = [("l_returnflag", 1, 1), ("l_linestatus", 1, 1), ("l_shipinstruct", 4, 17), ("l_shipmode", 3, 7),
               ("l_comment", 10, 43)]


class Shard:
    """nb buckets x K files of lineitem-shaped columns in HBM (file-major inside a bucket)."""

    def __init__(self, nb: int, seed: int):
        import torch
        self.torch = torch
        g = torch.Generator(device="cuda")
        g.manual_seed(seed)
        dev = "cuda"
        nd = K - 1
        per_bucket = BASE_ROWS + nd * DELTA_ROWS
        self.nb, self.rows_per_bucket = nb, per_bucket
        self.n = nb * per_bucket
        # base keys: strictly increasing multiples of 4 (room for inserts in between)
        gaps = torch.randint(1, 4, (nb, BASE_ROWS), generator=g, device=dev, dtype=torch.int64)
        base = torch.cumsum(gaps, dim=1) * 4
        hi = base[:, -1:].clone()
        # deltas: 50 % updates of base keys, 49 % inserts, 1 % in-file duplicates
        n_upd = DELTA_ROWS // 2
        n_dup = DELTA_ROWS // 100
        n_ins = DELTA_ROWS - n_upd - n_dup
        idx = torch.randint(0, BASE_ROWS, (nb, nd, n_upd), generator=g, device=dev)
        upd = torch.gather(base.unsqueeze(1).expand(nb, nd, BASE_ROWS), 2, idx)
        ins = (torch.rand((nb, nd, n_ins), generator=g, device=dev) * hi.unsqueeze(1)).to(torch.int64) // 4 * 4 + \
            torch.randint(1, 4, (nb, nd, n_ins), generator=g, device=dev)
        both = torch.cat([upd, ins], dim=2)
        didx = torch.randint(0, both.shape[2], (nb, nd, n_dup), generator=g, device=dev)
        dup = torch.gather(both, 2, didx)
        deltas, _ = torch.sort(torch.cat([both, dup], dim=2), dim=2)
        enc = torch.cat([base, deltas.reshape(nb, nd * DELTA_ROWS)], dim=1).reshape(-1)
        # composite PK (l_orderkey int64, l_linenumber int32): lexicographic order == enc order
        self.l_orderkey = (enc >> 3).contiguous()
        self.l_linenumber = (enc & 7).to(torch.int32).contiguous()
        del gaps, base, idx, upd, ins, both, didx, dup, deltas, enc
        n = self.n
        self.fixed = {}
        for name, w in FIXED_COLS:
            words = w // 4
            self.fixed[name] = torch.randint(-2**31, 2**31 - 1, (n, words), generator=g, device=dev, dtype=torch.int32)
        self.strings = {}
        for name, lo, hi_len in STRING_COLS:
            lens = torch.randint(lo, hi_len + 1, (n,), generator=g, device=dev, dtype=torch.int32)
            offs = torch.zeros(n + 1, dtype=torch.int64, device=dev)
            torch.cumsum(lens, dim=0, out=offs[1:])
            total = int(offs[-1].item())
            assert total < 2**31 - 1, "string column exceeds int32 offsets in one shard"
            data = torch.randint(97, 123, (total,), generator=g, device=dev, dtype=torch.uint8)
            self.strings[name] = (offs.to(torch.int32), data, total)
            del lens, offs
        self.bucket_file_begin = np.arange(nb + 1, dtype=np.int32) * K
        rows = [BASE_ROWS] + [DELTA_ROWS] * nd
        self.file_rows = np.tile(np.array(rows, dtype=np.int64), nb)
        self.zero_base = torch.zeros(nb, dtype=torch.int64, device=dev)
        self.out = None
        torch.cuda.synchronize()

    def alloc_outputs(self, U: int, TB: int, vstride: int):
        torch = self.torch
        if self.out is not None and self.out["U"] >= U and self.out["TB"] >= TB:
            return self.out
        o = {"U": U, "TB": TB}
        o["l_orderkey"] = torch.empty(U, dtype=torch.int64, device="cuda")
        o["l_linenumber"] = torch.empty(U, dtype=torch.int32, device="cuda")
        for name, w in FIXED_COLS:
            o[name] = torch.empty((U, w // 4), dtype=torch.int32, device="cuda")
        for name, _, _ in STRING_COLS:
            total = self.strings[name][2]
            o[name] = (torch.empty(U + TB, dtype=torch.int32, device="cuda"),
                       torch.empty(total + 16, dtype=torch.uint8, device="cuda"),
                       torch.empty(TB + 1, dtype=torch.int64, device="cuda"))
        self.out = o
        return o

    def step(self, record=None):
        """one pass: plan + merge + counts + gather of every column. Returns (rows_in, rows_out)."""
        from lakesoul_b200.engine import PK_SIGNED, ScanPlan
        torch = self.torch
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)] if record is not None else None
        plan = ScanPlan(self.bucket_file_begin, self.file_rows, key_bits=96, batch_size=BATCH_SIZE)
        if ev:
            ev[0].record()
        plan.merge([self.l_orderkey, self.l_linenumber], [PK_SIGNED, PK_SIGNED])
        ob, bb = plan.sync_counts()
        if ev:
            ev[1].record()
        U, TB = int(ob[-1]), int(bb[-1])
        o = self.alloc_outputs(U, TB, plan.valid_stride)
        srcs = [self.l_orderkey, self.l_linenumber] + [self.fixed[n] for n, _ in FIXED_COLS]
        dsts = [o["l_orderkey"], o["l_linenumber"]] + [o[n] for n, _ in FIXED_COLS]
        widths = [8, 4] + [w for _, w in FIXED_COLS]
        vcols = []
        for name, _, _ in STRING_COLS:
            offs, data, _ = self.strings[name]
            do, dd, db = o[name]
            vcols.append(dict(src_offsets=offs, src_data=data, src_data_base=self.zero_base, dst_offsets=do, dst_data=dd,
                              batch_data_start=db))
        plan.gather_all(srcs, dsts, widths, vcols)
        if ev:
            ev[2].record()
            torch.cuda.synchronize()
            record["merge_ms"].append(ev[0].elapsed_time(ev[1]))
            record["gather_ms"].append(ev[1].elapsed_time(ev[2]))
        st = plan.stats()
        plan.close()
        return self.n, U, st

    def algorithmic_bytes(self, U: int):
        """N*w_pk + 2*U*sum(w_c) (+ no validity: columns are non-null)."""
        w_pk = 12
        w_fixed = 12 + sum(w for _, w in FIXED_COLS)
        str_bytes = 0.0
        for name, lo, hi in STRING_COLS:
            str_bytes += 4 + (lo + hi) / 2.0
        merge_bytes = self.n * w_pk
        gather_bytes = 2 * U * (w_fixed + str_bytes)
        return merge_bytes, gather_bytes


# ------------------------------------------------------------------------------------------------
# e2e leg: Parquet files -> C-ABI readers -> host Arrow batches
# ------------------------------------------------------------------------------------------------
def make_e2e_files(nb: int, root: str, encoding: str = "E1"):
    from lakesoul_b200 import datagen
    spec = datagen.config_spec(2, 1.0)
    spec.buckets = nb
    tables = datagen.generate_table(spec)
    paths = datagen.write_table(root, spec, tables, encoding=encoding)
    rows = [sum(t.num_rows for t in b) for b in tables]
    return spec, paths, rows


def run_e2e(spec, paths, threads: int):
    from concurrent.futures import ThreadPoolExecutor
    from lakesoul_b200 import LakeSoulArrowReader, NativeIOReader

    def read_one(files):
        r = NativeIOReader()
        for f in files:
            r.addFile(f)
        r.setPrimaryKeys(spec.primary_keys)
        r.setSchema(spec.schema)
        r.setBatchSize(BATCH_SIZE)
        r.initializeReader()
        rows = 0
        for batch in LakeSoulArrowReader(r):
            rows += batch.num_rows
        st = r.stats()
        r.close()
        return rows, st

    from lakesoul_b200 import load_library
    load_library()
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        res = list(ex.map(read_one, paths))
    dt = time.perf_counter() - t0
    out_rows = sum(r for r, _ in res)
    h2d = sum(s["h2d_bytes"] for _, s in res)
    d2h = sum(s["d2h_bytes"] for _, s in res)
    launches = sum(s["kernel_launches"] for _, s in res)
    run_e2e.last_decode = (sum(s["device_decoded_pages"] for _, s in res), sum(s["host_decoded_files"] for _, s in res))
    return dt, out_rows, h2d, d2h, launches


def run_e2e_native(spec, paths, threads: int, repeat: int = 3):
    """Same call sequence from a native task executor (tools/cabi_driver.cc): T threads, one reader per
    bucket through the C-ABI, every batch released by the consumer. No Python in the timed path."""
    drv = os.path.join(ROOT, "lakesoul_b200", "lib", "cabi_driver")
    if not os.path.exists(drv):
        return None
    inp = (f"{threads} {BATCH_SIZE} {repeat}\n" + "\t".join(spec.primary_keys) + "\n\n\n" +
           "\n".join("\t".join(p) for p in paths) + "\n")
    env = dict(os.environ)
    out = subprocess.run([drv], input=inp, capture_output=True, text=True, env=env)
    if out.returncode != 0:
        raise RuntimeError("cabi_driver failed: " + out.stderr[-500:])
    if os.environ.get("LAKESOUL_B200_TRACE") == "1":
        sys.stderr.write(out.stderr)
    return json.loads(out.stdout.strip().splitlines()[-1])


# ------------------------------------------------------------------------------------------------
# CPU baseline: the reference-structured CPU reader (oracle/) on a bounded sample
# ------------------------------------------------------------------------------------------------
def usable_cores() -> int:
    """Cores this process may run on (affinity mask, then the cgroup CPU quota)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = max(1, min(n, int(int(quota) / int(period))))
    except Exception:
        pass
    return n


def run_cpu_baseline(spec, paths, rows, nb: int, threads: int = 0):
    """The reference-structured CPU reader (oracle/ref_cpu) with one reader per usable host core.
    The bounded sample = `nb` distinct buckets, replicated so that every thread owns at least one."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_cpu_driver
    threads = threads or usable_cores()
    paths, rows = paths[:nb], rows[:nb]
    reps = max(1, -(-threads // len(paths)))
    res = ref_cpu_driver.run(paths * reps, spec.schema, spec.primary_keys, batch_size=BATCH_SIZE, threads=threads)
    total = sum(rows) * reps
    return {"value": total / res["seconds"], "unit": "rows/s", "cores": res["threads"], "kind": res["kind"],
            "sample": f"{len(paths)} distinct config-2 buckets x {K} files read {reps}x ({total} input rows, "
                      f"{res['seconds']:.2f} s), {res['impl']}, one reader per host thread "
                      f"({res['threads']} threads on {os.cpu_count()} logical CPUs)"}, res


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    config = {"workload": f"config2: {args.buckets} hash buckets x {K} delta files per GPU, TPC-H lineitem schema, "
                          f"PK (l_orderkey,l_linenumber), last-write-wins, batch_size {BATCH_SIZE}",
              "rows_in_per_gpu": args.buckets * (BASE_ROWS + (K - 1) * DELTA_ROWS),
              "l2_policy": "inputs (~10 GB per GPU) are far larger than the 126 MB L2; no explicit flush",
              "parallelism": f"bucket-sharded x{world}"}

    if args.impl == "reference":
        # the reference's own CPU path (C++ port of the loser-tree reader; the Rust crate cannot be built here)
        if rank != 0:
            return
        tmp = args.data_dir or tempfile.mkdtemp(prefix="lsb_bench_")
        spec, paths, rows = make_e2e_files(args.cpu_buckets, tmp)
        vals = []
        for i in range(args.warmup + args.steps):
            cb, res = run_cpu_baseline(spec, paths, rows, args.cpu_buckets, args.cpu_threads)
            if i >= args.warmup:
                vals.append((cb, res))
        cb = vals[-1][0]
        v = float(np.mean([c["value"] for c, _ in vals]))
        cb["value"] = v
        line = {"metric": "merge-on-read rows/sec", "value": v, "unit": "rows/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": float(np.mean([r["seconds"] for _, r in vals]) * 1e3),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
                "impl": "reference", "config": config, "cpu_baseline": cb,
                "e2e": {"value": v, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    os.environ["LAKESOUL_B200_DEVICE"] = str(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from lakesoul_b200 import load_library
    lib = load_library()
    if lib.lsb_init_device() < 0:
        raise RuntimeError(lib.lsb_last_error().decode())

    # ---- bucket ranges balanced by bytes (NCCL all-gather of per-bucket sizes) ------------------
    nb_total = args.buckets * world
    my_sizes = torch.full((args.buckets,), float(BASE_ROWS + (K - 1) * DELTA_ROWS) * 169.0, device="cuda")
    if world > 1:
        gathered = [torch.empty_like(my_sizes) for _ in range(world)]
        dist.all_gather(gathered, my_sizes)
        sizes = torch.cat(gathered).cpu().numpy()
    else:
        sizes = my_sizes.cpu().numpy()
    from lakesoul_b200.sharding import balanced_ranges
    ranges = balanced_ranges(sizes, world)
    lo, hi = ranges[rank]
    shard = Shard(hi - lo, seed=20260921 + 2000 + rank)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        shard.step()
    rec = {"merge_ms": [], "gather_ms": []}
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    ev0.record()
    rows_in = rows_out = 0
    launches = 0
    for _ in range(args.steps):
        n, u, st = shard.step(rec)
        rows_in += n
        rows_out += u
        launches += st["launches"] + 1 + 1  # merge kernels + fixed-width gather + string gather
    ev1.record()
    barrier()
    wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    ms = torch.tensor([ev0.elapsed_time(ev1)], device="cuda")
    tot = torch.tensor([rows_in, rows_out], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_total = float(ms.item())
    rows_in_all, rows_out_all = int(tot[0].item()), int(tot[1].item())
    value = rows_in_all / (ms_total / 1e3)

    if rank == 0:
        U = rows_out // args.steps
        mb, gb = shard.algorithmic_bytes(U)
        merge_ms = float(np.mean(rec["merge_ms"]))
        gather_ms = float(np.mean(rec["gather_ms"]))
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        ach = (mb + gb) / ((merge_ms + gather_ms) / 1e3) / 1e9
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
            traffic = tj["dram_bytes_per_step"]
        except Exception:
            pass
        roofline = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                    "peak_source": "MEASURED_PEAKS.json (of measured)" if peaks else "6650 GB/s (of fallback)",
                    "kernels": "PK merge (part+tile+scan+compact) + winner gather (fixed + var)",
                    "merge_ms": merge_ms, "gather_ms": gather_ms,
                    "merge_GBps": mb / (merge_ms / 1e3) / 1e9, "gather_GBps": gb / (gather_ms / 1e3) / 1e9,
                    "algorithmic_bytes": {"merge": mb, "gather": gb}}
        line = {"metric": "merge-on-read rows/sec", "value": value, "unit": "rows/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
                "config": config, "rows_out_per_step": rows_out_all // args.steps, "gpu_launches": launches,
                "clocks": clocks, "roofline": roofline, "wall_s": wall}
        e2e_ctx = None
        if not args.no_e2e or not args.no_cpu:
            tmp = args.data_dir or tempfile.mkdtemp(prefix="lsb_bench_")
            e2e_ctx = make_e2e_files(max(args.e2e_buckets, args.cpu_buckets), tmp)
        if not args.no_e2e:
            spec, paths, rows = e2e_ctx
            args.e2e_threads = args.e2e_threads or usable_cores()
            paths_e, rows_e = paths[:args.e2e_buckets], rows[:args.e2e_buckets]
            run_e2e(spec, paths_e[:min(4, len(paths_e))], args.e2e_threads)  # warm-up (pinned pool, page cache)
            best = None
            for _ in range(3):
                r = run_e2e(spec, paths_e, args.e2e_threads)
                if best is None or r[0] < best[0]:
                    best = r
            dt, out_rows, h2d, d2h, e_launch = best
            py_val = sum(rows_e) / dt
            dec = run_e2e.last_decode
            harness = "python ctypes harness (lakesoul_b200.NativeIOReader)"
            host_busy = None
            nat = run_e2e_native(spec, paths_e, args.e2e_threads)
            if nat is not None and nat["rows_out"] == out_rows:
                dt, h2d, d2h, e_launch = nat["seconds"], nat["h2d_bytes"], nat["d2h_bytes"], nat["kernel_launches"]
                dec = (nat["device_decoded_pages"], nat["host_decoded_files"])
                harness = "native task executor tools/cabi_driver.cc (C-ABI calls only)"
                host_busy = nat.get("cpu_seconds", 0.0) / nat["seconds"]
            line["e2e"] = {"value": sum(rows_e) / dt, "unit": "rows/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                           "rows_in": sum(rows_e), "rows_out": out_rows, "seconds": dt, "gpu_launches": e_launch,
                           "device_decoded_pages": dec[0], "host_decoded_files": dec[1],
                           "consumer": harness, "python_harness_value": py_val, "task_slots": args.e2e_threads,
                           "host_cores_busy": host_busy,
                           "how": f"{len(paths_e)} config-2 buckets (Parquet PLAIN, uncompressed) read through the "
                                  f"lakesoul-io-c C-ABI by {args.e2e_threads} concurrent per-bucket readers (one per task slot); "
                                  "pread into pinned memory, every file image sent to the GPU while the next one is read, "
                                  "device page decode + merge/gather kernels, one D2H of the Arrow buffers consumed by "
                                  "the caller; best of 3"}
        if not args.no_cpu:
            spec, paths, rows = e2e_ctx
            try:
                cb, _ = run_cpu_baseline(spec, paths, rows, args.cpu_buckets, args.cpu_threads)
                line["cpu_baseline"] = cb
            except Exception as e:  # the baseline is reporting only
                line["cpu_baseline"] = {"value": None, "unit": "rows/s", "cores": 0, "kind": "port", "sample": f"failed: {e}"}
        print(json.dumps(line))
    if world > 1:
        # rank 0 is still measuring the end-to-end and CPU legs on the host cores: the other ranks wait on the
        # rendezvous store (a blocking socket read) instead of spinning in an NCCL barrier
        store = dist.distributed_c10d._get_default_store()
        if rank == 0:
            store.set("lsb_bench_done", "1")
        else:
            store.wait(["lsb_bench_done"])
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
