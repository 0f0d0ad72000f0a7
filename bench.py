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
import shutil
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
TRAFFIC_FILE = "r03_traffic.json"  # ncu dram bytes of the step kernels (profiles/make_traffic.py)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--buckets", type=int, default=BUCKETS, help="hash buckets per GPU")
    ap.add_argument("--e2e-buckets", type=int, default=4096, help="hash buckets of the end-to-end table (config 5: 4096)")
    ap.add_argument("--e2e-distinct", type=int, default=256,
                    help="distinct bucket contents generated; the remaining bucket ids are hard links to them")
    ap.add_argument("--e2e-passes", type=int, default=5, help="timed passes over the whole table (median reported)")
    ap.add_argument("--e2e-threads", type=int, default=0,
                    help="concurrent per-bucket readers (task slots) per rank; 0 = --e2e-slot-cap")
    ap.add_argument("--e2e-slot-cap", type=int, default=24)
    ap.add_argument("--side-buckets", type=int, default=64, help="buckets of the E1/E2/E3 side comparison (1 GPU only)")
    ap.add_argument("--keep-data", action="store_true")
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
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0: float = None, t1: float = None) -> dict:
        """clocks over the samples taken inside [t0, t1] (the timed region); if the region was too short to catch
        one, over every sample since the sampler started (warm-up steps: same kernels, same load)"""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        inside = [ln for ts, ln in self.lines if t0 is not None and t0 <= ts <= t1 + 0.05]
        window = "timed region" if inside else "warm-up + timed region"
        for ln in (inside or [ln for _, ln in self.lines]):
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
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


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
            # no synchronisation here: the gather runs while the host plans the next step (the only wait of a step is the
            # winner-count read-back between merge and gather); the event pairs are read after the timed loop
            ev[2].record()
            record.setdefault("events", []).append(ev)
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
# host topology: cores this rank's readers may use
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


def _parse_cpulist(text: str):
    out = []
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        out.extend(range(int(a), int(b or a) + 1))
    return out


def gpu_numa_nodes(n_gpus: int):
    """NUMA node of every GPU 0..n_gpus-1 (sysfs, via the PCI bus id nvidia-smi reports); -1 when unknown."""
    nodes = [-1] * n_gpus
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=index,pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=20).stdout
        for ln in out.splitlines():
            idx, bus = [x.strip() for x in ln.split(",")]
            if int(idx) >= n_gpus:
                continue
            bus = bus.lower()
            if len(bus.split(":")[0]) == 8:  # 00000000:1B:00.0 -> 0000:1b:00.0
                bus = bus[4:]
            try:
                nodes[int(idx)] = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
            except Exception:
                pass
    except Exception:
        pass
    return nodes


def rank_cpus(local_rank: int, world: int):
    """CPUs for this rank's reader threads: whole physical cores of its GPU's NUMA node, shared evenly with the
    other ranks whose GPU sits on the same node (SCALE_r01: GPU0-3 on node 0, GPU4-7 on node 1)."""
    allowed = sorted(os.sched_getaffinity(0))
    nodes = gpu_numa_nodes(world)
    my = nodes[local_rank] if local_rank < len(nodes) else -1
    node_cpus = allowed
    if my >= 0:
        try:
            lst = set(_parse_cpulist(open(f"/sys/devices/system/node/node{my}/cpulist").read()))
            inter = [c for c in allowed if c in lst]
            if inter:
                node_cpus = inter
        except Exception:
            pass
    peers = [r for r in range(world) if (nodes[r] if r < len(nodes) else -1) == my] or [local_rank]
    cores = {}
    for c in node_cpus:
        try:
            key = (open(f"/sys/devices/system/cpu/cpu{c}/topology/physical_package_id").read().strip(),
                   open(f"/sys/devices/system/cpu/cpu{c}/topology/core_id").read().strip())
        except Exception:
            key = ("0", str(c))
        cores.setdefault(key, []).append(c)
    core_list = [cores[k] for k in sorted(cores, key=lambda k: (int(k[0]), int(k[1])))]
    i = peers.index(local_rank)
    per = max(1, len(core_list) // len(peers))  # a contiguous block of whole cores per rank
    mine = core_list[i * per:(i + 1) * per] if i < len(peers) - 1 else core_list[i * per:]
    cpus = sorted(c for grp in (mine or core_list) for c in grp)
    return cpus, my


# ------------------------------------------------------------------------------------------------
# e2e leg: the config-5-shaped table as Parquet files -> C-ABI readers -> host Arrow batches
# ------------------------------------------------------------------------------------------------
E2E_CONFIG = 5
E2E_COMMITS = 64
E2E_ROWS_PER_BUCKET = 146_484
E2E_PKS = ["l_orderkey", "l_linenumber"]


def table_root(args) -> str:
    if args.data_dir:
        return args.data_dir
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else tempfile.gettempdir()
    tag = os.environ.get("MASTER_PORT", "0") + "_" + os.environ.get("TORCHELASTIC_RUN_ID", str(os.getppid()))
    return os.path.join(base, f"lsb_bench_{tag}")


def make_table(root: str, buckets: int, distinct: int, encodings: str, part: int, parts: int, procs: int, links: bool):
    cmd = [sys.executable, os.path.join(ROOT, "tools", "make_table.py"), "--root", root, "--config", str(E2E_CONFIG),
           "--buckets", str(buckets), "--distinct", str(distinct), "--encodings", encodings, "--part", str(part),
           "--parts", str(parts), "--procs", str(max(1, procs))]
    if links:
        cmd.append("--links-only")
    out = subprocess.run(cmd, capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("make_table failed: " + out.stderr[-800:])
    return json.loads(out.stdout.strip().splitlines()[-1])


def bucket_paths(root: str, encoding: str, bucket: int):
    from lakesoul_b200 import datagen
    spec_seed = datagen.BASE_SEED + 1000 * E2E_CONFIG
    return datagen.bucket_file_paths(root, encoding, bucket, E2E_COMMITS, spec_seed)


class NativeReaders:
    """tools/cabi_driver.cc in --serve mode: T task slots, one C-ABI reader per bucket, pinned to `cpus`."""

    def __init__(self, paths, threads: int, cpus, env=None):
        self.drv = os.path.join(ROOT, "lakesoul_b200", "lib", "cabi_driver")
        if not os.path.exists(self.drv):
            raise RuntimeError("lakesoul_b200/lib/cabi_driver is not built (run __graft_entry__.build())")
        fd, self.job = tempfile.mkstemp(prefix="lsb_job_", suffix=".txt")
        with os.fdopen(fd, "w") as f:
            f.write(f"{threads} {BATCH_SIZE} 1\n" + "\t".join(E2E_PKS) + "\n\n\n")
            f.write("\n".join("\t".join(p) for p in paths) + "\n")
        e = dict(os.environ)
        e.update(env or {})

        def pin():
            if cpus:
                try:
                    os.sched_setaffinity(0, cpus)
                except Exception:
                    pass
        self.errf = tempfile.TemporaryFile(mode="w+")  # a pipe nobody drains could block the driver
        self.p = subprocess.Popen([self.drv, self.job, "--serve"], stdin=subprocess.PIPE, stdout=subprocess.PIPE,
                                  stderr=self.errf, text=True, env=e, preexec_fn=pin)
        line = self.p.stdout.readline().strip()
        if line != "READY":
            raise RuntimeError("cabi_driver did not start: " + line + self._stderr())

    def _stderr(self) -> str:
        self.errf.seek(0)
        return self.errf.read()[-1500:]

    def run_pass(self, n: int = 0) -> dict:
        self.p.stdin.write(f"GO {n}\n" if n else "GO\n")
        self.p.stdin.flush()
        line = self.p.stdout.readline()
        if not line or "error" in line:
            raise RuntimeError("cabi_driver pass failed: " + self._stderr())
        return json.loads(line)

    def close(self):
        try:
            self.p.stdin.write("QUIT\n")
            self.p.stdin.flush()
            self.p.wait(timeout=30)
        except Exception:
            self.p.kill()
        if os.environ.get("LAKESOUL_B200_TRACE") == "1":
            sys.stderr.write(self._stderr())
        try:
            os.unlink(self.job)
        except OSError:
            pass


def e2e_env(encoding: str):
    """Reader knobs per file layout."""
    env = {}
    if "LAKESOUL_B200_DEVICE_ZSTD_SHARE" in os.environ:
        env["LAKESOUL_B200_DEVICE_ZSTD_SHARE"] = os.environ["LAKESOUL_B200_DEVICE_ZSTD_SHARE"]
    return env


# ------------------------------------------------------------------------------------------------
# CPU baseline: the reference-structured CPU reader (oracle/) on a bounded sample
# ------------------------------------------------------------------------------------------------
def run_cpu_baseline(paths, rows, threads: int = 0):
    """The reference-structured CPU reader (oracle/ref_cpu) with one reader per usable host core.
    The bounded sample = the given buckets, replicated so that every thread owns at least one."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_cpu_driver
    from lakesoul_b200 import datagen
    threads = threads or usable_cores()
    reps = max(1, -(-threads // len(paths)))
    res = ref_cpu_driver.run(paths * reps, datagen.LINEITEM_SCHEMA, E2E_PKS, batch_size=BATCH_SIZE, threads=threads)
    total = sum(rows) * reps
    return {"value": total / res["seconds"], "unit": "rows/s", "cores": res["threads"], "kind": res["kind"],
            "sample": f"{len(paths)} distinct config-5 buckets x {E2E_COMMITS} ZSTD files read {reps}x ({total} input rows, "
                      f"{res['seconds']:.2f} s), {res['impl']}, one reader per host thread "
                      f"({res['threads']} threads on {os.cpu_count()} logical CPUs)"}, res


def workload_config(args, world: int):
    return {"workload": f"HBM-resident leg (value, roofline): config2, {args.buckets} hash buckets x {K} delta files per GPU, "
                        f"TPC-H lineitem schema, PK (l_orderkey,l_linenumber), last-write-wins, batch_size {BATCH_SIZE}; "
                        f"end-to-end leg (e2e): config5-shaped full scan, {args.e2e_buckets} hash buckets x {E2E_COMMITS} delta files "
                        f"x {E2E_ROWS_PER_BUCKET} rows per bucket ({args.e2e_buckets * E2E_ROWS_PER_BUCKET} input rows), PLAIN+ZSTD "
                        f"Parquet as the LakeSoul writer emits it, the SAME table at every N (strong scaling), buckets sharded by "
                        f"contiguous byte-balanced ranges; {min(args.e2e_distinct, args.e2e_buckets)} distinct bucket contents, the other "
                        f"bucket ids are hard links to them (bounds the generation time; every bucket is opened, read, decoded and merged on its own)",
            "rows_in_per_gpu": args.buckets * (BASE_ROWS + (K - 1) * DELTA_ROWS),
            "e2e_rows_in": args.e2e_buckets * E2E_ROWS_PER_BUCKET,
            "l2_policy": "inputs (~10 GB per GPU) are far larger than the 126 MB L2; no explicit flush",
            "parallelism": f"bucket-sharded x{world}"}


def reference_arm(args, rank: int, config):
    """The reference's own CPU path (C++ port of the loser-tree reader; the Rust crate cannot be built here) on a
    bounded sample of the same table: `cpu_buckets` config-5 buckets in the writer's layout, all host threads."""
    if rank != 0:
        return
    root = table_root(args) + "_ref"
    info = make_table(root, args.cpu_buckets, args.cpu_buckets, "E3", 0, 1, usable_cores(), links=False)
    paths = [bucket_paths(root, "E3", b) for b in range(args.cpu_buckets)]
    rows = [info["rows"][str(b)] for b in range(args.cpu_buckets)]
    vals = []
    for i in range(args.warmup + args.steps):
        cb, res = run_cpu_baseline(paths, rows, args.cpu_threads)
        if i >= args.warmup:
            vals.append((cb, res))
    cb = vals[-1][0]
    v = float(np.median([c["value"] for c, _ in vals]))
    cb["value"] = v
    line = {"metric": "merge-on-read rows/sec", "value": v, "unit": "rows/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": float(np.median([r["seconds"] for _, r in vals]) * 1e3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
            "impl": "reference", "config": config, "cpu_baseline": cb, "statistic": "median over the timed steps",
            "e2e": {"value": v, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    shutil.rmtree(root, ignore_errors=True)


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    config = workload_config(args, world)
    if args.impl == "reference":
        reference_arm(args, rank, config)
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
    from lakesoul_b200.sharding import balanced_ranges

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ============ HBM-resident leg: `value`, `roofline` ===============================================
    # bucket ranges balanced by bytes (NCCL all-gather of per-bucket sizes)
    my_sizes = torch.full((args.buckets,), float(BASE_ROWS + (K - 1) * DELTA_ROWS) * 169.0, device="cuda")
    if world > 1:
        gathered = [torch.empty_like(my_sizes) for _ in range(world)]
        dist.all_gather(gathered, my_sizes)
        sizes = torch.cat(gathered).cpu().numpy()
    else:
        sizes = my_sizes.cpu().numpy()
    lo, hi = balanced_ranges(sizes, world)[rank]
    shard = Shard(hi - lo, seed=20260921 + 2000 + rank)
    sampler = ClockSampler(local_rank)
    sampler.start()  # nvidia-smi needs a moment to start: it samples from the warm-up steps on (same kernels, same load)
    for _ in range(args.warmup):
        shard.step()
    rec = {"merge_ms": [], "gather_ms": []}
    barrier()
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
    t_wall1 = time.perf_counter()
    wall = t_wall1 - t_wall0
    clocks = sampler.stop(t_wall0, t_wall1)
    ms = torch.tensor([ev0.elapsed_time(ev1)], device="cuda")
    tot = torch.tensor([rows_in, rows_out], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_total = float(ms.item())
    rows_in_all, rows_out_all = int(tot[0].item()), int(tot[1].item())
    value = rows_in_all / (ms_total / 1e3)
    U = rows_out // args.steps
    mb, gb = shard.algorithmic_bytes(U)
    for e3 in rec.get("events", []):  # all steps are complete (barrier above)
        rec["merge_ms"].append(e3[0].elapsed_time(e3[1]))
        rec["gather_ms"].append(e3[1].elapsed_time(e3[2]))
    merge_ms = float(np.mean(rec["merge_ms"]))
    gather_ms = float(np.mean(rec["gather_ms"]))
    del shard
    torch.cuda.empty_cache()

    line = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        ach = (mb + gb) / ((merge_ms + gather_ms) / 1e3) / 1e9
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", TRAFFIC_FILE)))
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
                "clocks": clocks, "roofline": roofline, "wall_s": wall,
                "scaling_note": "value = HBM-resident kernels on one config-2 shard per GPU (weak: nothing is exchanged); "
                                "e2e = one fixed table read from files by all ranks (strong, e2e.scaling)"}

    # ============ end-to-end leg: files -> C-ABI -> host Arrow batches, every rank =====================
    if not args.no_e2e:
        root = table_root(args)
        B = args.e2e_buckets
        D = min(args.e2e_distinct, B)
        cores_total = usable_cores()
        procs = max(1, cores_total // world)
        t_gen0 = time.perf_counter()
        info = make_table(root, B, D, "E3", rank, world, procs, links=False)
        barrier()
        make_table(root, B, D, "E3", rank, world, procs, links=True)
        barrier()
        gen_s = time.perf_counter() - t_gen0
        # real per-bucket byte sizes: every rank stats a slice, NCCL all-reduce, same split everywhere
        mine = torch.zeros(B, dtype=torch.float64, device="cuda")
        for b in range(rank, B, world):
            mine[b] = float(sum(os.path.getsize(p) for p in bucket_paths(root, "E3", b)))
        if world > 1:
            dist.all_reduce(mine, op=dist.ReduceOp.SUM)
        bsizes = mine.cpu().numpy()
        lo, hi = balanced_ranges(bsizes, world)[rank]
        paths = [bucket_paths(root, "E3", b) for b in range(lo, hi)]
        cpus, numa = rank_cpus(local_rank, world)
        # task slots per rank: with page decompression on the GPU a slot mostly sleeps on its stream, so the number in
        # flight is what saturates the GPU (measured optimum 24 on one B200: 45 / 53 / 58 / 60 / 54 M rows/s at
        # 12 / 16 / 20 / 24 / 32 slots), not the core count
        slots = args.e2e_threads or args.e2e_slot_cap
        readers = NativeReaders(paths, slots, cpus, e2e_env("E3"))
        readers.run_pass(min(len(paths), 2 * slots))  # warm-up: pinned pool, arenas, page cache
        passes = []
        for _ in range(args.e2e_passes):
            barrier()
            r = readers.run_pass()
            t = torch.tensor([r["seconds"]], dtype=torch.float64, device="cuda")
            c = torch.tensor([r["rows_in"], r["rows_out"], r["h2d_bytes"], r["d2h_bytes"], r["kernel_launches"],
                              r["device_decoded_pages"], r["device_decompressed_pages"], r["host_decoded_files"],
                              int(r["cpu_seconds"] * 1e6)], dtype=torch.int64, device="cuda")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dist.all_reduce(c, op=dist.ReduceOp.SUM)
            passes.append((float(t.item()), [int(x) for x in c.tolist()], r["seconds"]))
        readers.close()
        if rank == 0:
            secs = sorted(p[0] for p in passes)
            med = secs[len(secs) // 2]
            cnt = passes[0][1]
            line["e2e"] = {"value": cnt[0] / med, "unit": "rows/s", "h2d_bytes_per_step": cnt[2], "d2h_bytes_per_step": cnt[3],
                           "scaling": "strong", "encoding": "E3 (PLAIN + ZSTD, dictionary off: multipart_writer.rs:121-133)",
                           "rows_in": cnt[0], "rows_out": cnt[1], "seconds": med, "pass_seconds": [p[0] for p in passes],
                           "statistic": f"median of {len(passes)} passes over all {B} buckets, each pass = max over ranks",
                           "gpu_launches": cnt[4], "launches_per_bucket": cnt[4] / B, "device_decoded_pages": cnt[5],
                           "device_decompressed_pages": cnt[6], "host_decoded_files": cnt[7],
                           "host_cores_busy": cnt[8] / 1e6 / med, "task_slots_per_rank": slots, "ranks": world,
                           "rank0_cpus": f"{len(cpus)} CPUs of NUMA node {numa}", "buckets": B, "distinct_buckets": D,
                           "table_bytes": float(bsizes.sum()), "generation_s": gen_s,
                           "consumer": "native task executor tools/cabi_driver.cc (C-ABI calls only), one process per GPU",
                           "how": "every rank reads its byte-balanced contiguous bucket range through the lakesoul-io-c C-ABI, one reader "
                                  "per bucket on a pool of task slots pinned to the GPU's NUMA node; pread into pinned memory, file images "
                                  "sent to the GPU while the next file is read, device page decompression + decode + merge + gather, "
                                  "D2H of the Arrow buffers consumed by the caller; barrier in front of every pass, time = max over ranks"}
        # other file layouts beside the headline (1 GPU only: the same first buckets in E1 / E2 / E3)
        if world == 1 and args.side_buckets > 0:
            nb_side = min(args.side_buckets, B)
            side = {}
            make_table(root, nb_side, nb_side, "E1,E2", 0, 1, procs, links=False)
            for enc in ("E3", "E1", "E2"):
                sp = [bucket_paths(root, enc, b) for b in range(nb_side)]
                rd = NativeReaders(sp, slots, cpus, e2e_env(enc))
                rd.run_pass(min(nb_side, 2 * slots))
                rs = [rd.run_pass() for _ in range(args.e2e_passes)]
                rd.close()
                rs.sort(key=lambda x: x["seconds"])
                m = rs[len(rs) // 2]
                side[enc] = {"value": m["rows_in"] / m["seconds"], "seconds": m["seconds"], "rows_in": m["rows_in"],
                             "device_decompressed_pages": m["device_decompressed_pages"],
                             "host_decoded_files": m["host_decoded_files"], "host_cores_busy": m["cpu_seconds"] / m["seconds"]}
            line["e2e"]["layouts"] = {"buckets": nb_side, "statistic": f"median of {args.e2e_passes}",
                                      "E1 = PLAIN uncompressed, E2 = dictionary + SNAPPY (pyarrow/Spark default), E3 = PLAIN + ZSTD": side}
        barrier()
        if rank == 0:
            if not args.no_cpu and world == 1:
                try:
                    nb_cpu = min(args.cpu_buckets, D)
                    cp = [bucket_paths(root, "E3", b) for b in range(nb_cpu)]
                    cb, _ = run_cpu_baseline(cp, [info["rows"].get(str(b), E2E_ROWS_PER_BUCKET) for b in range(nb_cpu)],
                                             args.cpu_threads)
                    line["cpu_baseline"] = cb
                except Exception as e:  # the baseline is reporting only
                    line["cpu_baseline"] = {"value": None, "unit": "rows/s", "cores": 0, "kind": "port", "sample": f"failed: {e}"}
            if not args.keep_data:
                shutil.rmtree(root, ignore_errors=True)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
