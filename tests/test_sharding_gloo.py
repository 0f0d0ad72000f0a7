# SPDX-License-Identifier: Apache-2.0
"""Multi-GPU host logic on CPU: world_size-2 gloo run of the bucket-range balancing and count reduction
 -- the only collectives the path needs."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from lakesoul_b200.sharding import balanced_ranges


def test_balanced_ranges_cover_and_balance():
    rng = np.random.default_rng(0)
    sizes = rng.integers(1, 1000, size=4096).astype(np.float64)
    for world in (1, 2, 4, 8):
        r = balanced_ranges(sizes, world)
        assert r[0][0] == 0 and r[-1][1] == len(sizes)
        assert all(a[1] == b[0] for a, b in zip(r[:-1], r[1:]))
        loads = [sizes[lo:hi].sum() for lo, hi in r]
        assert max(loads) <= sizes.sum() / world + sizes.max()
    assert balanced_ranges([5.0], 4) == [(0, 0), (0, 0), (0, 0), (0, 1)] or sum(hi - lo for lo, hi in balanced_ranges([5.0], 4)) == 1


def _worker(rank, world, port, out):
    import torch.distributed as dist
    from lakesoul_b200.sharding import all_gather_sizes, all_reduce_counts, balanced_ranges
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    local = np.arange(8, dtype=np.float64) + 100 * rank      # every rank stats its own 8 buckets
    sizes = all_gather_sizes(local)
    ranges = balanced_ranges(sizes, world)
    lo, hi = ranges[rank]
    rows_in, rows_out = int(sizes[lo:hi].sum()), hi - lo
    total = all_reduce_counts([rows_in, rows_out])
    out.put((rank, sizes.tolist(), ranges, total))
    dist.destroy_process_group()


def test_world_size_2_gloo_agrees_on_ranges_and_totals():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, s0, r0, t0), (_, s1, r1, t1) = res
    assert s0 == s1 and r0 == r1 and t0 == t1
    assert t0[1] == 16 and t0[0] == int(sum(s0))
