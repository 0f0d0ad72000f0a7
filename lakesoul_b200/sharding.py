# SPDX-License-Identifier: Apache-2.0
"""Hash-bucket -> GPU assignment: rank r owns a contiguous range of bucket ids whose
byte sizes are balanced by a prefix-sum split. Every rank computes the same split from the same
all-gathered size vector; no payload crosses NVLink. Host logic, covered by world_size-2 gloo tests."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def balanced_ranges(bucket_bytes: Sequence[float], world: int) -> List[Tuple[int, int]]:
    """Split [0, B) into `world` contiguous ranges with near-equal byte totals (empty ranges allowed
    when B < world: surplus GPUs idle)."""
    sizes = np.asarray(bucket_bytes, dtype=np.float64)
    B = len(sizes)
    prefix = np.concatenate([[0.0], np.cumsum(sizes)])
    total = prefix[-1]
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        # first boundary whose prefix reaches the target, never moving backwards
        c = int(np.searchsorted(prefix, target, side="left"))
        if c > 0 and abs(prefix[c - 1] - target) <= abs(prefix[min(c, B)] - target):
            c -= 1
        cuts.append(min(max(c, cuts[-1]), B))
    cuts.append(B)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def all_gather_sizes(local_sizes: np.ndarray, group=None) -> np.ndarray:
    """All-gather per-bucket {bytes} vectors (equal length per rank) with torch.distributed
    (NCCL on GPUs, gloo in CPU tests)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return np.asarray(local_sizes, dtype=np.float64)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.as_tensor(np.asarray(local_sizes, dtype=np.float64), device=dev)
    out = [torch.empty_like(t) for _ in range(dist.get_world_size(group))]
    dist.all_gather(out, t, group=group)
    return torch.cat(out).cpu().numpy()


def all_reduce_counts(counts: Sequence[int], group=None) -> List[int]:
    """Sum {rows_in, rows_out, checksum...} over ranks."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return [int(c) for c in counts]
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.as_tensor(list(counts), dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return [int(x) for x in t.cpu().tolist()]
