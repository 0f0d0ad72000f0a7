# SPDX-License-Identifier: Apache-2.0
"""Pinned host<->device copy bandwidth of this box (development probe)."""
import torch, time
n = 256 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory(); h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda"); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
a = t(lambda: d.copy_(h, non_blocking=True)); b = t(lambda: h2.copy_(d2, non_blocking=True))
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
c = t(both)
print("H2D %.1f GB/s  D2H %.1f GB/s  both %.1f + %.1f GB/s" % (n / a / 1e9, n / b / 1e9, n / c / 1e9, n / c / 1e9))
# 4 MiB pieces like the reader
m = 4 << 20
def pieces():
    for i in range(0, n, m): d[i:i + m].copy_(h[i:i + m], non_blocking=True)
e = t(pieces)
print("H2D in 4 MiB pieces %.1f GB/s" % (n / e / 1e9))
