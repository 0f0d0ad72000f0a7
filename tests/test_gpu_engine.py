# SPDX-License-Identifier: Apache-2.0
"""Device-level engine parity (through the C-ABI extension): many buckets per launch, adversarial
duplicate structure, multi-level sampling, 128-bit keys, string gather."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _expected_winners(keys_per_file, bases):
    """numpy oracle: last (file, row) of every key group, keys ascending (nulls not used here)."""
    k = len(keys_per_file)
    allk = np.concatenate(keys_per_file) if k else np.empty(0, dtype=np.int64)
    fid = np.concatenate([np.full(len(a), j) for j, a in enumerate(keys_per_file)]) if k else np.empty(0, dtype=np.int64)
    row = np.concatenate([np.arange(len(a)) for a in keys_per_file]) if k else np.empty(0, dtype=np.int64)
    if allk.ndim == 2:
        order = np.lexsort((row, fid) + tuple(allk[:, c] for c in range(allk.shape[1] - 1, -1, -1)))
        ks = allk[order]
        last = np.append(np.any(ks[1:] != ks[:-1], axis=1), True) if len(ks) else np.empty(0, dtype=bool)
    else:
        order = np.lexsort((row, fid, allk))
        ks = allk[order]
        last = np.append(ks[1:] != ks[:-1], True) if len(ks) else np.empty(0, dtype=bool)
    sel = order[last]
    return np.asarray(bases)[fid[sel]] + row[sel]


def _run(buckets, batch_size=4096, pad=False, composite=False):
    import torch
    from lakesoul_b200.engine import PK_SIGNED, ScanPlan
    bfb = [0]
    rows, all_keys = [], []
    for files in buckets:
        bfb.append(bfb[-1] + len(files))
        for a in files:
            rows.append(len(a))
            all_keys.append(a)
    F = len(rows)
    if pad:
        base, pos = [], 0
        for n in rows:
            base.append(pos)
            pos = (pos + n + 1 + 63) & ~63
        span = pos
    else:
        base = np.concatenate([[0], np.cumsum(rows)])[:-1].tolist() if F else []
        span = int(sum(rows))
    ncol = 2 if composite else 1
    host = [np.zeros(max(span, 1), dtype=np.int64), np.zeros(max(span, 1), dtype=np.int32)]
    for f, a in enumerate(all_keys):
        if composite:
            host[0][base[f]:base[f] + len(a)] = a[:, 0]
            host[1][base[f]:base[f] + len(a)] = a[:, 1]
        else:
            host[0][base[f]:base[f] + len(a)] = a
    dev = [torch.from_numpy(h).cuda() for h in host[:ncol]]
    plan = ScanPlan(bfb, rows, key_bits=96 if composite else 64, batch_size=batch_size,
                    file_row_base=base if pad else None, row_span=span)
    plan.merge(dev, [PK_SIGNED] * ncol)
    ob, bb = plan.sync_counts()
    win = plan.winners().cpu().numpy()
    f0 = 0
    for b, files in enumerate(buckets):
        want = _expected_winners(files, base[f0:f0 + len(files)])
        got = win[ob[b]:ob[b + 1]]
        assert len(got) == len(want), f"bucket {b}: {len(got)} winners, expected {len(want)}"
        bad = np.flatnonzero(got != want)
        assert bad.size == 0, f"bucket {b}: first mismatch at output row {bad[:5]}: got {got[bad[:5]]} want {want[bad[:5]]}"
        f0 += len(files)
    stats = plan.stats()
    plan.close()
    return stats, ob, bb


def _sorted_keys(rng, n, space):
    return np.sort(rng.integers(0, space, size=n)).astype(np.int64)


@pytest.mark.parametrize("pad", [False, True])
def test_small_buckets_various_k(pad):
    rng = np.random.default_rng(1)
    buckets = []
    for b in range(40):
        k = int(rng.integers(1, 9))
        buckets.append([_sorted_keys(rng, int(rng.integers(0, 300)), 500) for _ in range(k)])
    buckets.append([np.empty(0, dtype=np.int64)])          # empty bucket
    buckets.append([np.array([5, 5, 5], dtype=np.int64)])  # one key only
    _run(buckets, pad=pad)


def test_tiles_and_one_sampling_level():
    rng = np.random.default_rng(2)
    buckets = [[_sorted_keys(rng, n, 200_000) for n in (60_000, 9_000, 9_000, 8_000, 1, 0, 7_000, 12_000)] for _ in range(3)]
    stats, _, _ = _run(buckets)
    assert stats["levels"] == 1 and stats["tiles_level0"] > 3 * 10


def test_two_sampling_levels_large_bucket():
    rng = np.random.default_rng(3)
    buckets = [[_sorted_keys(rng, 1_000_000, 3_000_000), _sorted_keys(rng, 200_000, 3_000_000)]]
    stats, _, _ = _run(buckets, batch_size=8192)
    assert stats["levels"] >= 2


def test_heavy_duplicates_cannot_overflow_a_tile():
    """One key repeated far more often than a tile holds, in several files (bound is positional)."""
    rng = np.random.default_rng(4)
    hot = np.full(20_000, 777, dtype=np.int64)
    buckets = [[np.sort(np.concatenate([hot, _sorted_keys(rng, 5_000, 2_000)])) for _ in range(4)]]
    _run(buckets)


def test_64_deltas_per_bucket():
    rng = np.random.default_rng(5)
    buckets = [[_sorted_keys(rng, 40_000 if j == 0 else 700, 60_000) for j in range(64)] for _ in range(2)]
    stats, _, _ = _run(buckets)
    assert stats["log_s"] < 6


def test_negative_keys_signed_order():
    rng = np.random.default_rng(6)
    buckets = [[np.sort(rng.integers(-2**62, 2**62, size=5000)).astype(np.int64) for _ in range(3)]]
    _run(buckets)


def test_composite_key_128bit():
    rng = np.random.default_rng(7)
    buckets = []
    for _ in range(3):
        files = []
        for n in (30_000, 4_000, 4_000, 2_000):
            a = np.stack([rng.integers(0, 9_000, size=n), rng.integers(1, 8, size=n)], axis=1).astype(np.int64)
            a = a[np.lexsort((a[:, 1], a[:, 0]))]
            files.append(a)
        buckets.append(files)
    _run(buckets, composite=True)


def test_composite_key_wide_span_takes_the_generic_tile_kernel():
    """First key column spread over the whole int64 range: a tile's key span exceeds 2^64 in the packed 96-bit space,
    so the fast (tile-relative 64-bit) kernel hands those tiles to the generic one; dense ranges in the same scan stay
    on the fast path. Negative values in both key columns (sign-flipped images)."""
    rng = np.random.default_rng(17)
    buckets = []
    for wide in (True, False, True):
        files = []
        for n in (20_000, 6_000, 3_000, 3_000, 500):
            if wide:
                c0 = rng.integers(-2**62, 2**62, size=n)
                c0[: n // 4] = rng.choice(c0, size=n // 4)  # duplicates across and inside files
            else:
                c0 = rng.integers(-3_000, 3_000, size=n)
            a = np.stack([c0, rng.integers(-3, 4, size=n)], axis=1).astype(np.int64)
            a = a[np.lexsort((a[:, 1], a[:, 0]))]
            files.append(a)
        buckets.append(files)
    pool = np.stack([rng.integers(-2**62, 2**62, size=6_000), rng.integers(-3, 4, size=6_000)], axis=1).astype(np.int64)
    shared = []
    for _ in range(6):  # the same wide keys in many files: long equal-key groups across tile boundaries
        a = pool[rng.integers(0, len(pool), size=9_000)]
        shared.append(a[np.lexsort((a[:, 1], a[:, 0]))])
    buckets.append(shared)
    _run(buckets, composite=True, pad=True)


def test_gather_fixed_and_strings_match_numpy():
    import torch
    from lakesoul_b200.engine import PK_SIGNED, ScanPlan
    rng = np.random.default_rng(8)
    files = [_sorted_keys(rng, n, 50_000) for n in (30_000, 5_000, 5_000)]
    rows = [len(a) for a in files]
    base, pos = [], 0
    for n in rows:
        base.append(pos)
        pos = (pos + n + 1 + 63) & ~63
    span = pos
    keys = np.zeros(span, dtype=np.int64)
    v32 = rng.integers(-2**31, 2**31 - 1, size=span).astype(np.int32)
    v128 = rng.integers(0, 2**62, size=(span, 2)).astype(np.int64)
    valid_bits = rng.random(span) < 0.8
    lens = rng.integers(0, 20, size=span).astype(np.int32)
    offsets = np.zeros(span + 1, dtype=np.int32)
    for f, a in enumerate(files):
        keys[base[f]:base[f] + len(a)] = a
    np.cumsum(lens, out=offsets[1:])
    data = rng.integers(97, 123, size=int(offsets[-1]), dtype=np.uint8)
    bs = 1000
    plan = ScanPlan([0, 3], rows, key_bits=64, batch_size=bs, file_row_base=base, row_span=span)
    d_keys = torch.from_numpy(keys).cuda()
    plan.merge([d_keys], [PK_SIGNED])
    ob, bb = plan.sync_counts()
    U, TB = int(ob[-1]), int(bb[-1])
    win = plan.winners().cpu().numpy()
    vs = plan.valid_stride
    d_v32, d_v128 = torch.from_numpy(v32).cuda(), torch.from_numpy(v128).cuda()
    d_valid = torch.from_numpy(np.packbits(valid_bits, bitorder="little")).cuda()
    o32 = torch.zeros(U, dtype=torch.int32, device="cuda")
    o128 = torch.zeros((U, 2), dtype=torch.int64, device="cuda")
    ovalid = torch.zeros(TB * vs, dtype=torch.uint8, device="cuda")
    plan.gather_fixed([d_v32, d_v128], [o32, o128], [4, 16], src_valid=[d_valid, None], dst_valid=[ovalid, None])
    assert np.array_equal(o32.cpu().numpy(), v32[win])
    assert np.array_equal(o128.cpu().numpy(), v128[win])
    ov = ovalid.cpu().numpy()
    for i in range(TB):
        r = min(bs, U - i * bs)
        bits = np.unpackbits(ov[i * vs:(i + 1) * vs], bitorder="little")[:r]
        assert np.array_equal(bits.astype(bool), valid_bits[win[i * bs:i * bs + r]])
    # strings
    d_off, d_data = torch.from_numpy(offsets).cuda(), torch.from_numpy(data).cuda()
    d_base = torch.zeros(1, dtype=torch.int64, device="cuda")
    o_off = torch.zeros(U + TB, dtype=torch.int32, device="cuda")
    o_data = torch.zeros(len(data) + 16, dtype=torch.uint8, device="cuda")
    o_bstart = torch.zeros(TB + 1, dtype=torch.int64, device="cuda")
    plan.gather_var(d_off, d_data, d_base, o_off, o_data, o_bstart)
    oo, od, ob_ = o_off.cpu().numpy(), o_data.cpu().numpy(), o_bstart.cpu().numpy()
    for i in range(TB):
        r = min(bs, U - i * bs)
        off = oo[i * bs + i:i * bs + i + r + 1]
        assert off[0] == 0
        w = win[i * bs:i * bs + r]
        assert np.array_equal(np.diff(off), lens[w])
        chunk = od[ob_[i]:ob_[i + 1]]
        want = np.concatenate([data[offsets[g]:offsets[g + 1]] for g in w]) if r else np.empty(0, dtype=np.uint8)
        assert np.array_equal(chunk, want)
    plan.close()
