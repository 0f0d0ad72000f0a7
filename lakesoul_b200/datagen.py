# SPDX-License-Identifier: Apache-2.0
"""Deterministic LakeSoul-shaped synthetic tables.

A table is generated the way LakeSoul writes it: one "commit" (file index j) at a time at table
level, every row routed to bucket ``spark_murmur3_32(pk, 42) mod B`` (hashing.py), every
(bucket, commit) slice sorted by primary key and written as ``part-<16 alnum>_<bucket:04>.parquet``
(reference rust/lakesoul-io/src/utils/mod.rs:30-32).  File j of a bucket is commit j: higher j wins.

Host-side helper for tests and bench.py -- not part of the read path.
"""
from __future__ import annotations

import os
import string
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq

from . import hashing

BASE_SEED = 20260921

LINEITEM_SCHEMA = pa.schema([
    pa.field("l_orderkey", pa.int64(), False), pa.field("l_partkey", pa.int64(), False),
    pa.field("l_suppkey", pa.int64(), False), pa.field("l_linenumber", pa.int32(), False),
    pa.field("l_quantity", pa.decimal128(15, 2), False), pa.field("l_extendedprice", pa.decimal128(15, 2), False),
    pa.field("l_discount", pa.decimal128(15, 2), False), pa.field("l_tax", pa.decimal128(15, 2), False),
    pa.field("l_returnflag", pa.string(), False), pa.field("l_linestatus", pa.string(), False),
    pa.field("l_shipdate", pa.date32(), False), pa.field("l_commitdate", pa.date32(), False),
    pa.field("l_receiptdate", pa.date32(), False), pa.field("l_shipinstruct", pa.string(), False),
    pa.field("l_shipmode", pa.string(), False), pa.field("l_comment", pa.string(), False),
])  # rust/lakesoul-datafusion/src/tpch/schemas.rs:7-26

CONFIG1_SCHEMA = pa.schema([pa.field("pk", pa.int64(), False)] + [pa.field(f"c{i}", pa.int32(), True) for i in range(4)])

CDC_SCHEMA = pa.schema([
    pa.field("id", pa.int64(), False), pa.field("name", pa.string(), True),
    pa.field("amount", pa.decimal128(15, 2), True), pa.field("note", pa.string(), True),
    pa.field("ts", pa.timestamp("us", "UTC"), True), pa.field("rowKinds", pa.string(), True),
])


@dataclass
class TableSpec:
    name: str
    schema: pa.Schema
    primary_keys: List[str]
    buckets: int
    commits: int                  # files per bucket
    rows_per_bucket: int          # total input rows per bucket (approximate: hashing spreads rows)
    base_frac: float = 0.5        # share of rows in commit 0
    update_frac: float = 0.5      # share of delta rows that hit existing keys
    dup_frac: float = 0.01        # share of delta rows duplicating a key inside the same file
    kind: str = "config1"         # config1 | lineitem | cdc | wide
    seed: int = BASE_SEED
    options: Dict[str, str] = field(default_factory=dict)


def config_spec(config: int, scale: float = 1.0) -> TableSpec:
    """The five BASELINE.json configs; `scale` shrinks rows per bucket AND bucket count for tests."""
    s = scale
    if config == 1:
        return TableSpec("config1", CONFIG1_SCHEMA, ["pk"], 1, 2, max(int(1_200_000 * s), 64), base_frac=1_000_000 / 1_200_000,
                         kind="config1", seed=BASE_SEED + 1000)
    if config == 2:
        return TableSpec("config2", LINEITEM_SCHEMA, ["l_orderkey", "l_linenumber"], max(int(256 * min(1, s * 4)), 2), 8,
                         max(int(234_375 * s), 256), kind="lineitem", seed=BASE_SEED + 2000)
    if config == 3:
        return TableSpec("config3", CDC_SCHEMA, ["id"], max(int(1024 * min(1, s * 4)), 2), 32, max(int(65_536 * s), 512),
                         kind="cdc", seed=BASE_SEED + 3000, options={"cdc_column": "rowKinds"})
    if config == 4:
        fields = [pa.field("pk", pa.int64(), False)]
        for st in range(8):
            fields += [pa.field(f"s{st}_a", pa.int64(), True), pa.field(f"s{st}_b", pa.float64(), True),
                       pa.field(f"s{st}_c", pa.int32(), True), pa.field(f"s{st}_d", pa.string(), True)]
        return TableSpec("config4", pa.schema(fields), ["pk"], max(int(512 * min(1, s * 4)), 2), 8, max(int(200_000 * s), 256),
                         kind="wide", seed=BASE_SEED + 4000)
    if config == 5:
        return TableSpec("config5", LINEITEM_SCHEMA, ["l_orderkey", "l_linenumber"], max(int(4096 * min(1, s * 4)), 2), 64,
                         max(int(146_484 * s), 1024), kind="lineitem", seed=BASE_SEED + 5000)
    raise ValueError(config)


# ---- payload columns ----------------------------------------------------------------------------
_WORDS = None


def _vocab(rng) -> np.ndarray:
    global _WORDS
    if _WORDS is None:
        r = np.random.Generator(np.random.PCG64(7))
        letters = np.array(list(string.ascii_lowercase))
        _WORDS = np.array(["".join(r.choice(letters, size=r.integers(2, 9))) for _ in range(4096)])
    return _WORDS


def _strings(rng, n: int, lo: int, hi: int) -> pa.Array:
    """n random lowercase strings with length U[lo, hi], built without Python loops."""
    lens = rng.integers(lo, hi + 1, size=n).astype(np.int32)
    offsets = np.zeros(n + 1, dtype=np.int32)
    np.cumsum(lens, out=offsets[1:])
    data = rng.integers(97, 123, size=int(offsets[-1]), dtype=np.uint8)
    return pa.Array.from_buffers(pa.string(), n, [None, pa.py_buffer(offsets.tobytes()), pa.py_buffer(data.tobytes())])


def _choice_strings(rng, n: int, options: Sequence[str]) -> pa.Array:
    idx = rng.integers(0, len(options), size=n)
    return pa.array(np.array(options, dtype=object)[idx], type=pa.string())


def _decimal(rng, n: int, lo: int, hi: int) -> pa.Array:
    v = rng.integers(lo, hi + 1, size=n).astype(np.int64)
    buf = np.zeros((n, 2), dtype=np.int64)
    buf[:, 0] = v
    buf[:, 1] = np.where(v < 0, -1, 0)
    return pa.Array.from_buffers(pa.decimal128(15, 2), n, [None, pa.py_buffer(buf.tobytes())])


def _with_nulls(rng, arr: pa.Array, frac: float) -> pa.Array:
    if frac <= 0:
        return arr
    mask = rng.random(len(arr)) < frac
    return pa.compute.if_else(pa.array(~mask), arr, pa.nulls(len(arr), arr.type))


def _payload(spec: TableSpec, rng, keys: Dict[str, np.ndarray], commit: int, ops: Optional[np.ndarray]) -> pa.Table:
    n = len(next(iter(keys.values())))
    cols = {}
    if spec.kind == "config1":
        cols["pk"] = pa.array(keys["pk"])
        for i in range(4):
            cols[f"c{i}"] = _with_nulls(rng, pa.array(rng.integers(-2**31, 2**31 - 1, size=n).astype(np.int32)), 0.05)
    elif spec.kind == "lineitem":
        cols["l_orderkey"] = pa.array(keys["l_orderkey"])
        cols["l_partkey"] = pa.array(rng.integers(1, 20_000_000, size=n))
        cols["l_suppkey"] = pa.array(rng.integers(1, 1_000_000, size=n))
        cols["l_linenumber"] = pa.array(keys["l_linenumber"])
        cols["l_quantity"] = _decimal(rng, n, 100, 5000)
        cols["l_extendedprice"] = _decimal(rng, n, 90_000, 10_494_950)
        cols["l_discount"] = _decimal(rng, n, 0, 10)
        cols["l_tax"] = _decimal(rng, n, 0, 8)
        cols["l_returnflag"] = _choice_strings(rng, n, ["R", "A", "N"])
        cols["l_linestatus"] = _choice_strings(rng, n, ["O", "F"])
        ship = rng.integers(8036, 10591, size=n).astype(np.int32)
        cols["l_shipdate"] = pa.array(ship).cast(pa.date32())
        cols["l_commitdate"] = pa.array((ship + rng.integers(-30, 60, size=n)).astype(np.int32)).cast(pa.date32())
        cols["l_receiptdate"] = pa.array((ship + rng.integers(1, 31, size=n)).astype(np.int32)).cast(pa.date32())
        cols["l_shipinstruct"] = _choice_strings(rng, n, ["DELIVER IN PERSON", "COLLECT COD", "NONE", "TAKE BACK RETURN"])
        cols["l_shipmode"] = _choice_strings(rng, n, ["REG AIR", "AIR", "RAIL", "SHIP", "TRUCK", "MAIL", "FOB"])
        cols["l_comment"] = _strings(rng, n, 10, 43)
    elif spec.kind == "cdc":
        cols["id"] = pa.array(keys["id"])
        cols["name"] = _strings(rng, n, 4, 24)
        cols["amount"] = _decimal(rng, n, -10_000_000, 10_000_000)
        cols["note"] = _with_nulls(rng, _strings(rng, n, 0, 64), 0.10)
        cols["ts"] = pa.array(rng.integers(1_600_000_000_000_000, 1_800_000_000_000_000, size=n)).cast(pa.timestamp("us", "UTC"))
        kinds = np.array(["insert", "update", "delete"], dtype=object)
        cols["rowKinds"] = pa.array(kinds[ops if ops is not None else np.zeros(n, dtype=np.int64)], type=pa.string())
    elif spec.kind == "wide":
        st = commit
        cols["pk"] = pa.array(keys["pk"])
        cols[f"s{st}_a"] = pa.array(rng.integers(-2**62, 2**62, size=n))
        cols[f"s{st}_b"] = pa.array(rng.standard_normal(n))
        cols[f"s{st}_c"] = pa.array(rng.integers(-2**31, 2**31 - 1, size=n).astype(np.int32))
        cols[f"s{st}_d"] = _strings(rng, n, 1, 16)
    names = [f.name for f in spec.schema if f.name in cols]
    return pa.table({k: cols[k] for k in names}).cast(pa.schema([spec.schema.field(k) for k in names]))


# ---- table generation -----------------------------------------------------------------------------
def generate_table(spec: TableSpec) -> List[List[pa.Table]]:
    """Returns tables[bucket][commit] (each sorted by primary key)."""
    B, K = spec.buckets, spec.commits
    total = spec.rows_per_bucket * B
    rng = np.random.Generator(np.random.PCG64(spec.seed))
    composite = spec.kind == "lineitem"
    n_base = int(total * spec.base_frac) if spec.kind != "wide" else 0
    per_delta = (total - n_base) // max(K - 1, 1) if spec.kind != "wide" else 0
    out: List[List[pa.Table]] = [[None] * K for _ in range(B)]
    universe = np.empty(0, dtype=np.int64)  # encoded keys present so far
    next_new = 1

    def encode(orderkey, line):
        return orderkey * 8 + line

    for j in range(K):
        if spec.kind == "wide":
            # 30000-key universe per bucket; every stream holds 5/6 of it, no in-file duplicates
            if j == 0:
                universe = np.arange(1, int(spec.rows_per_bucket * B / K * 1.2) + 1, dtype=np.int64)
            take = rng.random(len(universe)) < (5.0 / 6.0)
            enc = universe[take]
            ops = None
        else:
            n = n_base if j == 0 else (per_delta if j < K - 1 else total - n_base - per_delta * (K - 2))
            if j == 0:
                if composite:
                    orders = np.arange(1, n // 4 + 2, dtype=np.int64)
                    enc = (orders[:, None] * 8 + np.arange(1, 5, dtype=np.int64)[None, :]).ravel()[:n]
                    next_new = int(orders[-1]) + 1
                else:
                    enc = np.arange(1, n + 1, dtype=np.int64) * 3
                    next_new = n + 1
                ops = np.zeros(n, dtype=np.int64)
            else:
                n_upd = int(n * spec.update_frac)
                n_dup = int(n * spec.dup_frac)
                n_new = n - n_upd - n_dup
                upd = universe[rng.integers(0, len(universe), size=n_upd)]
                if composite:
                    new_orders = np.arange(next_new, next_new + n_new // 4 + 1, dtype=np.int64)
                    new = (new_orders[:, None] * 8 + np.arange(1, 5, dtype=np.int64)[None, :]).ravel()[:n_new]
                    next_new += len(new_orders)
                else:
                    new = np.arange(next_new, next_new + n_new, dtype=np.int64) * 3
                    next_new += n_new
                both = np.concatenate([upd, new])
                dup = both[rng.integers(0, len(both), size=n_dup)] if len(both) else both[:0]
                enc = np.concatenate([both, dup])
                ops = np.concatenate([rng.choice([1, 2], size=n_upd, p=[0.8, 0.2]), np.zeros(n_new, dtype=np.int64),
                                      rng.choice([1, 2], size=n_dup, p=[0.8, 0.2])])
            universe = np.union1d(universe, enc)
        if composite:
            keys = {"l_orderkey": enc // 8, "l_linenumber": (enc % 8).astype(np.int32)}
            bucket = hashing.bucket_of([keys["l_orderkey"], keys["l_linenumber"]], B)
        else:
            kname = spec.primary_keys[0]
            keys = {kname: enc}
            bucket = hashing.bucket_of([enc], B)
        order = np.lexsort((np.arange(len(enc)), enc, bucket))  # bucket, key, arrival order
        bsorted = bucket[order]
        bounds = np.searchsorted(bsorted, np.arange(B + 1))
        for b in range(B):
            sel = order[bounds[b]:bounds[b + 1]]
            brng = np.random.Generator(np.random.PCG64(spec.seed + 7919 * j + b))
            ksel = {k: v[sel] for k, v in keys.items()}
            out[b][j] = _payload(spec, brng, ksel, j, None if ops is None else ops[sel])
    return out


def file_name(bucket: int, commit: int, seed: int) -> str:
    r = np.random.Generator(np.random.PCG64(seed * 1000003 + bucket * 257 + commit))
    alnum = np.array(list(string.ascii_lowercase + string.digits))
    return "part-" + "".join(r.choice(alnum, size=16)) + f"_{bucket:04d}.parquet"


ENCODINGS = {
    "E1": dict(use_dictionary=False, compression="NONE"),  # PLAIN, uncompressed: isolates the merge
    "E2": dict(use_dictionary=True, compression="SNAPPY"),  # pyarrow defaults: RLE_DICTIONARY + bit-packed
    "E3": dict(use_dictionary=False, compression="ZSTD"),   # what the LakeSoul writer emits (multipart_writer.rs:121-133)
}


def write_table(root: str, spec: TableSpec, tables: Optional[List[List[pa.Table]]] = None, encoding: str = "E1",
                row_group_size: int = 250_000) -> List[List[str]]:
    """Writes tables[bucket][commit] and returns the file paths in the same shape."""
    tables = tables if tables is not None else generate_table(spec)
    os.makedirs(root, exist_ok=True)
    opts = ENCODINGS[encoding]
    paths: List[List[str]] = []
    for b, commits in enumerate(tables):
        row = []
        for j, t in enumerate(commits):
            p = os.path.join(root, file_name(b, j, spec.seed))
            pq.write_table(t, p, data_page_version="1.0", data_page_size=1 << 20, row_group_size=row_group_size,
                           write_statistics=True, **opts)
            row.append(p)
        paths.append(row)
    return paths


# ---- bench-scale tables: full-size buckets generated independently ----------------------------------
def bucket_file_paths(root: str, encoding: str, bucket: int, commits: int, seed: int) -> List[str]:
    """Deterministic paths of the delta files of one hash bucket (every rank computes the same list)."""
    return [os.path.join(root, encoding, file_name(bucket, j, seed)) for j in range(commits)]


def write_full_bucket(root: str, config: int, bucket: int, encodings: Sequence[str] = ("E3",)) -> int:
    """One hash bucket of BASELINE config `config` at its full per-bucket size, generated on its own
    (seed = the config's seed + bucket; every key lands in this bucket) and written once per encoding under
    <root>/<encoding>/. Returns the bucket's input rows. Independent per bucket, so a pool of processes can build a
    table of thousands of buckets in parallel (tools/make_table.py)."""
    spec = config_spec(config, 1.0)
    spec.buckets = 1
    base_seed = spec.seed
    spec.seed = base_seed + 104729 * (bucket + 1)
    tables = generate_table(spec)[0]
    rows = 0
    for enc in encodings:
        os.makedirs(os.path.join(root, enc), exist_ok=True)
        opts = ENCODINGS[enc]
        for t, p in zip(tables, bucket_file_paths(root, enc, bucket, spec.commits, base_seed)):
            tmp = p + ".tmp"
            pq.write_table(t, tmp, data_page_version="1.0", data_page_size=1 << 20, row_group_size=250_000,
                           write_statistics=True, **opts)
            os.replace(tmp, p)
    for t in tables:
        rows += t.num_rows
    return rows
