# SPDX-License-Identifier: Apache-2.0
"""CPU oracle for the merge-on-read hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the
product (lakesoul_b200/) never does.  Two independent restatements of the reference algorithm:

* ``StreamingMerger`` -- a line-of-reasoning restatement of the reference's streaming machinery in
  plain Python (small inputs only): the cursor/range model of
  rust/lakesoul-io/src/physical_plan/merge/sorted/sort_key_range.rs:20-113,182-187,
  the driver loop of sorted_stream_merger.rs:390-524, the loser-tree combiner
  combiner.rs:406-525,680-795 (full and partial merge, sort_key_range.rs:344-438), the min-heap
  combiner combiner.rs:90-336 and every MergeOperator of merge_operator.rs:55-579 including its
  quirks (JoinedAll* never emits the delimiter, :402-406; Sum is null when any row is null, :501-509;
  SumLast/JoinedLast skip a range that ends at a batch boundary when the same stream continues,
  :545-563, :445-460).
* ``merge_tables`` -- a vectorised "concat -> stable sort by (pk, file, row) -> pick per key"
  restatement (numpy/pyarrow), used at sizes where Python loops are too slow.

Parity pin: both are checked against the golden tables transcribed from the reference's own unit
tests (tests/golden/*.json: G1 sorted_stream_merger.rs:601-659, G3 :853-1106, G4
v2/window_sliding_merger.rs:467-546, ...) in tests/test_oracle_golden.py, and against each other on
random inputs.  Parquet *decode* parity is unpinned in the reference (reader.rs:421-441 asserts only
a row count): pyarrow/Arrow C++ reading the identical file is the decode oracle.
"""
from __future__ import annotations

import heapq
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import pyarrow as pa
import pyarrow.compute as pc

MERGE_OPS = ("UseLast", "UseLastNotNull", "SumAll", "SumLast", "JoinedLastByComma", "JoinedLastBySemicolon",
             "JoinedAllByComma", "JoinedAllBySemicolon")

LAKESOUL_NULL_STRING = "__L@KE$OUL_NULL__"          # constant.rs:18
LAKESOUL_EMPTY_STRING = "__L@KE$OUL_EMPTY_STRING__"  # constant.rs:19


# ------------------------------------------------------------------------------------------------
# streaming restatement (small inputs)
# ------------------------------------------------------------------------------------------------
def _key_lt(a, b) -> bool:
    """Lexicographic compare of key tuples, nulls first, null == null (cursor.rs:444-476)."""
    for x, y in zip(a, b):
        if x is None and y is None:
            continue
        if x is None:
            return True
        if y is None:
            return False
        if x != y:
            return x < y
    return False


def _key_eq(a, b) -> bool:
    return all((x is None and y is None) or (x is not None and y is not None and x == y) for x, y in zip(a, b))


def _as_cmp(v):
    if isinstance(v, str):
        return v.encode("utf-8")  # byte-wise, unsigned (cursor.rs:314-316)
    return v


@dataclass
class _Batch:
    cols: Dict[str, list]      # column name -> python values (None = null)
    names: List[str]
    n: int
    keys: List[tuple]


@dataclass
class _Range:
    """SortKeyBatchRange: run of equal keys inside one batch (sort_key_range.rs:20-33)."""
    begin: int
    end: int
    stream: int
    batch_idx: int
    batch: _Batch

    def finished(self) -> bool:
        return self.begin >= self.batch.n

    def advance(self) -> "_Range":
        cur = _Range(self.begin, self.end, self.stream, self.batch_idx, self.batch)
        self.begin = self.end
        if not self.finished():
            while self.end < self.batch.n and _key_eq(self.batch.keys[self.end], self.batch.keys[self.begin]):
                self.end += 1
        return cur

    def key(self):
        return self.batch.keys[self.begin]

    def lt(self, other: "_Range") -> bool:  # Ord: (key, stream_idx) sort_key_range.rs:182-187
        if _key_lt(self.key(), other.key()):
            return True
        if _key_lt(other.key(), self.key()):
            return False
        return self.stream < other.stream


def _sum_native(dtype: pa.DataType):
    if pa.types.is_floating(dtype):
        np_t = np.float32 if dtype == pa.float32() else np.float64
    else:
        np_t = dtype.to_pandas_dtype()
    return np_t


def _apply_operator(op: str, dtype: pa.DataType, ranges: List[Tuple[_Range, list]]):
    """MergeOperator::merge for one output cell (merge_operator.rs:55-167).

    ``ranges`` = [(range, column_values_of_that_batch)] in pop order. Returns a python value."""
    if not ranges:
        return None
    last_r, last_v = ranges[-1]
    if op == "UseLast":
        return last_v[last_r.end - 1]
    if op == "UseLastNotNull":  # :170-195
        for r, vals in reversed(ranges):
            for row in range(r.end - 1, r.begin - 1, -1):
                if vals[row] is not None:
                    return vals[row]
        return None
    single_row = len(ranges) == 1 and ranges[0][0].end - ranges[0][0].begin == 1
    if single_row:  # Extend(batch, row): copies the cell, null included (:63-123)
        return last_v[last_r.end - 1]
    if op in ("SumAll", "SumLast"):
        np_t = _sum_native(dtype)
        with np.errstate(over="ignore"):
            res = np_t(0)
            if op == "SumAll":  # :490-532
                for r, vals in ranges:
                    seg = vals[r.begin:r.end]
                    if any(v is None for v in seg):
                        return None
                    part = np_t(0)
                    for v in seg:
                        part = np_t(part + np_t(v))
                    res = np_t(res + part)
            else:  # :534-579
                for i, (r, vals) in enumerate(ranges):
                    if r.end == r.batch.n:
                        if i == len(ranges) - 1 or ranges[i + 1][0].stream != r.stream:
                            if vals[r.end - 1] is None:
                                return None
                            res = np_t(res + np_t(vals[r.end - 1]))
                    else:
                        if vals[r.end - 1] is None:
                            return None
                        res = np_t(res + np_t(vals[r.end - 1]))
        return res.item()
    delim = "," if op.endswith("Comma") else ";"
    if op.startswith("JoinedAll"):  # :389-431 -- `first` is never cleared: no delimiter at all
        out = ""
        for r, vals in ranges:
            for row in range(r.begin, r.end):
                if vals[row] is None:
                    return None
                out += vals[row]
        return out
    if op.startswith("JoinedLast"):  # :433-487
        out, first = "", True
        for i, (r, vals) in enumerate(ranges):
            take = True
            if r.end == r.batch.n:
                take = i == len(ranges) - 1 or ranges[i + 1][0].stream != r.stream
            if not take:
                continue
            if vals[r.end - 1] is None:
                return None
            if not first:
                out += delim
            first = False
            out += vals[r.end - 1]
        return out
    raise ValueError("Invalid MergeOperator name")  # merge_operator.rs:51


class StreamingMerger:
    """build_sorted_stream_merger + SortedStreamMerger (sorted_stream_merger.rs:280-524)."""

    def __init__(self, streams: Sequence[Sequence[pa.RecordBatch]], primary_keys: Sequence[str], schema: pa.Schema,
                 batch_size: int, merge_ops: Optional[Sequence[str]] = None):
        self.schema = schema
        self.pks = list(primary_keys)
        self.batch_size = batch_size
        self.ops = list(merge_ops) if merge_ops else []
        self.streams = [[self._convert(b) for b in s] for s in streams]
        self.stream_names = [list(s[0].schema.names) if len(s) else [] for s in streams]
        self.k = len(streams)
        all_use_last = not self.ops or all(o == "UseLast" for o in self.ops)
        self.partial = any(len(n) != len(schema.names) for n in self.stream_names)  # :176-188
        self.mode = "loser" if all_use_last else "heap"

    def _convert(self, b: pa.RecordBatch) -> _Batch:
        cols = {n: b.column(i).to_pylist() for i, n in enumerate(b.schema.names)}
        keys = [tuple(_as_cmp(cols[p][r]) for p in self.pks) for r in range(b.num_rows)]
        return _Batch(cols, list(b.schema.names), b.num_rows, keys)

    # -- driver ------------------------------------------------------------------------------
    def run(self) -> List[pa.RecordBatch]:
        pos = [0] * self.k
        batch_counter = 0
        cur: List[Optional[_Range]] = [None] * self.k

        def poll(i) -> Optional[_Range]:  # maybe_poll_stream: skips empty batches (:399-448)
            nonlocal batch_counter
            while pos[i] < len(self.streams[i]):
                b = self.streams[i][pos[i]]
                pos[i] += 1
                if b.n > 0:
                    batch_counter += 1
                    r = _Range(0, 0, i, batch_counter, b)
                    r.advance()
                    return r
            return None

        for i in range(self.k):
            cur[i] = poll(i)

        out_batches: List[pa.RecordBatch] = []
        rows: List[dict] = []  # finished output rows: column -> value

        # per-key accumulation
        pending_key = None
        pending_ranges: List[_Range] = []

        def flush_key():
            nonlocal pending_key, pending_ranges
            if pending_key is None:
                return
            rows.append(self._build_row(pending_ranges))
            pending_key, pending_ranges = None, []
            if len(rows) == self.batch_size:
                out_batches.append(self._to_batch(rows))
                rows.clear()

        while True:
            # winner = smallest (key, stream); exhausted streams always lose (combiner.rs:695-699)
            w = None
            for i in range(self.k):
                if cur[i] is not None and (w is None or cur[i].lt(cur[w])):
                    w = i
            if w is None:
                break
            r = cur[w]
            if pending_key is not None and not _key_eq(pending_key, r.key()):
                flush_key()
            pending_key = r.key()
            pending_ranges.append(_Range(r.begin, r.end, r.stream, r.batch_idx, r.batch))
            r.advance()
            if r.finished():
                cur[w] = poll(w)
        flush_key()
        if rows:
            out_batches.append(self._to_batch(rows))
        return out_batches

    def _build_row(self, ranges: List[_Range]) -> dict:
        row = {}
        for ci, f in enumerate(self.schema):
            carrying = [(r, r.batch.cols[f.name]) for r in ranges if f.name in r.batch.cols]
            if self.mode == "loser":
                # full merge: one pointer for all columns; partial: per column (sort_key_range.rs:387-412)
                if not self.partial:
                    r = ranges[-1]
                    row[f.name] = r.batch.cols[f.name][r.end - 1]
                else:
                    row[f.name] = carrying[-1][1][carrying[-1][0].end - 1] if carrying else None
            else:
                row[f.name] = _apply_operator(self.ops[ci], f.type, carrying)
        return row

    def _to_batch(self, rows: List[dict]) -> pa.RecordBatch:
        arrays = [pa.array([r[f.name] for r in rows], type=f.type) for f in self.schema]
        return pa.RecordBatch.from_arrays(arrays, schema=self.schema)


def streaming_merge(streams, primary_keys, schema, batch_size=8192, merge_ops=None) -> List[pa.RecordBatch]:
    return StreamingMerger(streams, primary_keys, schema, batch_size, merge_ops).run()


# ------------------------------------------------------------------------------------------------
# vectorised restatement (any size)
# ------------------------------------------------------------------------------------------------
def _group_boundaries(sorted_tbl: pa.Table, pks: Sequence[str]) -> np.ndarray:
    n = sorted_tbl.num_rows
    new = np.zeros(n, dtype=bool)
    if n:
        new[0] = True
    for p in pks:
        col = sorted_tbl.column(p).combine_chunks()
        valid = np.asarray(col.is_valid())
        if n > 1:
            a, b = col.slice(1), col.slice(0, n - 1)
            neq = np.asarray(pc.fill_null(pc.not_equal(a, b), False))
            new[1:] |= (valid[1:] != valid[:-1]) | (valid[1:] & valid[:-1] & neq)
    return new


def merge_tables(files: Sequence[pa.Table], primary_keys: Sequence[str], schema: pa.Schema,
                 merge_ops: Optional[Dict[str, str]] = None, decode_batch_rows: Optional[int] = None) -> pa.Table:
    """Merge the delta files of one bucket (commit order). `schema` = physical schema in target order.
    Files may lack columns (partial merge). Returns the merged table ordered by primary key."""
    merge_ops = merge_ops or {}
    k = len(files)
    names = schema.names
    parts = []
    for j, t in enumerate(files):
        cols = {}
        for f in schema:
            if f.name in t.schema.names:
                cols[f.name] = t.column(f.name).cast(f.type)
            else:
                cols[f.name] = pa.nulls(t.num_rows, f.type)
        cols["__file"] = pa.array(np.full(t.num_rows, j, dtype=np.int32))
        cols["__row"] = pa.array(np.arange(t.num_rows, dtype=np.int64))
        parts.append(pa.table(cols))
    if not parts:
        return schema.empty_table()
    allt = pa.concat_tables(parts).combine_chunks()
    n = allt.num_rows
    if n == 0:
        return schema.empty_table()
    # floating-point keys compare by IEEE total order and are equal only bit for bit (cursor.rs:265-276:
    # ArrowNativeTypeOp::compare = total_cmp, is_eq = to_bits equality): order and group them by the integer
    # image total_cmp itself uses (bits ^ ((bits >> 63) as unsigned >> 1), compared as signed integers)
    key_names = []
    for p in primary_keys:
        t = allt.schema.field(p).type
        if pa.types.is_floating(t):
            col = allt.column(p).combine_chunks()
            it, ut, sh = (np.int64, np.uint64, 63) if t == pa.float64() else (np.int32, np.uint32, 31)
            bits = np.asarray(col.fill_null(0.0)).view(it)
            img = bits ^ ((bits >> sh).view(ut) >> ut(1)).view(it)
            allt = allt.append_column("__k_" + p, pa.array(img, mask=~np.asarray(col.is_valid())))
            key_names.append("__k_" + p)
        else:
            key_names.append(p)
    keys = [(p, "ascending") for p in key_names] + [("__file", "ascending"), ("__row", "ascending")]
    order = pc.sort_indices(allt, sort_keys=keys, null_placement="at_start")
    st = allt.take(order)
    new = _group_boundaries(st, key_names)
    starts = np.flatnonzero(new)
    ends = np.append(starts[1:], n) - 1
    gid = np.cumsum(new) - 1
    file_idx = np.asarray(st.column("__file").combine_chunks())
    has = {f.name: np.array([f.name in t.schema.names for t in files], dtype=bool) for f in schema}
    idx = np.arange(n, dtype=np.int64)
    # (group, file) subgroups
    sub_new = new.copy()
    sub_new[1:] |= file_idx[1:] != file_idx[:-1]
    sub_starts = np.flatnonzero(sub_new)
    sub_ends = np.append(sub_starts[1:], n) - 1
    out_cols = []
    for f in schema:
        op = merge_ops.get(f.name, "UseLast")
        col = st.column(f.name).combine_chunks()
        carry = has[f.name][file_idx]
        valid = np.asarray(col.is_valid())
        if op in ("UseLast", "UseLastNotNull"):
            mask = carry if op == "UseLast" else (carry & valid)
            cand = np.where(mask, idx, -1)
            win = np.maximum.reduceat(cand, starts)
            taken = col.take(pa.array(np.where(win < 0, 0, win)))
            nullmask = win < 0
            if nullmask.any():
                taken = pc.if_else(pa.array(~nullmask), taken, pa.nulls(len(taken), f.type))
            out_cols.append(taken)
            continue
        if op in ("SumAll", "SumLast"):
            np_t = _sum_native(f.type)
            vals = np.asarray(pc.fill_null(col, 0)).astype(np_t, copy=False)
            with np.errstate(over="ignore"):
                if op == "SumAll":
                    # per (key, file[, decode batch]) partial sums, then summed per key: the
                    # association of merge_operator.rs:512-514
                    bounds = sub_new.copy()
                    if decode_batch_rows:
                        rows_in_file = np.asarray(st.column("__row").combine_chunks())
                        bounds[1:] |= (rows_in_file[1:] // decode_batch_rows) != (rows_in_file[:-1] // decode_batch_rows)
                    bstarts = np.flatnonzero(bounds)
                    contrib = np.where(carry, vals, np_t(0))
                    part = np.add.reduceat(contrib, bstarts, dtype=np_t) if np_t not in (np.float32, np.float64) else \
                        np.array([_seq_sum(contrib[a:b], np_t) for a, b in zip(bstarts, np.append(bstarts[1:], n))], dtype=np_t)
                    part_gid = gid[bstarts]
                    gstarts = np.flatnonzero(np.append(True, part_gid[1:] != part_gid[:-1]))
                    total = np.array([_seq_sum(part[a:b], np_t) for a, b in zip(gstarts, np.append(gstarts[1:], len(part)))], dtype=np_t) \
                        if np_t in (np.float32, np.float64) else np.add.reduceat(part, gstarts, dtype=np_t)
                    anynull = np.maximum.reduceat((carry & ~valid).astype(np.int8), starts) > 0
                else:
                    lastrow = sub_ends
                    c2 = carry[lastrow]
                    contrib = np.where(c2, vals[lastrow], np_t(0))
                    sg = gid[lastrow]
                    gstarts = np.flatnonzero(np.append(True, sg[1:] != sg[:-1]))
                    total = np.array([_seq_sum(contrib[a:b], np_t) for a, b in zip(gstarts, np.append(gstarts[1:], len(contrib)))], dtype=np_t) \
                        if np_t in (np.float32, np.float64) else np.add.reduceat(contrib, gstarts, dtype=np_t)
                    anynull = np.maximum.reduceat((c2 & ~valid[lastrow]).astype(np.int8), gstarts) > 0
            none = np.maximum.reduceat(carry.astype(np.int8), starts) == 0
            # a key with exactly one carrying row copies the cell (Extend)
            cnt = np.add.reduceat(carry.astype(np.int64), starts)
            single = cnt == 1
            cand = np.where(carry, idx, -1)
            lastc = np.maximum.reduceat(cand, starts)
            total = np.where(single, vals[np.where(lastc < 0, 0, lastc)], total)
            arr = pa.array(total, type=f.type, mask=(anynull | none))
            out_cols.append(arr)
            continue
        # string joins: python loop over groups (small inputs)
        pyvals = col.to_pylist()
        rows_in_file = np.asarray(st.column("__row").combine_chunks())
        res = []
        delim = "," if op.endswith("Comma") else ";"
        for a, b in zip(starts, ends + 1):
            sel = [i for i in range(a, b) if carry[i]]
            if not sel:
                res.append(None)
                continue
            if len(sel) == 1:
                res.append(pyvals[sel[0]])
                continue
            if op.startswith("JoinedAll"):
                vs = [pyvals[i] for i in sel]
                res.append(None if any(v is None for v in vs) else "".join(vs))
            else:
                lasts = [i for i in sel if i == b - 1 or file_idx[i + 1] != file_idx[i] or not carry[i + 1]]
                vs = [pyvals[i] for i in lasts]
                res.append(None if any(v is None for v in vs) else delim.join(vs))
        out_cols.append(pa.array(res, type=f.type))
    return pa.Table.from_arrays([c if isinstance(c, (pa.Array, pa.ChunkedArray)) else pa.array(c) for c in out_cols],
                                schema=schema)


def _seq_sum(a: np.ndarray, np_t):
    s = np_t(0)
    for v in a:
        s = np_t(s + v)
    return s


# ------------------------------------------------------------------------------------------------
# reader-level restatement: decode -> merge -> default columns -> filter -> projection
# ------------------------------------------------------------------------------------------------
def make_default_array(dtype: pa.DataType, value: str, rows: int) -> pa.Array:
    """helpers/transform.rs:356-513."""
    if value == LAKESOUL_NULL_STRING:
        return pa.nulls(rows, dtype)
    if pa.types.is_string(dtype):
        v = "" if value == LAKESOUL_EMPTY_STRING else value
        return pa.array([v] * rows, type=dtype)
    if pa.types.is_int32(dtype) or pa.types.is_int64(dtype):
        return pa.array([int(value)] * rows, type=dtype)
    if pa.types.is_date32(dtype):
        try:
            return pa.array([int(value)] * rows, type=pa.int32()).cast(dtype)
        except ValueError:
            return pa.array([value] * rows).cast(dtype)
    if pa.types.is_timestamp(dtype):
        try:
            return pa.array([int(value)] * rows, type=pa.int64()).cast(dtype)
        except ValueError:
            return pa.array([value] * rows).cast(pa.timestamp(dtype.unit)).cast(dtype)
    if pa.types.is_boolean(dtype):
        return pa.array([{"true": True, "false": False}[value]] * rows, type=dtype)
    try:
        return pa.array([value] * rows).cast(dtype)
    except (pa.ArrowInvalid, pa.ArrowNotImplementedError):
        return pa.nulls(rows, dtype)


def _split_call(s: str):
    lp = s.index("(")
    op, inner = s[:lp], s[lp + 1:-1]
    depth, comma = 0, 0
    for i, ch in enumerate(inner):
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        elif ch == "," and depth == 0 and comma == 0:
            comma = i
    if op == "not":
        return op, inner[:comma].strip(), inner[comma:].strip()
    return op, inner[:comma].strip(), inner[comma + 1:].strip()


def parse_filter(s: str, schema: pa.Schema) -> pc.Expression:
    """filter/parser.rs:50-265 restated on pyarrow expressions."""
    op, left, right = _split_call(s)
    if op == "or":
        return parse_filter(left, schema) | parse_filter(right, schema)
    if op == "and":
        return parse_filter(left, schema) & parse_filter(right, schema)
    if op == "not":
        return ~parse_filter(right, schema)
    if left not in schema.names:
        return pc.scalar(False)
    f = schema.field(left)
    col = pc.field(left)
    if right == "null":
        return {"eq": col.is_null(), "noteq": ~col.is_null()}.get(op, pc.scalar(True))
    t = f.type
    if pa.types.is_string(t):
        lit = pa.scalar(right[8:-2], type=t)
    elif pa.types.is_binary(t):
        body = right[right.index("[") + 1:right.index("]")].replace(" ", "")
        lit = pa.scalar(bytes((int(x) + 256) % 256 for x in body.split(",")) if body else b"", type=t)
    elif pa.types.is_boolean(t):
        lit = pa.scalar(right == "true", type=t)
    elif pa.types.is_floating(t):
        lit = pa.scalar(float(right), type=t)
    elif pa.types.is_decimal(t):
        import decimal
        lit = pa.scalar(decimal.Decimal(int(right)).scaleb(-t.scale), type=t)
    elif pa.types.is_timestamp(t):
        lit = pa.scalar(int(right), type=pa.int64()).cast(pa.timestamp("us", "UTC"))
        col = col.cast(pa.timestamp("us", "UTC")) if t.unit != "ns" else col
        if t.unit == "ns":
            lit = lit.cast(t)
    elif pa.types.is_date32(t):
        lit = pa.scalar(int(right), type=pa.int32()).cast(t)
    else:
        lit = pa.scalar(int(right), type=t)
    return {"eq": col == lit, "noteq": col != lit, "gt": col > lit, "gteq": col >= lit, "lt": col < lit,
            "lteq": col <= lit}.get(op, pc.scalar(True))


def read_bucket(files: Sequence[str], schema: pa.Schema, primary_keys: Sequence[str] = (), *,
                default_columns: Optional[Dict[str, str]] = None, merge_ops: Optional[Dict[str, str]] = None,
                options: Optional[Dict[str, str]] = None, filters: Sequence[str] = (),
                partition_schema: Optional[pa.Schema] = None) -> pa.Table:
    """What the reader returns for one hash bucket (reference reader.rs:138-232 + file_format.rs:234-310
    + merge/mod.rs:283-374 + stream/default_column.rs:62-108 + session.rs:839-877)."""
    import os
    import pyarrow.parquet as pq
    default_columns = default_columns or {}
    options = options or {}
    merge_ops = merge_ops or {}
    pks = list(primary_keys)
    part_names = set(partition_schema.names) if partition_schema is not None else set()
    files = [f for f in files if os.path.getsize(f) >= 8]
    tables = []
    for f in files:
        t = pq.read_table(f)
        t = t.drop_columns([n for n in t.schema.names if n in part_names])
        tables.append(t)
    # table schema: target first, then file-only fields
    fields = list(schema)
    for t in tables:
        for fld in t.schema:
            if fld.name not in [x.name for x in fields]:
                fields.append(fld)
    table_schema = pa.schema(fields)
    cdc = options.get("cdc_column", "")
    exprs = [parse_filter(s, table_schema) for s in filters]
    filter_cols = set()
    for s in filters:
        filter_cols |= {n for n in table_schema.names if n in s}
    needed = [f for f in fields if f.name in schema.names or f.name in pks or f.name == cdc or f.name in filter_cols]
    in_any = lambda n: any(n in t.schema.names for t in tables)  # noqa: E731
    phys = pa.schema([f for f in needed if f.name not in default_columns and in_any(f.name)])
    skip = options.get("skip_merge_on_read") == "true" or not pks
    merge = not skip and not (len(files) == 1 and not merge_ops and options.get("is_compacted") == "true")
    if merge:
        merged = merge_tables(tables, pks, phys, merge_ops)
    else:
        parts = []
        for t in tables:
            cols = [t.column(f.name).cast(f.type) if f.name in t.schema.names else pa.nulls(t.num_rows, f.type)
                    for f in phys]
            parts.append(pa.Table.from_arrays(cols, schema=phys))
        merged = pa.concat_tables(parts) if parts else phys.empty_table()
    n = merged.num_rows
    cols, out_fields = [], []
    for f in needed:
        if f.name in default_columns:
            cols.append(make_default_array(f.type, default_columns[f.name], n))
        elif f.name in phys.names:
            cols.append(merged.column(f.name))
        else:
            cols.append(pa.nulls(n, f.type))
        out_fields.append(f)
    full = pa.Table.from_arrays(cols, names=[f.name for f in out_fields])
    for e in exprs:
        full = full.filter(e)  # keeps rows where the predicate is TRUE (null -> dropped)
    return full.select(schema.names) if schema.names else pa.table({"__rows": pa.nulls(full.num_rows)}).select([])


# ------------------------------------------------------------------------------------------------
# search bounds of one sorted run (v2/batch_range.rs:120-190; pinned by G8): the two questions the tile partition
# asks of every delta file -- how many rows are < / <= a splitter key (mor_engine.cu part_kernel)
# ------------------------------------------------------------------------------------------------
def find_ge_start_index(values: Sequence, target, begin: int = 0) -> Optional[int]:
    """first index >= begin whose value is >= target (BatchRange::find_ge_start_index), None when there is none"""
    lo, hi = begin, len(values)
    while lo < hi:
        mid = (lo + hi) // 2
        if values[mid] < target:
            lo = mid + 1
        else:
            hi = mid
    return lo if lo < len(values) else None


def find_le_start_index(values: Sequence, target, begin: int = 0, end_row_for_merge: Optional[int] = None) -> Optional[int]:
    """last index in [begin, end_row_for_merge] whose value is <= target (BatchRange::find_le_start_index)"""
    end = len(values) - 1 if end_row_for_merge is None else end_row_for_merge
    lo, hi = begin, end + 1
    while lo < hi:
        mid = (lo + hi) // 2
        if values[mid] <= target:
            lo = mid + 1
        else:
            hi = mid
    return lo - 1 if lo - 1 >= begin else None
