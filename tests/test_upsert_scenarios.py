# SPDX-License-Identifier: Apache-2.0
"""All 22 end-to-end upsert scenarios of the reference (rust/lakesoul-datafusion/src/tests/upsert_tests.rs:241-1961,
G9 of SURVEY.md 8c), transcribed by tests/golden/make_upsert_goldens.py: every upsert becomes one delta file per
range partition (commit order = upsert order; the range columns live in the partition description, not in the files,
exactly as LakeSoul stores them), every partition is read by one reader with the reference's filter string and column
selection, and the union over the partitions must equal the table the reference's test prints.
(Hash buckets of a partition are merged by one reader here: keys never cross buckets, so the union is the same.)

CPU suite: the oracle against the reference's expected tables (pins the oracle).  GPU suite: the CUDA reader, through
the C-ABI, against the same tables."""
import datetime
import json
import os

import pyarrow as pa
import pyarrow.parquet as pq
import pytest

import mor_oracle as mo
from conftest import ROOT

_T = {"int32": pa.int32(), "utf8": pa.string(), "timestamp_us": pa.timestamp("us")}


def _cases():
    out = []
    for name in ("reference_upsert_scenarios.json", "reference_cdc_goldens.json"):  # G9, G10 (CDC deletes)
        with open(os.path.join(ROOT, "tests", "golden", name)) as f:
            out += json.load(f)["cases"]
    return out


CASES = _cases()
CHECKS = [(c, i) for c in CASES for i in range(len(c["checks"]))]
IDS = [f"{c['name'].removeprefix('test_')}#{i}" for c, i in CHECKS]


def _partition_files(case, n_upserts, tmp_path):
    """{partition values (tuple of str)} -> file paths in commit order"""
    types = dict(case["schema"])
    ranges = case["range_partitions"]
    parts = {}
    for j, up in enumerate(case["upserts"][:n_upserts]):
        n = len(next(iter(up.values()))) if up else 0
        if n == 0:
            continue  # an empty upsert commits nothing
        keys = [tuple(str(up[r][i]) for r in ranges) for i in range(n)]
        for key in dict.fromkeys(keys):
            rows = [i for i in range(n) if keys[i] == key]
            # the LakeSoul writer sorts every file by primary key (writer/async_writer/sort_writer.rs:59-76; stable,
            # so duplicates inside one upsert keep their arrival order): the reader's contract is sorted input
            rows.sort(key=lambda i: tuple(_sort_cell(up[p][i]) for p in case["primary_keys"]))
            cols = {c: pa.array([v[i] for i in rows], type=_T[types[c]]) for c, v in up.items() if c not in ranges}
            d = tmp_path / ("p_" + "_".join(key) if key else "p")
            d.mkdir(exist_ok=True)
            path = str(d / f"part-{j:016d}_0000.parquet")
            pq.write_table(pa.table(cols), path)
            parts.setdefault(key, []).append(path)
    return parts


def _sort_cell(v):
    return (0, "") if v is None else ((1, v.encode()) if isinstance(v, str) else (1, v))


def _expected(case, check):
    types = dict(case["schema"])

    def conv(col, cell):
        if cell == "":
            return None
        t = types[col]
        if t == "int32":
            return int(cell)
        if t == "timestamp_us":
            dt = datetime.datetime.fromisoformat(cell).replace(tzinfo=datetime.timezone.utc)
            delta = dt - datetime.datetime(1970, 1, 1, tzinfo=datetime.timezone.utc)
            return delta.days * 86_400_000_000 + delta.seconds * 1_000_000 + delta.microseconds
        return cell
    cols = check["expected_columns"]
    return cols, sorted((tuple(conv(c, r[c]) for c in cols) for r in check["expected_rows"]), key=_key)


def _key(row):
    return tuple((v is None, v if v is not None else 0) if not isinstance(v, str) else (False, v) for v in row)


def _rows(table: pa.Table, cols):
    out = []
    arrays = []
    for c in cols:
        a = table.column(c)
        if pa.types.is_timestamp(a.type):
            a = a.cast(pa.int64())
        arrays.append(a.to_pylist())
    for i in range(table.num_rows):
        out.append(tuple(a[i] for a in arrays))
    return sorted(out, key=_key)


def _read_all(read_fn, case, check, tmp_path):
    types = dict(case["schema"])
    ranges = case["range_partitions"]
    selected = check["selected"] or [c for c, _ in case["schema"]]
    schema = pa.schema([pa.field(c, _T[types[c]], True) for c in selected])
    part_schema = pa.schema([pa.field(r, _T[types[r]], True) for r in ranges])
    filters = [check["filter"]] if check["filter"] else []
    rows = []
    for key, files in _partition_files(case, check["after_upserts"], tmp_path).items():
        defaults = dict(zip(ranges, key))
        t = read_fn(files, schema, case["primary_keys"], default_columns=defaults, filters=filters,
                    partition_schema=part_schema if ranges else None, options=case.get("options"))
        rows += _rows(t, selected)
    return selected, sorted(rows, key=_key)


@pytest.mark.parametrize("case,ci", CHECKS, ids=IDS)
def test_oracle_matches_reference_upsert_scenario(case, ci, tmp_path):
    check = case["checks"][ci]
    cols, want = _expected(case, check)
    selected, got = _read_all(mo.read_bucket, case, check, tmp_path)
    assert selected == cols
    assert got == want, f"{case['source']}"


@pytest.mark.gpu
@pytest.mark.parametrize("case,ci", CHECKS, ids=IDS)
def test_reader_matches_reference_upsert_scenario(case, ci, tmp_path):
    from lakesoul_b200 import read_bucket
    check = case["checks"][ci]
    cols, want = _expected(case, check)

    def read_fn(files, schema, pks, **kw):
        return read_bucket(files, schema, pks, batch_size=3, **kw)
    selected, got = _read_all(read_fn, case, check, tmp_path)
    assert selected == cols
    assert got == want, f"{case['source']}"
