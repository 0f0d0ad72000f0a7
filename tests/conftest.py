# SPDX-License-Identifier: Apache-2.0
import json
import os
import sys

import pyarrow as pa
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

# Host-logic tests of the file framing (test_lazy_files, test_page_pruning*, test_stream_passes) open files through the
# library on machines without a GPU: there -- and only there -- the library may hand out pageable instead of pinned
# memory. On a GPU box nothing is set and every test runs on the production allocator (the library ignores the switch
# when a CUDA device exists).
if not os.path.exists("/dev/nvidia0") and not os.path.exists("/dev/nvidiactl"):
    os.environ.setdefault("LAKESOUL_B200_TEST_HOST_ALLOC", "1")

_TYPES = {"int32": pa.int32(), "int64": pa.int64(), "float64": pa.float64(), "utf8": pa.string()}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def golden_cases():
    with open(os.path.join(ROOT, "tests", "golden", "reference_unit_goldens.json")) as f:
        return json.load(f)["cases"]


def golden_schema(case) -> pa.Schema:
    return pa.schema([pa.field(n, _TYPES[t], nullable) for n, t, nullable in case["schema"]])


def golden_streams(case):
    """[[RecordBatch, ...] per stream] with the stream's own column subset (partial merge cases)."""
    schema = golden_schema(case)
    streams = []
    for s in case["streams"]:
        batches = []
        for b in s:
            fields = [schema.field(n) for n in b.keys()]
            batches.append(pa.RecordBatch.from_arrays([pa.array(v, type=f.type) for v, f in zip(b.values(), fields)],
                                                      schema=pa.schema(fields)))
        streams.append(batches)
    return streams


def golden_expected(case) -> pa.Table:
    schema = golden_schema(case)
    return pa.table({n: pa.array(case["expected"][n], type=schema.field(n).type) for n in schema.names}).cast(schema)


def assert_tables_equal(got: pa.Table, want: pa.Table, what=""):
    assert got.schema.names == want.schema.names, f"{what}: columns {got.schema.names} != {want.schema.names}"
    assert got.num_rows == want.num_rows, f"{what}: rows {got.num_rows} != {want.num_rows}"
    for name in want.schema.names:
        g = got.column(name).combine_chunks()
        w = want.column(name).combine_chunks().cast(g.type)
        if not g.equals(w):
            import pyarrow.compute as pc
            bad = [i for i in range(len(g)) if g[i].as_py() != w[i].as_py()][:5]
            raise AssertionError(f"{what}: column {name} differs at rows {bad}: "
                                 f"got {[g[i].as_py() for i in bad]} want {[w[i].as_py() for i in bad]}")


@pytest.fixture(scope="session")
def lib():
    from lakesoul_b200 import load_library
    return load_library()


def upsert_cases():
    with open(os.path.join(ROOT, "tests", "golden", "reference_upsert_goldens.json")) as f:
        return json.load(f)["cases"]


def upsert_case_files(case, tmp_path):
    import pyarrow.parquet as pq
    paths = []
    for j, cols in enumerate(case["files"]):
        t = pa.table({k: pa.array(v, type=pa.int32()) for k, v in cols.items()})
        p = str(tmp_path / f"part-{j:016d}_0000.parquet")
        pq.write_table(t, p)
        paths.append(p)
    schema = pa.schema([pa.field(n, _TYPES[t], nl) for n, t, nl in case["schema"]])
    expected = pa.table({n: pa.array(case["expected"][n], type=schema.field(n).type) for n in schema.names})
    return paths, schema, expected
