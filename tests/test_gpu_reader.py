# SPDX-License-Identifier: Apache-2.0
"""GPU parity through the C-ABI: reference golden vectors and synthetic configs vs the CPU oracle."""
import os

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq
import pytest

import mor_oracle as mo
from conftest import assert_tables_equal, golden_cases, golden_expected, golden_schema, golden_streams

pytestmark = pytest.mark.gpu
CASES = golden_cases()


def _write_streams(tmp_path, case):
    paths = []
    for j, s in enumerate(golden_streams(case)):
        t = pa.Table.from_batches(s)
        p = str(tmp_path / f"part-{j:016d}_0000.parquet")
        pq.write_table(t, p, use_dictionary=False, compression="NONE")
        paths.append(p)
    return paths


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_reader_matches_reference_golden(case, tmp_path):
    from lakesoul_b200 import read_bucket
    schema = golden_schema(case)
    files = _write_streams(tmp_path, case)
    ops = dict(zip(schema.names, case["merge_ops"])) if case["merge_ops"] else None
    got = read_bucket(files, schema, case["primary_keys"], batch_size=case["batch_size"], merge_ops=ops)
    assert_tables_equal(got, golden_expected(case), case["name"])


def _roundtrip(spec, tmp_path, encoding="E1", batch_size=1000, buckets=None, nbuckets=4, **kw):
    from lakesoul_b200 import datagen, read_bucket
    spec.buckets = min(spec.buckets, nbuckets)  # keep the number of Parquet files written small
    tables = datagen.generate_table(spec)
    paths = datagen.write_table(str(tmp_path), spec, tables, encoding=encoding)
    for b in (buckets if buckets is not None else range(spec.buckets)):
        got, reader = read_bucket(paths[b], spec.schema, spec.primary_keys, batch_size=batch_size,
                                  options=spec.options, return_reader=True, **kw)
        stats = reader.stats()
        reader.close()
        want = mo.read_bucket(paths[b], spec.schema, spec.primary_keys, options=spec.options, **kw)
        assert_tables_equal(got, want, f"{spec.name} bucket {b}")
        assert stats["kernel_launches"] > 0, "no CUDA kernel ran: a CPU/library fallback is not acceptable"
        assert stats["rows_out"] == want.num_rows
        if not os.environ.get("LAKESOUL_B200_HOST_DECODE"):
            assert stats["device_decoded_pages"] > 0 and stats["host_decoded_files"] == 0, stats


def test_config1_int64_pk(tmp_path):
    from lakesoul_b200 import datagen
    _roundtrip(datagen.config_spec(1, 0.05), tmp_path, batch_size=8192)


def test_config1_full_size_multilevel(tmp_path):
    """1.2 M rows in one bucket: two sampling levels above the tiles."""
    from lakesoul_b200 import datagen
    _roundtrip(datagen.config_spec(1, 1.0), tmp_path, batch_size=8192)


@pytest.mark.parametrize("encoding", ["E1", "E2", "E3"])
def test_config2_lineitem_composite_pk(tmp_path, encoding):
    from lakesoul_b200 import datagen
    _roundtrip(datagen.config_spec(2, 0.05), tmp_path, encoding=encoding, buckets=[0, 3])


def test_config3_cdc_deletes_visible_and_filtered(tmp_path):
    from lakesoul_b200 import datagen
    spec = datagen.config_spec(3, 0.2)
    _roundtrip(spec, tmp_path, buckets=[0, 3])
    sub = tmp_path / "f"
    sub.mkdir()
    _roundtrip(spec, sub, buckets=[0, 3], filters=['not(eq(rowKinds, Binary{"delete"}))'])


def test_config4_partial_merge_wide_table(tmp_path):
    from lakesoul_b200 import datagen
    _roundtrip(datagen.config_spec(4, 0.05), tmp_path, buckets=[0, 3])


def test_config5_64_deltas(tmp_path):
    from lakesoul_b200 import datagen
    _roundtrip(datagen.config_spec(5, 0.1), tmp_path, buckets=[0, 1], nbuckets=2)


def test_default_partition_columns_and_projection(tmp_path):
    from lakesoul_b200 import datagen, read_bucket
    spec = datagen.config_spec(1, 0.01)
    paths = datagen.write_table(str(tmp_path), spec)[0]
    schema = pa.schema([pa.field("c1", pa.int32()), pa.field("dt", pa.string()), pa.field("region", pa.int32()),
                        pa.field("missing", pa.float64())])
    kw = dict(default_columns={"dt": "2026-09-21", "region": "7"})
    got = read_bucket(paths, schema, spec.primary_keys, batch_size=512, **kw)
    want = mo.read_bucket(paths, schema, spec.primary_keys, **kw)
    assert_tables_equal(got, want, "defaults")
    assert got.schema.names == ["c1", "dt", "region", "missing"]


def test_count_star_empty_schema(tmp_path):
    from lakesoul_b200 import datagen, read_bucket
    spec = datagen.config_spec(1, 0.01)
    paths = datagen.write_table(str(tmp_path), spec)[0]
    got = read_bucket(paths, pa.schema([]), spec.primary_keys, batch_size=100)
    want = mo.read_bucket(paths, spec.schema, spec.primary_keys)
    assert got.num_columns == 0 and got.num_rows == want.num_rows


def test_skip_merge_on_read_concatenates(tmp_path):
    from lakesoul_b200 import datagen, read_bucket
    spec = datagen.config_spec(1, 0.01)
    paths = datagen.write_table(str(tmp_path), spec)[0]
    opts = {"skip_merge_on_read": "true"}
    got = read_bucket(paths, spec.schema, spec.primary_keys, batch_size=333, options=opts)
    want = mo.read_bucket(paths, spec.schema, spec.primary_keys, options=opts)
    assert_tables_equal(got, want, "skip merge")


def test_merge_operators_sum_and_not_null(tmp_path):
    from lakesoul_b200 import datagen, read_bucket
    spec = datagen.config_spec(1, 0.02)
    paths = datagen.write_table(str(tmp_path), spec)[0]
    ops = {"c0": "SumAll", "c1": "UseLastNotNull", "c2": "SumLast"}
    got = read_bucket(paths, spec.schema, spec.primary_keys, batch_size=4096, merge_ops=ops)
    want = mo.read_bucket(paths, spec.schema, spec.primary_keys, merge_ops=ops)
    assert_tables_equal(got, want, "ops")


@pytest.mark.parametrize("with_filter", [False, True])
def test_merge_operators_joined_strings(tmp_path, with_filter):
    """JoinedLast* / JoinedAll* (merge_operator.rs:389-487) incl. the reference's quirk that JoinedAll never
    emits its delimiter, nulls poisoning the cell, in-file duplicate keys and a file that lacks the column."""
    from lakesoul_b200 import read_bucket
    rng = np.random.default_rng(11)
    schema = pa.schema([("id", pa.int64()), ("a", pa.string()), ("b", pa.string()), ("c", pa.string()), ("v", pa.int32())])
    paths = []
    for f in range(4):
        n = 3000
        ids = np.sort(rng.integers(0, 2500, n))  # duplicates inside a file
        def strs(p_null):
            vals = ["s%d_%d" % (f, x) * int(rng.integers(0, 3)) for x in rng.integers(0, 1000, n)]
            mask = rng.random(n) < p_null
            return pa.array([None if m else v for v, m in zip(vals, mask)], pa.string())
        cols = {"id": pa.array(ids, pa.int64()), "a": strs(0.02), "b": strs(0.0), "c": strs(0.3),
                "v": pa.array(rng.integers(0, 100, n), pa.int32())}
        names = [x for x in schema.names if not (f == 2 and x == "b")]  # file 2 lacks column b
        t = pa.table({k: cols[k] for k in names})
        path = os.path.join(str(tmp_path), "part-%016d_0000.parquet" % f)
        pq.write_table(t, path, row_group_size=1024)
        paths.append(path)
    ops = {"a": "JoinedLastByComma", "b": "JoinedAllBySemicolon", "c": "JoinedLastBySemicolon", "v": "SumAll"}
    filters = ["gt(v, 150)"] if with_filter else []
    got = read_bucket(paths, schema, ["id"], batch_size=500, merge_ops=ops, filters=filters)
    want = mo.read_bucket(paths, schema, ["id"], merge_ops=ops, filters=filters)
    assert want.num_rows > 0 and want.column("a").null_count > 0
    assert_tables_equal(got, want, "joined")


def test_reference_sample_parquet_fixture(tmp_path):
    """G12: Spark-written snappy files with PLAIN_DICTIONARY / RLE / BIT_PACKED pages (row count 1000 each,
    reader.rs:421-441). The fixtures travel as a copy under tests/golden/sample-parquet-files if present."""
    from lakesoul_b200 import read_bucket
    d = os.path.join(os.path.dirname(__file__), "golden", "sample-parquet-files")
    if not os.path.isdir(d):
        pytest.skip("sample parquet fixtures not available")
    for name in sorted(os.listdir(d)):
        p = os.path.join(d, name)
        want = pq.read_table(p)
        got = read_bucket([p], want.schema, [], batch_size=256)
        assert got.num_rows == 1000
        assert_tables_equal(got, want, name)


def test_host_decode_fallback_gives_the_same_table(tmp_path, monkeypatch):
    """Files the device decoders cannot frame go through Arrow C++ on the host; forcing that path
    must not change the result."""
    from lakesoul_b200 import datagen
    monkeypatch.setenv("LAKESOUL_B200_HOST_DECODE", "1")
    _roundtrip(datagen.config_spec(3, 0.2), tmp_path, buckets=[0])


def test_device_decode_nulls_dictionary_and_v2_pages(tmp_path):
    """Optional columns with nulls, dictionary-encoded strings/ints, booleans, data page v2, several
    row groups and small pages: every device decoder against pyarrow's decode of the same file."""
    from lakesoul_b200 import datagen, read_bucket
    rng = np.random.default_rng(11)
    n = 50_000
    words = np.array(["alpha", "beta", "", "gamma-delta-epsilon", "z"], dtype=object)
    t = pa.table({
        "id": pa.array(np.arange(n, dtype=np.int64)),
        "i32": pa.array(rng.integers(-5, 5, n).astype(np.int32), mask=rng.random(n) < 0.3),
        "i8": pa.array(rng.integers(-100, 100, n).astype(np.int8)),
        "f64": pa.array(rng.standard_normal(n), mask=rng.random(n) < 0.1),
        "f32": pa.array(rng.standard_normal(n).astype(np.float32)),
        "flag": pa.array(rng.random(n) < 0.5, mask=rng.random(n) < 0.2),
        "s": pa.array(words[rng.integers(0, 5, n)], type=pa.string(), mask=rng.random(n) < 0.25),
        "u": pa.array([("k%d" % v) * (v % 7) for v in rng.integers(0, 10**6, n)], type=pa.string()),
        "dec": datagen._decimal(rng, n, -10**9, 10**9),
        "d": pa.array(rng.integers(0, 20000, n).astype(np.int32)).cast(pa.date32()),
        "ts": pa.array(rng.integers(0, 2**50, n)).cast(pa.timestamp("us", "UTC")),
    })
    variants = {
        "v1_dict_snappy": dict(data_page_version="1.0", use_dictionary=True, compression="SNAPPY", data_page_size=4096, row_group_size=17_000),
        "v2_dict_zstd": dict(data_page_version="2.0", use_dictionary=True, compression="ZSTD", data_page_size=8192, row_group_size=50_000),
        "v1_plain_none": dict(data_page_version="1.0", use_dictionary=False, compression="NONE", data_page_size=1 << 20, row_group_size=20_000),
        "v2_plain_none": dict(data_page_version="2.0", use_dictionary=False, compression="NONE", data_page_size=1000, row_group_size=50_000),
    }
    for name, opts in variants.items():
        p = str(tmp_path / f"part-{name}_0000.parquet")
        pq.write_table(t, p, **opts)
        got, reader = read_bucket([p], t.schema, ["id"], batch_size=4096, return_reader=True)
        st = reader.stats()
        reader.close()
        assert st["device_decoded_pages"] > 0 and st["host_decoded_files"] == 0, (name, st)
        assert_tables_equal(got, pq.read_table(p), name)


def test_device_decode_long_strings_and_all_null_blocks(tmp_path):
    """PLAIN BYTE_ARRAY values longer than the walker's 4 KiB window and than the gather's 24 KiB stage,
    empty strings, blocks and whole pages of nulls, one-row row groups; plain and dictionary pages."""
    from lakesoul_b200 import read_bucket
    rng = np.random.default_rng(5)
    n = 6000
    lens = rng.choice([0, 1, 7, 300, 5000, 30000], size=n, p=[0.2, 0.3, 0.3, 0.15, 0.04, 0.01])
    alphabet = np.frombuffer(b"abcdefghijklmnopqrstuvwxyz0123456789", dtype=np.uint8)
    big = [bytes(alphabet[rng.integers(0, len(alphabet), int(l))]).decode() for l in lens]
    null_run = np.zeros(n, dtype=bool)
    null_run[1000:4100] = True  # three whole 1024-row blocks of nulls inside one page
    t = pa.table({
        "id": pa.array(np.arange(n, dtype=np.int64)),
        "big": pa.array(big, type=pa.string(), mask=rng.random(n) < 0.3),
        "sparse": pa.array(["x%d" % v for v in rng.integers(0, 50, n)], type=pa.string(), mask=null_run),
        "n32": pa.array(rng.integers(0, 1000, n).astype(np.int32), mask=null_run),
        "allnull": pa.array([None] * n, type=pa.int64()),
    })
    variants = {
        "plain_big_pages": dict(use_dictionary=False, compression="NONE", data_page_size=8 << 20, row_group_size=n),
        "dict_small_groups": dict(use_dictionary=True, compression="SNAPPY", data_page_size=2048, row_group_size=1500),
        "plain_v2_one_row_groups": dict(use_dictionary=False, compression="NONE", data_page_version="2.0", row_group_size=997),
    }
    for name, opts in variants.items():
        p = str(tmp_path / f"part-{name}_0000.parquet")
        pq.write_table(t, p, **opts)
        # one more tiny file in front: a single row that is overridden by the big file
        p0 = str(tmp_path / f"part-{name}_tiny_0000.parquet")
        pq.write_table(t.slice(17, 1), p0, row_group_size=1)
        got, reader = read_bucket([p0, p], t.schema, ["id"], batch_size=257, return_reader=True)
        st = reader.stats()
        reader.close()
        assert st["device_decoded_pages"] > 0 and st["host_decoded_files"] == 0, (name, st)
        assert_tables_equal(got, pq.read_table(p), name)


def test_corrupt_files_fail_cleanly(tmp_path):
    """A truncated file and a file that is not Parquet at all are start_reader errors (status < 0 + message),
    never a crash; the reader object stays usable for free_lakesoul_reader."""
    from lakesoul_b200 import datagen, read_bucket
    spec = datagen.config_spec(1, 0.01)
    paths = datagen.write_table(str(tmp_path), spec)[0]
    raw = open(paths[0], "rb").read()
    cut = str(tmp_path / "part-cut_0000.parquet")
    open(cut, "wb").write(raw[: len(raw) // 2])
    junk = str(tmp_path / "part-junk_0000.parquet")
    open(junk, "wb").write(b"PAR1" + bytes(range(256)) * 8 + b"PAR1")
    for bad in (cut, junk):
        with pytest.raises(Exception) as ei:
            read_bucket([paths[1], bad], spec.schema, spec.primary_keys)
        assert "parquet" in str(ei.value).lower() or "invalid" in str(ei.value).lower(), str(ei.value)
    # the good files still read after the failures (slots and arenas were released)
    got = read_bucket(paths, spec.schema, spec.primary_keys)
    want = mo.read_bucket(paths, spec.schema, spec.primary_keys)
    assert_tables_equal(got, want, "after failures")


def test_record_batch_reader_and_arrow_c_stream(tmp_path):
    """the bucket as a streaming pyarrow.RecordBatchReader / Arrow C stream."""
    from lakesoul_b200 import datagen, open_bucket
    spec = datagen.config_spec(1, 0.01)
    paths = datagen.write_table(str(tmp_path), spec)[0]
    want = mo.read_bucket(paths, spec.schema, spec.primary_keys)
    rdr = open_bucket(paths, spec.schema, spec.primary_keys, batch_size=1000)
    assert isinstance(rdr, pa.RecordBatchReader) and rdr.schema.names == spec.schema.names
    first = rdr.read_next_batch()
    assert first.num_rows == min(1000, want.num_rows)
    rest = rdr.read_all()
    assert_tables_equal(pa.concat_tables([pa.Table.from_batches([first]), rest]), want, "stream")
    # the same through the C stream interface
    got = pa.table(open_bucket(paths, spec.schema, spec.primary_keys, batch_size=777))
    assert_tables_equal(got, want, "c stream")


@pytest.mark.parametrize("key_type", ["int64", "string"])
def test_key_filters_prune_row_groups(tmp_path, key_type):
    """conjuncts on primary-key columns skip row groups by min/max statistics; the result is the
    one the post-merge filter alone gives (the oracle never prunes)."""
    from lakesoul_b200 import read_bucket
    rng = np.random.default_rng(3)
    def keys(v):
        return pa.array(v.astype(np.int64)) if key_type == "int64" else pa.array(["k%08d" % x for x in v], pa.string())
    # string literals arrive in parquet-mr's toString form (filter/parser.rs:220-223)
    lit = (lambda x: str(x)) if key_type == "int64" else (lambda x: 'Binary{"k%08d"}' % x)
    schema = pa.schema([pa.field("id", pa.int64() if key_type == "int64" else pa.string(), False), ("v", pa.int32()), ("s", pa.string())])
    paths, total_rows = [], 0
    for f, n in enumerate((40_000, 9_000, 9_000)):
        ids = np.sort(rng.choice(100_000, n, replace=False))
        t = pa.table({"id": keys(ids), "v": pa.array(rng.integers(0, 10, n).astype(np.int32)),
                      "s": pa.array(["f%d_%d" % (f, x) for x in ids], pa.string())}).cast(schema)
        p = os.path.join(str(tmp_path), "part-%016d_0000.parquet" % f)
        pq.write_table(t, p, row_group_size=2_000, use_dictionary=(f == 1))
        paths.append(p)
        total_rows += n
    cases = [f"gt(id, {lit(90_000)})", f"and(gteq(id, {lit(30_000)}), lt(id, {lit(33_000)}))", f"eq(id, {lit(int(1234))})",
             f"or(lt(id, {lit(2_000)}), gt(id, {lit(98_000)}))", f"not(lt(id, {lit(95_000)}))",
             f"and(gt(id, {lit(50_000)}), gt(v, 5))", f"and(noteq(id, {lit(7)}), lteq(id, {lit(1_000)}))",
             f"gt(id, {lit(200_000)})"]
    for flt in cases:
        got, reader = read_bucket(paths, schema, ["id"], batch_size=1024, filters=[flt], return_reader=True)
        st = reader.stats()
        reader.close()
        want = mo.read_bucket(paths, schema, ["id"], filters=[flt])
        assert_tables_equal(got, want, flt)
        assert st["pruned_row_groups"] > 0 and st["rows_in"] < total_rows, (flt, st)
    # a predicate on a payload column alone must not prune anything
    got, reader = read_bucket(paths, schema, ["id"], filters=["gt(v, 8)"], return_reader=True)
    assert reader.stats()["pruned_row_groups"] == 0
    reader.close()
    assert_tables_equal(got, mo.read_bucket(paths, schema, ["id"], filters=["gt(v, 8)"]), "payload filter")


def test_zstd_pages_are_expanded_on_the_device(tmp_path):
    """PLAIN + ZSTD without dictionaries is what the LakeSoul writer emits (multipart_writer.rs:128-129): with
    LAKESOUL_B200_DEVICE_ZSTD_SHARE=1 every page payload stays compressed on the host and are expanded by
    zstd_pages_kernel; same table as the host codec path. (The variable is read once per process: the test runs
    the reader in a child process.)"""
    import subprocess
    import sys
    import textwrap
    if os.environ.get("LAKESOUL_B200_DEVICE_ZSTD_SHARE") != "1":
        here = os.path.dirname(os.path.abspath(__file__))
        root = os.path.dirname(here)
        env = dict(os.environ, LAKESOUL_B200_DEVICE_ZSTD_SHARE="1")
        code = textwrap.dedent(f"""
            import sys, pathlib
            sys.path[:0] = [{here!r}, {os.path.join(root, "oracle")!r}, {root!r}]
            import test_gpu_reader as t
            t.test_zstd_pages_are_expanded_on_the_device(pathlib.Path({str(tmp_path)!r}))
            print("child ok")
        """)
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
        assert out.returncode == 0 and "child ok" in out.stdout, out.stdout[-1500:] + out.stderr[-3000:]
        return
    from lakesoul_b200 import datagen, read_bucket
    spec = datagen.config_spec(2, 0.05)
    spec.buckets = 2
    paths = datagen.write_table(str(tmp_path), spec, encoding="E3")
    for b in range(2):
        got, reader = read_bucket(paths[b], spec.schema, spec.primary_keys, batch_size=4096, return_reader=True)
        st = reader.stats()
        reader.close()
        assert st["device_decompressed_pages"] > 0 and st["host_decoded_files"] == 0, st
        assert_tables_equal(got, mo.read_bucket(paths[b], spec.schema, spec.primary_keys), f"zstd bucket {b}")
    # large pages (several 128 KiB ZSTD blocks per page), several compression levels, nulls
    rng = np.random.default_rng(9)
    n = 120_000
    t = pa.table({"id": pa.array(np.arange(n, dtype=np.int64)),
                  "txt": pa.array(["row %d %s" % (i, "lorem ipsum dolor sit amet " * int(rng.integers(0, 4))) for i in range(n)], pa.string()),
                  "f": pa.array(rng.standard_normal(n), mask=rng.random(n) < 0.2),
                  "k": pa.array(rng.integers(0, 50, n).astype(np.int32))})
    for level in (1, 3, 12):
        p = str(tmp_path / f"part-lvl{level}_0000.parquet")
        pq.write_table(t, p, use_dictionary=False, compression="ZSTD", compression_level=level, data_page_size=2 << 20,
                       row_group_size=60_000)
        got, reader = read_bucket([p], t.schema, ["id"], return_reader=True)
        st = reader.stats()
        reader.close()
        assert st["device_decompressed_pages"] > 0, st
        assert_tables_equal(got, pq.read_table(p), f"zstd level {level}")


from conftest import upsert_case_files, upsert_cases  # noqa: E402

UPSERTS = upsert_cases()


@pytest.mark.parametrize("case", UPSERTS, ids=[c["name"] for c in UPSERTS])
def test_reader_matches_reference_upsert_golden(case, tmp_path):
    """G9 through the C-ABI: default (partition) column, partial merge, pushed string filters."""
    from lakesoul_b200 import read_bucket
    paths, schema, expected = upsert_case_files(case, tmp_path)
    got = read_bucket(paths, schema, case["primary_keys"], batch_size=2, default_columns=case["partition"],
                      filters=case["filters"])
    assert_tables_equal(got, expected, case["name"])


def test_nullable_primary_key_nulls_first_and_equal(tmp_path):
    """cursor.rs:444-476: null keys sort first and compare equal to each other."""
    from lakesoul_b200 import read_bucket
    schema = pa.schema([pa.field("k", pa.int32(), True), pa.field("v", pa.int32(), True)])
    f0 = pa.table({"k": pa.array([None, None, 1, 5], type=pa.int32()), "v": pa.array([1, 2, 3, 4], type=pa.int32())})
    f1 = pa.table({"k": pa.array([None, 5, 7], type=pa.int32()), "v": pa.array([10, 50, 70], type=pa.int32())})
    paths = []
    for j, t in enumerate((f0, f1)):
        p = str(tmp_path / f"part-{j:016d}_0000.parquet")
        pq.write_table(t.cast(schema), p)
        paths.append(p)
    got = read_bucket(paths, schema, ["k"], batch_size=3)
    want = mo.read_bucket(paths, schema, ["k"])
    assert_tables_equal(got, want, "null pk")
    assert got.column("k").to_pylist() == [None, 1, 5, 7] and got.column("v").to_pylist() == [10, 3, 50, 70]


def test_schema_evolution_cast_and_missing_column(tmp_path):
    """a2: a file written with an older schema (int32 column, one column missing) is cast / null-filled
    (helpers/transform.rs:107-354); needs the host decoder for the cast."""
    from lakesoul_b200 import read_bucket
    old = pa.table({"id": pa.array([1, 2, 3], type=pa.int64()), "qty": pa.array([10, 20, 30], type=pa.int32())})
    new = pa.table({"id": pa.array([2, 4], type=pa.int64()), "qty": pa.array([200, 400], type=pa.int64()),
                    "tag": pa.array(["b", "d"])})
    paths = []
    for j, t in enumerate((old, new)):
        p = str(tmp_path / f"part-{j:016d}_0000.parquet")
        pq.write_table(t, p)
        paths.append(p)
    schema = pa.schema([pa.field("id", pa.int64(), False), pa.field("qty", pa.int64()), pa.field("tag", pa.string())])
    got = read_bucket(paths, schema, ["id"], batch_size=2)
    assert got.column("id").to_pylist() == [1, 2, 3, 4]
    assert got.column("qty").to_pylist() == [10, 200, 30, 400]
    assert got.column("tag").to_pylist() == [None, "b", None, "d"]


@pytest.mark.parametrize("binary", [False, True])
def test_string_primary_key(tmp_path, binary):
    """Utf8 / Binary single-column PK (reference typed fast path sorted_stream_merger.rs:311-315):
    byte-wise unsigned order, long common prefixes, empty and null keys, duplicates inside a file."""
    from lakesoul_b200 import read_bucket
    rng = np.random.default_rng(21)
    typ = pa.binary() if binary else pa.string()
    schema = pa.schema([pa.field("k", typ, True), pa.field("v", pa.int64()), pa.field("s", pa.string())])

    def keys(n):
        pool = ["user_%08d" % i for i in range(3000)] + ["", "a", "ab", "ab\x00", "user_", "zz", "été", "user_00000001x"]
        ks = [pool[i] for i in rng.integers(0, len(pool), n)]
        if binary:
            return sorted(k.encode() for k in ks)
        return sorted(ks, key=lambda x: x.encode())
    paths = []
    for j, n in enumerate((4000, 700, 700, 50)):
        ks = keys(n)
        if j == 1:
            ks = [None, None] + ks
        t = pa.table({"k": pa.array(ks, type=typ), "v": pa.array(rng.integers(0, 10**9, len(ks))),
                      "s": pa.array(["f%d_%d" % (j, i) for i in range(len(ks))])}).cast(schema)
        p = str(tmp_path / f"part-{j:016d}_0000.parquet")
        pq.write_table(t, p)
        paths.append(p)
    got = read_bucket(paths, schema, ["k"], batch_size=1000)
    want = mo.read_bucket(paths, schema, ["k"])
    assert_tables_equal(got, want, "string pk")
