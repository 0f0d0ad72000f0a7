# SPDX-License-Identifier: Apache-2.0
"""Parity floor: float primary keys in IEEE total order, exported schema
nullability, the no-merge path with a column missing from a file, timestamp-unit key pruning and the callback
entry points of the C-ABI."""
import ctypes
import struct

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq
import pytest

import mor_oracle as mo
from conftest import assert_tables_equal

pytestmark = pytest.mark.gpu


def _bits_equal(got: pa.ChunkedArray, want: pa.ChunkedArray, np_u):
    g = got.combine_chunks()
    w = want.combine_chunks()
    assert np.array_equal(np.asarray(g.is_valid()), np.asarray(w.is_valid()))
    gv = np.asarray(g.fill_null(0)).view(np_u)
    wv = np.asarray(w.fill_null(0)).view(np_u)
    m = np.asarray(g.is_valid())
    assert np.array_equal(gv[m], wv[m]), (gv[m][:8], wv[m][:8])


def _f64(bits):
    return struct.unpack("<d", struct.pack("<Q", bits))[0]


def _f32(bits):
    return struct.unpack("<f", struct.pack("<I", bits))[0]


@pytest.mark.parametrize("ftype", ["f64", "f32"])
def test_float_primary_key_total_order(tmp_path, ftype):
    """cursor.rs:265-276: float keys compare with total_cmp and are equal only bit for bit: -0.0 < +0.0 are two keys,
    NaNs order by sign and payload, -NaN < -inf and +inf < +NaN."""
    if ftype == "f64":
        t, np_f, np_u = pa.float64(), np.float64, np.uint64
        specials = [_f64(0xfff8000000000001), _f64(0xfff8000000000000), -np.inf, -1.5, -5e-324, -0.0, 0.0, 5e-324, 1.5,
                    np.inf, _f64(0x7ff8000000000000), _f64(0x7ff8000000000001), _f64(0x7ff0000000000001)]
    else:
        t, np_f, np_u = pa.float32(), np.float32, np.uint32
        specials = [_f32(0xffc00001), _f32(0xffc00000), -np.inf, -1.5, -0.0, 0.0, 1.5, np.inf, _f32(0x7fc00000),
                    _f32(0x7fc00001)]
    rng = np.random.Generator(np.random.PCG64(5))

    def total_image(a):
        if np_f == np.float64:
            b = a.view(np.int64)
            return b ^ ((b >> 63).view(np.uint64) >> np.uint64(1)).view(np.int64)
        b = a.view(np.int32)
        return b ^ ((b >> 31).view(np.uint32) >> np.uint32(1)).view(np.int32)

    files = []
    for j in range(4):
        body = rng.standard_normal(3000).astype(np_f) * np_f(100)
        pick = np.array([s for s in specials if rng.random() < 0.8], dtype=np_f)
        pick_bits = pick.view(np_u)  # keep NaN payloads through the concatenation
        keys = np.concatenate([body.view(np_u), pick_bits, pick_bits[: len(pick_bits) // 2]]).view(np_f)
        keys = keys[np.argsort(total_image(keys), kind="stable")]
        tbl = pa.table({"k": pa.array(keys, type=t), "v": pa.array(np.arange(len(keys), dtype=np.int32) + 100000 * j)})
        p = str(tmp_path / f"part-{j:016d}_0000.parquet")
        pq.write_table(tbl, p, use_dictionary=False, compression="NONE")
        files.append(p)
    from lakesoul_b200 import read_bucket
    schema = pa.schema([pa.field("k", t), pa.field("v", pa.int32())])
    got = read_bucket(files, schema, ["k"], batch_size=777)
    want = mo.read_bucket(files, schema, ["k"])
    assert got.num_rows == want.num_rows
    _bits_equal(got.column("k"), want.column("k"), np_u)
    assert got.column("v").combine_chunks().equals(want.column("v").combine_chunks())
    # the oracle's order is itself total_cmp order with bitwise-distinct keys
    img = total_image(np.asarray(want.column("k").combine_chunks()))
    assert np.all(img[1:] > img[:-1])


def test_exported_schema_nullability(tmp_path):
    """merge/mod.rs:75-97: an output field is nullable when the target field is, when any file declares it
    nullable, or when a file lacks it; otherwise it stays non-nullable."""
    from lakesoul_b200 import NativeIOReader
    f0 = pa.table({"pk": pa.array([1, 2, 3], type=pa.int64()), "a": pa.array([1, 2, 3], type=pa.int32()),
                   "b": pa.array([1, 2, 3], type=pa.int32()), "c": pa.array(["x", "y", "z"])},
                  schema=pa.schema([pa.field("pk", pa.int64(), False), pa.field("a", pa.int32(), False),
                                    pa.field("b", pa.int32(), True), pa.field("c", pa.string(), False)]))
    f1 = pa.table({"pk": pa.array([2, 4], type=pa.int64()), "a": pa.array([20, 40], type=pa.int32()),
                   "b": pa.array([20, 40], type=pa.int32())},
                  schema=pa.schema([pa.field("pk", pa.int64(), False), pa.field("a", pa.int32(), False),
                                    pa.field("b", pa.int32(), False)]))
    paths = []
    for j, t in enumerate((f0, f1)):
        p = str(tmp_path / f"part-{j:016d}_0000.parquet")
        pq.write_table(t, p)
        paths.append(p)
    target = pa.schema([pa.field("pk", pa.int64(), False), pa.field("a", pa.int32(), False), pa.field("b", pa.int32(), False),
                        pa.field("c", pa.string(), False), pa.field("d", pa.float64(), False)])
    r = NativeIOReader()
    for p in paths:
        r.addFile(p)
    r.setPrimaryKeys(["pk"])
    r.setSchema(target)
    r.initializeReader()
    s = r.getSchema()
    r.close()
    assert s.names == ["pk", "a", "b", "c", "d"]
    # pk, a: non-nullable everywhere; b: nullable in file 0; c: missing from file 1; d: in no file
    assert [s.field(n).nullable for n in s.names] == [False, False, True, True, True]
    assert [s.field(n).type for n in s.names] == [f.type for f in target]


@pytest.mark.parametrize("with_pk", [False, True])
def test_concat_path_null_fills_a_column_missing_from_a_file(tmp_path, with_pk):
    """tables without a primary key / skip_merge_on_read take the concatenation path, where the
    rows of a file that lacks a column must come out null (DefaultColumnStream, default_column.rs:62-108) -- also
    when every file that has the column is proven non-null, and for Utf8 columns."""
    n0, n1, n2 = 5000, 3000, 4000
    nn = lambda name, t: pa.field(name, t, False)  # noqa: E731
    f0 = pa.table({"id": pa.array(np.arange(n0, dtype=np.int64)), "x": pa.array(np.arange(n0, dtype=np.int32)),
                   "s": pa.array([f"s{i}" for i in range(n0)])},
                  schema=pa.schema([nn("id", pa.int64()), nn("x", pa.int32()), nn("s", pa.string())]))
    f1 = pa.table({"id": pa.array(np.arange(n1, dtype=np.int64) + n0)}, schema=pa.schema([nn("id", pa.int64())]))
    f2 = pa.table({"id": pa.array(np.arange(n2, dtype=np.int64) + n0 + n1), "x": pa.array(np.arange(n2, dtype=np.int32) * 3),
                   "s": pa.array([f"t{i}" * (i % 4) for i in range(n2)])},
                  schema=pa.schema([nn("id", pa.int64()), nn("x", pa.int32()), nn("s", pa.string())]))
    paths = []
    for j, t in enumerate((f0, f1, f2)):
        p = str(tmp_path / f"part-{j:016d}_0000.parquet")
        pq.write_table(t, p, use_dictionary=False, compression="NONE")
        paths.append(p)
    from lakesoul_b200 import read_bucket
    schema = pa.schema([pa.field("id", pa.int64()), pa.field("x", pa.int32()), pa.field("s", pa.string())])
    pks = ["id"] if with_pk else []
    opts = {"skip_merge_on_read": "true"} if with_pk else {}
    got, rd = read_bucket(paths, schema, pks, batch_size=1024, options=opts, return_reader=True)
    st = rd.stats()
    rd.close()
    want = mo.read_bucket(paths, schema, pks, options=opts)
    assert_tables_equal(got, want, "concat with a missing column")
    assert got.num_rows == n0 + n1 + n2
    assert got.column("x").null_count == n1 and got.column("s").null_count == n1
    assert st["kernel_launches"] > 0


@pytest.mark.parametrize("unit", ["s", "ms", "us", "ns"])
def test_timestamp_key_filter_prunes_in_the_columns_unit(tmp_path, unit):
    """filter literals are microseconds, row-group statistics are in the column's unit."""
    from lakesoul_b200 import read_bucket
    per_us = {"s": 1e-6, "ms": 1e-3, "us": 1, "ns": 1000}[unit]
    n = 40_000
    base_us = 1_700_000_000_000_000
    step_us = 10_000_000  # 10 s between keys: exact in every unit
    ts_us = base_us + np.arange(n, dtype=np.int64) * step_us
    raw = (ts_us * per_us).astype(np.int64) if per_us >= 1 else (ts_us // int(round(1 / per_us)))
    t = pa.timestamp(unit, "UTC")
    files = []
    for j in range(2):
        sel = np.arange(j, n, 2)
        tbl = pa.table({"ts": pa.array(raw[sel], type=t), "v": pa.array((sel + 7 * j).astype(np.int32))})
        p = str(tmp_path / f"part-{j:016d}_0000.parquet")
        pq.write_table(tbl, p, row_group_size=2000, use_dictionary=False, compression="NONE",
                       coerce_timestamps=None, use_deprecated_int96_timestamps=False, version="2.6")
        files.append(p)
    schema = pa.schema([pa.field("ts", t), pa.field("v", pa.int32())])
    lo = int(ts_us[n // 2])
    hi = int(ts_us[n // 2 + 3000])
    flt = [f"and(gteq(ts, {lo}), lt(ts, {hi}))"]
    got, rd = read_bucket(files, schema, ["ts"], batch_size=4096, filters=flt, return_reader=True)
    st = rd.stats()
    rd.close()
    want = mo.read_bucket(files, schema, ["ts"], filters=flt)
    assert_tables_equal(got, want, f"timestamp[{unit}] key filter")
    assert got.num_rows == 3000
    assert st["pruned_row_groups"] > 0, st


def test_callback_entry_points(tmp_path, lib):
    """start_reader_with_data / next_record_batch / next_record_batch_with_data (lib.rs:791,823,931): same batches as
    the blocked variant, status delivered through the callback, schema exported when schema_addr != 0."""
    from lakesoul_b200 import datagen
    from lakesoul_b200.native_io import NativeIOReader, _ArrowArrayStruct, _ArrowSchemaStruct
    spec = datagen.config_spec(1, 0.01)
    paths = datagen.write_table(str(tmp_path), spec)[0]
    want = mo.read_bucket(paths, spec.schema, spec.primary_keys)

    def configured():
        r = NativeIOReader()
        for p in paths:
            r.addFile(p)
        r.setPrimaryKeys(spec.primary_keys)
        r.setSchema(spec.schema)
        r.setBatchSize(1000)
        runtime = lib.create_tokio_runtime_from_builder(r.tokioRuntimeBuilder)
        r.tokioRuntimeBuilder = None
        config = lib.create_lakesoul_io_config_from_builder(r.ioConfigBuilder)
        r.ioConfigBuilder = None
        r.reader = lib.create_lakesoul_reader_from_config(config, runtime)
        assert lib.check_reader_created(r.reader) is None
        return r

    DATA_CB = ctypes.CFUNCTYPE(None, ctypes.c_bool, ctypes.c_char_p, ctypes.c_void_p)
    I32_CB = ctypes.CFUNCTYPE(None, ctypes.c_int32, ctypes.c_char_p)
    I32_DATA_CB = ctypes.CFUNCTYPE(None, ctypes.c_int32, ctypes.c_char_p, ctypes.c_void_p)

    for with_data in (False, True):
        r = configured()
        started = {}
        token = ctypes.c_int64(4242)

        def on_start(ok, err, data):
            started["ok"], started["err"], started["data"] = ok, err, data

        cb0 = DATA_CB(on_start)
        lib.start_reader_with_data(r.reader, ctypes.addressof(token), ctypes.cast(cb0, ctypes.c_void_p))
        assert started["ok"] is True and started["err"] is None and started["data"] == ctypes.addressof(token)
        batches, schema = [], None
        while True:
            arr, sch = _ArrowArrayStruct(), _ArrowSchemaStruct()
            res = {}
            if with_data:
                def on_next(n, err, data):
                    res["n"], res["err"], res["data"] = n, err, data
                cb = I32_DATA_CB(on_next)
                lib.next_record_batch_with_data(r.reader, ctypes.addressof(sch), ctypes.addressof(arr),
                                                ctypes.addressof(token), ctypes.cast(cb, ctypes.c_void_p))
                assert res["data"] == ctypes.addressof(token)
            else:
                def on_next2(n, err):
                    res["n"], res["err"] = n, err
                cb = I32_CB(on_next2)
                lib.next_record_batch(r.reader, ctypes.addressof(sch), ctypes.addressof(arr), cb)
            assert res["err"] is None and res["n"] >= 0
            if res["n"] == 0:
                break
            schema = pa.Schema._import_from_c(ctypes.addressof(sch))
            batches.append(pa.RecordBatch._import_from_c(ctypes.addressof(arr), schema))
            assert batches[-1].num_rows == res["n"] <= 1000
        r.close()
        got = pa.Table.from_batches(batches, schema=schema)
        assert_tables_equal(got, want, "callback variants")


def _keyed_files(tmp_path, key_cols, payload_rows=4000, nfiles=4, seed=3):
    """files sorted by the key columns (pyarrow sort = the oracle's order: ascending, nulls first, strings bytewise)"""
    import pyarrow.compute as pc
    rng = np.random.Generator(np.random.PCG64(seed))
    files = []
    for j in range(nfiles):
        n = payload_rows if j == 0 else payload_rows // 3
        cols = {name: gen(rng, n) for name, gen in key_cols.items()}
        cols["v"] = pa.array(np.arange(n, dtype=np.int64) + 1_000_000 * j)
        cols["s"] = pa.array([f"f{j}r{i}" for i in range(n)])
        t = pa.table(cols)
        order = pc.sort_indices(t, sort_keys=[(k, "ascending") for k in key_cols], null_placement="at_start")
        t = t.take(order)
        p = str(tmp_path / f"part-{j:016d}_0000.parquet")
        pq.write_table(t, p, use_dictionary=False, compression="NONE")
        files.append(p)
    return files, pa.table(cols).schema


def _with_nulls(arr, rng, frac=0.03):
    mask = rng.random(len(arr)) < frac
    return pa.array(arr.to_pylist(), type=arr.type, mask=mask)


@pytest.mark.parametrize("shape", ["int_string", "string_string", "decimal", "five_columns", "wide_160_bits", "string_with_nuls_and_nulls"])
def test_general_primary_key_shapes_use_row_encoded_keys(tmp_path, shape):
    """Key shapes beyond the packed 64/128-bit fast paths: the reference takes them through
    arrow-row (sorted_stream_merger.rs:322-355, stream/mod.rs:104-122); here they are merged on row-encoded keys."""
    import decimal
    from lakesoul_b200 import read_bucket
    words = ["", "a", "ab", "abc", "b", "ba", "zz", "a\x00", "a\x00b", "\x00", "\u00e9t\u00e9", "long-key-" * 5]
    gi = lambda lo, hi, t: (lambda rng, n: pa.array(rng.integers(lo, hi, size=n), type=t))  # noqa: E731
    gs = lambda pool: (lambda rng, n: pa.array([pool[i] for i in rng.integers(0, len(pool), size=n)], type=pa.string()))  # noqa: E731
    if shape == "int_string":
        keys = {"k1": gi(0, 40, pa.int32()), "k2": gs(words[:7])}
    elif shape == "string_string":
        keys = {"k1": gs(words[:7]), "k2": gs(["x", "y", "xy", ""])}
    elif shape == "decimal":
        keys = {"k1": lambda rng, n: pa.array([decimal.Decimal(int(x)) / 100 for x in rng.integers(-5000, 5000, size=n)],
                                              type=pa.decimal128(15, 2))}
    elif shape == "five_columns":
        keys = {f"k{i}": gi(-3, 4, pa.int16()) for i in range(5)}
    elif shape == "wide_160_bits":
        keys = {"k1": gi(-2**40, -2**40 + 60, pa.int64()), "k2": gi(0, 3, pa.int64()), "k3": gi(-2, 3, pa.int32())}
    else:
        keys = {"k1": lambda rng, n: _with_nulls(pa.array(rng.integers(0, 6, size=n), type=pa.int32()), rng, 0.1),
                "k2": lambda rng, n: _with_nulls(pa.array([words[i] for i in rng.integers(0, len(words), size=n)]), rng, 0.1)}
    files, schema = _keyed_files(tmp_path, keys)
    pks = list(keys)
    got, rd = read_bucket(files, schema, pks, batch_size=1000, return_reader=True)
    st = rd.stats()
    rd.close()
    want = mo.read_bucket(files, schema, pks)
    assert_tables_equal(got, want, f"primary key shape {shape}")
    assert 0 < got.num_rows < sum(pq.read_metadata(f).num_rows for f in files)
    assert st["kernel_launches"] > 0


def test_oversized_scan_group_is_refused_up_front(tmp_path):
    """the whole scan group is resident at once: a group beyond LAKESOUL_B200_MAX_SCAN_BYTES fails with a message, not in
    an allocator (the variable is read once per process: a child process carries it)"""
    import subprocess
    import sys
    t = pa.table({"id": pa.array(range(5000), pa.int64()), "v": pa.array([1.5] * 5000)})
    p = str(tmp_path / "part-aaaaaaaaaaaaaaaa_0000.parquet")
    pq.write_table(t, p)
    code = ("import sys, pyarrow as pa\n"
            "sys.path.insert(0, %r)\n"
            "from lakesoul_b200 import read_bucket\n"
            "s = pa.schema([('id', pa.int64()), ('v', pa.float64())])\n"
            "try:\n"
            "    read_bucket([%r], s, ['id'])\n"
            "    print('READ')\n"
            "except IOError as e:\n"
            "    print('REFUSED' if 'LAKESOUL_B200_MAX_SCAN_BYTES' in str(e) else 'OTHER ' + str(e))\n") % (
        __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))), p)
    env = dict(__import__("os").environ, LAKESOUL_B200_MAX_SCAN_BYTES="1000")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env)
    assert out.stdout.strip().endswith("REFUSED"), (out.stdout, out.stderr[-300:])
