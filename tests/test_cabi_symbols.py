# SPDX-License-Identifier: Apache-2.0
"""The drop-in library loads and exports every symbol include/*.h declares (no compute calls)."""
import ctypes
import os
import re

from conftest import ROOT


def _declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set()
    for m in re.finditer(r"\b([a-z_0-9]+)\s*\(", text):
        n = m.group(1)
        if n.startswith(("lakesoul_", "lsb_", "new_", "create_", "check_", "start_", "next_", "free_", "write_",
                         "export_", "flush_", "abort_", "apply_", "tokio_", "rust_logger")):
            names.add(n)
    return names


def test_library_exports_every_declared_symbol(lib):
    declared = _declared("lakesoul_c_bindings.h") | _declared("lakesoul_b200_ext.h")
    assert len(declared) >= 60
    missing = [n for n in sorted(declared) if not hasattr(lib, n)]
    assert not missing, f"symbols declared in include/*.h but not exported: {missing}"


def test_python_binding_table_matches_header():
    from lakesoul_b200 import _lib
    declared = _declared("lakesoul_c_bindings.h")
    assert set(_lib.REFERENCE_SYMBOLS) == declared


def test_struct_layouts_match_reference(lib):
    from lakesoul_b200._lib import CResult, CStatus
    # CResult { ptr, err } (lib.rs:40-44); CStatus { err, status } (lib.rs:77-81)
    assert ctypes.sizeof(CResult) == 16 and CResult.ptr.offset == 0 and CResult.err.offset == 8
    assert ctypes.sizeof(CStatus) == 16 and CStatus.err.offset == 0 and CStatus.status.offset == 8


def test_builder_lifecycle_without_gpu(lib):
    b = lib.new_lakesoul_io_config_builder()
    b = lib.lakesoul_config_builder_add_single_file(b, b"/nonexistent/part-0000.parquet")
    b = lib.lakesoul_config_builder_set_batch_size(b, 128)
    b = lib.lakesoul_config_builder_set_option(b, b"is_compacted", b"true")
    cfg = lib.create_lakesoul_io_config_from_builder(b)
    rt = lib.create_tokio_runtime_from_builder(lib.new_tokio_runtime_builder())
    reader = lib.create_lakesoul_reader_from_config(cfg, rt)
    assert lib.check_reader_created(reader) is None
    st = lib.start_reader(reader)  # no GPU here or missing file: must be an error, never a silent success
    assert st.contents.status == -1 and st.contents.err
    lib.free_c_status(st)
    lib.free_lakesoul_reader(reader)


def test_empty_file_list_is_a_constructor_error(lib):
    # reader.rs:108-110
    cfg = lib.create_lakesoul_io_config_from_builder(lib.new_lakesoul_io_config_builder())
    reader = lib.create_lakesoul_reader_from_config_with_global_runtime(cfg)
    assert lib.check_reader_created(reader) == b"LakeSoulReader has the wrong number of files"
    lib.free_lakesoul_reader(reader)


def test_writer_entry_points_report_errors(lib):
    cfg = lib.create_lakesoul_io_config_from_builder(lib.new_lakesoul_io_config_builder())
    rt = lib.create_tokio_runtime_from_builder(lib.new_tokio_runtime_builder())
    w = lib.create_lakesoul_writer_from_config(cfg, rt)
    st = lib.check_writer_created(w)
    assert st.contents.status == -1
    lib.free_c_status(st)
    lib.free_lakesoul_writer(w)


def test_library_exports_nothing_but_the_declared_symbols():
    """Link hygiene (csrc/exports.map): no libstdc++ / Arrow / CUDA-runtime symbols leak into the host JVM."""
    import subprocess
    so = os.path.join(ROOT, "lakesoul_b200", "lib", "liblakesoul_io_c.so")
    out = subprocess.run(["nm", "-D", "--defined-only", so], capture_output=True, text=True, check=True).stdout
    exported = {ln.split()[-1].split("@")[0] for ln in out.splitlines() if ln.strip()}
    declared = _declared("lakesoul_c_bindings.h") | _declared("lakesoul_b200_ext.h")
    assert exported == declared, (sorted(exported - declared)[:10], sorted(declared - exported)[:10])


def test_hash_bucket_elimination_needs_no_gpu_and_matches_the_table_hash(tmp_path, lib):
    """N2: eq filters on the whole primary key select one bucket (spark murmur3, utils/hash/mod.rs; KATs G11). A
    reader whose files belong to another bucket returns an empty stream without touching the files or the GPU; for
    the matching bucket it must go on (and, without a GPU, fail on the device -- never silently return nothing)."""
    import pyarrow as pa
    from lakesoul_b200 import hashing
    from lakesoul_b200.native_io import NativeIOReader

    def reader_for(bucket, nbuckets, schema, pks, filters):
        r = NativeIOReader()
        r.addFile(str(tmp_path / f"part-abcdefghijklmnop_{bucket:04d}.parquet"))  # never opened when pruned
        r.setPrimaryKeys(pks)
        r.setSchema(schema)
        r.ioConfigBuilder = lib.lakesoul_config_builder_set_hash_bucket_num(r.ioConfigBuilder, nbuckets)
        for f in filters:
            r.addFilter(f)
        return r

    nb = 16
    schema = pa.schema([pa.field("k", pa.int64()), pa.field("l", pa.int32()), pa.field("v", pa.string())])
    for k, l in ((1, 1), (123456789, 7), (-5, 0)):
        import numpy as np
        want = int(hashing.bucket_of([np.array([k], dtype=np.int64), np.array([l], dtype=np.int32)], nb)[0])
        flt = [f"and(eq(k, {k}), and(eq(l, {l}), noteq(v, null)))"]
        for bucket in range(nb):
            r = reader_for(bucket, nb, schema, ["k", "l"], flt)
            if bucket != want:
                r.initializeReader()  # no file access, no GPU
                assert r.stats()["pruned_buckets"] == 1
                holder = (ctypes.c_byte * 80)()
                assert r.nextBatchBlocked(ctypes.addressof(holder)) == 0
                assert r.getSchema().names == ["k", "l", "v"]
            else:
                try:
                    r.initializeReader()
                    raise AssertionError("the selected bucket must be read, not pruned")
                except IOError as e:
                    assert "not found" in str(e) or "CUDA" in str(e), e
            r.close()
    # string keys: the reference's own known answers (hash_tests.rs:48-90): "321" chained after "321" -> -218318595
    sschema = pa.schema([pa.field("a", pa.string()), pa.field("b", pa.string())])
    want = (-218318595 & 0xFFFFFFFF) % 7
    for bucket in range(7):
        r = reader_for(bucket, 7, sschema, ["a", "b"], ['and(eq(a, Binary{"321"}), eq(b, Binary{"321"}))'])
        try:
            r.initializeReader()
            pruned = r.stats()["pruned_buckets"] == 1
        except IOError:
            pruned = False
        r.close()
        assert pruned == (bucket != want), (bucket, want)
    # a filter that does not pin the whole key prunes nothing
    r = reader_for(3, nb, schema, ["k", "l"], ["eq(k, 1)"])
    with_gpu_or_file_error = False
    try:
        r.initializeReader()
    except IOError:
        with_gpu_or_file_error = True
    r.close()
    assert with_gpu_or_file_error
