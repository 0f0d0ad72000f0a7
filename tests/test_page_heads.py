# SPDX-License-Identifier: Apache-2.0
"""The library's own Parquet page-header reader (thrift compact protocol, page_plan.cc) against what the footer says about
every column chunk: the headers must chain exactly from the chunk's first page to its last byte, the data pages must carry
the chunk's values, encodings must be among the chunk's. Files: the three layouts the reader meets (E1 PLAIN uncompressed,
E2 dictionary + SNAPPY, E3 PLAIN + ZSTD), v2 data pages with statistics and nulls, and the Spark-written reference fixtures."""
import ctypes
import glob
import os

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq
import pytest

from lakesoul_b200 import load_library

ENC = {0: "PLAIN", 2: "PLAIN_DICTIONARY", 3: "RLE", 4: "BIT_PACKED", 5: "DELTA_BINARY_PACKED", 6: "DELTA_LENGTH_BYTE_ARRAY",
       7: "DELTA_BYTE_ARRAY", 8: "RLE_DICTIONARY", 9: "BYTE_STREAM_SPLIT"}


def _walk(path):
    lib = load_library()
    img = np.fromfile(path, dtype=np.uint8)
    md = pq.ParquetFile(path).metadata
    pages = 0
    for rg in range(md.num_row_groups):
        g = md.row_group(rg)
        for c in range(g.num_columns):
            cc = g.column(c)
            first = cc.dictionary_page_offset if cc.has_dictionary_page and cc.dictionary_page_offset else cc.data_page_offset
            out = np.zeros((4096, 8), dtype=np.int32)
            n = lib.lsb_test_walk_page_heads(img.ctypes.data, first, cc.total_compressed_size, out.ctypes.data, 4096)
            assert n > 0, (path, rg, c, n)
            heads = out[:n]
            data = heads[(heads[:, 0] == 0) | (heads[:, 0] == 3)]
            assert int(data[:, 3].sum()) == cc.num_values, (path, rg, c)
            assert all(ENC.get(int(e), "?") in cc.encodings or ENC.get(int(e)) == "PLAIN_DICTIONARY" for e in data[:, 4]), (cc.encodings, data[:, 4])
            if cc.compression == "UNCOMPRESSED":
                assert (heads[:, 1] == heads[:, 2]).all()
            assert int(heads[:, 2].sum() + heads[:, 7].sum()) == cc.total_compressed_size
            dicts = heads[heads[:, 0] == 2]
            # (parquet-mr 1.x does not record dictionary_page_offset: the chunk then simply starts with its dictionary page)
            assert len(dicts) == 1 if cc.has_dictionary_page else len(dicts) <= 1
            pages += n
    return pages


@pytest.mark.parametrize("encoding", ["E1", "E2", "E3"])
def test_generated_layouts(tmp_path, encoding):
    from lakesoul_b200 import datagen
    for cfg in (2, 3):
        spec = datagen.config_spec(cfg, 0.02)
        spec.buckets = 1
        for p in datagen.write_table(str(tmp_path / f"c{cfg}"), spec, encoding=encoding)[0]:
            assert _walk(p) > 0


def test_v2_pages_statistics_and_nulls(tmp_path):
    rng = np.random.default_rng(0)
    n = 50_000
    t = pa.table({"a": pa.array(rng.integers(0, 1000, n), mask=rng.random(n) < 0.1),
                  "s": pa.array([f"v{i % 311}" for i in range(n)], mask=rng.random(n) < 0.2),
                  "f": rng.standard_normal(n), "b": rng.random(n) < 0.5})
    for i, kw in enumerate([dict(data_page_version="2.0", compression="zstd", use_dictionary=True, data_page_size=4096),
                            dict(data_page_version="2.0", compression="none", use_dictionary=False, write_statistics=True),
                            dict(data_page_version="1.0", compression="snappy", use_dictionary=["s"], data_page_size=1 << 14),
                            dict(data_page_version="1.0", compression="none", write_page_index=True, row_group_size=7000)]):
        p = str(tmp_path / f"f{i}.parquet")
        pq.write_table(t, p, **kw)
        assert _walk(p) > 4


def test_reference_fixture_files():
    files = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "sample-parquet-files", "*.parquet")))
    assert files
    for p in files:
        assert _walk(p) > 0


def test_truncated_and_garbage_headers_are_errors():
    lib = load_library()
    out = np.zeros((16, 8), dtype=np.int32)
    junk = np.frombuffer(bytes([0x15, 0x00, 0x15, 0x90, 0x01]), dtype=np.uint8).copy()  # truncated PageHeader
    assert lib.lsb_test_walk_page_heads(junk.ctypes.data, 0, len(junk), out.ctypes.data, 16) == -1
    rnd = np.random.default_rng(1).integers(0, 256, 4096).astype(np.uint8)
    assert lib.lsb_test_walk_page_heads(rnd.ctypes.data, 0, len(rnd), out.ctypes.data, 16) < 0


def test_malformed_headers_are_rejected_not_followed(tmp_path):
    """truncated and bit-flipped page headers: the walker answers -1 / -2 (or a consistent walk), never reads outside the
    byte range it was given (the buffer ends exactly there: an over-read would fault under the guard page below)"""
    import mmap
    lib = load_library()
    t = pa.table({"a": pa.array(np.arange(5000, dtype=np.int64)), "s": pa.array(["x" * (i % 50) for i in range(5000)])})
    path = str(tmp_path / "f.parquet")
    pq.write_table(t, path, data_page_size=2 << 10, compression="ZSTD", use_dictionary=False, write_statistics=True)
    md = pq.ParquetFile(path).metadata
    cc = md.row_group(0).column(1)
    raw = np.fromfile(path, dtype=np.uint8)[cc.data_page_offset:cc.data_page_offset + cc.total_compressed_size].copy()
    page = mmap.PAGESIZE
    rng = np.random.default_rng(9)
    out = np.zeros((4096, 8), dtype=np.int32)
    for trial in range(300):
        n = int(rng.integers(1, len(raw) + 1)) if trial % 2 else len(raw)
        buf = raw[:n].copy()
        for _ in range(int(rng.integers(0, 4))):
            buf[int(rng.integers(0, min(n, 64)))] ^= np.uint8(1 << int(rng.integers(0, 8)))
        # place the bytes so that they END at a page boundary followed by an inaccessible page
        size = (n + page - 1) // page * page + page
        m = mmap.mmap(-1, size)
        start = size - page - n
        m[start:start + n] = buf.tobytes()
        addr = ctypes.addressof(ctypes.c_char.from_buffer(m))
        libc = ctypes.CDLL(None, use_errno=True)
        assert libc.mprotect(ctypes.c_void_p(addr + size - page), ctypes.c_size_t(page), 0) == 0
        r = lib.lsb_test_walk_page_heads(ctypes.c_void_p(addr), start, n, out.ctypes.data, 4096)
        assert r >= -2
        if trial % 2 == 0 and (buf == raw).all():
            assert r > 0
        libc.mprotect(ctypes.c_void_p(addr + size - page), ctypes.c_size_t(page), 3)
        del addr
        m.close()
