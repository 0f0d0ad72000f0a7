# SPDX-License-Identifier: Apache-2.0
"""Partial file reads (page_plan.cc OpenLazy / LoadKept): a reader that streams a bucket by key range opens footers only
and loads the byte span of the row groups a pass keeps. The framing of that span must describe exactly the pages -- same
order, same sizes, same payload bytes -- that the framing of the whole file image describes for those row groups."""
import ctypes
import os

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq
import pytest

from lakesoul_b200 import load_library  # noqa: E402


def _digest(path, lazy, lo, hi):
    lib = load_library()
    buf = ctypes.create_string_buffer(1 << 22)
    n = lib.lsb_test_plan_digest(path.encode(), lazy, lo, hi, buf, len(buf))
    assert n >= 0, buf.value.decode()
    return buf.value.decode()


def _table(n, seed=0):
    rng = np.random.default_rng(seed)
    return pa.table({
        "k": pa.array(np.arange(n, dtype=np.int64)),
        "v": pa.array(rng.integers(0, 50, n).astype(np.int32)),
        "s": pa.array([("row%05d" % i) * (1 + i % 3) if i % 11 else None for i in range(n)], pa.string()),
        "f": pa.array(rng.random(n)),
    })


@pytest.mark.parametrize("layout", ["plain", "zstd", "snappy_dict", "v2"])
def test_span_of_kept_row_groups_frames_like_the_whole_image(tmp_path, layout):
    kw = dict(plain=dict(compression="NONE", use_dictionary=False),
              zstd=dict(compression="ZSTD", use_dictionary=False),
              snappy_dict=dict(compression="SNAPPY", use_dictionary=True),
              v2=dict(compression="ZSTD", use_dictionary=True, data_page_version="2.0"))[layout]
    path = str(tmp_path / f"{layout}.parquet")
    pq.write_table(_table(20000), path, row_group_size=3000, data_page_size=8 << 10, **kw)
    nrg = pq.ParquetFile(path).metadata.num_row_groups
    assert nrg == 7
    for lo, hi in [(0, nrg), (0, 1), (2, 5), (nrg - 1, nrg), (3, 3)]:
        eager = _digest(path, 0, lo, hi).splitlines()
        lazy = _digest(path, 1, lo, hi).splitlines()
        assert eager[1:] == lazy[1:], (layout, lo, hi)
        assert len(eager) > 1 or lo == hi
        # the lazy image is the byte span of the kept row groups, not the file
        rows, _, window, _, total = eager[0].split()[1:6]
        lrows, _, lwindow, _, ltotal = lazy[0].split()[1:6]
        assert rows == lrows and total == ltotal
        md = pq.ParquetFile(path).metadata
        want = sum(md.row_group(g).column(c).total_compressed_size for g in range(lo, hi) for c in range(md.num_columns))
        assert int(lwindow) == want, (layout, lo, hi)
        assert int(window) == int(total)


def test_reference_fixture_frames_lazily():
    here = os.path.dirname(os.path.abspath(__file__))
    import glob
    files = sorted(glob.glob(os.path.join(here, "golden", "sample-parquet-files", "*.parquet")))
    assert files
    for f in files:
        nrg = pq.ParquetFile(f).metadata.num_row_groups
        lib = load_library()
        buf = ctypes.create_string_buffer(1 << 22)
        if lib.lsb_test_plan_digest(f.encode(), 0, 0, nrg, buf, len(buf)) < 0:
            continue  # encodings the device decoders do not cover: host-decoded either way
        assert _digest(f, 0, 0, nrg).splitlines()[1:] == _digest(f, 1, 0, nrg).splitlines()[1:]
