# SPDX-License-Identifier: Apache-2.0
"""Full-size parity: one bucket of BASELINE configs 2..5 at the sizes SURVEY.md 8d fixes (234 375 / 65 536 /
200 000 / 146 484 input rows per bucket; 8 / 32 / 8 / 64 delta files) through the C-ABI against the CPU
oracle, in the uncompressed PLAIN layout (E1), the pyarrow/Spark default (E2: dictionary + SNAPPY) and the
layout the LakeSoul writer emits (E3: PLAIN + ZSTD, multipart_writer.rs:121-133)."""
import pyarrow as pa
import pyarrow.parquet  # noqa: F401
import pytest

import mor_oracle as mo
from conftest import assert_tables_equal

pytestmark = pytest.mark.gpu

FULL_ROWS = {2: 234_375, 3: 65_536, 4: 200_000, 5: 146_484}


def _full_bucket(cfg, tmp_path, encoding):
    from lakesoul_b200 import datagen
    spec = datagen.config_spec(cfg, 1.0)
    assert spec.rows_per_bucket == FULL_ROWS[cfg]
    spec.buckets = 1  # every key hashes to bucket 0: one bucket at the full per-bucket size
    paths = datagen.write_table(str(tmp_path), spec, encoding=encoding)[0]
    assert len(paths) == spec.commits
    return spec, paths


@pytest.mark.parametrize("encoding", ["E1", "E2", "E3"])
@pytest.mark.parametrize("cfg", [2, 3, 4, 5])
def test_full_size_bucket_matches_oracle(cfg, encoding, tmp_path):
    from lakesoul_b200 import read_bucket
    spec, paths = _full_bucket(cfg, tmp_path, encoding)
    got, reader = read_bucket(paths, spec.schema, spec.primary_keys, batch_size=8192, options=spec.options,
                              return_reader=True)
    stats = reader.stats()
    reader.close()
    want = mo.read_bucket(paths, spec.schema, spec.primary_keys, options=spec.options)
    assert_tables_equal(got, want, f"config {cfg} {encoding} full size")
    assert stats["rows_in"] == sum(pa.parquet.read_metadata(p).num_rows for p in paths)
    assert stats["rows_out"] == want.num_rows
    assert stats["kernel_launches"] > 0 and stats["device_decoded_pages"] > 0 and stats["host_decoded_files"] == 0, stats


def test_full_size_cdc_bucket_with_delete_filter(tmp_path):
    """config 3 at full size, deletes removed by the pushed CDC predicate (metadata_format.rs:310-320)."""
    from lakesoul_b200 import read_bucket
    spec, paths = _full_bucket(3, tmp_path, "E3")
    flt = ['not(eq(rowKinds, Binary{"delete"}))']
    got = read_bucket(paths, spec.schema, spec.primary_keys, batch_size=8192, options=spec.options, filters=flt)
    want = mo.read_bucket(paths, spec.schema, spec.primary_keys, options=spec.options, filters=flt)
    assert_tables_equal(got, want, "config 3 full size, post-merge CDC filter")
    assert 0 < got.num_rows
