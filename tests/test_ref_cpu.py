# SPDX-License-Identifier: Apache-2.0
"""The CPU baseline (oracle/ref_cpu, the `--impl reference` arm of bench.py) must itself be right: its merged
output (verify mode) is compared with the pinned oracle on config-2 and config-5 shaped buckets. A baseline that
dropped or duplicated rows would inflate every reported ratio."""
import os

import pyarrow as pa
import pyarrow.ipc
import pyarrow.parquet  # noqa: F401
import pytest

import mor_oracle as mo
import ref_cpu_driver
from conftest import assert_tables_equal

pytestmark = pytest.mark.skipif(not os.path.exists(ref_cpu_driver.REF_CPU), reason="oracle/_build/ref_cpu not built")


@pytest.mark.parametrize("cfg,scale,encoding", [(2, 0.25, "E1"), (2, 0.1, "E3"), (5, 0.2, "E2"), (1, 0.1, "E1")])
def test_ref_cpu_output_matches_oracle(cfg, scale, encoding, tmp_path):
    from lakesoul_b200 import datagen
    spec = datagen.config_spec(cfg, scale)
    spec.buckets = min(spec.buckets, 3)
    paths = datagen.write_table(str(tmp_path / "t"), spec, encoding=encoding)
    out = tmp_path / "verify"
    out.mkdir()
    res = ref_cpu_driver.run(paths, spec.schema, spec.primary_keys, batch_size=1000, threads=2, verify_dir=str(out))
    assert res["kind"] == "port" and "C++ port" in res["impl"]
    rows_in = rows_out = 0
    for b, files in enumerate(paths):
        want = mo.read_bucket(files, spec.schema, spec.primary_keys)
        got = pa.ipc.open_file(str(out / f"bucket_{b}.arrow")).read_all()
        assert_tables_equal(got, want, f"ref_cpu config {cfg} bucket {b}")
        assert all(batch.num_rows <= 1000 for batch in got.to_batches())
        rows_out += want.num_rows
        rows_in += sum(pa.parquet.read_metadata(f).num_rows for f in files)
    assert res["rows_in"] == rows_in and res["rows_out"] == rows_out
