# SPDX-License-Identifier: Apache-2.0
"""Many hash buckets behind ONE reader (option b200_file_groups, VERDICT r1 item 7 / SURVEY.md 7.2 "tiny units at
scale"): the file list carries several buckets back to back; every bucket is merged on its own in one device scan and the
stream returns the buckets one after the other -- exactly the concatenation of per-bucket readers, with one launch set."""
import pyarrow as pa
import pytest

import mor_oracle as mo
from conftest import assert_tables_equal

pytestmark = pytest.mark.gpu


def _table(cfg, tmp_path, scale, buckets, encoding):
    from lakesoul_b200 import datagen
    spec = datagen.config_spec(cfg, scale)
    spec.buckets = buckets
    return spec, datagen.write_table(str(tmp_path), spec, encoding=encoding)


@pytest.mark.parametrize("cfg,encoding,filters", [(5, "E3", []), (2, "E2", []), (4, "E1", []),
                                                    (3, "E3", ['not(eq(rowKinds, Binary{"delete"}))'])])
def test_grouped_reader_equals_per_bucket_readers(cfg, encoding, filters, tmp_path):
    from lakesoul_b200 import read_bucket
    nb = 5
    spec, paths = _table(cfg, tmp_path, 0.02, nb, encoding)
    files = [p for b in range(nb) for p in paths[b]]
    opts = dict(spec.options)
    opts["b200_file_groups"] = ",".join(str(len(paths[b])) for b in range(nb))
    got, rd = read_bucket(files, spec.schema, spec.primary_keys, batch_size=1024, options=opts, filters=filters, return_reader=True)
    grouped = rd.stats()
    rd.close()
    singles = []
    launches = 0
    for b in range(nb):
        t, r1 = read_bucket(paths[b], spec.schema, spec.primary_keys, batch_size=1024, options=spec.options, filters=filters,
                            return_reader=True)
        launches += r1.stats()["kernel_launches"]
        r1.close()
        want = mo.read_bucket(paths[b], spec.schema, spec.primary_keys, options=spec.options, filters=filters)
        assert_tables_equal(t, want, f"bucket {b}")
        singles.append(t)
    assert_tables_equal(got, pa.concat_tables(singles), f"config {cfg} {encoding}: {nb} buckets behind one reader")
    assert grouped["rows_out"] == sum(t.num_rows for t in singles)
    assert grouped["kernel_launches"] < launches / 2, (grouped["kernel_launches"], launches)  # one launch set, not five


def test_bad_group_list_is_an_error(tmp_path):
    from lakesoul_b200 import read_bucket
    spec, paths = _table(5, tmp_path, 0.01, 2, "E1")
    files = paths[0] + paths[1]
    with pytest.raises(IOError, match="b200_file_groups"):
        read_bucket(files, spec.schema, spec.primary_keys, options={"b200_file_groups": "3,1"})
