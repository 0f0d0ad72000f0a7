# SPDX-License-Identifier: Apache-2.0
"""The host framing code (page_plan.cc + filter.cc) built with AddressSanitizer and UBSan and run -- whole-file and lazy,
unpruned and pruned by nine key predicates -- over six page layouts with odd page boundaries and null key pages: no
out-of-range access, no undefined behaviour, every hook call succeeds."""
import os
import shutil
import subprocess

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N = 30000

FILTERS = ["and(gteq(k, 12000), lt(k, 13500))", "eq(k, 20000)", "lt(k, 5)", "or(eq(k, null), gt(k, 29000))", "gt(k, 999999)",
           'and(gteq(k2, Binary{"id0012000"}), lt(k2, Binary{"id0012100"}))', "and(gteq(k, 6990), lteq(k, 7010))", "lt(v, 100)",
           "not(lt(k, 25000))"]
LAYOUTS = {"zstd": {}, "plain": dict(compression="NONE"), "snappy_dict": dict(compression="SNAPPY", use_dictionary=True),
           "v2": dict(data_page_version="2.0"), "noindex": dict(write_page_index=False), "zstd_dict": dict(use_dictionary=["v", "s"])}


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs the system g++ (libasan)")
def test_framing_under_sanitizers(tmp_path):
    import pyarrow
    pa_dir = os.path.dirname(pyarrow.__file__)
    csrc = os.path.join(ROOT, "lakesoul_b200", "csrc")
    exe = str(tmp_path / "framing_check")
    cmd = ["g++", "-std=c++20", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
           "-I", csrc, "-I", os.path.join(ROOT, "include"), "-I", os.path.join(pa_dir, "include"),
           os.path.join(ROOT, "tests", "native", "framing_check.cc"), os.path.join(csrc, "page_plan.cc"), os.path.join(csrc, "filter.cc"),
           "-o", exe, "-L", pa_dir, "-l:libarrow.so.2400", "-l:libparquet.so.2400", f"-Wl,-rpath,{pa_dir}"]
    build = subprocess.run(cmd, capture_output=True, text=True)
    if build.returncode != 0 and "asan" in build.stderr:
        pytest.skip("no libasan for this g++")
    assert build.returncode == 0, build.stderr[-2000:]
    rng = np.random.default_rng(7)
    t = pa.table({"k": pa.array([None] * 300 + list(range(N - 300)), pa.int64()),
                  "k2": pa.array(["id%07d" % i for i in range(N)], pa.string()),
                  "v": pa.array(rng.integers(0, 1 << 40, N).astype(np.int64)),
                  "s": pa.array([("x" * (i % 40)) if i % 7 else None for i in range(N)], pa.string()),
                  "b": pa.array(rng.integers(0, 2, N).astype(bool))})
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0")
    for name, kw in LAYOUTS.items():
        args = dict(row_group_size=7000, max_rows_per_page=111, write_batch_size=13, write_page_index=True, compression="ZSTD",
                    use_dictionary=False)
        args.update(kw)
        path = str(tmp_path / f"{name}.parquet")
        pq.write_table(t, path, **args)
        run = subprocess.run([exe, path, "k,k2"] + FILTERS, capture_output=True, text=True, env=env)
        assert run.returncode == 0 and "done bad=0" in run.stdout, (name, run.stdout[-500:], run.stderr[-1500:])
