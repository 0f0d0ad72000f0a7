# SPDX-License-Identifier: Apache-2.0
"""The from-scratch ZSTD frame decoder the GPU runs (lakesoul_b200/csrc/zstd_decode.h), compiled for the host with
sanitizers and checked against frames produced by Arrow's ZSTD codec (levels 1/3/9/19; text, random, structured,
constant and Parquet-page-like inputs), plus mutated frames that must never cause an out-of-range access."""
import os
import struct
import subprocess

import numpy as np
import pyarrow as pa
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cases():
    rng = np.random.default_rng(1)
    words = [b"alpha", b"beta", b"gamma", b"delta", b"the quick brown fox", b"lakesoul", b"0123456789", b" "]

    def text(n):
        parts, tot = [], 0
        while tot < n:
            w = words[rng.integers(0, len(words))]
            parts.append(w)
            tot += len(w)
        return b"".join(parts)[:n]

    datas = [b"", b"a", b"abc" * 5, bytes(1000), text(100), text(5000), text(200_000), text(700_000),
             rng.integers(0, 256, 50_000, dtype=np.uint8).tobytes(), rng.integers(0, 4, 300_000, dtype=np.uint8).tobytes(),
             np.arange(100_000, dtype=np.int64).tobytes(), rng.integers(0, 1000, 150_000).astype(np.int32).tobytes(),
             (rng.standard_normal(60_000) * 100).astype(np.float64).tobytes(), b"x" * 500_000,
             b"".join(struct.pack("<I", len(w)) + w for w in [text(int(rng.integers(0, 44))) for _ in range(30000)])]
    for d in datas:
        for lvl in (1, 3, 9, 19):
            yield d, pa.Codec("zstd", compression_level=lvl).compress(d, asbytes=True)


def test_zstd_decoder_matches_arrow_codec(tmp_path):
    exe = str(tmp_path / "zstd_check")
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=all",
           "-I", os.path.join(ROOT, "lakesoul_b200", "csrc"), "-o", exe, os.path.join(ROOT, "tests", "native", "zstd_check.cc")]
    subprocess.run(cmd, check=True)
    cases = list(_cases())
    path = str(tmp_path / "cases.bin")
    with open(path, "wb") as f:
        f.write(struct.pack("<I", len(cases)))
        for d, c in cases:
            f.write(struct.pack("<QQ", len(d), len(c)))
            f.write(d)
            f.write(c)
    out = subprocess.run([exe, path], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert " 0 wrong" in out.stdout, out.stdout
