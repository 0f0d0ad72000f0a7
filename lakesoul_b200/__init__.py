# SPDX-License-Identifier: Apache-2.0
"""lakesoul_b200 -- B200-native merge-on-read engine behind LakeSoul's native-IO C-ABI.

Python here is host-side plumbing only: it loads ``lib/liblakesoul_io_c.so`` (hand-written
sm_100a CUDA + C++ host code) through ctypes and mirrors the reference's Java bridge
(``native-io/lakesoul-io-java``: ``NativeIOBase`` / ``NativeIOReader`` / ``LakeSoulArrowReader``)
so the parity tests read like the reference's own.  There is no Python or CPU implementation of
the merge in this package: without the compiled library and a GPU every reader fails loudly.
"""
from ._lib import load_library, library_path  # noqa: F401
from .native_io import NativeIOReader, LakeSoulArrowReader, open_bucket, read_bucket  # noqa: F401

__all__ = ["load_library", "library_path", "NativeIOReader", "LakeSoulArrowReader", "open_bucket", "read_bucket"]
