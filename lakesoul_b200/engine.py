# SPDX-License-Identifier: Apache-2.0
"""ctypes driver for the device-level scan plan (include/lakesoul_b200_ext.h).

Inputs and outputs are torch CUDA tensors used purely as HBM allocations (data_ptr());
all kernels live in liblakesoul_io_c.so.  Used by bench.py (`value` leg: inputs resident in HBM)
and by the `-m gpu` kernel parity tests.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence

import numpy as np

from ._lib import load_library

PK_UNSIGNED, PK_SIGNED, PK_FLOAT = 0, 1, 2


def _ptr_array(ptrs: Sequence[Optional[int]]):
    arr = (ctypes.c_void_p * len(ptrs))()
    for i, p in enumerate(ptrs):
        arr[i] = p if p else None
    return arr


def _stream_ptr(stream) -> int:
    if stream is None:
        import torch
        return torch.cuda.current_stream().cuda_stream
    return stream.cuda_stream if hasattr(stream, "cuda_stream") else int(stream)


class ScanPlan:
    """nb buckets, bucket b = files [bucket_file_begin[b], bucket_file_begin[b+1]) in commit order."""

    def __init__(self, bucket_file_begin: Sequence[int], file_rows: Sequence[int], key_bits: int = 64,
                 batch_size: int = 8192, file_row_base: Optional[Sequence[int]] = None, row_span: int = 0,
                 stream=None):
        self.lib = load_library()
        self.nb = len(bucket_file_begin) - 1
        self._bfb = np.ascontiguousarray(bucket_file_begin, dtype=np.int32)
        self._rows = np.ascontiguousarray(file_rows, dtype=np.int64)
        self._base = None if file_row_base is None else np.ascontiguousarray(file_row_base, dtype=np.int64)
        self.batch_size = batch_size
        self.stream = stream
        self.h = self.lib.lsb_plan_create(
            self.nb, self._bfb.ctypes.data, self._rows.ctypes.data,
            None if self._base is None else self._base.ctypes.data, int(row_span), key_bits, batch_size,
            _stream_ptr(stream))
        if not self.h:
            raise RuntimeError(self.lib.lsb_last_error().decode())
        self.out_base = None
        self.batch_base = None

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(self.lib.lsb_last_error().decode())

    def merge(self, pk_tensors, pk_kinds: Sequence[int], pk_validity=None):
        n = len(pk_tensors)
        vals = _ptr_array([t.data_ptr() for t in pk_tensors])
        valid = _ptr_array([(v.data_ptr() if v is not None else None) for v in (pk_validity or [None] * n)])
        widths = (ctypes.c_int32 * n)(*[t.element_size() for t in pk_tensors])
        kinds = (ctypes.c_int32 * n)(*pk_kinds)
        self._check(self.lib.lsb_plan_merge(self.h, n, vals, valid, widths, kinds, _stream_ptr(self.stream)))

    def identity(self):
        self._check(self.lib.lsb_plan_identity(self.h, _stream_ptr(self.stream)))

    def sync_counts(self):
        ob = np.zeros(self.nb + 1, dtype=np.int64)
        bb = np.zeros(self.nb + 1, dtype=np.int64)
        self._check(self.lib.lsb_plan_sync_counts(self.h, ob.ctypes.data, bb.ctypes.data, _stream_ptr(self.stream)))
        self.out_base, self.batch_base = ob, bb
        return ob, bb

    def winners(self):
        """torch view (uint32 reinterpreted as int32 storage -> returned as int64 copy) of the winner list."""
        import torch
        n = int(self.out_base[-1])
        ptr = self.lib.lsb_plan_winners(self.h)
        if n == 0:
            return torch.empty(0, dtype=torch.int64, device="cuda")
        # wrap the device pointer without owning it
        class _Holder:
            pass
        h = _Holder()
        h.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (ptr, False), "version": 2}
        t = torch.as_tensor(h, device="cuda")
        return t.to(torch.int64) & 0xFFFFFFFF

    @property
    def valid_stride(self) -> int:
        return int(self.lib.lsb_plan_valid_stride(self.h))

    def stats(self) -> dict:
        s = np.zeros(8, dtype=np.int64)
        self.lib.lsb_plan_stats(self.h, s.ctypes.data)
        return dict(rows_in=int(s[0]), rows_out=int(s[1]), tiles_level0=int(s[2]), levels=int(s[3]),
                    log_s=int(s[4]), tile_cap=int(s[5]), launches=int(s[6]))

    def gather_fixed(self, srcs, dsts, widths: Sequence[int], src_valid=None, dst_valid=None):
        n = len(srcs)
        self._check(self.lib.lsb_plan_gather_fixed(
            self.h, n, _ptr_array([t.data_ptr() for t in srcs]),
            _ptr_array([(v.data_ptr() if v is not None else None) for v in (src_valid or [None] * n)]),
            _ptr_array([t.data_ptr() for t in dsts]),
            _ptr_array([(v.data_ptr() if v is not None else None) for v in (dst_valid or [None] * n)]),
            (ctypes.c_int32 * n)(*widths), _stream_ptr(self.stream)))

    def gather_var(self, src_offsets, src_data, src_data_base, dst_offsets, dst_data, batch_data_start,
                   src_valid=None, dst_valid=None):
        p = lambda t: t.data_ptr() if t is not None else None  # noqa: E731
        self._check(self.lib.lsb_plan_gather_var(self.h, p(src_offsets), p(src_data), p(src_data_base), p(src_valid),
                                                 p(dst_offsets), p(dst_data), p(dst_valid), p(batch_data_start),
                                                 _stream_ptr(self.stream)))

    def gather_var_multi(self, cols):
        """cols: list of dicts with src_offsets, src_data, src_data_base, dst_offsets, dst_data,
        batch_data_start and optional src_valid / dst_valid tensors."""
        n = len(cols)
        arr = lambda key: _ptr_array([(c.get(key).data_ptr() if c.get(key) is not None else None) for c in cols])  # noqa: E731
        self._check(self.lib.lsb_plan_gather_var_multi(
            self.h, n, arr("src_offsets"), arr("src_data"), arr("src_data_base"), arr("src_valid"), arr("dst_offsets"),
            arr("dst_data"), arr("dst_valid"), arr("batch_data_start"), _stream_ptr(self.stream)))

    def gather_all(self, srcs, dsts, widths, vcols, src_valid=None, dst_valid=None):
        """Fixed-width columns (as in gather_fixed) + string columns (as in gather_var_multi), 2 launches."""
        n = len(srcs)
        nv = len(vcols)
        arr = lambda key: _ptr_array([(c.get(key).data_ptr() if c.get(key) is not None else None) for c in vcols])  # noqa: E731
        self._check(self.lib.lsb_plan_gather_all(
            self.h, n, _ptr_array([t.data_ptr() for t in srcs]),
            _ptr_array([(v.data_ptr() if v is not None else None) for v in (src_valid or [None] * n)]),
            _ptr_array([t.data_ptr() for t in dsts]),
            _ptr_array([(v.data_ptr() if v is not None else None) for v in (dst_valid or [None] * n)]),
            (ctypes.c_int32 * n)(*widths), nv, arr("src_offsets"), arr("src_data"), arr("src_data_base"),
            arr("src_valid"), arr("dst_offsets"), arr("dst_data"), arr("dst_valid"), arr("batch_data_start"),
            _stream_ptr(self.stream)))

    def close(self):
        if self.h:
            self.lib.lsb_plan_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
