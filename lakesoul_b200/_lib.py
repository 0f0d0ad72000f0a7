# SPDX-License-Identifier: Apache-2.0
"""ctypes loader for liblakesoul_io_c.so (the drop-in C-ABI library)."""
from __future__ import annotations

import ctypes
import os
from ctypes import (POINTER, Structure, c_bool, c_char_p, c_double, c_int32, c_int64, c_size_t, c_ssize_t,
                    c_void_p)

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def library_path() -> str:
    return os.environ.get("LAKESOUL_B200_LIB", os.path.join(_HERE, "lib", "liblakesoul_io_c.so"))


class CResult(Structure):  # rust/lakesoul-io-c/src/lib.rs:40-44
    _fields_ = [("ptr", c_void_p), ("err", c_char_p)]


class CStatus(Structure):  # lib.rs:77-81 ; Java mirror LibLakeSoulIO.java:86-94
    _fields_ = [("err", c_char_p), ("status", c_int32)]


class LsbReaderStats(Structure):
    _fields_ = [("rows_in", c_int64), ("rows_out", c_int64), ("h2d_bytes", c_int64), ("d2h_bytes", c_int64),
                ("kernel_launches", c_int32), ("stream_passes", c_int32), ("decode_ms", c_double), ("upload_ms", c_double),
                ("merge_ms", c_double), ("gather_ms", c_double), ("download_ms", c_double),
                ("device_decoded_pages", c_int64), ("host_decoded_files", c_int64), ("pruned_row_groups", c_int64),
                ("device_decompressed_pages", c_int64), ("pruned_buckets", c_int64), ("pruned_pages", c_int64)]


class LsbWriterStats(Structure):
    _fields_ = [("rows_in", c_int64), ("rows_out", c_int64), ("files", c_int64), ("flush_cycles", c_int64),
                ("h2d_bytes", c_int64), ("d2h_bytes", c_int64), ("kernel_launches", c_int32), ("pad", c_int32),
                ("sort_ms", c_double), ("encode_ms", c_double)]


I32_CALLBACK = ctypes.CFUNCTYPE(None, c_int32, c_char_p)
BOOL_CALLBACK = ctypes.CFUNCTYPE(None, c_bool, c_char_p)

# name -> (restype, argtypes): exactly the symbols include/lakesoul_c_bindings.h declares
REFERENCE_SYMBOLS = {
    "new_lakesoul_io_config_builder": (c_void_p, []),
    "lakesoul_config_builder_with_prefix": (c_void_p, [c_void_p, c_char_p]),
    "lakesoul_config_builder_add_single_file": (c_void_p, [c_void_p, c_char_p]),
    "lakesoul_config_builder_add_single_column": (c_void_p, [c_void_p, c_char_p]),
    "lakesoul_config_builder_add_single_aux_sort_column": (c_void_p, [c_void_p, c_char_p]),
    "lakesoul_config_builder_add_filter": (c_void_p, [c_void_p, c_char_p]),
    "lakesoul_config_builder_add_filter_proto": (c_void_p, [c_void_p, c_ssize_t, c_int32]),
    "lakesoul_config_builder_set_schema": (c_void_p, [c_void_p, c_ssize_t]),
    "lakesoul_config_builder_set_partition_schema": (c_void_p, [c_void_p, c_ssize_t]),
    "lakesoul_config_builder_set_thread_num": (c_void_p, [c_void_p, c_size_t]),
    "lakesoul_config_builder_set_dynamic_partition": (c_void_p, [c_void_p, c_bool]),
    "lakesoul_config_builder_set_inferring_schema": (c_void_p, [c_void_p, c_bool]),
    "lakesoul_config_builder_set_batch_size": (c_void_p, [c_void_p, c_size_t]),
    "lakesoul_config_builder_set_max_row_group_size": (c_void_p, [c_void_p, c_size_t]),
    "lakesoul_config_builder_set_max_row_group_num_values": (c_void_p, [c_void_p, c_size_t]),
    "lakesoul_config_builder_set_buffer_size": (c_void_p, [c_void_p, c_size_t]),
    "lakesoul_config_builder_set_hash_bucket_num": (c_void_p, [c_void_p, c_size_t]),
    "lakesoul_config_builder_set_object_store_option": (c_void_p, [c_void_p, c_char_p, c_char_p]),
    "lakesoul_config_builder_set_option": (c_void_p, [c_void_p, c_char_p, c_char_p]),
    "lakesoul_config_builder_add_files": (c_void_p, [c_void_p, POINTER(c_char_p), c_size_t]),
    "lakesoul_config_builder_add_single_primary_key": (c_void_p, [c_void_p, c_char_p]),
    "lakesoul_config_builder_add_single_range_partition": (c_void_p, [c_void_p, c_char_p]),
    "lakesoul_config_builder_add_merge_op": (c_void_p, [c_void_p, c_char_p, c_char_p]),
    "lakesoul_config_builder_add_primary_keys": (c_void_p, [c_void_p, POINTER(c_char_p), c_size_t]),
    "lakesoul_config_builder_set_default_column_value": (c_void_p, [c_void_p, c_char_p, c_char_p]),
    "create_lakesoul_io_config_from_builder": (c_void_p, [c_void_p]),
    "new_tokio_runtime_builder": (c_void_p, []),
    "tokio_runtime_builder_set_thread_num": (c_void_p, [c_void_p, c_size_t]),
    "create_tokio_runtime_from_builder": (c_void_p, [c_void_p]),
    "create_lakesoul_reader_from_config": (POINTER(CResult), [c_void_p, c_void_p]),
    "create_lakesoul_reader_from_config_with_global_runtime": (POINTER(CResult), [c_void_p]),
    "check_reader_created": (c_char_p, [POINTER(CResult)]),
    "start_reader": (POINTER(CStatus), [POINTER(CResult)]),
    "start_reader_with_data": (None, [POINTER(CResult), c_void_p, c_void_p]),
    "next_record_batch": (None, [POINTER(CResult), c_ssize_t, c_ssize_t, I32_CALLBACK]),
    "next_record_batch_blocked": (POINTER(CStatus), [POINTER(CResult), c_ssize_t]),
    "next_record_batch_with_data": (None, [POINTER(CResult), c_ssize_t, c_ssize_t, c_void_p, c_void_p]),
    "lakesoul_reader_get_schema": (None, [POINTER(CResult), c_ssize_t]),
    "free_lakesoul_reader": (None, [POINTER(CResult)]),
    "free_lakesoul_writer": (None, [c_void_p]),
    "free_lakesoul_io_config_builder": (None, [c_void_p]),
    "free_lakesoul_io_config": (None, [c_void_p]),
    "free_tokio_runtime_builder": (None, [c_void_p]),
    "free_tokio_runtime": (None, [c_void_p]),
    "free_c_status": (None, [POINTER(CStatus)]),
    "free_bytes_result": (None, [c_void_p]),
    "create_lakesoul_writer_from_config": (POINTER(CResult), [c_void_p, c_void_p]),
    "check_writer_created": (POINTER(CStatus), [POINTER(CResult)]),
    "write_record_batch": (None, [c_void_p, c_ssize_t, c_ssize_t, BOOL_CALLBACK]),
    "write_record_batch_blocked": (POINTER(CStatus), [c_void_p, c_ssize_t, c_ssize_t]),
    "write_record_batch_ipc_blocked": (POINTER(CStatus), [c_void_p, c_ssize_t, c_int64]),
    "export_bytes_result": (None, [BOOL_CALLBACK, c_void_p, c_int32, c_ssize_t]),
    "flush_and_close_writer": (c_void_p, [c_void_p, I32_CALLBACK]),
    "abort_and_close_writer": (None, [c_void_p, BOOL_CALLBACK]),
    "apply_partition_filter": (c_void_p, [I32_CALLBACK, c_int32, c_ssize_t, c_ssize_t, c_int32, c_ssize_t]),
    "rust_logger_init": (None, []),
}

EXT_SYMBOLS = {
    "lsb_last_error": (c_char_p, []),
    "lsb_init_device": (c_int32, []),
    "lsb_version": (c_char_p, []),
    "lsb_reader_get_stats": (c_int32, [POINTER(CResult), POINTER(LsbReaderStats)]),
    "lsb_writer_get_stats": (c_int32, [c_void_p, POINTER(LsbWriterStats)]),
    "lsb_test_walk_page_heads": (c_int64, [c_void_p, c_int64, c_int64, c_void_p, c_int64]),
    "lsb_test_plan_digest": (c_int64, [c_char_p, c_int32, c_int32, c_int32, c_void_p, c_int64]),
    "lsb_test_page_prune": (c_int64, [c_char_p, c_char_p, c_char_p, c_int32, c_void_p, c_int64]),
    "lsb_test_stream_passes": (c_int64, [c_char_p, c_char_p, c_char_p, c_int64, c_void_p, c_int64]),
    "lsb_plan_create": (c_void_p, [c_int32, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p]),
    "lsb_plan_destroy": (None, [c_void_p]),
    "lsb_plan_merge": (c_int32, [c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "lsb_plan_identity": (c_int32, [c_void_p, c_void_p]),
    "lsb_plan_sync_counts": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "lsb_plan_winners": (c_void_p, [c_void_p]),
    "lsb_plan_valid_stride": (c_int64, [c_void_p]),
    "lsb_plan_stats": (c_int32, [c_void_p, c_void_p]),
    "lsb_plan_gather_fixed": (c_int32, [c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "lsb_plan_gather_var": (c_int32, [c_void_p] * 10),
    "lsb_plan_gather_var_multi": (c_int32, [c_void_p, c_int32] + [c_void_p] * 9),
    "lsb_plan_gather_all": (c_int32, [c_void_p, c_int32] + [c_void_p] * 5 + [c_int32] + [c_void_p] * 9),
}


def load_library() -> ctypes.CDLL:
    """Load the C-ABI library once; raises (never falls back) when it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no Python/CPU fallback for the merge-on-read engine)")
    lib = ctypes.CDLL(path)
    for table in (REFERENCE_SYMBOLS, EXT_SYMBOLS):
        for name, (res, args) in table.items():
            fn = getattr(lib, name)  # AttributeError here == a symbol include/*.h declares is missing
            fn.restype = res
            fn.argtypes = args
    _LIB = lib
    return lib
