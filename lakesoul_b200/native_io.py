# SPDX-License-Identifier: Apache-2.0
"""Python mirror of the reference's Java bridge over the lakesoul-io-c C-ABI.

``NativeIOReader`` replays, call for call, what ``NativeIOBase`` / ``NativeIOReader``
(native-io/lakesoul-io-java/src/main/java/com/dmetasoul/lakesoul/lakesoul/io/NativeIOBase.java:87-187,
NativeIOReader.java:29-129) do through jnr-ffi; ``LakeSoulArrowReader`` is the batch iterator of
native-io/lakesoul-io-java/src/main/scala/com/dmetasoul/lakesoul/LakeSoulArrowReader.scala:32-51
(rowCount == 0 means end of stream).  Everything below the ctypes calls is the CUDA engine.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Iterable, Iterator, List, Optional

import pyarrow as pa

from ._lib import CStatus, LsbReaderStats, load_library


class _ArrowSchemaStruct(ctypes.Structure):
    _fields_ = [("_", ctypes.c_byte * 72)]


class _ArrowArrayStruct(ctypes.Structure):
    _fields_ = [("_", ctypes.c_byte * 80)]


class NativeIOReader:
    """One reader per (range partition, hash bucket): files in commit order, later file wins."""

    DEFAULT_BATCH_SIZE = 10240  # NativeIOBase.java:63
    DEFAULT_THREAD_NUM = 2      # NativeIOBase.java:64

    def __init__(self):
        self.lib = load_library()
        self.ioConfigBuilder = self.lib.new_lakesoul_io_config_builder()
        self.tokioRuntimeBuilder = self.lib.new_tokio_runtime_builder()
        self.reader = None
        self.readerSchema: Optional[pa.Schema] = None
        self.setBatchSize(self.DEFAULT_BATCH_SIZE)
        self.setThreadNum(self.DEFAULT_THREAD_NUM)

    # ---- NativeIOBase ---------------------------------------------------------------------
    def addFile(self, file: str):
        self.ioConfigBuilder = self.lib.lakesoul_config_builder_add_single_file(self.ioConfigBuilder, file.encode())

    def withPrefix(self, prefix: str):
        self.ioConfigBuilder = self.lib.lakesoul_config_builder_with_prefix(self.ioConfigBuilder, prefix.encode())

    def addColumn(self, column: str):
        self.ioConfigBuilder = self.lib.lakesoul_config_builder_add_single_column(self.ioConfigBuilder, column.encode())

    def setPrimaryKeys(self, primaryKeys: Iterable[str]):
        for pk in primaryKeys:
            self.ioConfigBuilder = self.lib.lakesoul_config_builder_add_single_primary_key(self.ioConfigBuilder, pk.encode())

    def setRangePartitions(self, rangePartitions: Iterable[str]):
        for col in rangePartitions:
            self.ioConfigBuilder = self.lib.lakesoul_config_builder_add_single_range_partition(self.ioConfigBuilder, col.encode())

    def _export_schema(self, schema: pa.Schema) -> int:
        holder = _ArrowSchemaStruct()
        addr = ctypes.addressof(holder)
        schema._export_to_c(addr)
        self._keep = getattr(self, "_keep", []) + [holder]  # memory of the struct itself stays ours
        return addr

    def setSchema(self, schema: pa.Schema):
        self.ioConfigBuilder = self.lib.lakesoul_config_builder_set_schema(self.ioConfigBuilder, self._export_schema(schema))

    def setPartitionSchema(self, schema: pa.Schema):
        self.ioConfigBuilder = self.lib.lakesoul_config_builder_set_partition_schema(self.ioConfigBuilder, self._export_schema(schema))

    def setThreadNum(self, threadNum: int):
        self.ioConfigBuilder = self.lib.lakesoul_config_builder_set_thread_num(self.ioConfigBuilder, threadNum)

    def useDynamicPartition(self, enable: bool):
        self.ioConfigBuilder = self.lib.lakesoul_config_builder_set_dynamic_partition(self.ioConfigBuilder, enable)

    def setInferringSchema(self, enable: bool):
        self.ioConfigBuilder = self.lib.lakesoul_config_builder_set_inferring_schema(self.ioConfigBuilder, enable)

    def setBatchSize(self, batchSize: int):
        self.ioConfigBuilder = self.lib.lakesoul_config_builder_set_batch_size(self.ioConfigBuilder, batchSize)

    def setBufferSize(self, bufferSize: int):
        self.ioConfigBuilder = self.lib.lakesoul_config_builder_set_buffer_size(self.ioConfigBuilder, bufferSize)

    def setObjectStoreOption(self, key: str, value: str):
        self.ioConfigBuilder = self.lib.lakesoul_config_builder_set_object_store_option(self.ioConfigBuilder, key.encode(), value.encode())

    def setOption(self, key: str, value: str):
        self.ioConfigBuilder = self.lib.lakesoul_config_builder_set_option(self.ioConfigBuilder, key.encode(), value.encode())

    # ---- NativeIOReader -------------------------------------------------------------------
    def addFilter(self, filter: str):
        self.ioConfigBuilder = self.lib.lakesoul_config_builder_add_filter(self.ioConfigBuilder, filter.encode())

    def addFilterProto(self, plan_bytes: bytes):
        buf = ctypes.create_string_buffer(plan_bytes, len(plan_bytes))
        self.ioConfigBuilder = self.lib.lakesoul_config_builder_add_filter_proto(
            self.ioConfigBuilder, ctypes.addressof(buf), len(plan_bytes))

    def addMergeOps(self, mergeOps: Dict[str, str]):
        for field, op in mergeOps.items():
            self.ioConfigBuilder = self.lib.lakesoul_config_builder_add_merge_op(self.ioConfigBuilder, field.encode(), op.encode())

    def setDefaultColumnValue(self, column: str, value: str):
        self.ioConfigBuilder = self.lib.lakesoul_config_builder_set_default_column_value(self.ioConfigBuilder, column.encode(), value.encode())

    def initializeReader(self):
        """NativeIOReader.java:60-90."""
        assert self.tokioRuntimeBuilder is not None
        assert self.ioConfigBuilder is not None
        runtime = self.lib.create_tokio_runtime_from_builder(self.tokioRuntimeBuilder)
        self.tokioRuntimeBuilder = None
        config = self.lib.create_lakesoul_io_config_from_builder(self.ioConfigBuilder)
        self.ioConfigBuilder = None
        self.reader = self.lib.create_lakesoul_reader_from_config(config, runtime)
        err = self.lib.check_reader_created(self.reader)
        if err is not None:
            msg = err.decode()
            self.lib.free_lakesoul_reader(self.reader)
            self.reader = None
            raise IOError("Init native reader failed with error: " + msg)
        status = self.lib.start_reader(self.reader)
        try:
            if status.contents.status != 0:
                raise IOError("Init native reader failed with error: " + (status.contents.err or b"").decode())
        finally:
            self.lib.free_c_status(status)
        holder = _ArrowSchemaStruct()
        self.lib.lakesoul_reader_get_schema(self.reader, ctypes.addressof(holder))
        self.readerSchema = pa.Schema._import_from_c(ctypes.addressof(holder))

    def getSchema(self) -> pa.Schema:
        return self.readerSchema

    def nextBatchBlocked(self, arrayAddr: int) -> int:
        """NativeIOReader.java:118-129: > 0 rows, 0 end of stream, raises on error."""
        status = self.lib.next_record_batch_blocked(self.reader, arrayAddr)
        try:
            rows = status.contents.status
            if rows < 0:
                raise IOError("next batch failed with error: " + (status.contents.err or b"").decode())
            return rows
        finally:
            self.lib.free_c_status(status)

    def stats(self) -> dict:
        st = LsbReaderStats()
        if self.lib.lsb_reader_get_stats(self.reader, ctypes.byref(st)) != 0:
            raise RuntimeError(self.lib.lsb_last_error().decode())
        return {name: getattr(st, name) for name, _ in LsbReaderStats._fields_ if name != "pad"}

    def close(self):
        if self.reader is not None:
            self.lib.free_lakesoul_reader(self.reader)
            self.reader = None
        if self.tokioRuntimeBuilder is not None:
            self.lib.free_tokio_runtime_builder(self.tokioRuntimeBuilder)
            self.tokioRuntimeBuilder = None
        if self.ioConfigBuilder is not None:
            self.lib.free_lakesoul_io_config_builder(self.ioConfigBuilder)
            self.ioConfigBuilder = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class LakeSoulArrowReader:
    """Iterator of pyarrow RecordBatches (LakeSoulArrowReader.scala:32-51)."""

    def __init__(self, reader: NativeIOReader):
        self.reader = reader
        self.finished = False

    def __iter__(self) -> Iterator[pa.RecordBatch]:
        schema = self.reader.getSchema()
        while not self.finished:
            consumer = _ArrowArrayStruct()
            rows = self.reader.nextBatchBlocked(ctypes.addressof(consumer))
            if rows <= 0:
                self.finished = True
                break
            batch = pa.RecordBatch._import_from_c(ctypes.addressof(consumer), schema)
            assert batch.num_rows == rows
            yield batch

    def read_all(self) -> pa.Table:
        batches = list(self)
        return pa.Table.from_batches(batches, schema=self.reader.getSchema())

    def to_reader(self) -> pa.RecordBatchReader:
        """pyarrow RecordBatchReader over the native reader (what python/src/dataset.rs:32-70 hands to
        pyarrow.dataset through PyO3 in the reference); it also provides __arrow_c_stream__."""
        return pa.RecordBatchReader.from_batches(self.reader.getSchema(), iter(self))

    def __arrow_c_stream__(self, requested_schema=None):
        return self.to_reader().__arrow_c_stream__(requested_schema)

    def close(self):
        self.finished = True
        self.reader.close()


def open_bucket(files: List[str], schema: pa.Schema, primary_keys: Iterable[str] = (), *, batch_size: int = 8192,
                default_columns: Optional[Dict[str, str]] = None, merge_ops: Optional[Dict[str, str]] = None,
                options: Optional[Dict[str, str]] = None, filters: Iterable[str] = (),
                partition_schema: Optional[pa.Schema] = None) -> pa.RecordBatchReader:
    """One hash bucket as a streaming pyarrow.RecordBatchReader (Arrow C stream capable): the Python surface the
    reference exposes with `python/src/dataset.rs`. The native reader is freed when the stream ends."""
    r = _configure(files, schema, primary_keys, batch_size, default_columns, merge_ops, options, filters, partition_schema)
    r.initializeReader()
    ar = LakeSoulArrowReader(r)

    def batches():
        try:
            yield from ar
        finally:
            r.close()

    return pa.RecordBatchReader.from_batches(r.getSchema(), batches())


def _configure(files, schema, primary_keys, batch_size, default_columns, merge_ops, options, filters,
               partition_schema) -> NativeIOReader:
    """the builder calls of lakesoul-spark NativeVectorizedReader.java:224-269"""
    r = NativeIOReader()
    for f in files:
        r.addFile(f)
    r.setPrimaryKeys(primary_keys)
    r.setSchema(schema)
    if partition_schema is not None:
        r.setPartitionSchema(partition_schema)
    for k, v in (default_columns or {}).items():
        r.setDefaultColumnValue(k, v)
    r.setBatchSize(batch_size)
    for k, v in (options or {}).items():
        r.setOption(k, v)
    for flt in filters:
        r.addFilter(flt)
    if merge_ops:
        r.addMergeOps(merge_ops)
    return r


def read_bucket(files: List[str], schema: pa.Schema, primary_keys: Iterable[str] = (), *,
                batch_size: int = 8192, default_columns: Optional[Dict[str, str]] = None,
                merge_ops: Optional[Dict[str, str]] = None, options: Optional[Dict[str, str]] = None,
                filters: Iterable[str] = (), partition_schema: Optional[pa.Schema] = None,
                return_reader: bool = False):
    """Convenience wrapper: configure a NativeIOReader the way the connectors do
    (lakesoul-spark NativeVectorizedReader.java:224-269) and read the whole bucket."""
    r = _configure(files, schema, primary_keys, batch_size, default_columns, merge_ops, options, filters, partition_schema)
    r.initializeReader()
    ar = LakeSoulArrowReader(r)
    table = ar.read_all()
    if return_reader:
        return table, r
    r.close()
    return table
