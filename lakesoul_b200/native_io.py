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

from ._lib import BOOL_CALLBACK, I32_CALLBACK, CStatus, LsbReaderStats, LsbWriterStats, load_library


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


class FlushResult:
    """NativeIOWriter.java FlushResult: one data file of a flush."""

    def __init__(self, filePath: str, fileSize: int, fileExistCols: str):
        self.filePath, self.fileSize, self.fileExistCols = filePath, fileSize, fileExistCols

    def __repr__(self):
        return f"FlushResult(filePath={self.filePath!r}, fileSize={self.fileSize}, fileExistCols={self.fileExistCols!r})"


class NativeIOWriter(NativeIOReader):
    """Mirror of NativeIOWriter.java (native-io/lakesoul-io-java/.../NativeIOWriter.java): the NativeIOBase setters are
    shared with the reader; write() hands a batch over through the Arrow C data interface, flush() decodes the
    summary of flush_and_close_writer into {partition_desc: [FlushResult]}."""

    def __init__(self, schema: pa.Schema):
        super().__init__()
        self.writer = None
        self.setSchema(schema)

    def setAuxSortColumns(self, cols: Iterable[str]):
        for c in cols:
            self.ioConfigBuilder = self.lib.lakesoul_config_builder_add_single_aux_sort_column(self.ioConfigBuilder, c.encode())

    def setHashBucketNum(self, n: int):
        self.ioConfigBuilder = self.lib.lakesoul_config_builder_set_hash_bucket_num(self.ioConfigBuilder, max(int(n), 1))

    def setRowGroupRowNumber(self, n: int):
        self.ioConfigBuilder = self.lib.lakesoul_config_builder_set_max_row_group_size(self.ioConfigBuilder, n)

    def setRowGroupValueNumber(self, n: int):
        self.ioConfigBuilder = self.lib.lakesoul_config_builder_set_max_row_group_num_values(self.ioConfigBuilder, n)

    def initializeWriter(self):
        runtime = self.lib.create_tokio_runtime_from_builder(self.tokioRuntimeBuilder)
        self.tokioRuntimeBuilder = None
        config = self.lib.create_lakesoul_io_config_from_builder(self.ioConfigBuilder)
        self.ioConfigBuilder = None
        w = self.lib.create_lakesoul_writer_from_config(config, runtime)
        status = self.lib.check_writer_created(w)
        try:
            if status.contents.status < 0:
                msg = (status.contents.err or b"unknown").decode()
                self.lib.free_lakesoul_writer(w)
                raise IOError("Init native writer failed with error: " + msg)
        finally:
            self.lib.free_c_status(status)
        self.writer = w

    def write(self, batch: pa.RecordBatch):
        arr, sch = _ArrowArrayStruct(), _ArrowSchemaStruct()
        batch._export_to_c(ctypes.addressof(arr), ctypes.addressof(sch))
        status = self.lib.write_record_batch_blocked(self.writer, ctypes.addressof(sch), ctypes.addressof(arr))
        try:
            if status.contents.status < 0:
                raise IOError("Native writer write batch failed with error: " + (status.contents.err or b"unknown").decode())
        finally:
            self.lib.free_c_status(status)

    def writeIpc(self, encoded: bytes) -> int:
        buf = ctypes.create_string_buffer(encoded, len(encoded))
        status = self.lib.write_record_batch_ipc_blocked(self.writer, ctypes.addressof(buf), len(encoded))
        try:
            if status.contents.status < 0:
                raise IOError("Native writer write batch failed with error: " + (status.contents.err or b"unknown").decode())
            return status.contents.status
        finally:
            self.lib.free_c_status(status)

    def writer_stats(self) -> dict:
        st = LsbWriterStats()
        if self.lib.lsb_writer_get_stats(self.writer, ctypes.byref(st)) != 0:
            raise RuntimeError(self.lib.lsb_last_error().decode())
        return {name: getattr(st, name) for name, _ in LsbWriterStats._fields_ if name != "pad"}

    def flush(self) -> Dict[str, List[FlushResult]]:
        """NativeIOWriter.java flush(): flush_and_close_writer + export_bytes_result, then the summary is decoded."""
        got = {}

        def on_len(n, err):
            got["len"], got["err"] = n, (err.decode() if err else None)

        cb = I32_CALLBACK(on_len)
        res = self.lib.flush_and_close_writer(self.writer, cb)
        self.writer = None
        if got.get("len", -1) < 0:
            self.lib.free_bytes_result(res)
            raise IOError("Flush native writer failed with error: " + str(got.get("err")))
        n = got["len"]
        out: Dict[str, List[FlushResult]] = {}
        if n == 0:
            self.lib.free_bytes_result(res)
            return out
        buf = ctypes.create_string_buffer(n + 1)
        ok = {}

        def on_done(status, err):
            ok["ok"], ok["err"] = status, (err.decode() if err else None)

        cb2 = BOOL_CALLBACK(on_done)
        self.lib.export_bytes_result(cb2, res, n, ctypes.addressof(buf))
        if not ok.get("ok"):
            raise IOError("Export flush result failed: " + str(ok.get("err")))
        text = buf.raw[:n].decode()
        parts = text.split("\x01")
        count = int(parts[0])
        assert count == len(parts) - 1
        for p in parts[1:]:
            fields = p.split("\x02")
            desc = fields[0]
            files = []
            for enc in fields[1:]:
                f = enc.split("\x03")
                assert len(f) == 3
                files.append(FlushResult(f[0], int(f[1]), f[2]))
            out[desc] = files
        return out

    def abort(self):
        if self.writer is None:
            return
        got = {}

        def on_done(status, err):
            got["ok"], got["err"] = status, (err.decode() if err else None)

        cb = BOOL_CALLBACK(on_done)
        self.lib.abort_and_close_writer(self.writer, cb)
        self.writer = None
        if not got.get("ok"):
            raise IOError("Abort native writer failed with error: " + str(got.get("err")))

    def close(self):
        if self.writer is not None:
            self.abort()
        super().close()
