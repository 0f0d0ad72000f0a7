/* SPDX-License-Identifier: Apache-2.0
 *
 * lakesoul_c_bindings.h -- the C-ABI of liblakesoul_io_c.so, re-implemented by the B200
 * merge-on-read engine. Every symbol keeps the name, argument list, ownership rules and error
 * convention of the reference's cdylib `lakesoul-io-c` (cbindgen header of
 * rust/lakesoul-io-c/src/lib.rs; the line after each prototype is the reference definition it
 * replaces). The JVM side (native-io/lakesoul-io-java/.../LibLakeSoulIO.java:12-134, jnr-ffi)
 * binds these by name and needs no change.
 *
 * Ownership (reference lib.rs:107-130):
 *   - every lakesoul_config_builder_* call CONSUMES the builder pointer and returns a new one;
 *   - create_lakesoul_io_config_from_builder consumes the builder;
 *   - create_lakesoul_reader_from_config consumes the config AND the runtime;
 *   - CStatus* results are heap allocated: release with free_c_status (frees `err` too);
 *   - C strings passed in are borrowed, UTF-8, NUL terminated;
 *   - lakesoul_config_builder_set_schema takes ownership of the ArrowSchema (calls release);
 *   - next_record_batch_blocked moves a populated ArrowArray (struct array, one child per column)
 *     into caller memory; the consumer calls its release callback.
 * Status convention (lib.rs:883-909): status < 0 error (err set), 0 end of stream, > 0 row count.
 */
#ifndef LAKESOUL_C_BINDINGS_H
#define LAKESOUL_C_BINDINGS_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef size_t c_size_t;
typedef ptrdiff_t c_ptrdiff_t;

/* opaque handles (lib.rs:132-160, 1371-1381) */
typedef struct IOConfigBuilder IOConfigBuilder;
typedef struct IOConfig IOConfig;
typedef struct Reader Reader;
typedef struct Writer Writer;
typedef struct BytesResult BytesResult;
typedef struct TokioRuntimeBuilder TokioRuntimeBuilder;
typedef struct TokioRuntime TokioRuntime;

/* lib.rs:40-44  CResult<T> { ptr, err } */
typedef struct CResult_Reader { Reader *ptr; const char *err; } CResult_Reader;
typedef struct CResult_Writer { Writer *ptr; const char *err; } CResult_Writer;
typedef struct CResult_BytesResult { BytesResult *ptr; const char *err; } CResult_BytesResult;

/* lib.rs:77-81 */
typedef struct CStatus { const char *err; int32_t status; } CStatus;

typedef void (*ResultCallback)(bool, const char *);                          /* lib.rs:694 */
typedef void (*DataResultCallback)(bool, const char *, const void *);        /* lib.rs:696 */
typedef void (*I32ResultCallback)(int32_t, const char *);                    /* lib.rs:726 */
typedef void (*I32DataResultCallback)(int32_t, const char *, const void *);  /* lib.rs:728 */

/* ---- config builder ------------------------------------------------------------------------- */
IOConfigBuilder *new_lakesoul_io_config_builder(void);                                              /* lib.rs:164 */
IOConfigBuilder *lakesoul_config_builder_with_prefix(IOConfigBuilder *builder, const char *prefix); /* lib.rs:175 */
IOConfigBuilder *lakesoul_config_builder_add_single_file(IOConfigBuilder *builder, const char *file); /* lib.rs:193 */
IOConfigBuilder *lakesoul_config_builder_add_single_column(IOConfigBuilder *builder, const char *column); /* lib.rs:211 */
IOConfigBuilder *lakesoul_config_builder_add_single_aux_sort_column(IOConfigBuilder *builder, const char *column); /* lib.rs:229 */
IOConfigBuilder *lakesoul_config_builder_add_filter(IOConfigBuilder *builder, const char *filter);  /* lib.rs:247 */
IOConfigBuilder *lakesoul_config_builder_add_filter_proto(IOConfigBuilder *builder, c_ptrdiff_t proto_addr, int32_t len); /* lib.rs:267 */
IOConfigBuilder *lakesoul_config_builder_set_schema(IOConfigBuilder *builder, c_ptrdiff_t schema_addr); /* lib.rs:290 */
IOConfigBuilder *lakesoul_config_builder_set_partition_schema(IOConfigBuilder *builder, c_ptrdiff_t schema_addr); /* lib.rs:311 */
IOConfigBuilder *lakesoul_config_builder_set_thread_num(IOConfigBuilder *builder, c_size_t thread_num); /* lib.rs:331 */
IOConfigBuilder *lakesoul_config_builder_set_dynamic_partition(IOConfigBuilder *builder, bool enable); /* lib.rs:347 */
IOConfigBuilder *lakesoul_config_builder_set_inferring_schema(IOConfigBuilder *builder, bool enable); /* lib.rs:363 */
IOConfigBuilder *lakesoul_config_builder_set_batch_size(IOConfigBuilder *builder, c_size_t batch_size); /* lib.rs:380 */
IOConfigBuilder *lakesoul_config_builder_set_max_row_group_size(IOConfigBuilder *builder, c_size_t max_row_group_size); /* lib.rs:396 */
IOConfigBuilder *lakesoul_config_builder_set_max_row_group_num_values(IOConfigBuilder *builder, c_size_t max_row_group_num_values); /* lib.rs:412 */
IOConfigBuilder *lakesoul_config_builder_set_buffer_size(IOConfigBuilder *builder, c_size_t buffer_size); /* lib.rs:428 */
IOConfigBuilder *lakesoul_config_builder_set_hash_bucket_num(IOConfigBuilder *builder, c_size_t hash_bucket_num); /* lib.rs:444 */
IOConfigBuilder *lakesoul_config_builder_set_object_store_option(IOConfigBuilder *builder, const char *key, const char *value); /* lib.rs:461 */
IOConfigBuilder *lakesoul_config_builder_set_option(IOConfigBuilder *builder, const char *key, const char *value); /* lib.rs:482 */
IOConfigBuilder *lakesoul_config_builder_add_files(IOConfigBuilder *builder, const char *const *files, c_size_t file_num); /* lib.rs:504 */
IOConfigBuilder *lakesoul_config_builder_add_single_primary_key(IOConfigBuilder *builder, const char *pk); /* lib.rs:531 */
IOConfigBuilder *lakesoul_config_builder_add_single_range_partition(IOConfigBuilder *builder, const char *col); /* lib.rs:551 */
IOConfigBuilder *lakesoul_config_builder_add_merge_op(IOConfigBuilder *builder, const char *field, const char *merge_op); /* lib.rs:571 */
IOConfigBuilder *lakesoul_config_builder_add_primary_keys(IOConfigBuilder *builder, const char *const *pks, c_size_t pk_num); /* lib.rs:593 */
IOConfigBuilder *lakesoul_config_builder_set_default_column_value(IOConfigBuilder *builder, const char *field, const char *value); /* lib.rs:620 */
IOConfig *create_lakesoul_io_config_from_builder(IOConfigBuilder *builder);                         /* lib.rs:637 */

/* ---- runtime (kept for ABI compatibility; the GPU engine needs no tokio runtime) ------------ */
TokioRuntimeBuilder *new_tokio_runtime_builder(void);                                               /* lib.rs:1385 */
TokioRuntimeBuilder *tokio_runtime_builder_set_thread_num(TokioRuntimeBuilder *builder, c_size_t thread_num); /* lib.rs:1400 */
TokioRuntime *create_tokio_runtime_from_builder(TokioRuntimeBuilder *builder);                      /* lib.rs:1415 */

/* ---- reader --------------------------------------------------------------------------------- */
CResult_Reader *create_lakesoul_reader_from_config(IOConfig *config, TokioRuntime *runtime);       /* lib.rs:648 */
CResult_Reader *create_lakesoul_reader_from_config_with_global_runtime(IOConfig *config);          /* lib.rs:666 */
const char *check_reader_created(CResult_Reader *reader);                                          /* lib.rs:681 */
CStatus *start_reader(CResult_Reader *reader);                                                     /* lib.rs:768 */
void start_reader_with_data(CResult_Reader *reader, const void *data, DataResultCallback callback); /* lib.rs:791 */
void next_record_batch(CResult_Reader *reader, c_ptrdiff_t schema_addr, c_ptrdiff_t array_addr, I32ResultCallback callback); /* lib.rs:823 */
CStatus *next_record_batch_blocked(CResult_Reader *reader, c_ptrdiff_t array_addr);                /* lib.rs:883 */
void next_record_batch_with_data(CResult_Reader *reader, c_ptrdiff_t schema_addr, c_ptrdiff_t array_addr, const void *data, I32DataResultCallback callback); /* lib.rs:931 */
void lakesoul_reader_get_schema(CResult_Reader *reader, c_ptrdiff_t schema_addr);                  /* lib.rs:999 */

/* ---- free ----------------------------------------------------------------------------------- */
void free_lakesoul_reader(CResult_Reader *reader);                                                  /* lib.rs:1020 */
void free_lakesoul_writer(CResult_Writer *writer);                                                  /* lib.rs:1028 */
void free_lakesoul_io_config_builder(IOConfigBuilder *builder);                                     /* lib.rs:1036 */
void free_lakesoul_io_config(IOConfig *io_config);                                                  /* lib.rs:1043 */
void free_tokio_runtime_builder(TokioRuntimeBuilder *builder);                                      /* lib.rs:1049 */
void free_tokio_runtime(TokioRuntime *runtime);                                                     /* lib.rs:1060 */
void free_c_status(CStatus *status);                                                                /* lib.rs:1070 */
void free_bytes_result(CResult_BytesResult *bytes);                                                 /* lib.rs:1477 */

/* ---- writer side: murmur3 bucketing + sort on the GPU, Parquet (ZSTD, no dictionaries) encoded on
 *      the host; same writers, file names and flush summary as writer/mod.rs:61-400 ------------------------------- */
CResult_Writer *create_lakesoul_writer_from_config(IOConfig *io_config, TokioRuntime *runtime);    /* lib.rs:1081 */
CStatus *check_writer_created(CResult_Writer *writer);                                              /* lib.rs:1101 */
void write_record_batch(CResult_Writer *writer, c_ptrdiff_t schema_addr, c_ptrdiff_t array_addr, ResultCallback callback); /* lib.rs:1125 */
CStatus *write_record_batch_blocked(CResult_Writer *writer, c_ptrdiff_t schema_addr, c_ptrdiff_t array_addr); /* lib.rs:1170 */
CStatus *write_record_batch_ipc_blocked(CResult_Writer *writer, c_ptrdiff_t ipc_addr, int64_t len); /* lib.rs:1211 */
void export_bytes_result(void (*callback)(bool, const char *), CResult_BytesResult *bytes, int32_t len, c_ptrdiff_t addr); /* lib.rs:1262 */
CResult_BytesResult *flush_and_close_writer(CResult_Writer *writer, I32ResultCallback callback);   /* lib.rs:1309 */
void abort_and_close_writer(CResult_Writer *writer, ResultCallback callback);                       /* lib.rs:1347 */
CResult_BytesResult *apply_partition_filter(void (*callback)(int32_t, const char *), int32_t len, c_ptrdiff_t jni_wrapper_addr, c_ptrdiff_t schema_addr, int32_t filter_len, c_ptrdiff_t filter_addr); /* lib.rs:1434 */

void rust_logger_init(void);                                                                        /* lib.rs:1484 */

#ifdef __cplusplus
}
#endif
#endif /* LAKESOUL_C_BINDINGS_H */
