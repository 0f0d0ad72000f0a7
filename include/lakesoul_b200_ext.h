/* SPDX-License-Identifier: Apache-2.0
 *
 * lakesoul_b200_ext.h -- extensions exported by liblakesoul_io_c.so next to the reference C-ABI
 * (include/lakesoul_c_bindings.h). Nothing here exists in the reference; these entry points expose
 *   (1) per-reader counters (bytes over PCIe, kernel launches) for bench.py's `e2e` leg, and
 *   (2) the device-level scan plan: the same CUDA engine the reader uses, driven with columns that
 *       are already resident in HBM (plain device pointers, no torch types) -- bench.py's `value`
 *       leg and the `-m gpu` kernel parity tests call it through ctypes.
 * All functions return 0 on success and -1 on error unless stated; lsb_last_error() gives the
 * message for the calling thread.
 */
#ifndef LAKESOUL_B200_EXT_H
#define LAKESOUL_B200_EXT_H

#include <stdint.h>

#include "lakesoul_c_bindings.h"

#ifdef __cplusplus
extern "C" {
#endif

const char *lsb_last_error(void);
/* selects the CUDA device for this process (LAKESOUL_B200_DEVICE, else LOCAL_RANK, else 0);
 * returns the device id or -1 when no GPU is usable (the engine has no CPU fallback) */
int32_t lsb_init_device(void);
const char *lsb_version(void);

typedef struct LsbReaderStats {
  int64_t rows_in, rows_out, h2d_bytes, d2h_bytes;
  int32_t kernel_launches;
  int32_t stream_passes;        /* > 1: the bucket was read in that many scans over ascending primary-key ranges */
  double decode_ms, upload_ms, merge_ms, gather_ms, download_ms;
  int64_t device_decoded_pages; /* Parquet data pages decoded by the GPU kernels */
  int64_t host_decoded_files;   /* files that needed the Arrow C++ host decoder (unsupported encodings) */
  int64_t pruned_row_groups;    /* row groups skipped by primary-key min/max statistics (pushed key filters) */
  int64_t device_decompressed_pages; /* ZSTD page payloads expanded by the GPU (the rest is decompressed on the host) */
  int64_t pruned_buckets;       /* 1: an equality filter on the whole primary key selects another hash bucket: nothing was read */
  int64_t pruned_pages;         /* data pages skipped by the key columns' page index (pushed key filters) */
} LsbReaderStats;
int32_t lsb_reader_get_stats(CResult_Reader *reader, LsbReaderStats *out);

/* write side (create_lakesoul_writer_from_config .. flush_and_close_writer): what the GPU did for this writer */
typedef struct LsbWriterStats {
  int64_t rows_in, rows_out, files, flush_cycles; /* a flush cycle = one device pass (hash + sort + gather) over the buffered rows */
  int64_t h2d_bytes, d2h_bytes;
  int32_t kernel_launches, pad;
  double sort_ms, encode_ms;                      /* device pass incl. copies; host Parquet encode */
} LsbWriterStats;
int32_t lsb_writer_get_stats(CResult_Writer *writer, LsbWriterStats *out);

/* test hook: walks the Parquet page headers of the byte range [off, off + len) of a file image with the library's own
 * compact-thrift reader; out[8 * i ..] = {page type, uncompressed size, compressed size, num_values, encoding,
 * definition-level encoding, v2 definition-levels byte length, header length}. Returns the page count, -1 on a
 * malformed header, -2 when the pages do not end exactly at off + len. */
int64_t lsb_test_walk_page_heads(const uint8_t *image, int64_t off, int64_t len, int32_t *out, int64_t max_pages);
/* test hook: frames row groups [rg_lo, rg_hi) of a file from the whole image (lazy = 0) or from the byte span of those row
 * groups (lazy = 1) and prints one position-independent line per page; returns the text length, -1 on error. */
int64_t lsb_test_plan_digest(const char *path, int32_t lazy, int32_t rg_lo, int32_t rg_hi, char *out, int64_t cap);
/* test hook: prunes row groups, then pages, of a file by a filter (parquet-mr string form) over the key columns `pks` (comma
 * separated) and prints the live row range and the framed pages of every column */
int64_t lsb_test_page_prune(const char *path, const char *filter, const char *pks, int32_t lazy, char *out, int64_t cap);
/* test hook: the key-range filters of the passes a bucket with these files (newline separated) is streamed in, one per line */
int64_t lsb_test_stream_passes(const char *files, const char *pk, const char *opts, int64_t pass_bytes, char *out, int64_t cap);

/* ---- device-level plan --------------------------------------------------------------------- */
typedef struct LsbPlan LsbPlan;

/* nb hash buckets; bucket b owns files [bucket_file_begin[b], bucket_file_begin[b+1]) in commit
 * order; file f has file_rows[f] rows sorted by primary key and starts at global row
 * file_row_base[f] of every column buffer (NULL = files packed back to back). */
LsbPlan *lsb_plan_create(int32_t nb, const int32_t *bucket_file_begin, const int64_t *file_rows,
                         const int64_t *file_row_base, int64_t row_span, int32_t key_bits,
                         int32_t batch_size, void *cuda_stream);
void lsb_plan_destroy(LsbPlan *plan);

/* pk_kind: 0 unsigned, 1 signed, 2 float (IEEE total order); pk_width in bytes (1,2,4,8);
 * pk_validity[i] may be NULL. key_bits passed to lsb_plan_create must be
 * sum(8*width + (validity ? 1 : 0)). */
int32_t lsb_plan_merge(LsbPlan *plan, int32_t n_pk, const void *const *pk_values,
                       const void *const *pk_validity, const int32_t *pk_width, const int32_t *pk_kind,
                       void *cuda_stream);
int32_t lsb_plan_identity(LsbPlan *plan, void *cuda_stream);
/* blocks until the merge is done; fills out_base[nb+1] and batch_base[nb+1] (host arrays) */
int32_t lsb_plan_sync_counts(LsbPlan *plan, int64_t *out_base, int64_t *batch_base, void *cuda_stream);
const void *lsb_plan_winners(LsbPlan *plan); /* device uint32[total_out] of global row ids */
int64_t lsb_plan_valid_stride(LsbPlan *plan);
/* stats[8] = {rows_in, rows_out, tiles_level0, levels, log2(S), tile_capacity, launches, 0} */
int32_t lsb_plan_stats(LsbPlan *plan, int64_t *stats);

int32_t lsb_plan_gather_fixed(LsbPlan *plan, int32_t ncols, const void *const *src,
                              const void *const *src_valid, void *const *dst, void *const *dst_valid,
                              const int32_t *width, void *cuda_stream);
int32_t lsb_plan_gather_var(LsbPlan *plan, const void *src_offsets, const void *src_data,
                            const void *src_data_base, const void *src_valid, void *dst_offsets,
                            void *dst_data, void *dst_valid, void *batch_data_start, void *cuda_stream);

/* n variable-width columns in one launch (per 16 columns); every argument is an array of n device pointers */
int32_t lsb_plan_gather_var_multi(LsbPlan *plan, int32_t ncols, const void *const *src_offsets,
                                  const void *const *src_data, const void *const *src_data_base,
                                  const void *const *src_valid, void *const *dst_offsets, void *const *dst_data,
                                  void *const *dst_valid, void *const *batch_data_start, void *cuda_stream);

/* fixed-width + variable-width columns of one scan (2 launches) */
int32_t lsb_plan_gather_all(LsbPlan *plan, int32_t nfixed, const void *const *src, const void *const *src_valid,
                            void *const *dst, void *const *dst_valid, const int32_t *width, int32_t nvars,
                            const void *const *v_src_offsets, const void *const *v_src_data,
                            const void *const *v_src_data_base, const void *const *v_src_valid,
                            void *const *v_dst_offsets, void *const *v_dst_data, void *const *v_dst_valid,
                            void *const *v_batch_data_start, void *cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* LAKESOUL_B200_EXT_H */
