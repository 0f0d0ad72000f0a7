// SPDX-License-Identifier: Apache-2.0
//
// mor_engine.h -- device-level merge-on-read engine (internal C++ API).
//
// One MergePlan describes a *scan batch*: nb hash buckets, bucket b holding k_b
// delta files (commit order = file order), every file sorted by primary key.
// All columns of all files of the batch live in HBM as ONE concatenated array per
// column ("global rows": file f occupies rows [file_row_base[f], +file_rows[f])).
//
// plan_merge()   : PK k-way merge + last-writer-wins selection for every bucket in
//                  a handful of launches -> dense winner list win[] of global row ids
//                  (replaces UseLastRangeCombiner::poll / loser tree,
//                  reference rust/lakesoul-io/src/physical_plan/merge/sorted/combiner.rs:478-525,680-785
//                  and SortKeyBatchRange::advance, sort_key_range.rs:99-113,182-187).
// gather_*()     : out[u] = src[win[u]] into batch-segmented Arrow buffers
//                  (replaces build_record_batch_full_merge + arrow interleave, combiner.rs:529-605).
//
// No torch / arrow types here: plain device pointers and sizes.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include <string>
#include <vector>

namespace lsb {

enum PkKind : uint8_t { PK_UNSIGNED = 0, PK_SIGNED = 1, PK_FLOAT = 2, PK_BYTES = 3 };

struct PkCol {
  const void* values;       // `width` bytes per global row; PK_BYTES: int32 offsets (bucket relative)
  const uint8_t* validity;  // Arrow validity bitmap over global rows, or nullptr (no nulls)
  uint8_t width;            // 1, 2, 4, 8 (PK_BYTES: 0)
  uint8_t kind;             // PkKind
  uint8_t pad[6];
  const uint8_t* data;          // PK_BYTES: byte buffer of the column
  const int64_t* bucket_data;   // PK_BYTES: data base of every bucket (device, nb entries)
};

constexpr int kMaxPkCols = 4;
struct PkDesc {
  int32_t ncols;
  int32_t key_bits;  // sum(width*8 + (validity?1:0)); <=64 -> 64-bit keys, <=128 -> 128-bit keys;
                     // a single Utf8/Binary key column (PK_BYTES) uses 8-byte prefixes + full compares: pass 128
  PkCol col[kMaxPkCols];
};

constexpr int kMaxLevels = 6;   // sampling levels above level 0
constexpr int kMaxFilesPerBucket = 256;

// Per-bucket plan record (host-built, copied to the device once per plan).
struct BucketDesc {
  int32_t file_begin;               // first global file id
  int32_t k;                        // number of files
  int32_t levels;                   // L_b: top sampling level (0 = bucket fits one tile)
  int32_t m;                        // splitter stride, in samples of the level above
  int64_t row_base;                 // global row of the bucket's first row
  int64_t n_rows;                   // total input rows
  int64_t split_ofs[kMaxLevels + 1];  // offset into splits[] of the (T_l+1) x k table
  int64_t mlist_ofs[kMaxLevels + 1];  // offset into mlist[] of the merged level-l list (l>=1)
  int32_t ntiles[kMaxLevels + 1];     // T_{b,l}
  int32_t tile_base[kMaxLevels + 1];  // index of the bucket's first tile inside level l
};

struct MergeStats {
  int64_t n_rows = 0;       // input rows
  int64_t n_out = 0;        // output rows (valid after plan_sync_counts)
  int32_t tiles_l0 = 0;
  int32_t levels = 0;
  int32_t log_s = 0;
  int32_t tile_cap = 0;
  int32_t launches = 0;     // kernels launched by the last plan_merge()
};

// Gather descriptors ---------------------------------------------------------
constexpr uint32_t kNoRow = 0xffffffffu;  // winner slot "no source row" -> output null

struct FixedCol {
  const void* src;          // global-row indexed values (width bytes; bit-packed when width==0)
  const uint8_t* src_valid; // bitmap over global rows or nullptr
  void* dst;                // dense output values, (total_out) * width; width==0: batch-segmented bits
  uint8_t* dst_valid;       // batch-segmented bitmap or nullptr (then no validity is produced)
  const uint32_t* win;      // per-column winner list (partial merge / operators) or nullptr = plan winners
  int32_t width;            // 0 (boolean bits),1,2,4,8,16,32
  int32_t pad;
};

struct MergePlan;  // opaque

// Build the plan (host work + descriptor upload on `stream`). Throws std::runtime_error on
// unsupported shapes (too many files per bucket, > 2^32-2 rows per batch).
// file_row_base (optional): global row of every file's first row (F entries, increasing, a file
// never overlaps the next); nullptr = files packed back to back. row_span = one past the last
// global row (ignored when file_row_base is null).
MergePlan* plan_create(int nb, const int32_t* bucket_file_begin, const int64_t* file_rows,
                       const int64_t* file_row_base, int64_t row_span, int key_bits, int batch_size,
                       cudaStream_t stream);
void plan_destroy(MergePlan*);

// Run partition + tile merge + scan + compaction. Asynchronous on `stream`.
void plan_merge(MergePlan*, const PkDesc& pk, cudaStream_t stream);

// Primary keys the fixed-width fast paths do not cover (a string inside a composite key, decimal keys, more than
// 128 key bits, more than 4 columns): every row's key is materialised as ONE byte string whose bytewise order is
// the reference's row order (arrow-row through RowCursorStream, stream/mod.rs:104-122: columns in order, nulls
// first, signed / float / decimal values as order-preserving big-endian images, strings bytewise), and the merge
// runs on that column like on a single Binary key. Per column: 1 validity byte, then the fixed-width image or the
// escaped string (0x00 -> 0x00 0xFF) closed by 0x00 0x00.
enum EncKind : int32_t { ENC_UNSIGNED = 0, ENC_SIGNED = 1, ENC_FLOAT = 2, ENC_BYTES = 3, ENC_BOOL_BITS = 4 };
struct EncKeyCol {
  const void* values;          // fixed width: `width` bytes per global row (little endian); ENC_BYTES: int32 offsets
  const uint8_t* validity;     // or nullptr
  const uint8_t* data;         // ENC_BYTES
  const int64_t* bucket_data;  // ENC_BYTES: device, per bucket
  int32_t width;               // 1, 2, 4, 8, 16, 32 (ENC_SIGNED covers the decimals)
  int32_t kind;
};
// Fills `pk` (one PK_BYTES column, key_bits 128) with device buffers owned by the plan. Blocks once on `stream`.
void plan_encode_keys(MergePlan*, const EncKeyCol* cols, int ncols, PkDesc* pk, cudaStream_t stream);

// Write side. plan_hash_buckets: d_out[global row] = spark murmur3_32 of the key columns (chained,
// seed 42, nulls skipped; utils/hash/mod.rs:22-178) mod nbuckets. plan_sort: after plan_identity + plan_sync_counts,
// reorders the plan's row list by `pk` (ties keep the input order); the gathers then emit the rows sorted.
// plan_take_u32: d_out[u] = d_src[row list[u]].
void plan_hash_buckets(MergePlan*, const EncKeyCol* cols, int ncols, uint32_t nbuckets, uint32_t* d_out, cudaStream_t stream);
void plan_sort(MergePlan*, const PkDesc& pk, cudaStream_t stream);
void plan_take_u32(MergePlan*, const uint32_t* d_src, uint32_t* d_out, cudaStream_t stream);

// "No merge" variant (skip_merge_on_read / no primary key, merge/mod.rs:297-312): every input row
// survives, in file order.
void plan_identity(MergePlan*, cudaStream_t stream);

// Keep only output rows with keep[u] != 0 (device array over the current output rows): compacts
// the winner list and every list in `extra_lists` (device, same length), refreshes the per-bucket
// bases. Call plan_sync_counts() afterwards.
void plan_compact(MergePlan*, const uint8_t* keep, const std::vector<uint32_t*>& extra_lists,
                  cudaStream_t stream);

// Wait for the merge and fetch per-bucket output row counts (host vector of nb entries).
const std::vector<int64_t>& plan_sync_counts(MergePlan*, cudaStream_t stream);

const MergeStats& plan_stats(const MergePlan*);
int64_t plan_total_rows(const MergePlan*);
int plan_num_buckets(const MergePlan*);
int plan_batch_size(const MergePlan*);

// Device views (valid after plan_merge on the same stream).
const uint32_t* plan_winners(const MergePlan*);      // dense winners, bucket b at [out_base[b], out_base[b+1])
const int64_t* plan_out_base_dev(const MergePlan*);  // nb+1 entries (device)
const int32_t* plan_error_flag_dev(const MergePlan*);  // device int32: 0 = no kernel of the plan reported an error

// Layout helpers for batch-segmented outputs, valid after plan_sync_counts():
//   validity: every batch owns valid_stride_bytes() bytes; batch (b,i) is batch number
//             batch_base(b)+i;  offsets: batch (b,i) starts at entry out_base[b]+i*bs + batch_base(b)+i.
int64_t plan_valid_stride_bytes(const MergePlan*);
const std::vector<int64_t>& plan_out_base(const MergePlan*);    // host copy, nb+1
const std::vector<int64_t>& plan_batch_base(const MergePlan*);  // host, nb+1 (prefix of batches per bucket)

// out[u] = src[win[u]] for a set of fixed-width columns in one launch.
void gather_fixed(MergePlan*, const FixedCol* cols, int ncols, cudaStream_t stream);

// Variable-width (Utf8/Binary) gather. src offsets are int32 indexed by global row (entries g and
// g+1 delimit row g, so a file needs one spare global row after its last row, or the next file
// must continue the same data run), relative to the bucket's data base `src_data_base[b]`
// (int64, device array of nb).
// src_data must be readable up to the next 16-byte boundary after its last byte and start 16-byte aligned (true
// for any cudaMalloc'd buffer): the windowed gather moves whole 16-byte aligned windows with TMA bulk copies.
// Produces: dst_offsets (int32, batch-rebased, (rows+1) per batch), dst_data (dense bytes),
// batch_data_start (int64 per batch + 1, device) = byte offset of each batch inside dst_data.
struct VarCol {
  const int32_t* src_offsets;
  const uint8_t* src_data;
  const int64_t* src_data_base;  // device, nb entries
  const uint8_t* src_valid;      // or nullptr
  const uint32_t* win;           // per-column winner list or nullptr
  int32_t* dst_offsets;
  uint8_t* dst_data;             // capacity must be >= total source bytes of the column
  uint8_t* dst_valid;            // or nullptr
  int64_t* batch_data_start;     // device, (total_batches + 1)
  int32_t direct;                // 1: src rows are NOT file rows of the plan (computed columns) -> no source windows
  int32_t pad;
};
void gather_var(MergePlan*, const VarCol& col, cudaStream_t stream);
// same for several columns: one launch per 16 columns (chunk positions by a look-back scan inside the kernel)
void gather_var_multi(MergePlan*, const VarCol* cols, int ncols, cudaStream_t stream);

// Everything at once: 1 launch for the fixed-width columns + 1 for the string columns.
void gather_all(MergePlan*, const FixedCol* fixed, int nfixed, const VarCol* vars, int nvars, cudaStream_t stream);

// Scratch management for gather_var (lengths + scanned offsets); grows on demand.
void plan_reserve_scratch(MergePlan*, cudaStream_t stream);

std::string last_error();

// Scratch arena for one reader at a time: device memory is handed out by bumping a pointer, nothing
// is returned until reset(). While a thread has an arena installed (arena_install), every device
// allocation of the engine on that thread comes from it and frees are no-ops -- per-bucket readers
// then make no cudaMallocAsync / cudaFreeAsync calls at all (measured ~100 us each under 16 concurrent
// readers, the largest host-side cost of a bucket).
class DeviceArena {
 public:
  ~DeviceArena();
  void* alloc(size_t bytes);   // 256-byte aligned; grows with cudaMalloc when the current block is full
  void reset();                // all GPU work on the arena must have completed
  size_t capacity() const;
 private:
  struct Block { uint8_t* p; size_t cap; };
  std::vector<Block> blocks_;
  size_t cur_ = 0, off_ = 0, high_ = 0, used_ = 0, recent_ = 0;
  int idle_ = 0;
};
void arena_install(DeviceArena* a);   // nullptr restores cudaMallocAsync / cudaFreeAsync
DeviceArena* arena_current();
void* device_alloc_bytes(size_t bytes, cudaStream_t stream);  // arena if installed, else cudaMallocAsync
void device_free_bytes(void* p, cudaStream_t stream);

// Pinned host staging for the small descriptor uploads of a scan (plans, page lists, column tables).
// cudaMemcpyAsync from PAGEABLE memory first synchronises the stream and holds the calling thread (and the
// context) until everything queued before it has run -- with 16 readers in flight that serialised the whole
// process (r2 trace: 48 ms spent "enqueueing" per bucket). While a thread has a HostStage installed,
// h2d_staged() copies the bytes into pinned staging memory and issues a truly asynchronous copy; the staging
// memory is recycled by reset(), once the scan's stream has drained.
class HostStage {
 public:
  ~HostStage();
  void* alloc(size_t bytes);  // 64-byte aligned pinned memory, valid until reset()
  void reset();
 private:
  struct Block { uint8_t* p; size_t cap; };
  std::vector<Block> blocks_;
  size_t cur_ = 0, off_ = 0;
};
void host_stage_install(HostStage* s);  // nullptr: h2d_staged falls back to a plain cudaMemcpyAsync
HostStage* host_stage_current();
cudaError_t h2d_staged(void* dst, const void* src, size_t bytes, cudaStream_t stream);

// Wait for `stream` without spinning: the calling thread sleeps on a blocking-sync event, so task
// slots that wait for the GPU leave their core to the slots that are reading files
// (LAKESOUL_B200_SPIN_SYNC=1 restores cudaStreamSynchronize).
cudaError_t stream_wait(cudaStream_t stream);

}  // namespace lsb
