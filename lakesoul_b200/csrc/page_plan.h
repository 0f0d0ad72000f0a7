// SPDX-License-Identifier: Apache-2.0
//
// page_plan.h -- host-side framing of Parquet files for the device decoders. A file is read once into pinned memory --
// whole, or (lazily opened files: key-filtered readers, passes of a streamed bucket) footer first and then only the byte
// segments of the row groups / pages / columns that are needed --, the footer is parsed by libparquet, page headers by a
// purpose-built thrift-compact reader, row groups and pages are pruned by the key columns' statistics, and every data
// page is described by a POD record that points into the raw image. Page payloads stay compressed (ZSTD, SNAPPY: expanded
// on the GPU) or are used in place (uncompressed); only chunks the device cannot expand go through Arrow's codecs on
// the host. Values are never touched on the host.
#pragma once
#include <arrow/api.h>
#include <arrow/io/interfaces.h>

#include <map>
#include <memory>
#include <set>
#include <string>
#include <vector>

#include "decode_types.h"
#include "filter.h"

namespace lsb {

std::shared_ptr<uint8_t> pinned_alloc(size_t bytes);  // recycled cudaMallocHost blocks (device_scan.cc)

constexpr uint8_t kCodecZstd = 1, kCodecSnappy = 2;

struct RawColumn {                    // one requested column of one file
  std::string name;
  bool present = false;
  bool optional = false;              // max definition level 1
  bool may_have_nulls = false;        // optional and statistics do not prove null_count == 0
  PqChunk proto{};                    // phys/type_length/conv/out_width/max_def filled; col/bucket set later
  struct Page {
    int64_t src_off = 0;              // inside the file image (aux = false) or the aux arena (aux = true)
    bool aux = false;
    uint8_t codec = 0;                // 0: src_off/size = plain payload; kCodecZstd / kCodecSnappy: src_off/comp_size =
                                      // compressed payload inside the file image, expanded on the device to `size` bytes
    int32_t comp_size = 0;
    int32_t size = 0, num_values = 0, encoding = 0, def_encoding = 0, v2_def_len = -1;
    int32_t row_group = 0;
    int64_t first_row = 0;            // row inside the file
  };
  struct Dict {
    int64_t src_off = 0;
    bool aux = false;
    uint8_t codec = 0;  // like Page::codec
    int32_t comp_size = 0;
    int32_t size = 0, num_values = 0;
  };
  std::vector<Page> pages;
  std::vector<Dict> dicts;            // one per row group (or size/num_values 0)
  std::vector<int32_t> page_dict;     // page -> index into dicts, -1 none
};

// min / max statistics of one row group's chunk of a key column (streaming by key range, reader.cc)
struct RowGroupKeyRange {
  enum Fam { kNone, kInt, kBytes } fam = kNone;  // kNone: no usable statistics
  int64_t imin = 0, imax = 0;                    // kInt (timestamps in microseconds, like filter literals)
  std::string bmin, bmax;                        // kBytes (unsigned byte order)
  int64_t rows = 0, nulls = -1;                  // nulls: -1 unknown
  int64_t bytes = 0;                             // uncompressed bytes of the whole row group
};

class RawParquetFile {
 public:
  // reads the whole file into pinned memory and parses the footer; throws on IO / format errors
  static std::shared_ptr<RawParquetFile> Open(const std::string& path);
  // reads the footer only; the image stays empty until LoadKept() reads the byte span of the kept row groups
  // (files are sorted by primary key: the row groups a key range keeps are neighbours). Used when a bucket is
  // streamed by key range (reader.cc) so that pinned memory holds one pass, not the bucket.
  static std::shared_ptr<RawParquetFile> OpenLazy(const std::string& path);
  // no-op for files opened with Open(). After PrunePages() found a live row range and the file has an offset index, only
  // the pages that overlap it (plus dictionary pages) of the columns in `columns` (nullptr = all) are read; the image is
  // then a concatenation of byte segments of the file.
  void LoadKept(const std::vector<std::string>* columns = nullptr);
  int64_t file_bytes() const { return file_bytes_; }
  // the file as an Arrow source (host-decode fallback): the image when it is the whole file, else image + pread
  std::shared_ptr<arrow::io::RandomAccessFile> source() const;
  // one entry per kept row group; empty when the file has no such (flat) column
  std::vector<RowGroupKeyRange> KeyRanges(const std::string& column) const;

  int64_t num_rows() const { return rows_; }
  const std::shared_ptr<arrow::Schema>& schema() const { return schema_; }  // full file schema (Arrow types)

  // PK-filter pushdown (reference file_format.rs:234-265 pushes PK-only conjuncts below the merge as Inexact
  // filters): drops the row groups in which `filter` cannot be TRUE judging by the min/max
  // statistics of the `pk_columns` only -- every version of a key carries the same key, so a row group pruned
  // by a key predicate is pruned in every file alike and the merge result is unchanged. Call before Plan();
  // num_rows() then counts the kept row groups. Returns the number of row groups dropped.
  int PruneRowGroups(const FilterExpr& filter, const std::set<std::string>& pk_columns);
  // Page-level refinement (reference session.rs:685-717 enables the page index for exactly this): the min / max of the
  // key columns' PAGES (ColumnIndex + OffsetIndex, written by parquet-rs by default) narrow the rows in which `filter`
  // can be TRUE to one contiguous range of the kept rows -- contiguous because a file is sorted by primary key; for
  // predicates that are not a range the range is the hull. Plan() then frames only the pages that overlap it, the merge
  // takes rows [live_offset(), live_offset() + live_rows()) of the file's (smaller) row space. Call after
  // PruneRowGroups(), before Plan(). Returns the number of kept rows outside the range (0 without a page index).
  int64_t PrunePages(const FilterExpr& filter, const std::set<std::string>& pk_columns);
  // After Plan(): the rows of this file in the scan's row space = pages framed (layout_rows), of which the merge reads
  // [live_offset, live_offset + live_rows). Without page pruning: layout_rows = live_rows = num_rows, offset 0.
  int64_t layout_rows() const { return layout_rows_; }
  int64_t live_offset() const { return live_lo_ - origin_; }
  int64_t live_rows() const { return live_hi_ - live_lo_; }
  // the live range in kept-row coordinates (host-decoded files are sliced by it)
  int64_t live_begin() const { return live_lo_; }
  int64_t pruned_pages() const { return pruned_pages_; }
  const std::vector<int>& kept_row_groups() const { return kept_rg_; }  // indices, ascending (all by default)
  void KeepRowGroups(int lo, int hi);  // keeps the kept row groups with lo <= index < hi (before LoadKept / Plan)

  // Frames the pages of the requested columns (missing columns are allowed: present = false).
  // Returns false and fills `why` when some column chunk needs a decoder the device does not have;
  // the caller then decodes this file on the host.
  bool Plan(const std::vector<std::shared_ptr<arrow::Field>>& want, std::string* why);

  const std::vector<RawColumn>& columns() const { return cols_; }
  // the bytes of the file in memory -- all of it, or the segments LoadKept() read, back to back; page records hold
  // offsets into image()
  const uint8_t* image() const { return image_.get(); }
  int64_t image_bytes() const { return image_bytes_; }
  struct Segment { int64_t file_off = 0, len = 0, img_off = 0; };
  // where `len` bytes at `file_off` of the file sit inside image(), or -1 when they are not (all) loaded
  int64_t image_offset_of(int64_t file_off, int64_t len) const;
  const uint8_t* aux() const { return aux_.get(); }
  int64_t aux_bytes() const { return aux_used_; }
  int64_t host_decompressed_bytes() const { return aux_used_; }
  // device copy of image() when the reader already sent it (session_upload), else nullptr
  uint8_t* device_image() const { return device_image_; }
  void set_device_image(uint8_t* p) { device_image_ = p; }

 private:
  int64_t aux_append(const uint8_t* p, int64_t n);
  std::string path_;
  std::shared_ptr<uint8_t> image_, aux_;
  int64_t image_bytes_ = 0, aux_cap_ = 0, aux_used_ = 0, rows_ = 0;
  int64_t file_bytes_ = 0;
  std::vector<Segment> segs_;  // ascending file offsets, disjoint
  struct PageSpan {            // (row group, column) of a lazily loaded, page-pruned file: what LoadKept() read of the chunk
    int64_t dict_off = 0, dict_len = 0, data_off = 0, data_len = 0;
    int64_t first_row = 0, end_row = 0;  // rows of the row group the data pages cover
    int32_t skipped_pages = 0;
  };
  std::map<std::pair<int, int>, PageSpan> spans_;
  int64_t live_lo_ = 0, live_hi_ = 0, origin_ = 0, layout_rows_ = 0, pruned_pages_ = 0;
  bool page_pruned_ = false;
  bool lazy_ = false, loaded_ = true;
  std::shared_ptr<arrow::io::RandomAccessFile> source_;
  std::shared_ptr<arrow::Schema> schema_;
  std::shared_ptr<void> reader_;  // parquet::ParquetFileReader
  std::vector<RawColumn> cols_;
  std::vector<int> kept_rg_;
  uint8_t* device_image_ = nullptr;
};

}  // namespace lsb
