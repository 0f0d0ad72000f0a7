// SPDX-License-Identifier: Apache-2.0
#include "page_plan.h"

#include <arrow/io/memory.h>
#include <fcntl.h>
#include <parquet/arrow/schema.h>
#include <parquet/column_page.h>
#include <parquet/column_reader.h>
#include <parquet/file_reader.h>
#include <parquet/metadata.h>
#include <parquet/page_index.h>
#include <parquet/statistics.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <stdexcept>

#include "zstd_decode.h"

namespace lsb {

namespace {
[[noreturn]] void fail(const std::string& m) { throw std::runtime_error(m); }

// ZSTD page payloads are expanded on the GPU (zstd_pages_kernel_v2) by default: with 32 hardware work queues the
// all-device path reads config-5 buckets at 51 M rows/s on 16 task slots with 5 host cores busy, the host codec
// (Arrow) at 40 M rows/s with 12 cores busy, and splits by chunk size at 36 M rows/s (r2 sweep).
// LAKESOUL_B200_DEVICE_ZSTD_SHARE=<0..1> sends that share of a file's compressed-chunk bytes (smallest chunks first)
// to the GPU (0 = host codec only); LAKESOUL_B200_DEVICE_ZSTD_MIN_BYTES=<n> instead sends the chunks of at least n
// uncompressed bytes.
double device_zstd_share() {
  static const double share = [] {
    const char* v = getenv("LAKESOUL_B200_DEVICE_ZSTD_SHARE");
    if (!v) v = getenv("LAKESOUL_B200_DEVICE_ZSTD");  // 0 / 1
    if (!v || !*v) return getenv("LAKESOUL_B200_DEVICE_ZSTD_MIN_BYTES") ? -1.0 : 1.0;  // default: everything on the GPU
    const double x = atof(v);
    return x < 0 ? 0.0 : (x > 1 ? 1.0 : x);
  }();
  return share;
}

// Optional split of the ZSTD work by chunk size: column chunks of at least this many uncompressed bytes are expanded
// on the device, smaller ones by the host codec.
int64_t device_zstd_min_bytes() {
  static const int64_t v = [] {
    const char* e = getenv("LAKESOUL_B200_DEVICE_ZSTD_MIN_BYTES");
    return e && *e ? (int64_t)atoll(e) : (int64_t)(128 << 10);
  }();
  return v;
}

// SNAPPY page payloads (pyarrow / Spark default files) are expanded by snappy_pages_kernel unless
// LAKESOUL_B200_DEVICE_SNAPPY=0 sends them to the host codec.
bool device_snappy() {
  static const bool on = [] { const char* v = getenv("LAKESOUL_B200_DEVICE_SNAPPY"); return !(v && *v == '0'); }();
  return on;
}

// uncompressed size announced by a compressed payload itself, or -1
int64_t announced_size(uint8_t codec, const uint8_t* p, int64_t n) {
  if (codec == kCodecZstd) return zstd_frame_content_size(p, n);
  uint64_t v = 0;  // snappy preamble: little-endian base-128 varint
  for (int i = 0, sh = 0; i < 5 && i < n; ++i, sh += 7) {
    v |= (uint64_t)(p[i] & 0x7f) << sh;
    if (!(p[i] & 0x80)) return v <= 0x7fffffffull ? (int64_t)v : -1;
  }
  return -1;
}

int32_t map_encoding(parquet::Encoding::type e) {
  switch (e) {
    case parquet::Encoding::PLAIN: return ENC_PLAIN;
    case parquet::Encoding::PLAIN_DICTIONARY: return ENC_PLAIN_DICT;
    case parquet::Encoding::RLE: return ENC_RLE;
    case parquet::Encoding::BIT_PACKED: return ENC_BIT_PACKED;
    case parquet::Encoding::RLE_DICTIONARY: return ENC_RLE_DICT;
    default: return -1;
  }
}
}  // namespace

// ---------------------------------------------------------------------------------------
// Page headers read straight from the file image (thrift compact protocol, parquet.thrift PageHeader): the framing of a
// config-5 bucket walks 1 072 of them, and parquet::PageReader costs ~1 us of allocations and statistics copies per
// page on top of the decode (r3 host profile). Used for chunks whose payloads stay where they are (uncompressed, or
// expanded on the GPU); chunks libparquet has to decompress keep going through parquet::PageReader.
// ---------------------------------------------------------------------------------------
namespace {
struct PageHead {
  int type = -1;                  // parquet::PageType: 0 data, 1 index, 2 dictionary, 3 data v2
  int32_t uncompressed = -1, compressed = -1;
  int32_t num_values = -1;
  int encoding = -1, def_encoding = -1;
  int32_t v2_def_len = 0, v2_rep_len = 0;
  bool have_sub = false;
  int64_t header_len = 0;
};

class CompactReader {
 public:
  CompactReader(const uint8_t* p, const uint8_t* end) : p_(p), end_(end) {}
  bool ok() const { return ok_; }
  const uint8_t* pos() const { return p_; }
  uint64_t varint() {
    uint64_t v = 0;
    for (int sh = 0; sh < 64; sh += 7) {
      if (p_ >= end_) { ok_ = false; return 0; }
      const uint8_t b = *p_++;
      v |= (uint64_t)(b & 0x7f) << sh;
      if (!(b & 0x80)) return v;
    }
    ok_ = false;
    return 0;
  }
  int64_t zigzag() { const uint64_t v = varint(); return (int64_t)(v >> 1) ^ -(int64_t)(v & 1); }
  // next field of the current struct: returns false at STOP; *type = compact type, *id = field id
  bool field(int* type, int* id, int* last_id) {
    if (p_ >= end_) { ok_ = false; return false; }
    const uint8_t b = *p_++;
    if (b == 0) return false;
    *type = b & 0x0f;
    const int delta = b >> 4;
    if (delta) *id = *last_id + delta;
    else *id = (int)zigzag();
    *last_id = *id;
    return ok_;
  }
  void skip(int type, int depth = 0) {
    if (!ok_ || depth > 8) { ok_ = false; return; }
    switch (type) {
      case 1: case 2: break;                  // bool in the field header
      case 3: if (p_ < end_) ++p_; else ok_ = false; break;
      case 4: case 5: case 6: (void)varint(); break;
      case 7: if (end_ - p_ >= 8) p_ += 8; else ok_ = false; break;
      case 8: { const uint64_t n = varint(); if ((uint64_t)(end_ - p_) >= n) p_ += n; else ok_ = false; } break;
      case 9: case 10: {
        if (p_ >= end_) { ok_ = false; break; }
        const uint8_t h = *p_++;
        uint64_t n = h >> 4;
        if (n == 15) n = varint();
        const int et = h & 0x0f;
        for (uint64_t i = 0; i < n && ok_; ++i) skip_elem(et, depth + 1);
      } break;
      case 11: {
        const uint64_t n = varint();
        if (n == 0) break;
        if (p_ >= end_) { ok_ = false; break; }
        const uint8_t kv = *p_++;
        for (uint64_t i = 0; i < n && ok_; ++i) { skip_elem(kv >> 4, depth + 1); skip_elem(kv & 0x0f, depth + 1); }
      } break;
      case 12: {
        int t, id, last = 0;
        while (field(&t, &id, &last)) skip(t, depth + 1);
      } break;
      default: ok_ = false;
    }
  }
 private:
  void skip_elem(int type, int depth) {  // inside a container a bool takes a byte
    if (type == 1 || type == 2) { if (p_ < end_) ++p_; else ok_ = false; }
    else skip(type, depth);
  }
  const uint8_t* p_;
  const uint8_t* end_;
  bool ok_ = true;
};

// parses one PageHeader at `p`; false = malformed / truncated
bool read_page_head(const uint8_t* p, const uint8_t* end, PageHead* h) {
  CompactReader r(p, end);
  int t, id, last = 0;
  while (r.field(&t, &id, &last)) {
    if (!r.ok()) return false;
    if (id == 1 && t == 5) h->type = (int)r.zigzag();
    else if (id == 2 && t == 5) h->uncompressed = (int32_t)r.zigzag();
    else if (id == 3 && t == 5) h->compressed = (int32_t)r.zigzag();
    else if ((id == 5 || id == 7 || id == 8) && t == 12) {
      h->have_sub = true;
      int t2, id2, last2 = 0;
      while (r.field(&t2, &id2, &last2)) {
        if (!r.ok()) return false;
        const bool i32 = t2 == 5;
        if (id == 5) {          // DataPageHeader
          if (id2 == 1 && i32) h->num_values = (int32_t)r.zigzag();
          else if (id2 == 2 && i32) h->encoding = (int)r.zigzag();
          else if (id2 == 3 && i32) h->def_encoding = (int)r.zigzag();
          else r.skip(t2);
        } else if (id == 7) {   // DictionaryPageHeader
          if (id2 == 1 && i32) h->num_values = (int32_t)r.zigzag();
          else if (id2 == 2 && i32) h->encoding = (int)r.zigzag();
          else r.skip(t2);
        } else {                // DataPageHeaderV2
          if (id2 == 1 && i32) h->num_values = (int32_t)r.zigzag();
          else if (id2 == 4 && i32) h->encoding = (int)r.zigzag();
          else if (id2 == 5 && i32) h->v2_def_len = (int32_t)r.zigzag();
          else if (id2 == 6 && i32) h->v2_rep_len = (int32_t)r.zigzag();
          else r.skip(t2);
        }
      }
    } else {
      r.skip(t);
    }
    if (!r.ok()) return false;
  }
  if (!r.ok() || h->type < 0 || h->uncompressed < 0 || h->compressed < 0) return false;
  h->header_len = r.pos() - p;
  return true;
}
}  // namespace

// test hook (tests/test_page_heads.py through ctypes): page headers of the byte range [off, off + len) of a file image
extern "C" __attribute__((visibility("default"))) int64_t lsb_test_walk_page_heads(const uint8_t* image, int64_t off, int64_t len, int32_t* out,
                                                                                   int64_t max_pages) {
  int64_t n = 0, p = off;
  while (p < off + len && n < max_pages) {
    PageHead h;
    if (!read_page_head(image + p, image + off + len, &h)) return -1;
    int32_t* o = out + n * 8;
    o[0] = h.type; o[1] = h.uncompressed; o[2] = h.compressed; o[3] = h.num_values; o[4] = h.encoding; o[5] = h.def_encoding;
    o[6] = h.v2_def_len; o[7] = (int32_t)h.header_len;
    p += h.header_len + h.compressed;
    ++n;
  }
  return p == off + len ? n : -2;
}

// A file of which only a window [base, base + len) sits in (pinned) memory: reads inside the window are served from it
// without a copy -- Plan() recognises such pages as part of the image --, everything else (the footer, a chunk the
// caller did not load) by pread.
namespace {
class WindowFile : public arrow::io::RandomAccessFile {
 public:
  WindowFile(int fd, int64_t size) : fd_(fd), size_(size) {}
  ~WindowFile() override { if (fd_ >= 0) ::close(fd_); }
  void set_segments(const uint8_t* image, std::vector<RawParquetFile::Segment> segs) { win_ = image; segs_ = std::move(segs); }
  // image position of [position, position + nbytes) when one loaded segment holds all of it, else nullptr
  const uint8_t* loaded(int64_t position, int64_t nbytes) const {
    if (!win_ || segs_.empty()) return nullptr;
    auto it = std::upper_bound(segs_.begin(), segs_.end(), position,
                               [](int64_t p, const RawParquetFile::Segment& s) { return p < s.file_off; });
    if (it == segs_.begin()) return nullptr;
    --it;
    if (position + nbytes > it->file_off + it->len) return nullptr;
    return win_ + it->img_off + (position - it->file_off);
  }
  arrow::Status Close() override {
    if (fd_ >= 0) ::close(fd_);
    fd_ = -1;
    return arrow::Status::OK();
  }
  bool closed() const override { return fd_ < 0; }
  arrow::Result<int64_t> Tell() const override { return pos_; }
  arrow::Status Seek(int64_t position) override {
    if (position < 0 || position > size_) return arrow::Status::IOError("seek out of bounds");
    pos_ = position;
    return arrow::Status::OK();
  }
  arrow::Result<int64_t> GetSize() override { return size_; }
  arrow::Result<int64_t> Read(int64_t nbytes, void* out) override {
    ARROW_ASSIGN_OR_RAISE(int64_t n, ReadAt(pos_, nbytes, out));
    pos_ += n;
    return n;
  }
  arrow::Result<std::shared_ptr<arrow::Buffer>> Read(int64_t nbytes) override {
    ARROW_ASSIGN_OR_RAISE(auto b, ReadAt(pos_, nbytes));
    pos_ += b->size();
    return b;
  }
  arrow::Result<int64_t> ReadAt(int64_t position, int64_t nbytes, void* out) override {
    if (position < 0 || nbytes < 0) return arrow::Status::IOError("bad read");
    nbytes = std::max<int64_t>(0, std::min(nbytes, size_ - position));
    if (nbytes == 0) return 0;
    if (const uint8_t* src = loaded(position, nbytes)) {
      memcpy(out, src, (size_t)nbytes);
      return nbytes;
    }
    if (fd_ < 0) return arrow::Status::IOError("file is closed");
    int64_t got = 0;
    while (got < nbytes) {
      const ssize_t n = ::pread(fd_, static_cast<uint8_t*>(out) + got, (size_t)(nbytes - got), position + got);
      if (n <= 0) return arrow::Status::IOError("short read");
      got += n;
    }
    return got;
  }
  arrow::Result<std::shared_ptr<arrow::Buffer>> ReadAt(int64_t position, int64_t nbytes) override {
    if (position < 0 || nbytes < 0) return arrow::Status::IOError("bad read");
    nbytes = std::max<int64_t>(0, std::min(nbytes, size_ - position));
    if (nbytes > 0)
      if (const uint8_t* src = loaded(position, nbytes)) return std::make_shared<arrow::Buffer>(src, nbytes);  // the owner outlives every reader
    ARROW_ASSIGN_OR_RAISE(auto buf, arrow::AllocateResizableBuffer(nbytes));
    ARROW_ASSIGN_OR_RAISE(int64_t n, ReadAt(position, nbytes, buf->mutable_data()));
    if (n != nbytes) ARROW_RETURN_NOT_OK(buf->Resize(n));
    return std::shared_ptr<arrow::Buffer>(std::move(buf));
  }

 private:
  int fd_;
  int64_t size_, pos_ = 0;
  const uint8_t* win_ = nullptr;
  std::vector<RawParquetFile::Segment> segs_;
};

// footer -> row count, Arrow schema, row-group list
void parse_footer(const std::shared_ptr<arrow::io::RandomAccessFile>& src, std::shared_ptr<void>* reader_out, int64_t* rows,
                  std::shared_ptr<arrow::Schema>* schema, std::vector<int>* kept) {
  std::unique_ptr<parquet::ParquetFileReader> rd;
  try {
    rd = parquet::ParquetFileReader::Open(src);
  } catch (const std::exception& e) {
    fail(std::string("Parquet error: ") + e.what());
  }
  auto md = rd->metadata();
  *rows = md->num_rows();
  // The Arrow schema of a file written by an Arrow writer (parquet-rs ArrowWriter of the LakeSoul writer, Arrow C++) is
  // fully described by its ARROW:schema key-value entry: the 64 delta files of a bucket carry the same one, and
  // rebuilding it costs as much as parsing the footer (23 of 73 us per file, tools: r3 host profile). One cache per process.
  std::string schema_key;
  if (auto kv = md->key_value_metadata()) {
    auto v = kv->Get("ARROW:schema");
    if (v.ok()) schema_key = std::to_string(md->schema()->num_columns()) + ":" + *v;
  }
  static std::mutex cache_mu;
  static std::map<std::string, std::shared_ptr<arrow::Schema>> cache;
  if (!schema_key.empty()) {
    std::lock_guard<std::mutex> g(cache_mu);
    auto it = cache.find(schema_key);
    if (it != cache.end()) *schema = it->second;
  }
  if (!*schema) {
    auto st2 = parquet::arrow::FromParquetSchema(md->schema(), parquet::default_arrow_reader_properties(),
                                                 md->key_value_metadata(), schema);
    if (!st2.ok()) fail(st2.ToString());
    if (!schema_key.empty()) {
      std::lock_guard<std::mutex> g(cache_mu);
      if (cache.size() > 256) cache.clear();
      cache[schema_key] = *schema;
    }
  }
  kept->resize((size_t)md->num_row_groups());
  for (int rg = 0; rg < md->num_row_groups(); ++rg) (*kept)[(size_t)rg] = rg;
  *reader_out = std::shared_ptr<parquet::ParquetFileReader>(std::move(rd));
}
}  // namespace

std::shared_ptr<RawParquetFile> RawParquetFile::Open(const std::string& path) {
  auto f = std::make_shared<RawParquetFile>();
  f->path_ = path;
  const int fd = ::open(path.c_str(), O_RDONLY);
  if (fd < 0) fail("Object at location " + path + " not found");
  struct stat st;
  if (fstat(fd, &st) != 0) { ::close(fd); fail("cannot stat " + path); }
  f->image_bytes_ = f->file_bytes_ = st.st_size;
  f->image_ = pinned_alloc((size_t)st.st_size + 64);
  int64_t got = 0;
  while (got < st.st_size) {
    const ssize_t n = ::pread(fd, f->image_.get() + got, (size_t)(st.st_size - got), got);
    if (n <= 0) { ::close(fd); fail("short read on " + path); }
    got += n;
  }
  ::close(fd);
  f->segs_.push_back(Segment{0, f->image_bytes_, 0});
  auto buf = std::make_shared<arrow::Buffer>(f->image_.get(), f->image_bytes_);
  f->source_ = std::make_shared<arrow::io::BufferReader>(buf);
  parse_footer(f->source_, &f->reader_, &f->rows_, &f->schema_, &f->kept_rg_);
  f->live_hi_ = f->layout_rows_ = f->rows_;
  return f;
}

std::shared_ptr<RawParquetFile> RawParquetFile::OpenLazy(const std::string& path) {
  auto f = std::make_shared<RawParquetFile>();
  f->path_ = path;
  const int fd = ::open(path.c_str(), O_RDONLY);
  if (fd < 0) fail("Object at location " + path + " not found");
  struct stat st;
  if (fstat(fd, &st) != 0) { ::close(fd); fail("cannot stat " + path); }
  f->file_bytes_ = st.st_size;
  f->lazy_ = true;
  f->loaded_ = false;
  f->source_ = std::make_shared<WindowFile>(fd, (int64_t)st.st_size);  // owns fd
  parse_footer(f->source_, &f->reader_, &f->rows_, &f->schema_, &f->kept_rg_);
  f->live_hi_ = f->layout_rows_ = f->rows_;
  return f;
}

int64_t RawParquetFile::image_offset_of(int64_t file_off, int64_t len) const {
  auto it = std::upper_bound(segs_.begin(), segs_.end(), file_off, [](int64_t p, const Segment& s) { return p < s.file_off; });
  if (it == segs_.begin()) return -1;
  --it;
  if (file_off + len > it->file_off + it->len) return -1;
  return it->img_off + (file_off - it->file_off);
}

void RawParquetFile::LoadKept(const std::vector<std::string>* columns) {
  if (!lazy_ || loaded_) return;
  auto* rd = static_cast<parquet::ParquetFileReader*>(reader_.get());
  auto md = rd->metadata();
  const bool flat = md->schema()->num_columns() == schema_->num_fields();
  std::vector<char> wanted((size_t)md->schema()->num_columns(), 1);
  if (columns && flat) {
    std::fill(wanted.begin(), wanted.end(), 0);
    for (const auto& n : *columns) {
      const int ci = schema_->GetFieldIndex(n);
      if (ci >= 0) wanted[(size_t)ci] = 1;
    }
  }
  std::shared_ptr<parquet::PageIndexReader> pir;
  if (page_pruned_ && flat) {
    try {
      pir = rd->GetPageIndexReader();
    } catch (const std::exception&) {
      pir = nullptr;
    }
  }
  std::vector<Segment> segs;
  auto add = [&](int64_t off, int64_t len) {
    if (len > 0 && off >= 0 && off + len <= file_bytes_) segs.push_back(Segment{off, len, 0});
  };
  int64_t first_row = 0;
  for (int rg : kept_rg_) {
    auto rgm = md->RowGroup(rg);
    const int64_t rg_rows = rgm->num_rows();
    const bool live_rg = !page_pruned_ || (first_row < live_hi_ && first_row + rg_rows > live_lo_);
    std::shared_ptr<parquet::RowGroupPageIndexReader> rgi;
    if (pir && live_rg) {
      try {
        rgi = pir->RowGroup(rg);
      } catch (const std::exception&) {
        rgi = nullptr;
      }
    }
    for (int ci = 0; ci < rgm->num_columns() && live_rg; ++ci) {
      if (!wanted[(size_t)ci]) continue;
      auto ccm = rgm->ColumnChunk(ci);
      int64_t first = ccm->data_page_offset();
      if (ccm->has_dictionary_page() && ccm->dictionary_page_offset() > 0) first = std::min(first, ccm->dictionary_page_offset());
      if (first <= 0) continue;
      const int64_t bytes = ccm->total_compressed_size();
      std::shared_ptr<parquet::OffsetIndex> oi;
      if (rgi) {
        try {
          oi = rgi->GetOffsetIndex(ci);
        } catch (const std::exception&) {
          oi = nullptr;
        }
      }
      const int64_t lo_row = live_lo_ - first_row, hi_row = live_hi_ - first_row;  // inside this row group
      if (!oi || oi->page_locations().empty() || oi->page_locations()[0].offset < first ||
          oi->page_locations()[0].first_row_index != 0) {
        add(first, bytes);  // no (usable) offset index: the whole chunk
        continue;
      }
      const auto& locs = oi->page_locations();
      const size_t np = locs.size();
      size_t a = np, b = 0;
      for (size_t i = 0; i < np; ++i) {
        const int64_t start = locs[i].first_row_index;
        const int64_t end = i + 1 < np ? locs[i + 1].first_row_index : rg_rows;
        if (end > lo_row && start < hi_row) { a = std::min(a, i); b = i; }
      }
      if (a == np) continue;  // (cannot happen for a live row group)
      PageSpan sp;
      sp.dict_off = first;
      sp.dict_len = locs[0].offset - first;
      sp.data_off = locs[a].offset;
      sp.data_len = locs[b].offset + locs[b].compressed_page_size - locs[a].offset;
      sp.first_row = locs[a].first_row_index;
      sp.end_row = b + 1 < np ? locs[b + 1].first_row_index : rg_rows;
      sp.skipped_pages = (int32_t)(np - (b - a + 1));
      if (sp.data_off + sp.data_len > first + bytes || sp.data_len <= 0) {  // an index that disagrees with the footer
        add(first, bytes);
        continue;
      }
      spans_[{rg, ci}] = sp;
      add(sp.dict_off, sp.dict_len);
      add(sp.data_off, sp.data_len);
    }
    first_row += rg_rows;
  }
  // sorted, merged, laid out back to back
  std::sort(segs.begin(), segs.end(), [](const Segment& x, const Segment& y) { return x.file_off < y.file_off; });
  for (const auto& sg : segs) {
    if (!segs_.empty() && sg.file_off <= segs_.back().file_off + segs_.back().len)
      segs_.back().len = std::max(segs_.back().len, sg.file_off + sg.len - segs_.back().file_off);
    else
      segs_.push_back(sg);
  }
  int64_t total = 0;
  for (auto& sg : segs_) { sg.img_off = total; total += sg.len; }
  image_bytes_ = total;
  image_ = pinned_alloc((size_t)image_bytes_ + 64);
  auto* wf = static_cast<WindowFile*>(source_.get());
  for (const auto& sg : segs_) {
    auto r = wf->ReadAt(sg.file_off, sg.len, image_.get() + sg.img_off);
    if (!r.ok() || *r != sg.len) fail("short read on " + path_);
  }
  if (image_bytes_ > 0) wf->set_segments(image_.get(), segs_);
  loaded_ = true;
}

std::shared_ptr<arrow::io::RandomAccessFile> RawParquetFile::source() const { return source_; }

int64_t RawParquetFile::aux_append(const uint8_t* p, int64_t n) {
  if (aux_used_ + n + 64 > aux_cap_) {
    const int64_t cap = std::max<int64_t>((aux_used_ + n + 64) * 2, 1 << 20);
    auto nb = pinned_alloc((size_t)cap);
    if (aux_used_ > 0) memcpy(nb.get(), aux_.get(), (size_t)aux_used_);
    aux_ = nb;
    aux_cap_ = cap;
  }
  const int64_t off = aux_used_;
  memcpy(aux_.get() + off, p, (size_t)n);
  aux_used_ = (aux_used_ + n + 15) & ~15ll;
  return off;
}


// ---------------------------------------------------------------------------------------
// row-group pruning by primary-key statistics
// ---------------------------------------------------------------------------------------
namespace {

// what a predicate can evaluate to over the rows of one row group (NULL results never keep a row)
struct Possible { bool t, f; };

struct Bound {          // min / max of one column chunk, comparable with a literal of the same family
  enum Fam { kNone, kInt, kReal, kBytes } fam = kNone;
  int64_t imin = 0, imax = 0;
  double rmin = 0, rmax = 0;
  std::string bmin, bmax;
  int64_t nulls = -1;   // null count of the chunk, -1 unknown
  bool all_null = false;  // a page of nulls only (column index null_pages)
};

// min / max as the statistics or the column index hold them, before the column's logical type is applied
struct RawMinMax {
  int64_t imin = 0, imax = 0;
  double rmin = 0, rmax = 0;
  std::string bmin, bmax;
};

// applies the logical type: which family the values compare in, unsigned reinterpretation, timestamp units
void finish_bound(Bound* bp, parquet::Type::type phys, const arrow::DataType& type, const RawMinMax& m) {
  Bound& b = *bp;
  using AT = arrow::Type;
  switch (phys) {
    case parquet::Type::INT32: {
      if (!arrow::is_integer(type.id()) && type.id() != AT::DATE32 && type.id() != AT::TIME32) return;
      const bool uns = arrow::is_unsigned_integer(type.id());
      b.imin = uns ? (int64_t)(uint32_t)(int32_t)m.imin : m.imin;
      b.imax = uns ? (int64_t)(uint32_t)(int32_t)m.imax : m.imax;
      b.fam = Bound::kInt;
    } break;
    case parquet::Type::INT64: {
      if (type.id() == AT::UINT64) return;  // would need unsigned 64-bit compares: not worth a special case
      if (!arrow::is_integer(type.id()) && type.id() != AT::TIMESTAMP && type.id() != AT::TIME64) return;
      b.imin = m.imin; b.imax = m.imax; b.fam = Bound::kInt;
      if (type.id() == AT::TIMESTAMP) {
        // filter literals are microseconds (filter.cc parse_literal); the statistics are in the column's own
        // unit. Bring the bounds to microseconds, widening (never narrowing) the interval; no bound on overflow.
        const auto unit = static_cast<const arrow::TimestampType&>(type).unit();
        auto mul = [&](int64_t f) {
          if (b.imin > INT64_MAX / f || b.imin < INT64_MIN / f || b.imax > INT64_MAX / f || b.imax < INT64_MIN / f) {
            b.fam = Bound::kNone;
            return;
          }
          b.imin *= f; b.imax *= f;
        };
        auto floor_div = [](int64_t v, int64_t d) { return v / d - ((v % d != 0 && v < 0) ? 1 : 0); };
        if (unit == arrow::TimeUnit::SECOND) mul(1000000);
        else if (unit == arrow::TimeUnit::MILLI) mul(1000);
        else if (unit == arrow::TimeUnit::NANO) {
          const int64_t hi = b.imax;
          b.imin = floor_div(b.imin, 1000);
          b.imax = floor_div(hi, 1000) + (hi % 1000 != 0 ? 1 : 0);
        }
      }
    } break;
    case parquet::Type::FLOAT:
    case parquet::Type::DOUBLE:
      b.rmin = m.rmin; b.rmax = m.rmax; b.fam = Bound::kReal;
      break;
    case parquet::Type::BYTE_ARRAY:
      if (type.id() != AT::STRING && type.id() != AT::BINARY) return;
      b.bmin = m.bmin; b.bmax = m.bmax; b.fam = Bound::kBytes;
      break;
    default: break;
  }
  if (b.fam == Bound::kReal && (b.rmin != b.rmin || b.rmax != b.rmax)) b.fam = Bound::kNone;  // NaN bounds
}

Bound chunk_bound(const parquet::ColumnChunkMetaData& ccm, const parquet::ColumnDescriptor& cd, const arrow::DataType& type) {
  Bound b;
  if (!ccm.is_stats_set()) return b;
  auto st = ccm.statistics();
  if (!st) return b;
  if (st->HasNullCount()) b.nulls = st->null_count();
  else if (cd.max_definition_level() == 0) b.nulls = 0;
  if (!st->HasMinMax()) return b;
  RawMinMax m;
  switch (cd.physical_type()) {
    case parquet::Type::INT32: {
      auto ts = std::static_pointer_cast<parquet::Int32Statistics>(st);
      m.imin = ts->min(); m.imax = ts->max();
    } break;
    case parquet::Type::INT64: {
      auto ts = std::static_pointer_cast<parquet::Int64Statistics>(st);
      m.imin = ts->min(); m.imax = ts->max();
    } break;
    case parquet::Type::FLOAT: {
      auto ts = std::static_pointer_cast<parquet::FloatStatistics>(st);
      m.rmin = ts->min(); m.rmax = ts->max();
    } break;
    case parquet::Type::DOUBLE: {
      auto ts = std::static_pointer_cast<parquet::DoubleStatistics>(st);
      m.rmin = ts->min(); m.rmax = ts->max();
    } break;
    case parquet::Type::BYTE_ARRAY: {
      auto ts = std::static_pointer_cast<parquet::ByteArrayStatistics>(st);
      m.bmin.assign(reinterpret_cast<const char*>(ts->min().ptr), ts->min().len);
      m.bmax.assign(reinterpret_cast<const char*>(ts->max().ptr), ts->max().len);
    } break;
    default: return b;
  }
  finish_bound(&b, cd.physical_type(), type, m);
  return b;
}

// the same for page `i` of a column index (min / max of every page of a column chunk)
Bound page_bound(const parquet::ColumnIndex& ci, size_t i, const parquet::ColumnDescriptor& cd, const arrow::DataType& type) {
  Bound b;
  if (ci.has_null_counts() && i < ci.null_counts().size()) b.nulls = ci.null_counts()[i];
  else if (cd.max_definition_level() == 0) b.nulls = 0;
  if (i < ci.null_pages().size() && ci.null_pages()[i]) { b.all_null = true; return b; }
  RawMinMax m;
  switch (cd.physical_type()) {
    case parquet::Type::INT32: {
      const auto& t = static_cast<const parquet::Int32ColumnIndex&>(ci);
      m.imin = t.min_values()[i]; m.imax = t.max_values()[i];
    } break;
    case parquet::Type::INT64: {
      const auto& t = static_cast<const parquet::Int64ColumnIndex&>(ci);
      m.imin = t.min_values()[i]; m.imax = t.max_values()[i];
    } break;
    case parquet::Type::FLOAT: {
      const auto& t = static_cast<const parquet::FloatColumnIndex&>(ci);
      m.rmin = t.min_values()[i]; m.rmax = t.max_values()[i];
    } break;
    case parquet::Type::DOUBLE: {
      const auto& t = static_cast<const parquet::DoubleColumnIndex&>(ci);
      m.rmin = t.min_values()[i]; m.rmax = t.max_values()[i];
    } break;
    case parquet::Type::BYTE_ARRAY: {
      const auto& t = static_cast<const parquet::ByteArrayColumnIndex&>(ci);
      const auto& lo = t.min_values()[i];
      const auto& hi = t.max_values()[i];
      m.bmin.assign(reinterpret_cast<const char*>(lo.ptr), lo.len);
      m.bmax.assign(reinterpret_cast<const char*>(hi.ptr), hi.len);
    } break;
    default: return b;
  }
  finish_bound(&b, cd.physical_type(), type, m);
  return b;
}

// sign of (bound value - literal) for min and max; false when the literal is not comparable
bool compare_literal(const Bound& b, const arrow::Scalar& lit, int* cmin, int* cmax) {
  if (!lit.is_valid) return false;
  using AT = arrow::Type;
  auto sgn = [](auto a, auto c) { return a < c ? -1 : (a > c ? 1 : 0); };
  if (b.fam == Bound::kInt) {
    int64_t v;
    switch (lit.type->id()) {
      case AT::INT8: v = static_cast<const arrow::Int8Scalar&>(lit).value; break;
      case AT::INT16: v = static_cast<const arrow::Int16Scalar&>(lit).value; break;
      case AT::INT32: v = static_cast<const arrow::Int32Scalar&>(lit).value; break;
      case AT::INT64: v = static_cast<const arrow::Int64Scalar&>(lit).value; break;
      case AT::UINT8: v = static_cast<const arrow::UInt8Scalar&>(lit).value; break;
      case AT::UINT16: v = static_cast<const arrow::UInt16Scalar&>(lit).value; break;
      case AT::UINT32: v = static_cast<const arrow::UInt32Scalar&>(lit).value; break;
      case AT::DATE32: v = static_cast<const arrow::Date32Scalar&>(lit).value; break;
      case AT::TIMESTAMP: v = static_cast<const arrow::TimestampScalar&>(lit).value; break;
      default: return false;
    }
    *cmin = sgn(b.imin, v); *cmax = sgn(b.imax, v);
    return true;
  }
  if (b.fam == Bound::kReal) {
    double v;
    if (lit.type->id() == AT::FLOAT) v = static_cast<const arrow::FloatScalar&>(lit).value;
    else if (lit.type->id() == AT::DOUBLE) v = static_cast<const arrow::DoubleScalar&>(lit).value;
    else return false;
    if (v != v) return false;
    *cmin = sgn(b.rmin, v); *cmax = sgn(b.rmax, v);
    return true;
  }
  if (b.fam == Bound::kBytes) {
    if (lit.type->id() != AT::STRING && lit.type->id() != AT::BINARY) return false;
    const auto& bs = static_cast<const arrow::BaseBinaryScalar&>(lit);
    if (!bs.value) return false;
    const std::string v = bs.value->ToString();
    *cmin = sgn(b.bmin.compare(v), 0); *cmax = sgn(b.bmax.compare(v), 0);
    return true;
  }
  return false;
}

using BoundFn = std::function<Bound(const std::string&)>;

Possible possible(const FilterExpr& e, const BoundFn& bound_of) {
  const Possible unknown{true, true};
  switch (e.kind) {
    case FilterExpr::kTrue: return {true, false};
    case FilterExpr::kFalse: return {false, true};
    case FilterExpr::kAnd: {
      const Possible l = possible(*e.left, bound_of), r = possible(*e.right, bound_of);
      return {l.t && r.t, l.f || r.f};
    }
    case FilterExpr::kOr: {
      const Possible l = possible(*e.left, bound_of), r = possible(*e.right, bound_of);
      return {l.t || r.t, l.f && r.f};
    }
    case FilterExpr::kNot: {
      const Possible x = possible(*e.right, bound_of);
      return {x.f, x.t};
    }
    case FilterExpr::kCmp: {
      if (!e.literal) return unknown;
      const Bound b = bound_of(e.column);
      if (b.all_null) return {false, false};  // NULL for every row: neither TRUE nor FALSE
      int cmin = 0, cmax = 0;
      if (b.fam == Bound::kNone || !compare_literal(b, *e.literal, &cmin, &cmax)) return unknown;
      // values of the chunk lie in [min, max]; cmin/cmax = sign(min - lit), sign(max - lit)
      switch (e.op) {
        case FilterExpr::kEq: return {cmin <= 0 && cmax >= 0, !(cmin == 0 && cmax == 0)};
        case FilterExpr::kNe: return {!(cmin == 0 && cmax == 0), cmin <= 0 && cmax >= 0};
        case FilterExpr::kGt: return {cmax > 0, cmin <= 0};
        case FilterExpr::kGe: return {cmax >= 0, cmin < 0};
        case FilterExpr::kLt: return {cmin < 0, cmax >= 0};
        case FilterExpr::kLe: return {cmin <= 0, cmax > 0};
      }
      return unknown;
    }
    case FilterExpr::kIsNull: {  // a chunk proven free of nulls cannot satisfy it
      const Bound b = bound_of(e.column);
      if (b.all_null) return {true, false};
      return b.nulls == 0 ? Possible{false, true} : unknown;
    }
    case FilterExpr::kIsNotNull: {
      const Bound b = bound_of(e.column);
      if (b.all_null) return {false, true};
      return b.nulls == 0 ? Possible{true, false} : unknown;
    }
    default: return unknown;
  }
}

}  // namespace

int RawParquetFile::PruneRowGroups(const FilterExpr& filter, const std::set<std::string>& pk_columns) {
  auto* rd = static_cast<parquet::ParquetFileReader*>(reader_.get());
  auto md = rd->metadata();
  const auto* sd = md->schema();
  if (sd->num_columns() != schema_->num_fields()) return 0;  // nested schema: leaf index != field index
  std::vector<int> kept;
  int64_t rows = 0;
  for (int rg : kept_rg_) {
    auto rgm = md->RowGroup(rg);
    BoundFn bound_of = [&](const std::string& name) -> Bound {
      if (!pk_columns.count(name)) return Bound{};
      const int ci = schema_->GetFieldIndex(name);
      if (ci < 0) return Bound{};
      return chunk_bound(*rgm->ColumnChunk(ci), *sd->Column(ci), *schema_->field(ci)->type());
    };
    if (possible(filter, bound_of).t) {
      kept.push_back(rg);
      rows += rgm->num_rows();
    }
  }
  const int dropped = (int)(kept_rg_.size() - kept.size());
  kept_rg_ = std::move(kept);
  rows_ = rows;
  live_lo_ = 0;
  live_hi_ = layout_rows_ = rows_;
  return dropped;
}

int64_t RawParquetFile::PrunePages(const FilterExpr& filter, const std::set<std::string>& pk_columns) {
  auto* rd = static_cast<parquet::ParquetFileReader*>(reader_.get());
  auto md = rd->metadata();
  const auto* sd = md->schema();
  if (sd->num_columns() != schema_->num_fields() || kept_rg_.empty()) return 0;
  std::set<std::string> fcols;
  filter_columns(filter, &fcols);
  std::vector<int> key_cis;
  for (const auto& n : fcols)
    if (pk_columns.count(n) && schema_->GetFieldIndex(n) >= 0) key_cis.push_back(schema_->GetFieldIndex(n));
  if (key_cis.empty()) return 0;
  std::shared_ptr<parquet::PageIndexReader> pir;
  try {
    pir = rd->GetPageIndexReader();
  } catch (const std::exception&) {
    return 0;
  }
  if (!pir) return 0;
  int64_t lo = INT64_MAX, hi = -1;  // hull of the rows that can match, in kept-row coordinates
  int64_t rg_first = 0;
  bool any_index = false;
  for (int rg : kept_rg_) {
    const int64_t rg_rows = md->RowGroup(rg)->num_rows();
    int64_t a = 0, b = rg_rows;  // this row group: intersection over the key columns of the hulls of their possible pages
    std::shared_ptr<parquet::RowGroupPageIndexReader> rgi;
    try {
      rgi = pir->RowGroup(rg);
    } catch (const std::exception&) {
      rgi = nullptr;
    }
    for (int ci : key_cis) {
      if (!rgi) break;
      std::shared_ptr<parquet::ColumnIndex> cidx;
      std::shared_ptr<parquet::OffsetIndex> oidx;
      try {
        cidx = rgi->GetColumnIndex(ci);
        oidx = rgi->GetOffsetIndex(ci);
      } catch (const std::exception&) {
        continue;
      }
      if (!cidx || !oidx) continue;
      const auto& locs = oidx->page_locations();
      const size_t np = locs.size();
      if (np == 0 || cidx->null_pages().size() != np) continue;
      any_index = true;
      const std::string& cname = schema_->field(ci)->name();
      int64_t ca = rg_rows, cb = 0;
      for (size_t i = 0; i < np; ++i) {
        BoundFn bound_of = [&](const std::string& name) -> Bound {
          if (name != cname) return Bound{};
          return page_bound(*cidx, i, *sd->Column(ci), *schema_->field(ci)->type());
        };
        if (!possible(filter, bound_of).t) continue;
        const int64_t start = locs[i].first_row_index;
        const int64_t end = i + 1 < np ? locs[i + 1].first_row_index : rg_rows;
        ca = std::min(ca, start);
        cb = std::max(cb, end);
      }
      a = std::max(a, ca);
      b = std::min(b, cb);
      if (a >= b) break;
    }
    if (a < b) {
      lo = std::min(lo, rg_first + a);
      hi = std::max(hi, rg_first + b);
    }
    rg_first += rg_rows;
  }
  if (!any_index) return 0;
  if (hi < 0) lo = hi = 0;  // no page can hold a match
  page_pruned_ = true;
  live_lo_ = lo;
  live_hi_ = hi;
  return rows_ - (hi - lo);
}

void RawParquetFile::KeepRowGroups(int lo, int hi) {
  auto* rd = static_cast<parquet::ParquetFileReader*>(reader_.get());
  auto md = rd->metadata();
  std::vector<int> kept;
  int64_t rows = 0;
  for (int rg : kept_rg_)
    if (rg >= lo && rg < hi) { kept.push_back(rg); rows += md->RowGroup(rg)->num_rows(); }
  kept_rg_ = std::move(kept);
  rows_ = rows;
  live_lo_ = 0;
  live_hi_ = layout_rows_ = rows_;
}

std::vector<RowGroupKeyRange> RawParquetFile::KeyRanges(const std::string& column) const {
  std::vector<RowGroupKeyRange> out;
  auto* rd = static_cast<parquet::ParquetFileReader*>(reader_.get());
  auto md = rd->metadata();
  const auto* sd = md->schema();
  if (sd->num_columns() != schema_->num_fields()) return out;
  const int ci = schema_->GetFieldIndex(column);
  if (ci < 0) return out;
  for (int rg : kept_rg_) {
    auto rgm = md->RowGroup(rg);
    const Bound b = chunk_bound(*rgm->ColumnChunk(ci), *sd->Column(ci), *schema_->field(ci)->type());
    RowGroupKeyRange r;
    r.rows = rgm->num_rows();
    r.nulls = b.nulls;
    r.bytes = rgm->total_byte_size();
    if (b.fam == Bound::kInt) { r.fam = RowGroupKeyRange::kInt; r.imin = b.imin; r.imax = b.imax; }
    else if (b.fam == Bound::kBytes) { r.fam = RowGroupKeyRange::kBytes; r.bmin = b.bmin; r.bmax = b.bmax; }
    out.push_back(std::move(r));
  }
  return out;
}

bool RawParquetFile::Plan(const std::vector<std::shared_ptr<arrow::Field>>& want, std::string* why) {
  if (lazy_ && !loaded_) LoadKept();
  // a chunk [first, first + bytes) of the file lies inside the image?
  // (what LoadKept() read of it: all of it, or -- page-pruned with an offset index -- its dictionary and live pages)
  auto in_window = [&](int rg, int ci, int64_t first, int64_t bytes) {
    return first > 0 && (spans_.count({rg, ci}) > 0 || image_offset_of(first, bytes) >= 0);
  };
  auto* rd = static_cast<parquet::ParquetFileReader*>(reader_.get());
  auto md = rd->metadata();
  const auto* sd = md->schema();
  cols_.clear();
  cols_.resize(want.size());
  pruned_pages_ = 0;
  int64_t emit_lo = INT64_MAX, emit_hi = -1;  // rows covered by the framed pages (kept-row coordinates)
  // flat schemas only: leaf index == field index
  if (sd->num_columns() != schema_->num_fields()) { *why = "nested schema"; return false; }
  // which compressed column chunks the GPU expands. SNAPPY: all of them; ZSTD: by the configured share / size policy
  // the metadata objects of every (kept row group, column) are built once (libparquet allocates one per call)
  const int ncols_file = sd->num_columns();
  const int nrg_all = md->num_row_groups();
  std::vector<std::unique_ptr<parquet::RowGroupMetaData>> rg_md((size_t)nrg_all);
  std::vector<std::unique_ptr<parquet::ColumnChunkMetaData>> cc_md((size_t)nrg_all * (size_t)ncols_file);
  auto chunk_md = [&](int rg, int ci) -> parquet::ColumnChunkMetaData* {
    auto& slot = cc_md[(size_t)rg * (size_t)ncols_file + (size_t)ci];
    if (!slot) {
      if (!rg_md[(size_t)rg]) rg_md[(size_t)rg] = md->RowGroup(rg);
      slot = rg_md[(size_t)rg]->ColumnChunk(ci);
    }
    return slot.get();
  };
  // (row group, column) -> expanded on the device? with which codec?
  struct ChunkTable {
    std::vector<uint8_t> device, codec;
    int ncols;
    size_t at(int rg, int ci) const { return (size_t)rg * (size_t)ncols + (size_t)ci; }
    struct Ref { int rg, ci; };
    void insert(const Ref& r) { device[at(r.rg, r.ci)] = 1; }
    size_t count(const Ref& r) const { return device[at(r.rg, r.ci)]; }
  } device_chunks{std::vector<uint8_t>((size_t)nrg_all * (size_t)ncols_file, 0), std::vector<uint8_t>((size_t)nrg_all * (size_t)ncols_file, 0), ncols_file};
  auto& chunk_codec_v = device_chunks.codec;
  if (device_snappy()) {
    for (size_t w = 0; w < want.size(); ++w) {
      const int ci = schema_->GetFieldIndex(want[w]->name());
      if (ci < 0) continue;
      for (int rg : kept_rg_) {
        auto* ccm = chunk_md(rg, ci);
        const int64_t first = ccm->has_dictionary_page() ? ccm->dictionary_page_offset() : ccm->data_page_offset();
        if (ccm->compression() != parquet::Compression::SNAPPY || !in_window(rg, ci, first, ccm->total_compressed_size())) continue;
        device_chunks.insert({rg, ci});
        chunk_codec_v[device_chunks.at(rg, ci)] = kCodecSnappy;
      }
    }
  }
  if (device_zstd_share() != 0) {
    std::vector<std::pair<int64_t, std::pair<int, int>>> cand;
    int64_t total = 0;
    for (size_t w = 0; w < want.size(); ++w) {
      const int ci = schema_->GetFieldIndex(want[w]->name());
      if (ci < 0) continue;
      for (int rg : kept_rg_) {
        auto* ccm = chunk_md(rg, ci);
        const int64_t first = ccm->has_dictionary_page() ? ccm->dictionary_page_offset() : ccm->data_page_offset();
        if (ccm->compression() != parquet::Compression::ZSTD || !in_window(rg, ci, first, ccm->total_compressed_size())) continue;
        chunk_codec_v[device_chunks.at(rg, ci)] = kCodecZstd;
        cand.push_back({ccm->total_uncompressed_size(), {rg, ci}});
        total += ccm->total_uncompressed_size();
      }
    }
    std::sort(cand.begin(), cand.end());
    if (device_zstd_share() < 0) {  // default policy: by chunk size
      for (const auto& c : cand)
        if (c.first >= device_zstd_min_bytes()) device_chunks.insert({c.second.first, c.second.second});
    } else {
      const double budget = device_zstd_share() * (double)total;
      int64_t acc = 0;
      for (const auto& c : cand) {
        if (device_zstd_share() < 1.0 && (double)(acc + c.first) > budget) break;
        device_chunks.insert({c.second.first, c.second.second});
        acc += c.first;
      }
    }
  }
  for (size_t w = 0; w < want.size(); ++w) {
    RawColumn& rc = cols_[w];
    rc.name = want[w]->name();
    const int ci = schema_->GetFieldIndex(rc.name);
    if (ci < 0) continue;
    rc.present = true;
    const auto& ftype = schema_->field(ci)->type();
    const auto& ttype = want[w]->type();
    if (!ftype->Equals(*ttype)) { *why = "column " + rc.name + " needs a cast (" + ftype->ToString() + " -> " + ttype->ToString() + ")"; return false; }
    const parquet::ColumnDescriptor* cd = sd->Column(ci);
    if (cd->max_repetition_level() != 0 || cd->max_definition_level() > 1) { *why = "nested column " + rc.name; return false; }
    PqChunk& pc = rc.proto;
    pc.max_def = cd->max_definition_level();
    rc.optional = pc.max_def > 0;
    pc.type_length = cd->type_length();
    pc.dict_nvals = 0;
    using AT = arrow::Type;
    const AT::type tid = ttype->id();
    auto fixed_copy = [&](int phys, int width) { pc.phys = phys; pc.conv = CONV_COPY; pc.out_width = width; };
    switch (cd->physical_type()) {
      case parquet::Type::BOOLEAN:
        if (tid != AT::BOOL) { *why = "boolean mapping"; return false; }
        pc.phys = PQ_BOOLEAN; pc.conv = CONV_BOOL; pc.out_width = 0; break;
      case parquet::Type::INT32:
        if (tid == AT::DECIMAL128) { pc.phys = PQ_INT32; pc.conv = CONV_DEC_INT; pc.out_width = 16; }
        else if (tid == AT::INT8 || tid == AT::UINT8) fixed_copy(PQ_INT32, 1);
        else if (tid == AT::INT16 || tid == AT::UINT16) fixed_copy(PQ_INT32, 2);
        else if (tid == AT::INT32 || tid == AT::UINT32 || tid == AT::DATE32 || tid == AT::TIME32) fixed_copy(PQ_INT32, 4);
        else { *why = "INT32 -> " + ttype->ToString(); return false; }
        break;
      case parquet::Type::INT64:
        if (tid == AT::DECIMAL128) { pc.phys = PQ_INT64; pc.conv = CONV_DEC_INT; pc.out_width = 16; }
        else if (tid == AT::INT64 || tid == AT::UINT64 || tid == AT::TIMESTAMP || tid == AT::TIME64 || tid == AT::DURATION)
          fixed_copy(PQ_INT64, 8);
        else { *why = "INT64 -> " + ttype->ToString(); return false; }
        break;
      case parquet::Type::FLOAT:
        if (tid != AT::FLOAT) { *why = "FLOAT mapping"; return false; }
        fixed_copy(PQ_FLOAT, 4); break;
      case parquet::Type::DOUBLE:
        if (tid != AT::DOUBLE) { *why = "DOUBLE mapping"; return false; }
        fixed_copy(PQ_DOUBLE, 8); break;
      case parquet::Type::BYTE_ARRAY:
        if (tid != AT::STRING && tid != AT::BINARY) { *why = "BYTE_ARRAY -> " + ttype->ToString(); return false; }
        pc.phys = PQ_BYTE_ARRAY; pc.conv = CONV_VAR; pc.out_width = 0; break;
      case parquet::Type::FIXED_LEN_BYTE_ARRAY:
        pc.phys = PQ_FLBA;
        if (tid == AT::DECIMAL128 && pc.type_length <= 16) { pc.conv = CONV_DEC_FLBA; pc.out_width = 16; }
        else if (tid == AT::FIXED_SIZE_BINARY) {
          const int wd = pc.type_length;
          if (!(wd == 1 || wd == 2 || wd == 4 || wd == 8 || wd == 16 || wd == 32)) { *why = "fixed_size_binary width"; return false; }
          pc.conv = CONV_FLBA_RAW; pc.out_width = wd;
        } else { *why = "FLBA -> " + ttype->ToString(); return false; }
        break;
      default: *why = "INT96 column " + rc.name; return false;
    }
    // pages
    int64_t first_row = 0;
    for (int rg : kept_rg_) {
      if (page_pruned_) {  // a row group outside the rows that can match: nothing of it is framed
        if (!rg_md[(size_t)rg]) rg_md[(size_t)rg] = md->RowGroup(rg);
        const int64_t n = rg_md[(size_t)rg]->num_rows();
        if (first_row + n <= live_lo_ || first_row >= live_hi_) { first_row += n; continue; }
      }
      auto* ccm = chunk_md(rg, ci);
      for (auto e : ccm->encodings())
        if (map_encoding(e) < 0) { *why = "encoding " + parquet::EncodingToString(e) + " in column " + rc.name; return false; }
      if (rc.optional) {
        bool proven = false;
        if (ccm->is_stats_set()) {
          auto stats = ccm->statistics();
          if (stats && stats->HasNullCount() && stats->null_count() == 0) proven = true;
        }
        if (!proven) rc.may_have_nulls = true;
      }
      // ZSTD (what the LakeSoul writer emits) and SNAPPY chunks chosen above: the pages stay compressed in the file
      // image and are expanded by the GPU (zstd_pages_kernel_v2 / snappy_pages_kernel); uncompressed chunks are used
      // in place; everything else is decompressed here by Arrow's codecs (parquet::PageReader).
      bool raw_zstd = device_chunks.count({rg, ci}) > 0;  // (name kept: "raw" = still compressed, either codec)
      const uint8_t codec = raw_zstd ? chunk_codec_v[device_chunks.at(rg, ci)] : (uint8_t)0;
      const int64_t chunk_first = ccm->has_dictionary_page() ? ccm->dictionary_page_offset() : ccm->data_page_offset();
      const int64_t chunk_bytes = ccm->total_compressed_size();
      const bool in_image = in_window(rg, ci, chunk_first, chunk_bytes);
      const PageSpan* span = nullptr;
      if (auto it = spans_.find({rg, ci}); it != spans_.end()) span = &it->second;
      // own header walk (read_page_head) when the payloads are not touched on the host
      bool walk = in_image && (raw_zstd || ccm->compression() == parquet::Compression::UNCOMPRESSED);
      // the walk covers the chunk, or the two loaded pieces of it: dictionary page(s), then the live data pages
      const uint8_t* wp = nullptr;
      const uint8_t* wend = nullptr;
      int walk_phase = 0;  // spans only: 0 = dictionary piece, 1 = data piece
      int64_t row_in_rg = 0;
      if (walk && !span) {
        wp = image_.get() + image_offset_of(chunk_first, chunk_bytes);
        wend = wp + chunk_bytes;
      } else if (walk) {
        const int64_t o = span->dict_len > 0 ? image_offset_of(span->dict_off, span->dict_len) : 0;
        wp = image_.get() + o;
        wend = wp + span->dict_len;
      }
      std::unique_ptr<parquet::PageReader> pr;
      auto open_libparquet = [&]() -> bool {
        try {
          pr = rd->RowGroup(rg)->GetColumnPageReader(ci);
        } catch (const std::exception& e) {
          *why = std::string("page reader: ") + e.what();
          return false;
        }
        walk = false;
        return true;
      };
      if (!walk) {
        raw_zstd = false;  // (a device chunk outside the image cannot happen: the candidates were checked above)
        if (!open_libparquet()) return false;
      }
      struct PageView {
        int type = -1;
        const uint8_t* data = nullptr;
        int64_t size = 0, uncompressed = 0;
        int32_t num_values = 0, v2_def_len = 0, v2_rep_len = 0;
        parquet::Encoding::type enc = parquet::Encoding::PLAIN, def_enc = parquet::Encoding::RLE;
      };
      std::shared_ptr<parquet::Page> hold;
      // 1 = a page, 0 = end of the chunk, -1 = error (*why set)
      auto next_page = [&](PageView* v) -> int {
        if (walk) {
          while (true) {
            if (wp >= wend && span && walk_phase == 0) {  // on to the live data pages
              walk_phase = 1;
              wp = image_.get() + image_offset_of(span->data_off, span->data_len);
              wend = wp + span->data_len;
              row_in_rg = span->first_row;
            }
            if (wp >= wend) return 0;
            PageHead h;
            if (!read_page_head(wp, wend, &h) || wp + h.header_len + h.compressed > wend) { *why = "malformed page header in column " + rc.name; return -1; }
            const uint8_t* payload = wp + h.header_len;
            wp = payload + h.compressed;
            if (h.type == 1) continue;  // index page: skipped, as libparquet does
            if (!h.have_sub || h.num_values < 0 || h.encoding < 0) { *why = "page header without its page-type section in column " + rc.name; return -1; }
            v->type = h.type;
            v->data = payload;
            v->size = h.compressed;
            v->uncompressed = h.uncompressed;
            v->num_values = h.num_values;
            v->enc = (parquet::Encoding::type)h.encoding;
            v->def_enc = h.def_encoding >= 0 ? (parquet::Encoding::type)h.def_encoding : parquet::Encoding::RLE;
            v->v2_def_len = h.v2_def_len;
            v->v2_rep_len = h.v2_rep_len;
            return 1;
          }
        }
        try {
          hold = pr->NextPage();
        } catch (const std::exception& e) {
          *why = std::string("page decode: ") + e.what();
          return -1;
        }
        if (!hold) return 0;
        v->data = hold->data();
        v->size = hold->size();
        v->type = (int)hold->type();
        if (hold->type() == parquet::PageType::DICTIONARY_PAGE) {
          auto* dp = static_cast<parquet::DictionaryPage*>(hold.get());
          v->num_values = dp->num_values();
          v->enc = dp->encoding();
        } else if (hold->type() == parquet::PageType::DATA_PAGE) {
          auto* p1 = static_cast<parquet::DataPageV1*>(hold.get());
          v->num_values = p1->num_values();
          v->enc = p1->encoding();
          v->def_enc = p1->definition_level_encoding();
          v->uncompressed = p1->uncompressed_size();
        } else if (hold->type() == parquet::PageType::DATA_PAGE_V2) {
          auto* p2 = static_cast<parquet::DataPageV2*>(hold.get());
          v->num_values = p2->num_values();
          v->enc = p2->encoding();
          v->v2_def_len = p2->definition_levels_byte_length();
          v->v2_rep_len = p2->repetition_levels_byte_length();
          v->uncompressed = p2->uncompressed_size();
        }
        return 1;
      };
      // a chunk the device cannot expand after all is started over with libparquet's decompressing reader
      auto restart_on_host = [&]() -> bool {
        raw_zstd = false;
        return open_libparquet();
      };
      int dict_idx = -1;
      int chunk_pages = 0;  // data pages of this chunk framed so far
      while (true) {
        PageView page;
        const int got = next_page(&page);
        if (got < 0) return false;
        if (got == 0) break;
        const uint8_t* data = page.data;
        const int64_t size = page.size;
        auto place = [&](int64_t* off, bool* aux) {
          if (data >= image_.get() && data + size <= image_.get() + image_bytes_) {
            *off = data - image_.get();
            *aux = false;
          } else {
            *off = aux_append(data, size);
            *aux = true;
          }
        };
        if (page.type == (int)parquet::PageType::DICTIONARY_PAGE) {
          if (page.enc != parquet::Encoding::PLAIN && page.enc != parquet::Encoding::PLAIN_DICTIONARY) {
            *why = "dictionary page encoding"; return false;
          }
          RawColumn::Dict d;
          place(&d.src_off, &d.aux);
          d.size = (int32_t)size;
          d.num_values = page.num_values;
          if (raw_zstd) {
            const int64_t usz = d.aux ? -1 : announced_size(codec, data, size);
            if (usz < 0 || usz > ((int64_t)1 << 30)) {
              // the payload does not say how large it gets: let libparquet decompress this chunk
              if (!restart_on_host()) return false;
              continue;
            }
            d.codec = codec;
            d.comp_size = (int32_t)size;
            d.size = (int32_t)usz;
          }
          rc.dicts.push_back(d);
          dict_idx = (int)rc.dicts.size() - 1;
        } else if (page.type == (int)parquet::PageType::DATA_PAGE || page.type == (int)parquet::PageType::DATA_PAGE_V2) {
          const bool v2 = page.type == (int)parquet::PageType::DATA_PAGE_V2;
          RawColumn::Page pg;
          pg.num_values = page.num_values;
          const parquet::Encoding::type enc = page.enc;
          if (!v2) {
            pg.def_encoding = map_encoding(page.def_enc);
            pg.v2_def_len = -1;
          } else {
            pg.def_encoding = ENC_RLE;
            pg.v2_def_len = page.v2_def_len;
            if (page.v2_rep_len != 0) { *why = "repetition levels"; return false; }
          }
          pg.encoding = map_encoding(enc);
          const bool is_dict = pg.encoding == ENC_PLAIN_DICT || pg.encoding == ENC_RLE_DICT;
          const bool rle_bool = pg.encoding == ENC_RLE && pc.conv == CONV_BOOL;
          if (!(pg.encoding == ENC_PLAIN || is_dict || rle_bool)) { *why = "value encoding " + parquet::EncodingToString(enc) + " in column " + rc.name; return false; }
          if (is_dict && dict_idx < 0) { *why = "dictionary page missing"; return false; }
          if (pc.conv == CONV_BOOL && is_dict) { *why = "dictionary coded boolean"; return false; }
          if (raw_zstd && v2 && chunk_pages == 0) {
            // v2 pages keep their levels outside the compressed section: let libparquet decompress this chunk
            // (from its first page: nothing of it was framed yet)
            if (!restart_on_host()) return false;
            row_in_rg = 0;
            continue;
          }
          if (page_pruned_ && (first_row + row_in_rg + pg.num_values <= live_lo_ || first_row + row_in_rg >= live_hi_)) {
            row_in_rg += pg.num_values;  // no row of this page can match the pushed key predicate
            pruned_pages_ += 1;
            continue;
          }
          place(&pg.src_off, &pg.aux);
          pg.size = (int32_t)size;
          if (raw_zstd) {
            if (v2 || pg.aux || page.uncompressed <= 0 || page.uncompressed > (int64_t)1 << 30) {
              *why = "ZSTD page shape in column " + rc.name;
              return false;
            }
            pg.codec = codec;
            pg.comp_size = (int32_t)size;
            pg.size = (int32_t)page.uncompressed;
          }
          pg.row_group = rg;
          pg.first_row = first_row + row_in_rg;
          emit_lo = std::min(emit_lo, pg.first_row);
          emit_hi = std::max(emit_hi, pg.first_row + pg.num_values);
          row_in_rg += pg.num_values;
          chunk_pages += 1;
          rc.pages.push_back(pg);
          rc.page_dict.push_back(is_dict ? dict_idx : -1);
        }
      }
      if (walk && span) pruned_pages_ += span->skipped_pages;  // pages that were never read
      if (row_in_rg != (walk && span ? span->end_row : rg_md[(size_t)rg]->num_rows())) { *why = "page row count mismatch in column " + rc.name; return false; }
      first_row += rg_md[(size_t)rg]->num_rows();
    }
  }
  // the file's rows in the scan = what was framed; pages count from there
  if (page_pruned_) {
    origin_ = emit_lo <= emit_hi ? emit_lo : live_lo_;
    // the rows the merge reads start on a 64-row boundary of the scan's row space, like an unpruned file's do (files are
    // laid out on 64-row boundaries, device_scan.cc): up to 63 unused rows in front of the first framed page
    if (const int64_t mis = (live_lo_ - origin_) % 64) origin_ -= 64 - mis;
    layout_rows_ = emit_lo <= emit_hi ? emit_hi - origin_ : 0;
    for (auto& c : cols_)
      for (auto& pg : c.pages) pg.first_row -= origin_;
  } else {
    origin_ = 0;
    layout_rows_ = rows_;
  }
  return true;
}

// test hook (tests/test_lazy_files.py through ctypes): frames the row groups [rg_lo, rg_hi) of every flat column of a
// file, from a whole-file image (lazy = 0) or from the loaded byte span of those row groups (lazy = 1), and prints
// one line per page -- position-independent, with a hash of the payload bytes the device would read -- so that the two
// can be compared. Returns the text length, -1 on error (message in `out`).
extern "C" __attribute__((visibility("default"))) int64_t lsb_test_plan_digest(const char* path, int32_t lazy, int32_t rg_lo, int32_t rg_hi,
                                                                               char* out, int64_t cap) {
  std::string text;
  try {
    auto f = lazy ? RawParquetFile::OpenLazy(path) : RawParquetFile::Open(path);
    f->KeepRowGroups(rg_lo, rg_hi);
    f->LoadKept();
    std::vector<std::shared_ptr<arrow::Field>> want = f->schema()->fields();
    std::string why;
    if (!f->Plan(want, &why)) throw std::runtime_error("not framed: " + why);
    auto hash = [&](bool aux, int64_t off, int64_t n) {
      const uint8_t* p = (aux ? f->aux() : f->image()) + off;
      uint64_t h = 1469598103934665603ull;
      for (int64_t i = 0; i < n; ++i) h = (h ^ p[i]) * 1099511628211ull;
      return h;
    };
    text += "rows " + std::to_string(f->num_rows()) + " window " + std::to_string(f->image_bytes()) + " of " + std::to_string(f->file_bytes()) + "\n";
    for (const auto& c : f->columns()) {
      for (size_t i = 0; i < c.dicts.size(); ++i) {
        const auto& d = c.dicts[i];
        text += "D " + c.name + " " + std::to_string(i) + " " + std::to_string(d.num_values) + " " + std::to_string(d.size) + " " +
                std::to_string((int)d.codec) + " " + std::to_string(hash(d.aux, d.src_off, d.codec ? d.comp_size : d.size)) + "\n";
      }
      for (size_t i = 0; i < c.pages.size(); ++i) {
        const auto& g = c.pages[i];
        text += "P " + c.name + " " + std::to_string(g.row_group) + " " + std::to_string(g.first_row) + " " + std::to_string(g.num_values) + " " +
                std::to_string(g.size) + " " + std::to_string(g.encoding) + " " + std::to_string((int)g.codec) + " " + std::to_string(c.page_dict[i]) + " " +
                std::to_string(hash(g.aux, g.src_off, g.codec ? g.comp_size : g.size)) + "\n";
      }
    }
  } catch (const std::exception& e) {
    snprintf(out, (size_t)cap, "%s", e.what());
    return -1;
  }
  if ((int64_t)text.size() + 1 > cap) { snprintf(out, (size_t)cap, "digest needs %zu bytes", text.size() + 1); return -1; }
  memcpy(out, text.c_str(), text.size() + 1);
  return (int64_t)text.size();
}

// test hook (tests/test_page_pruning.py): row groups, then pages, pruned by `filter` (parquet-mr string form) over the key
// columns `pks` (comma separated); prints the live range and the framed pages of every column (with a hash of the
// payload bytes the device would read). lazy = 1: footer first, then only the needed byte segments of the file.
extern "C" __attribute__((visibility("default"))) int64_t lsb_test_page_prune(const char* path, const char* filter, const char* pks, int32_t lazy,
                                                                              char* out, int64_t cap) {
  std::string text;
  try {
    auto f = lazy ? RawParquetFile::OpenLazy(path) : RawParquetFile::Open(path);
    std::set<std::string> pk;
    std::string cur;
    for (const char* p = pks; ; ++p) {
      if (*p == ',' || *p == 0) { if (!cur.empty()) pk.insert(cur); cur.clear(); if (*p == 0) break; }
      else cur.push_back(*p);
    }
    auto fe = parse_filter(filter, *f->schema());
    const int dropped = f->PruneRowGroups(*fe, pk);
    const int64_t outside = f->PrunePages(*fe, pk);
    f->LoadKept();
    std::vector<std::shared_ptr<arrow::Field>> want = f->schema()->fields();
    std::string why;
    if (!f->Plan(want, &why)) throw std::runtime_error("not framed: " + why);
    auto hash = [&](bool aux, int64_t off, int64_t n) {
      const uint8_t* p = (aux ? f->aux() : f->image()) + off;
      uint64_t h = 1469598103934665603ull;
      for (int64_t i = 0; i < n; ++i) h = (h ^ p[i]) * 1099511628211ull;
      return h;
    };
    text += "dropped_rg " + std::to_string(dropped) + " first_kept_rg " + std::to_string(f->kept_row_groups().empty() ? -1 : f->kept_row_groups()[0]) +
            " kept_rows " + std::to_string(f->num_rows()) + " outside " + std::to_string(outside) + " live_begin " +
            std::to_string(f->live_begin()) + " live_rows " + std::to_string(f->live_rows()) + " live_offset " + std::to_string(f->live_offset()) +
            " layout_rows " + std::to_string(f->layout_rows()) + " pruned_pages " + std::to_string(f->pruned_pages()) + " image_bytes " +
            std::to_string(f->image_bytes()) + " file_bytes " + std::to_string(f->file_bytes()) + "\n";
    text += "K";
    for (int rg : f->kept_row_groups()) text += " " + std::to_string(rg);
    text += "\n";
    for (const auto& c : f->columns()) {
      text += "C " + c.name;
      for (const auto& g : c.pages)
        text += " " + std::to_string(g.first_row) + ":" + std::to_string(g.num_values) + ":" + std::to_string(hash(g.aux, g.src_off, g.codec ? g.comp_size : g.size));
      text += "\n";
    }
  } catch (const std::exception& e) {
    snprintf(out, (size_t)cap, "%s", e.what());
    return -1;
  }
  if ((int64_t)text.size() + 1 > cap) { snprintf(out, (size_t)cap, "needs %zu bytes", text.size() + 1); return -1; }
  memcpy(out, text.c_str(), text.size() + 1);
  return (int64_t)text.size();
}

}  // namespace lsb
