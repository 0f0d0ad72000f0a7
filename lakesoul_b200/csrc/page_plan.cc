// SPDX-License-Identifier: Apache-2.0
#include "page_plan.h"

#include <arrow/io/memory.h>
#include <fcntl.h>
#include <parquet/arrow/schema.h>
#include <parquet/column_page.h>
#include <parquet/column_reader.h>
#include <parquet/file_reader.h>
#include <parquet/metadata.h>
#include <parquet/statistics.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstring>
#include <stdexcept>

namespace lsb {

namespace {
[[noreturn]] void fail(const std::string& m) { throw std::runtime_error(m); }

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

std::shared_ptr<RawParquetFile> RawParquetFile::Open(const std::string& path) {
  auto f = std::make_shared<RawParquetFile>();
  f->path_ = path;
  const int fd = ::open(path.c_str(), O_RDONLY);
  if (fd < 0) fail("Object at location " + path + " not found");
  struct stat st;
  if (fstat(fd, &st) != 0) { ::close(fd); fail("cannot stat " + path); }
  f->image_bytes_ = st.st_size;
  f->image_ = pinned_alloc((size_t)st.st_size + 64);
  int64_t got = 0;
  while (got < st.st_size) {
    const ssize_t n = ::pread(fd, f->image_.get() + got, (size_t)(st.st_size - got), got);
    if (n <= 0) { ::close(fd); fail("short read on " + path); }
    got += n;
  }
  ::close(fd);
  auto buf = std::make_shared<arrow::Buffer>(f->image_.get(), f->image_bytes_);
  auto src = std::make_shared<arrow::io::BufferReader>(buf);
  std::unique_ptr<parquet::ParquetFileReader> rd;
  try {
    rd = parquet::ParquetFileReader::Open(src);
  } catch (const std::exception& e) {
    fail(std::string("Parquet error: ") + e.what());
  }
  auto md = rd->metadata();
  f->rows_ = md->num_rows();
  auto st2 = parquet::arrow::FromParquetSchema(md->schema(), parquet::default_arrow_reader_properties(),
                                               md->key_value_metadata(), &f->schema_);
  if (!st2.ok()) fail(st2.ToString());
  f->reader_ = std::shared_ptr<parquet::ParquetFileReader>(std::move(rd));
  return f;
}

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

bool RawParquetFile::Plan(const std::vector<std::shared_ptr<arrow::Field>>& want, std::string* why) {
  auto* rd = static_cast<parquet::ParquetFileReader*>(reader_.get());
  auto md = rd->metadata();
  const auto* sd = md->schema();
  cols_.clear();
  cols_.resize(want.size());
  // flat schemas only: leaf index == field index
  if (sd->num_columns() != schema_->num_fields()) { *why = "nested schema"; return false; }
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
    for (int rg = 0; rg < md->num_row_groups(); ++rg) {
      auto rgm = md->RowGroup(rg);
      auto ccm = rgm->ColumnChunk(ci);
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
      std::unique_ptr<parquet::PageReader> pr;
      try {
        pr = rd->RowGroup(rg)->GetColumnPageReader(ci);
      } catch (const std::exception& e) {
        *why = std::string("page reader: ") + e.what();
        return false;
      }
      int dict_idx = -1;
      int64_t row_in_rg = 0;
      while (true) {
        std::shared_ptr<parquet::Page> page;
        try {
          page = pr->NextPage();
        } catch (const std::exception& e) {
          *why = std::string("page decode: ") + e.what();
          return false;
        }
        if (!page) break;
        const uint8_t* data = page->data();
        const int64_t size = page->size();
        auto place = [&](int64_t* off, bool* aux) {
          if (data >= image_.get() && data + size <= image_.get() + image_bytes_) {
            *off = data - image_.get();
            *aux = false;
          } else {
            *off = aux_append(data, size);
            *aux = true;
          }
        };
        if (page->type() == parquet::PageType::DICTIONARY_PAGE) {
          auto* dp = static_cast<parquet::DictionaryPage*>(page.get());
          if (dp->encoding() != parquet::Encoding::PLAIN && dp->encoding() != parquet::Encoding::PLAIN_DICTIONARY) {
            *why = "dictionary page encoding"; return false;
          }
          RawColumn::Dict d;
          place(&d.src_off, &d.aux);
          d.size = (int32_t)size;
          d.num_values = dp->num_values();
          rc.dicts.push_back(d);
          dict_idx = (int)rc.dicts.size() - 1;
        } else if (page->type() == parquet::PageType::DATA_PAGE || page->type() == parquet::PageType::DATA_PAGE_V2) {
          RawColumn::Page pg;
          parquet::Encoding::type enc;
          if (page->type() == parquet::PageType::DATA_PAGE) {
            auto* p1 = static_cast<parquet::DataPageV1*>(page.get());
            pg.num_values = p1->num_values();
            enc = p1->encoding();
            pg.def_encoding = map_encoding(p1->definition_level_encoding());
            pg.v2_def_len = -1;
          } else {
            auto* p2 = static_cast<parquet::DataPageV2*>(page.get());
            pg.num_values = p2->num_values();
            enc = p2->encoding();
            pg.def_encoding = ENC_RLE;
            pg.v2_def_len = p2->definition_levels_byte_length();
            if (p2->repetition_levels_byte_length() != 0) { *why = "repetition levels"; return false; }
          }
          pg.encoding = map_encoding(enc);
          const bool is_dict = pg.encoding == ENC_PLAIN_DICT || pg.encoding == ENC_RLE_DICT;
          const bool rle_bool = pg.encoding == ENC_RLE && pc.conv == CONV_BOOL;
          if (!(pg.encoding == ENC_PLAIN || is_dict || rle_bool)) { *why = "value encoding " + parquet::EncodingToString(enc) + " in column " + rc.name; return false; }
          if (is_dict && dict_idx < 0) { *why = "dictionary page missing"; return false; }
          if (pc.conv == CONV_BOOL && is_dict) { *why = "dictionary coded boolean"; return false; }
          place(&pg.src_off, &pg.aux);
          pg.size = (int32_t)size;
          pg.row_group = rg;
          pg.first_row = first_row + row_in_rg;
          row_in_rg += pg.num_values;
          rc.pages.push_back(pg);
          rc.page_dict.push_back(is_dict ? dict_idx : -1);
        }
      }
      if (row_in_rg != rgm->num_rows()) { *why = "page row count mismatch in column " + rc.name; return false; }
      first_row += rgm->num_rows();
    }
  }
  return true;
}

}  // namespace lsb
