// SPDX-License-Identifier: Apache-2.0
//
// decode_types.h -- POD descriptors shared by the host page planner (page_plan.cc, g++) and the
// device page decoders (decode_kernels.cu, nvcc). Replaces the Arrow array materialisation of the
// third-party `parquet` crate on the reference's scan (rust/lakesoul-io/src/physical_plan/merge/mod.rs:66-72,
// file_format.rs:242-252): Parquet data pages are decoded on the GPU straight into the HBM column
// buffers the merge and gather kernels read.
#pragma once
#include <cstdint>

namespace lsb {

enum PqPhys : int32_t { PQ_BOOLEAN = 0, PQ_INT32 = 1, PQ_INT64 = 2, PQ_INT96 = 3, PQ_FLOAT = 4, PQ_DOUBLE = 5,
                        PQ_BYTE_ARRAY = 6, PQ_FLBA = 7 };
// parquet.thrift Encoding
enum PqEnc : int32_t { ENC_PLAIN = 0, ENC_PLAIN_DICT = 2, ENC_RLE = 3, ENC_BIT_PACKED = 4, ENC_RLE_DICT = 8 };

enum PqConv : int32_t {
  CONV_COPY = 0,       // fixed width: copy the low `out_width` bytes of the little-endian physical value
  CONV_BOOL = 1,       // boolean -> bit
  CONV_DEC_FLBA = 2,   // FIXED_LEN_BYTE_ARRAY(type_length) big-endian two's complement -> 16-byte LE
  CONV_DEC_INT = 3,    // INT32/INT64 unscaled decimal -> sign-extended 16-byte LE
  CONV_VAR = 4,        // BYTE_ARRAY -> int32 offsets + bytes
  CONV_FLBA_RAW = 5    // FIXED_LEN_BYTE_ARRAY -> same bytes (fixed_size_binary)
};

struct PqColOut {        // one entry per column slot of the scan group
  void* values;          // fixed: width*rows ; bool: bits ; var: int32 offsets (global row indexed)
  uint8_t* validity;     // zero-initialised bitmap over global rows, or nullptr
  uint8_t* data;         // var: byte buffer of the column
  const int64_t* bucket_data;  // var: data base of every bucket inside `data`
};

struct PqChunk {         // one column chunk (file x column x row group)
  int32_t col;           // column slot
  int32_t phys;          // PqPhys
  int32_t type_length;   // FLBA
  int32_t conv;          // PqConv
  int32_t out_width;     // bytes per output value (fixed), 0 for bool / var
  int32_t max_def;       // 0 required, 1 optional
  int32_t bucket;
  int32_t dict_nvals;    // 0 = no dictionary page
  int32_t dict_size;     // payload bytes of the dictionary page
  int32_t pad;
  int64_t dict_off;      // payload offset of the dictionary page inside the raw image
  int64_t dict_index;    // BYTE_ARRAY dictionaries: first entry of this chunk in the dict_pos table
};

constexpr int kDecodeBlockRows = 1024;  // rows per decode work unit (one warp)

struct PqPage {
  int64_t src_off;       // payload (after the thrift header, decompressed) inside the raw image
  int32_t size;
  int32_t num_values;    // rows of the page (flat columns)
  int32_t encoding;      // PqEnc of the values
  int32_t def_encoding;  // ENC_RLE or ENC_BIT_PACKED
  int32_t v2_def_len;    // DATA_PAGE_V2: bytes of definition levels (no length prefix); -1 for v1 pages
  int32_t chunk;
  int32_t closes_file;   // 1: also write the closing string offset after the page's last row
  int32_t first_block;   // the page's first entry in the block table (ceil(num_values / kDecodeBlockRows) entries)
  int64_t row_start;     // global row of the page's first value
  int64_t vpos_base;     // PLAIN BYTE_ARRAY: first entry of the page in the vpos table (nonnull + 1 entries)
  int32_t nonnull;       // index pass: non-null values of the page
  int32_t values_off;    // index pass: offset of the value section inside the page payload
};

// State of one RLE / bit-packed hybrid stream at a block boundary (all offsets relative to the page payload).
struct PqStreamCk {
  int32_t pos;           // next run header
  int32_t base;          // payload of the current bit-packed run
  int32_t run_left;
  int32_t mode;          // 0 rle, 1 bit-packed, 2 legacy BIT_PACKED
  int32_t idx;           // position inside the bit-packed run
  uint32_t val;          // value of the current RLE run
};

// One decode work unit: rows [row0, row0 + kDecodeBlockRows) of a page. Filled by the index pass
// (one warp per page walks the run headers / length prefixes once), consumed by the decode pass
// (one warp per block, all blocks of all pages in parallel).
struct PqBlock {
  int32_t page;
  int32_t row0;          // first row of the block inside the page
  int32_t vbase;         // non-null values of the page before the block
  int32_t pad;
  PqStreamCk levels;     // definition levels at row0
  PqStreamCk index;      // dictionary indices / RLE booleans at value vbase
  int64_t str_bytes;     // var: bytes produced by the block
  int64_t str_base;      // var: first byte of the block inside the bucket's data (after the scan)
};

// One ZSTD-compressed page payload to expand on the device before the index pass.
struct PqZstdPage {
  int64_t src_off;       // compressed bytes (absolute device address, like PqPage.src_off)
  int64_t dst_off;       // where the decompressed payload goes (absolute device address)
  int32_t comp_size;
  int32_t raw_size;      // uncompressed_page_size of the page header
  int32_t first_block;   // the frame's blocks in the block list (scanned by the host: three header bytes per block)
  int32_t nblocks;
  int64_t lit_off;       // the page's slice of the literal scratch (bytes)
  int64_t seq_off;       // the page's slice of the sequence scratch (records)
};

struct DecodeArgs {
  const uint8_t* raw;        // device image of the files / decompressed pages
  const PqColOut* cols;
  const PqChunk* chunks;
  PqPage* pages;
  int32_t* dict_pos;         // start of every dictionary string inside its dictionary page (+1 closing entry per chunk)
  PqBlock* blocks;
  int32_t* vpos;             // PLAIN BYTE_ARRAY pages: string bytes of the page before every non-null value
  int32_t npages;
  int32_t nchunks;
  int32_t nblocks;
  int32_t* err;              // device error flag (unsupported page shape met at decode time)
};

}  // namespace lsb
