// SPDX-License-Identifier: Apache-2.0
//
// zstd_decode.h -- a from-scratch Zstandard frame decoder (RFC 8878 format) that compiles for the host
// (g++, used by the CPU unit test against frames produced by Arrow's codec) and for the device (nvcc,
// decode_kernels.cu: the LakeSoul writer emits PLAIN pages compressed with ZSTD,
// reference rust/lakesoul-io/src/writer/async_writer/multipart_writer.rs:128).
//
// Scope: single frames without a dictionary (what Parquet page compression produces): raw / RLE / compressed
// blocks, raw / RLE / Huffman (1 or 4 streams) / treeless literals, predefined / RLE / FSE / repeat sequence
// tables, repeat offsets, optional content checksum (skipped, not verified). Every read is bounds-checked;
// malformed input returns a negative error instead of reading or writing out of range.
//
// No allocation, no recursion, no library calls: all working memory is the caller's ZstdWork (tables) plus a
// literals buffer of kZstdBlockMax bytes.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define LSB_HD __host__ __device__ __forceinline__
#define LSB_HD_NOINLINE __host__ __device__
#else
#define LSB_HD inline
#define LSB_HD_NOINLINE inline
#endif

namespace lsb {

constexpr int kZstdBlockMax = 128 * 1024;
constexpr int kHufLogMax = 11;
constexpr int kLLLogMax = 9, kMLLogMax = 9, kOFLogMax = 8;

enum ZstdError : int64_t {
  ZSTD_E_TRUNCATED = -1, ZSTD_E_MAGIC = -2, ZSTD_E_UNSUPPORTED = -3, ZSTD_E_CORRUPT = -4, ZSTD_E_DST_SMALL = -5
};

struct FseEntry {
  uint16_t baseline;
  uint8_t symbol;
  uint8_t nbits;
};

struct ZstdWork {
  uint16_t huf[1 << kHufLogMax];   // symbol | nbits << 8
  int huf_log;                     // 0 = no table yet
  FseEntry ll[1 << kLLLogMax], ml[1 << kMLLogMax], of[1 << kOFLogMax];
  int ll_log, ml_log, of_log;      // -1 = no table yet
  uint32_t rep[3];
  int16_t norm[256];               // scratch: normalised counts
  uint16_t next[256];              // scratch: FSE symbolNext
  uint8_t weights[256];            // scratch: Huffman weights
  FseEntry wtab[64];               // scratch: FSE table of the Huffman weights (accuracy log <= 6)
};

namespace zstd_detail {

LSB_HD int highbit(uint32_t v) {  // position of the highest set bit, v > 0
  int r = 0;
  while (v >>= 1) ++r;
  return r;
}

// ---- forward bit reader (FSE table descriptions) --------------------------------------------------
struct FwdBits {
  const uint8_t* p;
  int64_t nbytes;
  int64_t bitpos;
  LSB_HD uint32_t peek(int n) const {  // n <= 25; bits beyond the end read as 0
    const int64_t byte = bitpos >> 3;
    uint64_t v = 0;
    for (int i = 0; i < 5; ++i)
      if (byte + i < nbytes) v |= (uint64_t)p[byte + i] << (8 * i);
    return (uint32_t)((v >> (bitpos & 7)) & ((1u << n) - 1));
  }
  LSB_HD uint32_t read(int n) {
    const uint32_t v = peek(n);
    bitpos += n;
    return v;
  }
};

// ---- backward bit reader (Huffman streams, FSE-coded data) ----------------------------------------
// The stream is read from its last byte towards its first; the highest set bit of the last byte is a marker.
struct BackBits {
  const uint8_t* p;
  int64_t bits_left;  // bits not consumed yet (may go negative: over-read, zeros are returned)
  LSB_HD bool init(const uint8_t* src, int64_t size) {
    p = src;
    if (size <= 0 || src[size - 1] == 0) return false;
    bits_left = size * 8 - (8 - highbit(src[size - 1]));
    return true;
  }
  // the next n bits (n <= 24), most significant first, without consuming them
  LSB_HD uint32_t peek(int n) const {
    if (n == 0) return 0;
    const int64_t hi = bits_left;  // exclusive index of the first unread bit, counted from the stream start
    const int64_t lo = hi - n;     // inclusive
    // gather bytes covering [lo, hi)
    uint64_t v = 0;
    const int64_t b0 = lo >= 0 ? (lo >> 3) : -(((-lo) + 7) >> 3);
    for (int i = 0; i < 5; ++i) {
      const int64_t b = b0 + i;
      const uint64_t byte = (b >= 0 && b * 8 < hi) ? p[b] : 0;
      v |= byte << (8 * i);
    }
    const int sh = (int)(lo - b0 * 8);
    return (uint32_t)((v >> sh) & ((1u << n) - 1));
  }
  LSB_HD uint32_t read(int n) {
    const uint32_t v = peek(n);
    bits_left -= n;
    return v;
  }
};

// ---- FSE ---------------------------------------------------------------------------------------------
// Reads a normalised-count table description; returns bytes consumed or < 0.
LSB_HD int64_t fse_read_ncount(const uint8_t* src, int64_t size, int max_symbol, int max_log, int16_t* norm, int* nsym,
                               int* table_log) {
  if (size < 1) return ZSTD_E_TRUNCATED;
  FwdBits br{src, size, 0};
  const int al = (int)br.read(4) + 5;
  if (al > max_log) return ZSTD_E_CORRUPT;
  int remaining = (1 << al) + 1;
  int threshold = 1 << al;
  int nbits = al + 1;
  int sym = 0;
  while (remaining > 1 && sym <= max_symbol) {
    const int max = (2 * threshold - 1) - remaining;
    int count;
    const uint32_t low = br.peek(nbits - 1);
    if ((int)low < max) {
      count = (int)low;
      br.bitpos += nbits - 1;
    } else {
      count = (int)br.peek(nbits);
      if (count >= threshold) count -= max;
      br.bitpos += nbits;
    }
    --count;  // -1 = "less than one"
    remaining -= count < 0 ? -count : count;
    norm[sym++] = (int16_t)count;
    if (count == 0) {
      // repeat flags: 2 bits each, 3 = "three more zeros and another flag follows"
      while (true) {
        const int r = (int)br.read(2);
        for (int i = 0; i < r && sym <= max_symbol; ++i) norm[sym++] = 0;
        if (r != 3) break;
      }
    }
    while (remaining < threshold && threshold > 1) {
      --nbits;
      threshold >>= 1;
    }
    if (br.bitpos > size * 8 + 16) return ZSTD_E_TRUNCATED;
  }
  if (remaining != 1 || sym > max_symbol + 1) return ZSTD_E_CORRUPT;
  *nsym = sym;
  *table_log = al;
  const int64_t used = (br.bitpos + 7) >> 3;
  return used > size ? (int64_t)ZSTD_E_TRUNCATED : used;
}

LSB_HD void fse_build(FseEntry* table, int table_log, const int16_t* norm, int nsym, uint16_t* next) {
  const int size = 1 << table_log;
  int high = size - 1;
  for (int s = 0; s < nsym; ++s) {
    if (norm[s] == -1) {
      table[high--].symbol = (uint8_t)s;
      next[s] = 1;
    } else {
      next[s] = (uint16_t)norm[s];
    }
  }
  const int step = (size >> 1) + (size >> 3) + 3;
  const int mask = size - 1;
  int pos = 0;
  for (int s = 0; s < nsym; ++s) {
    for (int i = 0; i < norm[s]; ++i) {
      table[pos].symbol = (uint8_t)s;
      do { pos = (pos + step) & mask; } while (pos > high);
    }
  }
  for (int i = 0; i < size; ++i) {
    const int s = table[i].symbol;
    const uint32_t ns = next[s]++;
    const int nb = table_log - highbit(ns);
    table[i].nbits = (uint8_t)nb;
    table[i].baseline = (uint16_t)((ns << nb) - size);
  }
}

LSB_HD void fse_build_rle(FseEntry* table, uint8_t symbol) {
  table[0].symbol = symbol;
  table[0].nbits = 0;
  table[0].baseline = 0;
}

// predefined distributions (RFC 8878 3.1.1.3.2.2)
LSB_HD int16_t ll_default(int i) {
  static const int8_t t[36] = {4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 2, 1, 1, 1, 1, 1, -1, -1, -1, -1};
  return t[i];
}
LSB_HD int16_t ml_default(int i) {
  static const int8_t t[53] = {1, 4, 3, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
                        1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1, -1, -1};
  return t[i];
}
LSB_HD int16_t of_default(int i) {
  static const int8_t t[29] = {1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1};
  return t[i];
}

LSB_HD uint32_t ll_base(int c) {
  static const uint32_t t[36] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 18, 20, 22, 24, 28, 32, 40,
                          48, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536};
  return t[c];
}
LSB_HD int ll_bits(int c) {
  static const uint8_t t[36] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
  return t[c];
}
LSB_HD uint32_t ml_base(int c) {
  static const uint32_t t[53] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30,
                          31, 32, 33, 34, 35, 37, 39, 41, 43, 47, 51, 59, 67, 83, 99, 131, 259, 515, 1027, 2051, 4099, 8195,
                          16387, 32771, 65539};
  return t[c];
}
LSB_HD int ml_bits(int c) {
  static const uint8_t t[53] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
                         1, 1, 1, 1, 2, 2, 3, 3, 4, 4, 5, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
  return t[c];
}

// ---- Huffman -----------------------------------------------------------------------------------------
// Reads the tree description, fills W.huf / W.huf_log; returns bytes consumed or < 0.
LSB_HD_NOINLINE int64_t huf_read_table(ZstdWork& W, const uint8_t* src, int64_t size) {
  if (size < 1) return ZSTD_E_TRUNCATED;
  const int hb = src[0];
  int nw = 0;
  int64_t used;
  if (hb >= 128) {  // direct: 4 bits per weight
    nw = hb - 127;
    used = 1 + ((nw + 1) >> 1);
    if (used > size) return ZSTD_E_TRUNCATED;
    for (int i = 0; i < nw; ++i) {
      const uint8_t b = src[1 + (i >> 1)];
      W.weights[i] = (i & 1) ? (b & 15) : (b >> 4);
    }
  } else {  // FSE-compressed weights, two interleaved states
    used = 1 + hb;
    if (hb == 0 || used > size) return ZSTD_E_TRUNCATED;
    int nsym = 0, tlog = 0;
    const int64_t hdr = fse_read_ncount(src + 1, hb, 12, 6, W.norm, &nsym, &tlog);
    if (hdr < 0) return hdr;
    FseEntry* tab = W.wtab;
    fse_build(tab, tlog, W.norm, nsym, W.next);
    BackBits bb;
    if (!bb.init(src + 1 + hdr, hb - hdr)) return ZSTD_E_CORRUPT;
    uint32_t s1 = bb.read(tlog), s2 = bb.read(tlog);
    while (true) {
      if (nw > 253) return ZSTD_E_CORRUPT;
      W.weights[nw++] = tab[s1].symbol;
      s1 = tab[s1].baseline + bb.read(tab[s1].nbits);
      if (bb.bits_left < 0) { W.weights[nw++] = tab[s2].symbol; break; }
      W.weights[nw++] = tab[s2].symbol;
      s2 = tab[s2].baseline + bb.read(tab[s2].nbits);
      if (bb.bits_left < 0) { W.weights[nw++] = tab[s1].symbol; break; }
    }
  }
  // implied last weight
  uint32_t total = 0;
  for (int i = 0; i < nw; ++i) {
    if (W.weights[i] > kHufLogMax) return ZSTD_E_CORRUPT;
    if (W.weights[i]) total += 1u << (W.weights[i] - 1);
  }
  if (total == 0) return ZSTD_E_CORRUPT;
  const int max_bits = highbit(total) + 1;
  if (max_bits > kHufLogMax) return ZSTD_E_CORRUPT;
  const uint32_t rest = (1u << max_bits) - total;
  if (rest == 0 || (rest & (rest - 1))) return ZSTD_E_CORRUPT;  // must be a power of two
  W.weights[nw++] = (uint8_t)(highbit(rest) + 1);
  // table: weights ascending, symbols in natural order inside a weight
  uint32_t pos = 0;
  for (int w = 1; w <= max_bits; ++w) {
    const uint32_t span = 1u << (w - 1);
    const uint16_t nb = (uint16_t)(max_bits + 1 - w);
    for (int s = 0; s < nw; ++s) {
      if (W.weights[s] != w) continue;
      for (uint32_t i = 0; i < span; ++i) W.huf[pos + i] = (uint16_t)(s | (nb << 8));
      pos += span;
    }
  }
  if (pos != (1u << max_bits)) return ZSTD_E_CORRUPT;
  W.huf_log = max_bits;
  return used;
}

// one Huffman stream -> n symbols
LSB_HD int64_t huf_decode_stream(const ZstdWork& W, const uint8_t* src, int64_t size, uint8_t* dst, int64_t n) {
  BackBits bb;
  if (!bb.init(src, size)) return ZSTD_E_CORRUPT;
  const int log = W.huf_log;
  for (int64_t i = 0; i < n; ++i) {
    const uint16_t e = W.huf[bb.peek(log)];
    dst[i] = (uint8_t)e;
    bb.bits_left -= e >> 8;
  }
  return bb.bits_left == 0 ? n : (int64_t)ZSTD_E_CORRUPT;
}

}  // namespace zstd_detail

LSB_HD void zstd_work_init(ZstdWork& W) {
  W.huf_log = 0;
  W.ll_log = W.ml_log = W.of_log = -1;
  W.rep[0] = 1; W.rep[1] = 4; W.rep[2] = 8;
}

// Decompressed size announced by the frame header, or -1 when the header does not carry it / is malformed.
LSB_HD int64_t zstd_frame_content_size(const uint8_t* src, int64_t size) {
  if (size < 6) return -1;
  if (!(src[0] == 0x28 && src[1] == 0xB5 && src[2] == 0x2F && src[3] == 0xFD)) return -1;
  const int fhd = src[4];
  const int fcs_flag = fhd >> 6, single = (fhd >> 5) & 1, did = fhd & 3;
  int64_t pos = 5 + (single ? 0 : 1) + (did == 0 ? 0 : (did == 1 ? 1 : (did == 2 ? 2 : 4)));
  const int fcs_bytes = fcs_flag == 0 ? (single ? 1 : 0) : (fcs_flag == 1 ? 2 : (fcs_flag == 2 ? 4 : 8));
  if (fcs_bytes == 0 || pos + fcs_bytes > size) return -1;
  uint64_t v = 0;
  for (int i = 0; i < fcs_bytes; ++i) v |= (uint64_t)src[pos + i] << (8 * i);
  if (fcs_bytes == 2) v += 256;
  return (int64_t)v;
}

// Decodes one frame. `lit` = scratch of kZstdBlockMax bytes. Returns the number of bytes written to dst or < 0.
LSB_HD_NOINLINE int64_t zstd_decompress_frame(ZstdWork& W, const uint8_t* src, int64_t size, uint8_t* dst, int64_t dst_cap,
                                              uint8_t* lit) {
  using namespace zstd_detail;
  zstd_work_init(W);
  if (size < 5) return ZSTD_E_TRUNCATED;
  if (!(src[0] == 0x28 && src[1] == 0xB5 && src[2] == 0x2F && src[3] == 0xFD)) return ZSTD_E_MAGIC;
  const int fhd = src[4];
  const int fcs_flag = fhd >> 6, single = (fhd >> 5) & 1, has_checksum = (fhd >> 2) & 1, did = fhd & 3;
  if (fhd & 0x08) return ZSTD_E_UNSUPPORTED;  // reserved bit
  if (did != 0) return ZSTD_E_UNSUPPORTED;    // dictionaries are not used by page compression
  int64_t ip = 5 + (single ? 0 : 1);
  ip += fcs_flag == 0 ? (single ? 1 : 0) : (fcs_flag == 1 ? 2 : (fcs_flag == 2 ? 4 : 8));
  if (ip > size) return ZSTD_E_TRUNCATED;
  int64_t op = 0;
  while (true) {
    if (ip + 3 > size) return ZSTD_E_TRUNCATED;
    const uint32_t bh = (uint32_t)src[ip] | ((uint32_t)src[ip + 1] << 8) | ((uint32_t)src[ip + 2] << 16);
    ip += 3;
    const int last = bh & 1, type = (bh >> 1) & 3;
    const int64_t bsize = bh >> 3;
    if (type == 0) {  // raw
      if (ip + bsize > size) return ZSTD_E_TRUNCATED;
      if (op + bsize > dst_cap) return ZSTD_E_DST_SMALL;
      for (int64_t i = 0; i < bsize; ++i) dst[op + i] = src[ip + i];
      ip += bsize; op += bsize;
    } else if (type == 1) {  // RLE
      if (ip + 1 > size) return ZSTD_E_TRUNCATED;
      if (op + bsize > dst_cap) return ZSTD_E_DST_SMALL;
      const uint8_t b = src[ip++];
      for (int64_t i = 0; i < bsize; ++i) dst[op + i] = b;
      op += bsize;
    } else if (type == 2) {  // compressed
      if (bsize > kZstdBlockMax || ip + bsize > size) return ZSTD_E_TRUNCATED;
      const uint8_t* b = src + ip;
      const int64_t bend = bsize;
      int64_t bp = 0;
      // ---- literals section --------------------------------------------------------------------
      if (bend < 1) return ZSTD_E_CORRUPT;
      const int ltype = b[0] & 3, lfmt = (b[0] >> 2) & 3;
      int64_t regen = 0, lcomp = 0;
      int streams = 1;
      const uint8_t* lits = nullptr;  // where the block's literals are
      if (ltype <= 1) {  // raw / RLE
        int hs;
        if (lfmt == 0 || lfmt == 2) { hs = 1; regen = b[0] >> 3; }
        else if (lfmt == 1) { hs = 2; if (bend < 2) return ZSTD_E_TRUNCATED; regen = (b[0] >> 4) | ((int64_t)b[1] << 4); }
        else { hs = 3; if (bend < 3) return ZSTD_E_TRUNCATED; regen = (b[0] >> 4) | ((int64_t)b[1] << 4) | ((int64_t)b[2] << 12); }
        bp = hs;
        if (regen > kZstdBlockMax) return ZSTD_E_CORRUPT;
        if (ltype == 0) {
          if (bp + regen > bend) return ZSTD_E_TRUNCATED;
          lits = b + bp;
          bp += regen;
        } else {
          if (bp + 1 > bend) return ZSTD_E_TRUNCATED;
          const uint8_t v = b[bp++];
          for (int64_t i = 0; i < regen; ++i) lit[i] = v;
          lits = lit;
        }
      } else {  // Huffman coded (2) or treeless (3)
        int hs;
        if (lfmt == 0 || lfmt == 1) {
          hs = 3; if (bend < 3) return ZSTD_E_TRUNCATED;
          const uint32_t v = (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16);
          regen = (v >> 4) & 0x3ff; lcomp = (v >> 14) & 0x3ff;
          streams = lfmt == 0 ? 1 : 4;
        } else if (lfmt == 2) {
          hs = 4; if (bend < 4) return ZSTD_E_TRUNCATED;
          const uint32_t v = (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24);
          regen = (v >> 4) & 0x3fff; lcomp = (v >> 18) & 0x3fff;
          streams = 4;
        } else {
          hs = 5; if (bend < 5) return ZSTD_E_TRUNCATED;
          const uint64_t v = (uint64_t)b[0] | ((uint64_t)b[1] << 8) | ((uint64_t)b[2] << 16) | ((uint64_t)b[3] << 24) | ((uint64_t)b[4] << 32);
          regen = (v >> 4) & 0x3ffff; lcomp = (v >> 22) & 0x3ffff;
          streams = 4;
        }
        bp = hs;
        if (regen > kZstdBlockMax || bp + lcomp > bend) return ZSTD_E_CORRUPT;
        const uint8_t* lp = b + bp;
        int64_t lsize = lcomp;
        if (ltype == 2) {
          const int64_t t = huf_read_table(W, lp, lsize);
          if (t < 0) return t;
          lp += t; lsize -= t;
        } else if (W.huf_log == 0) {
          return ZSTD_E_CORRUPT;  // treeless without a previous table
        }
        if (streams == 1) {
          const int64_t r = huf_decode_stream(W, lp, lsize, lit, regen);
          if (r < 0) return r;
        } else {
          if (lsize < 6) return ZSTD_E_CORRUPT;
          const int64_t s1 = lp[0] | (lp[1] << 8), s2 = lp[2] | (lp[3] << 8), s3 = lp[4] | (lp[5] << 8);
          const int64_t s4 = lsize - 6 - s1 - s2 - s3;
          if (s4 < 1) return ZSTD_E_CORRUPT;
          const int64_t per = (regen + 3) / 4;
          const int64_t n4 = regen - 3 * per;
          if (n4 < 0) return ZSTD_E_CORRUPT;
          const uint8_t* sp = lp + 6;
          int64_t r = huf_decode_stream(W, sp, s1, lit, per); if (r < 0) return r;
          r = huf_decode_stream(W, sp + s1, s2, lit + per, per); if (r < 0) return r;
          r = huf_decode_stream(W, sp + s1 + s2, s3, lit + 2 * per, per); if (r < 0) return r;
          r = huf_decode_stream(W, sp + s1 + s2 + s3, s4, lit + 3 * per, n4); if (r < 0) return r;
        }
        lits = lit;
        bp += lcomp;
      }
      // ---- sequences section -------------------------------------------------------------------
      if (bp + 1 > bend) return ZSTD_E_TRUNCATED;
      int64_t nseq = b[bp++];
      if (nseq >= 128) {
        if (nseq == 255) {
          if (bp + 2 > bend) return ZSTD_E_TRUNCATED;
          nseq = (int64_t)b[bp] + ((int64_t)b[bp + 1] << 8) + 0x7F00;
          bp += 2;
        } else {
          if (bp + 1 > bend) return ZSTD_E_TRUNCATED;
          nseq = ((nseq - 128) << 8) + b[bp++];
        }
      }
      int64_t lpos = 0;  // literals consumed
      if (nseq > 0) {
        if (bp + 1 > bend) return ZSTD_E_TRUNCATED;
        const int modes = b[bp++];
        if (modes & 3) return ZSTD_E_CORRUPT;
        const int mode[3] = {(modes >> 6) & 3, (modes >> 4) & 3, (modes >> 2) & 3};  // LL, OF, ML
        for (int t = 0; t < 3; ++t) {
          FseEntry* tab = t == 0 ? W.ll : (t == 1 ? W.of : W.ml);
          int* tlog = t == 0 ? &W.ll_log : (t == 1 ? &W.of_log : &W.ml_log);
          const int max_sym = t == 0 ? 35 : (t == 1 ? 31 : 52);
          const int max_log = t == 0 ? kLLLogMax : (t == 1 ? kOFLogMax : kMLLogMax);
          if (mode[t] == 0) {  // predefined
            const int n = t == 0 ? 36 : (t == 1 ? 29 : 53);
            for (int i = 0; i < n; ++i) W.norm[i] = t == 0 ? ll_default(i) : (t == 1 ? of_default(i) : ml_default(i));
            *tlog = t == 1 ? 5 : 6;
            fse_build(tab, *tlog, W.norm, n, W.next);
          } else if (mode[t] == 1) {  // RLE
            if (bp + 1 > bend) return ZSTD_E_TRUNCATED;
            if (b[bp] > max_sym) return ZSTD_E_CORRUPT;
            fse_build_rle(tab, b[bp++]);
            *tlog = 0;
          } else if (mode[t] == 2) {  // FSE description
            int nsym = 0, al = 0;
            const int64_t u = fse_read_ncount(b + bp, bend - bp, max_sym, max_log, W.norm, &nsym, &al);
            if (u < 0) return u;
            bp += u;
            fse_build(tab, al, W.norm, nsym, W.next);
            *tlog = al;
          } else if (*tlog < 0) {  // repeat without a previous table
            return ZSTD_E_CORRUPT;
          }
        }
        BackBits bb;
        if (!bb.init(b + bp, bend - bp)) return ZSTD_E_CORRUPT;
        uint32_t sl = bb.read(W.ll_log), so = bb.read(W.of_log), sm = bb.read(W.ml_log);
        for (int64_t q = 0; q < nseq; ++q) {
          const int oc = W.of[so].symbol, mc = W.ml[sm].symbol, lc = W.ll[sl].symbol;
          if (oc > 31 || mc > 52 || lc > 35) return ZSTD_E_CORRUPT;
          // extra bits: offset, match length, literals length (an offset code can need more than 24 bits)
          uint32_t ov;
          if (oc > 24) {
            const uint32_t hi = bb.read(oc - 16);
            ov = (1u << oc) + (hi << 16) + bb.read(16);
          } else {
            ov = (1u << oc) + bb.read(oc);
          }
          const uint32_t ml = ml_base(mc) + bb.read(ml_bits(mc));
          const uint32_t ll = ll_base(lc) + bb.read(ll_bits(lc));
          if (q + 1 < nseq) {
            sl = W.ll[sl].baseline + bb.read(W.ll[sl].nbits);
            sm = W.ml[sm].baseline + bb.read(W.ml[sm].nbits);
            so = W.of[so].baseline + bb.read(W.of[so].nbits);
          }
          if (bb.bits_left < 0) return ZSTD_E_CORRUPT;
          // repeat offsets
          uint32_t offset;
          if (ov > 3) {
            offset = ov - 3;
            W.rep[2] = W.rep[1]; W.rep[1] = W.rep[0]; W.rep[0] = offset;
          } else {
            uint32_t idx = ov - 1 + (ll == 0 ? 1 : 0);  // 0..3
            if (idx == 0) {
              offset = W.rep[0];
            } else {
              offset = idx == 3 ? W.rep[0] - 1 : W.rep[idx];
              if (idx > 1) W.rep[2] = W.rep[1];
              W.rep[1] = W.rep[0];
              W.rep[0] = offset;
            }
          }
          if (offset == 0) return ZSTD_E_CORRUPT;
          // execute
          if (lpos + ll > regen) return ZSTD_E_CORRUPT;
          if (op + ll + ml > dst_cap) return ZSTD_E_DST_SMALL;
          for (uint32_t i = 0; i < ll; ++i) dst[op + i] = lits[lpos + i];
          op += ll; lpos += ll;
          if ((int64_t)offset > op) return ZSTD_E_CORRUPT;
          const uint8_t* m = dst + op - offset;
          for (uint32_t i = 0; i < ml; ++i) dst[op + i] = m[i];
          op += ml;
        }
        if (bb.bits_left != 0) return ZSTD_E_CORRUPT;
      }
      // remaining literals
      const int64_t rest = regen - lpos;
      if (op + rest > dst_cap) return ZSTD_E_DST_SMALL;
      for (int64_t i = 0; i < rest; ++i) dst[op + i] = lits[lpos + i];
      op += rest;
      ip += bsize;
    } else {
      return ZSTD_E_CORRUPT;
    }
    if (last) break;
  }
  if (has_checksum) ip += 4;
  if (ip > size) return ZSTD_E_TRUNCATED;
  return op;
}

// ---------------------------------------------------------------------------------------------------------
// Phased decoding: the same format logic cut into steps that a thread block can run side by side.
//   scan    : block headers of the frame                                   (host, a few bytes per block)
//   plan    : per compressed block, in order: section headers + entropy tables (tables may be inherited)
//   entropy : per block, independent: the Huffman streams -> literals, the FSE bit stream -> sequences
//   execute : in order: repeat offsets, literal runs and matches
// zstd_decompress_frame_phased() runs them one after the other on the host (unit test); zstd_pages_kernel maps
// "entropy" onto the threads of a block.
// ---------------------------------------------------------------------------------------------------------
struct ZstdBlockDesc {      // one block of a frame (scan)
  int32_t src_off;          // first byte of the block content inside the frame
  int32_t size;             // content bytes (RLE: 1)
  int32_t type;             // 0 raw, 1 RLE, 2 compressed
  int32_t out_size;         // raw / RLE: bytes produced; compressed: unknown (-1)
};

struct ZstdTables {
  uint16_t huf[1 << kHufLogMax];
  FseEntry ll[1 << kLLLogMax], ml[1 << kMLLogMax], of[1 << kOFLogMax];
};

struct ZstdBlockPlan {      // one compressed block (plan)
  int32_t lit_type;         // 0 raw, 1 RLE, 2 Huffman (own or inherited table)
  int32_t regen;            // literal bytes of the block
  int32_t lit_src;          // block-relative: raw literals / the RLE byte / the first Huffman stream
  int32_t streams;          // 1 or 4
  int32_t stream_size[4];
  int32_t huf_log;
  int32_t nseq;
  int32_t seq_src, seq_size;  // block-relative FSE bit stream
  int32_t ll_log, ml_log, of_log;
  int32_t lit_base;         // where the block's literals start inside the page's literal scratch
  int32_t seq_base;         // where the block's sequences start inside the page's sequence scratch
  // block-parallel execution (zstd_pages_kernel_v2): the repeat-offset history is the only state that crosses
  // blocks. The sequence decoder resolves every offset relative to the block's INITIAL history (symbolically);
  // the final history is a function of the initial one: rep_final[j] = rep_src[j] == 3 ? rep_val[j]
  //                                                                   : rep_init[rep_src[j]] + rep_val[j]
  int32_t out_size;         // bytes the block produces (regen + sum of match lengths)
  int32_t out_base;         // where they start inside the page
  int32_t rep_src[3];
  int32_t rep_val[3];
  uint32_t rep_init[3];     // history at the block's start (filled once the blocks before it are chained)
};

struct ZstdPlanScratch {
  int16_t norm[256];
  uint16_t next[256];
  uint8_t weights[256];
  FseEntry wtab[64];
};

// sequence record: literal length (17 bits) | match length (17 bits) << 17 | raw offset value << 34
LSB_HD uint64_t zstd_pack_seq(uint32_t ll, uint32_t ml, uint32_t ov) { return (uint64_t)ll | ((uint64_t)ml << 17) | ((uint64_t)ov << 34); }

// Frame header + block headers. Returns the number of blocks or < 0; *end = bytes of the frame.
LSB_HD int64_t zstd_scan_frame(const uint8_t* src, int64_t size, ZstdBlockDesc* out, int maxb, int64_t* end) {
  if (size < 5) return ZSTD_E_TRUNCATED;
  if (!(src[0] == 0x28 && src[1] == 0xB5 && src[2] == 0x2F && src[3] == 0xFD)) return ZSTD_E_MAGIC;
  const int fhd = src[4];
  const int fcs_flag = fhd >> 6, single = (fhd >> 5) & 1, has_checksum = (fhd >> 2) & 1, did = fhd & 3;
  if ((fhd & 0x08) || did != 0) return ZSTD_E_UNSUPPORTED;
  int64_t ip = 5 + (single ? 0 : 1);
  ip += fcs_flag == 0 ? (single ? 1 : 0) : (fcs_flag == 1 ? 2 : (fcs_flag == 2 ? 4 : 8));
  int n = 0;
  while (true) {
    if (ip + 3 > size) return ZSTD_E_TRUNCATED;
    const uint32_t bh = (uint32_t)src[ip] | ((uint32_t)src[ip + 1] << 8) | ((uint32_t)src[ip + 2] << 16);
    ip += 3;
    const int last = bh & 1, type = (bh >> 1) & 3;
    const int64_t bsize = bh >> 3;
    if (type == 3) return ZSTD_E_CORRUPT;
    const int64_t content = type == 1 ? 1 : bsize;
    if (bsize > kZstdBlockMax || ip + content > size) return ZSTD_E_TRUNCATED;
    if (n >= maxb) return ZSTD_E_UNSUPPORTED;
    out[n].src_off = (int32_t)ip;
    out[n].size = (int32_t)content;
    out[n].type = type;
    out[n].out_size = type == 2 ? -1 : (int32_t)bsize;
    ++n;
    ip += content;
    if (last) break;
  }
  if (has_checksum) ip += 4;
  if (ip > size) return ZSTD_E_TRUNCATED;
  *end = ip;
  return n;
}

// Section headers and entropy tables of one compressed block. `prev`/`prevT` = the previous COMPRESSED block of the
// frame (nullptr for the first): treeless literals and "repeat" sequence tables copy its tables.
LSB_HD_NOINLINE int64_t zstd_plan_block(const uint8_t* b, int64_t bend, ZstdBlockPlan& P, ZstdTables& T, const ZstdBlockPlan* prev,
                                        const ZstdTables* prevT, ZstdPlanScratch& S) {
  using namespace zstd_detail;
  if (bend < 1) return ZSTD_E_CORRUPT;
  int64_t bp = 0;
  const int ltype = b[0] & 3, lfmt = (b[0] >> 2) & 3;
  int64_t regen = 0, lcomp = 0;
  P.streams = 1;
  P.huf_log = prev ? prev->huf_log : 0;
  P.ll_log = prev ? prev->ll_log : -1;
  P.ml_log = prev ? prev->ml_log : -1;
  P.of_log = prev ? prev->of_log : -1;
  if (ltype <= 1) {
    int hs;
    if (lfmt == 0 || lfmt == 2) { hs = 1; regen = b[0] >> 3; }
    else if (lfmt == 1) { hs = 2; if (bend < 2) return ZSTD_E_TRUNCATED; regen = (b[0] >> 4) | ((int64_t)b[1] << 4); }
    else { hs = 3; if (bend < 3) return ZSTD_E_TRUNCATED; regen = (b[0] >> 4) | ((int64_t)b[1] << 4) | ((int64_t)b[2] << 12); }
    bp = hs;
    if (regen > kZstdBlockMax) return ZSTD_E_CORRUPT;
    P.lit_type = ltype;
    P.lit_src = (int32_t)bp;
    if (ltype == 0) { if (bp + regen > bend) return ZSTD_E_TRUNCATED; bp += regen; }
    else { if (bp + 1 > bend) return ZSTD_E_TRUNCATED; bp += 1; }
    // a later treeless block still inherits the last Huffman table: carry it along
    if (prevT && P.huf_log > 0) for (int i = 0; i < (1 << P.huf_log); ++i) T.huf[i] = prevT->huf[i];
  } else {
    int hs;
    if (lfmt == 0 || lfmt == 1) {
      hs = 3; if (bend < 3) return ZSTD_E_TRUNCATED;
      const uint32_t v = (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16);
      regen = (v >> 4) & 0x3ff; lcomp = (v >> 14) & 0x3ff;
      P.streams = lfmt == 0 ? 1 : 4;
    } else if (lfmt == 2) {
      hs = 4; if (bend < 4) return ZSTD_E_TRUNCATED;
      const uint32_t v = (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24);
      regen = (v >> 4) & 0x3fff; lcomp = (v >> 18) & 0x3fff;
      P.streams = 4;
    } else {
      hs = 5; if (bend < 5) return ZSTD_E_TRUNCATED;
      const uint64_t v = (uint64_t)b[0] | ((uint64_t)b[1] << 8) | ((uint64_t)b[2] << 16) | ((uint64_t)b[3] << 24) | ((uint64_t)b[4] << 32);
      regen = (v >> 4) & 0x3ffff; lcomp = (v >> 22) & 0x3ffff;
      P.streams = 4;
    }
    bp = hs;
    if (regen > kZstdBlockMax || bp + lcomp > bend) return ZSTD_E_CORRUPT;
    const uint8_t* lp = b + bp;
    int64_t lsize = lcomp;
    P.lit_type = 2;
    if (ltype == 2) {
      // huf_read_table works on a ZstdWork: use a local view of its fields through a small adapter
      int64_t t;
      {
        // inline copy of the table reader on T.huf (same logic as huf_read_table)
        if (lsize < 1) return ZSTD_E_TRUNCATED;
        const int hb = lp[0];
        int nw = 0;
        int64_t used;
        if (hb >= 128) {
          nw = hb - 127;
          used = 1 + ((nw + 1) >> 1);
          if (used > lsize) return ZSTD_E_TRUNCATED;
          for (int i = 0; i < nw; ++i) {
            const uint8_t x = lp[1 + (i >> 1)];
            S.weights[i] = (i & 1) ? (x & 15) : (x >> 4);
          }
        } else {
          used = 1 + hb;
          if (hb == 0 || used > lsize) return ZSTD_E_TRUNCATED;
          int nsym = 0, tlog = 0;
          const int64_t hdr = fse_read_ncount(lp + 1, hb, 12, 6, S.norm, &nsym, &tlog);
          if (hdr < 0) return hdr;
          fse_build(S.wtab, tlog, S.norm, nsym, S.next);
          BackBits bb;
          if (!bb.init(lp + 1 + hdr, hb - hdr)) return ZSTD_E_CORRUPT;
          uint32_t s1 = bb.read(tlog), s2 = bb.read(tlog);
          while (true) {
            if (nw > 253) return ZSTD_E_CORRUPT;
            S.weights[nw++] = S.wtab[s1].symbol;
            s1 = S.wtab[s1].baseline + bb.read(S.wtab[s1].nbits);
            if (bb.bits_left < 0) { S.weights[nw++] = S.wtab[s2].symbol; break; }
            S.weights[nw++] = S.wtab[s2].symbol;
            s2 = S.wtab[s2].baseline + bb.read(S.wtab[s2].nbits);
            if (bb.bits_left < 0) { S.weights[nw++] = S.wtab[s1].symbol; break; }
          }
        }
        uint32_t total = 0;
        for (int i = 0; i < nw; ++i) {
          if (S.weights[i] > kHufLogMax) return ZSTD_E_CORRUPT;
          if (S.weights[i]) total += 1u << (S.weights[i] - 1);
        }
        if (total == 0) return ZSTD_E_CORRUPT;
        const int max_bits = highbit(total) + 1;
        if (max_bits > kHufLogMax) return ZSTD_E_CORRUPT;
        const uint32_t rest = (1u << max_bits) - total;
        if (rest == 0 || (rest & (rest - 1))) return ZSTD_E_CORRUPT;
        S.weights[nw++] = (uint8_t)(highbit(rest) + 1);
        uint32_t pos = 0;
        for (int w = 1; w <= max_bits; ++w) {
          const uint32_t span = 1u << (w - 1);
          const uint16_t nb = (uint16_t)(max_bits + 1 - w);
          for (int sy = 0; sy < nw; ++sy) {
            if (S.weights[sy] != w) continue;
            for (uint32_t i = 0; i < span; ++i) T.huf[pos + i] = (uint16_t)(sy | (nb << 8));
            pos += span;
          }
        }
        if (pos != (1u << max_bits)) return ZSTD_E_CORRUPT;
        P.huf_log = max_bits;
        t = used;
      }
      lp += t; lsize -= t;
    } else {
      if (P.huf_log == 0 || !prevT) return ZSTD_E_CORRUPT;  // treeless without a previous table
      for (int i = 0; i < (1 << P.huf_log); ++i) T.huf[i] = prevT->huf[i];
    }
    P.lit_src = (int32_t)(lp - b);
    if (P.streams == 1) {
      P.stream_size[0] = (int32_t)lsize; P.stream_size[1] = P.stream_size[2] = P.stream_size[3] = 0;
    } else {
      if (lsize < 6) return ZSTD_E_CORRUPT;
      const int64_t s1 = lp[0] | (lp[1] << 8), s2 = lp[2] | (lp[3] << 8), s3 = lp[4] | (lp[5] << 8);
      const int64_t s4 = lsize - 6 - s1 - s2 - s3;
      if (s4 < 1 || regen - 3 * ((regen + 3) / 4) < 0) return ZSTD_E_CORRUPT;
      P.lit_src += 6;
      P.stream_size[0] = (int32_t)s1; P.stream_size[1] = (int32_t)s2; P.stream_size[2] = (int32_t)s3; P.stream_size[3] = (int32_t)s4;
    }
    bp += lcomp;
  }
  P.regen = (int32_t)regen;
  // sequences header
  if (bp + 1 > bend) return ZSTD_E_TRUNCATED;
  int64_t nseq = b[bp++];
  if (nseq >= 128) {
    if (nseq == 255) {
      if (bp + 2 > bend) return ZSTD_E_TRUNCATED;
      nseq = (int64_t)b[bp] + ((int64_t)b[bp + 1] << 8) + 0x7F00;
      bp += 2;
    } else {
      if (bp + 1 > bend) return ZSTD_E_TRUNCATED;
      nseq = ((nseq - 128) << 8) + b[bp++];
    }
  }
  P.nseq = (int32_t)nseq;
  int modes = 0xFC;  // no sequences: every table is inherited as it is
  if (nseq > 0) {
    if (bp + 1 > bend) return ZSTD_E_TRUNCATED;
    modes = b[bp++];
    if (modes & 3) return ZSTD_E_CORRUPT;
  }
  const int mode[3] = {(modes >> 6) & 3, (modes >> 4) & 3, (modes >> 2) & 3};  // LL, OF, ML
  for (int t = 0; t < 3; ++t) {
    FseEntry* tab = t == 0 ? T.ll : (t == 1 ? T.of : T.ml);
    const FseEntry* ptab = prevT ? (t == 0 ? prevT->ll : (t == 1 ? prevT->of : prevT->ml)) : nullptr;
    int32_t* tlog = t == 0 ? &P.ll_log : (t == 1 ? &P.of_log : &P.ml_log);
    const int max_sym = t == 0 ? 35 : (t == 1 ? 31 : 52);
    const int max_log = t == 0 ? kLLLogMax : (t == 1 ? kOFLogMax : kMLLogMax);
    if (mode[t] == 0) {
      const int n = t == 0 ? 36 : (t == 1 ? 29 : 53);
      for (int i = 0; i < n; ++i) S.norm[i] = t == 0 ? ll_default(i) : (t == 1 ? of_default(i) : ml_default(i));
      *tlog = t == 1 ? 5 : 6;
      fse_build(tab, *tlog, S.norm, n, S.next);
    } else if (mode[t] == 1) {
      if (bp + 1 > bend) return ZSTD_E_TRUNCATED;
      if (b[bp] > max_sym) return ZSTD_E_CORRUPT;
      fse_build_rle(tab, b[bp++]);
      *tlog = 0;
    } else if (mode[t] == 2) {
      int nsym = 0, al = 0;
      const int64_t u = fse_read_ncount(b + bp, bend - bp, max_sym, max_log, S.norm, &nsym, &al);
      if (u < 0) return u;
      bp += u;
      fse_build(tab, al, S.norm, nsym, S.next);
      *tlog = al;
    } else {
      if (*tlog < 0 || !ptab) { if (nseq > 0) return ZSTD_E_CORRUPT; }
      else for (int i = 0; i < (1 << *tlog); ++i) tab[i] = ptab[i];
    }
  }
  P.seq_src = (int32_t)bp;
  P.seq_size = (int32_t)(bend - bp);
  if (nseq == 0 && bp != bend) return ZSTD_E_CORRUPT;
  return 0;
}

// one Huffman stream -> n literal bytes
LSB_HD int64_t zstd_huf_stream(const uint16_t* huf, int log, const uint8_t* src, int64_t size, uint8_t* dst, int64_t n) {
  zstd_detail::BackBits bb;
  if (!bb.init(src, size)) return ZSTD_E_CORRUPT;
  for (int64_t i = 0; i < n; ++i) {
    const uint16_t e = huf[bb.peek(log)];
    dst[i] = (uint8_t)e;
    bb.bits_left -= e >> 8;
  }
  return bb.bits_left == 0 ? n : (int64_t)ZSTD_E_CORRUPT;
}

// FSE bit stream of one block -> packed sequence records (offsets still as coded: repeat codes unresolved)
LSB_HD int64_t zstd_decode_sequences(const uint8_t* bits, int64_t size, const ZstdBlockPlan& P, const ZstdTables& T, uint64_t* out) {
  using namespace zstd_detail;
  BackBits bb;
  if (!bb.init(bits, size)) return ZSTD_E_CORRUPT;
  uint32_t sl = bb.read(P.ll_log), so = bb.read(P.of_log), sm = bb.read(P.ml_log);
  for (int32_t q = 0; q < P.nseq; ++q) {
    const int oc = T.of[so].symbol, mc = T.ml[sm].symbol, lc = T.ll[sl].symbol;
    if (oc > 29 || mc > 52 || lc > 35) return ZSTD_E_CORRUPT;  // offsets beyond 1 GiB are not representable here
    uint32_t ov;
    if (oc > 24) {
      const uint32_t hi = bb.read(oc - 16);
      ov = (1u << oc) + (hi << 16) + bb.read(16);
    } else {
      ov = (1u << oc) + bb.read(oc);
    }
    const uint32_t ml = ml_base(mc) + bb.read(ml_bits(mc));
    const uint32_t ll = ll_base(lc) + bb.read(ll_bits(lc));
    if (q + 1 < P.nseq) {
      sl = T.ll[sl].baseline + bb.read(T.ll[sl].nbits);
      sm = T.ml[sm].baseline + bb.read(T.ml[sm].nbits);
      so = T.of[so].baseline + bb.read(T.of[so].nbits);
    }
    if (bb.bits_left < 0) return ZSTD_E_CORRUPT;
    out[q] = zstd_pack_seq(ll, ml, ov);
  }
  return bb.bits_left == 0 ? 0 : (int64_t)ZSTD_E_CORRUPT;
}

// coded offset value + repeat history -> actual offset (updates the history); 0 = corrupt
LSB_HD uint32_t zstd_resolve_offset(uint32_t ov, uint32_t ll, uint32_t* rep) {
  uint32_t offset;
  if (ov > 3) {
    offset = ov - 3;
    rep[2] = rep[1]; rep[1] = rep[0]; rep[0] = offset;
  } else {
    const uint32_t idx = ov - 1 + (ll == 0 ? 1 : 0);
    if (idx == 0) return rep[0];
    offset = idx == 3 ? rep[0] - 1 : rep[idx];
    if (idx > 1) rep[2] = rep[1];
    rep[1] = rep[0];
    rep[0] = offset;
  }
  return offset;
}

}  // namespace lsb
