// SPDX-License-Identifier: Apache-2.0
//
// decode_kernels.cu -- K1: Parquet data-page decode on the GPU.
//   * definition levels (RLE / bit-packed hybrid, bit width 1; legacy BIT_PACKED) -> Arrow validity
//   * PLAIN fixed width (INT32/INT64/FLOAT/DOUBLE/FLBA, decimals re-encoded to 16-byte LE), PLAIN BOOLEAN
//   * PLAIN BYTE_ARRAY (length-prefixed) -> int32 offsets + bytes
//   * PLAIN_DICTIONARY / RLE_DICTIONARY (hybrid-coded indices + PLAIN dictionary page), RLE booleans
// Host side (page_plan.cc) frames the pages (thrift headers, decompression) and never touches values.
//
// Pages are big (1 MiB by default in parquet-mr / Arrow writers) and few, so the work is cut in
// two passes:
//   index pass  -- one warp per PAGE walks what is inherently sequential exactly once: the run headers
//                  of the level / index streams (payloads are skipped, level payloads are pop-counted
//                  32 bytes per step) and the length-prefix chain of PLAIN BYTE_ARRAY values. It leaves
//                  a checkpoint of every stream at each 1024-row block boundary (PqBlock) and, for
//                  PLAIN BYTE_ARRAY, the running string offset of every value (vpos).
//   decode pass -- one warp per BLOCK, all blocks of all pages in parallel, restarts the streams from
//                  its checkpoint and produces validity, values, offsets and string bytes.
// Every lane of a warp holds the same stream state; run headers are parsed redundantly (uniform
// loads), values inside a run are addressed directly by lane.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#include "decode_types.h"
#include "zstd_decode.h"

namespace lsb {

namespace {

constexpr int kErrBadPage = 1;
constexpr int kWalkWindow = 4096;  // bytes of a page staged per refill of the BYTE_ARRAY walk

__device__ __forceinline__ uint32_t load_le32(const uint8_t* p) {
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
__device__ __forceinline__ uint64_t load_le(const uint8_t* p, int n) {
  uint64_t v = 0;
  for (int i = 0; i < n; ++i) v |= (uint64_t)p[i] << (8 * i);
  return v;
}

struct Hybrid {
  const uint8_t* p;
  const uint8_t* end;
  const uint8_t* base;  // bit-packed run payload
  int bw;
  int run_left;
  int mode;  // 0 rle, 1 bit-packed (lsb first), 2 legacy BIT_PACKED (msb first, whole section)
  int idx;
  uint32_t rle_val;
};

__device__ __forceinline__ void hybrid_init(Hybrid& h, const uint8_t* p, const uint8_t* end, int bw) {
  h.p = p; h.end = end; h.base = p; h.bw = bw; h.run_left = 0; h.mode = 0; h.idx = 0; h.rle_val = 0;
}
__device__ __forceinline__ void legacy_init(Hybrid& h, const uint8_t* p, int nvalues, int bw) {
  h.p = p; h.end = p; h.base = p; h.bw = bw; h.run_left = nvalues; h.mode = 2; h.idx = 0; h.rle_val = 0;
}
__device__ __forceinline__ void hybrid_save(const Hybrid& h, const uint8_t* page, PqStreamCk& ck) {
  ck.pos = (int32_t)(h.p - page);
  ck.base = (int32_t)(h.base - page);
  ck.run_left = h.run_left;
  ck.mode = h.mode;
  ck.idx = h.idx;
  ck.val = h.rle_val;
}
__device__ __forceinline__ void hybrid_restore(Hybrid& h, const uint8_t* page, const PqStreamCk& ck) {
  h.p = page + ck.pos;
  h.base = page + ck.base;
  h.run_left = ck.run_left;
  h.mode = ck.mode;
  h.idx = ck.idx;
  h.rle_val = ck.val;
}

__device__ __forceinline__ bool hybrid_next_run(Hybrid& h) {
  if (h.p >= h.end) return false;
  uint32_t hd = 0;
  int shift = 0;
  while (h.p < h.end) {
    const uint8_t b = *h.p++;
    hd |= (uint32_t)(b & 0x7f) << shift;
    if (!(b & 0x80)) break;
    shift += 7;
  }
  if (hd & 1) {
    const int groups = (int)(hd >> 1);
    h.mode = 1;
    h.run_left = groups * 8;
    h.base = h.p;
    h.idx = 0;
    h.p += (int64_t)groups * h.bw;
  } else {
    h.mode = 0;
    h.run_left = (int)(hd >> 1);
    const int nb = (h.bw + 7) >> 3;
    h.rle_val = (uint32_t)load_le(h.p, nb);
    h.p += nb;
  }
  return h.run_left > 0 || h.p < h.end;
}

__device__ __forceinline__ uint32_t extract_bits(const uint8_t* base, int64_t bitpos, int bw) {
  const uint8_t* p = base + (bitpos >> 3);
  const int sh = (int)(bitpos & 7);
  const int nbytes = (sh + bw + 7) >> 3;  // <= 5
  const uint64_t v = load_le(p, nbytes);
  return (uint32_t)((v >> sh) & ((bw >= 32) ? 0xffffffffull : ((1ull << bw) - 1)));
}

// lane i < count receives the next i-th value; all lanes advance the (replicated) state together
__device__ __forceinline__ uint32_t hybrid_get(Hybrid& h, int count, int lane) {
  uint32_t out = 0;
  int done = 0;
  while (done < count) {
    if (h.run_left == 0) {
      if (h.mode == 2 || !hybrid_next_run(h)) break;
      if (h.run_left == 0) continue;
    }
    const int t = min(h.run_left, count - done);
    if (lane >= done && lane < done + t) {
      const int j = lane - done;
      if (h.mode == 0) out = h.rle_val;
      else if (h.mode == 1) out = extract_bits(h.base, (int64_t)(h.idx + j) * h.bw, h.bw);
      else {
        const int64_t bit = (int64_t)(h.idx + j) * h.bw;  // bw == 1 in practice (flat columns)
        out = (h.base[bit >> 3] >> (7 - (bit & 7))) & 1;
      }
    }
    h.run_left -= t;
    h.idx += t;
    done += t;
  }
  return out;
}

// advance by `count` values reading run headers only
__device__ __forceinline__ void hybrid_skip(Hybrid& h, int count) {
  while (count > 0) {
    if (h.run_left == 0) {
      if (h.mode == 2 || !hybrid_next_run(h)) return;
      if (h.run_left == 0) continue;
    }
    const int t = min(h.run_left, count);
    h.run_left -= t;
    h.idx += t;
    count -= t;
  }
}

// set bits among positions [b0, b1) of a bit string (lsb-first or msb-first inside a byte); warp-wide
__device__ __forceinline__ int count_bits(const uint8_t* base, int64_t b0, int64_t b1, bool msb_first, int lane) {
  int ones = 0;
  const int64_t byte0 = b0 >> 3, byte1 = (b1 + 7) >> 3;
  for (int64_t y = byte0 + lane; y < byte1; y += 32) {
    uint32_t v = base[y];
    if (msb_first) v = __brev(v) >> 24;
    const int64_t lo = y * 8;
    if (lo < b0) v &= 0xffu << (int)(b0 - lo);
    if (lo + 8 > b1) v &= 0xffu >> (int)(lo + 8 - b1);
    ones += __popc(v);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) ones += __shfl_xor_sync(0xffffffffu, ones, d);
  return ones;
}

// advance a bit-width-1 level stream by `count` values and return how many of them are 1
__device__ __forceinline__ int hybrid_count_ones(Hybrid& h, int count, int lane) {
  int ones = 0;
  while (count > 0) {
    if (h.run_left == 0) {
      if (h.mode == 2 || !hybrid_next_run(h)) break;
      if (h.run_left == 0) continue;
    }
    const int t = min(h.run_left, count);
    if (h.mode == 0) ones += (h.rle_val & 1) ? t : 0;
    else ones += count_bits(h.base, h.idx, (int64_t)h.idx + t, h.mode == 2, lane);
    h.run_left -= t;
    h.idx += t;
    count -= t;
  }
  return ones;
}

struct PageView {
  Hybrid levels;
  bool has_levels;
  const uint8_t* values;
  const uint8_t* values_end;
};

__device__ __forceinline__ bool open_page(const DecodeArgs& A, const PqPage& pg, const PqChunk& ch, PageView& v) {
  const uint8_t* p = A.raw + pg.src_off;
  const uint8_t* end = p + pg.size;
  v.has_levels = ch.max_def > 0;
  if (v.has_levels) {
    if (pg.v2_def_len >= 0) {
      hybrid_init(v.levels, p, p + pg.v2_def_len, 1);
      p += pg.v2_def_len;
    } else if (pg.def_encoding == ENC_RLE) {
      if (pg.size < 4) return false;
      const uint32_t len = load_le32(p);
      if ((int64_t)len + 4 > pg.size) return false;
      hybrid_init(v.levels, p + 4, p + 4 + len, 1);
      p += 4 + len;
    } else if (pg.def_encoding == ENC_BIT_PACKED) {
      legacy_init(v.levels, p, pg.num_values, 1);
      p += (pg.num_values + 7) >> 3;
    } else {
      return false;
    }
  }
  v.values = p;
  v.values_end = end;
  return p <= end;
}

// the dictionary-index / RLE-boolean stream of a page (false when the page has none)
__device__ __forceinline__ bool open_index_stream(const PqPage& pg, const PqChunk& ch, const PageView& v, Hybrid& idx) {
  const bool dict = pg.encoding == ENC_PLAIN_DICT || pg.encoding == ENC_RLE_DICT;
  const bool rle_bool = pg.encoding == ENC_RLE && ch.conv == CONV_BOOL;  // data page v2 booleans
  if (dict) {
    hybrid_init(idx, v.values + 1, v.values_end, v.values < v.values_end ? v.values[0] : 0);
    return true;
  }
  if (rle_bool) {  // <4-byte length> <hybrid runs, bit width 1>
    const uint32_t len = v.values + 4 <= v.values_end ? load_le32(v.values) : 0;
    hybrid_init(idx, v.values + 4, (v.values + 4 + len < v.values_end) ? (v.values + 4 + len) : v.values_end, 1);
    return true;
  }
  return false;
}

__device__ __forceinline__ void or_bits(uint8_t* bitmap, int64_t bit0, uint32_t word) {
  // bitmap is 4-byte aligned (cudaMalloc), zero initialised; neighbouring pages share edge words
  uint32_t* w = reinterpret_cast<uint32_t*>(bitmap) + (bit0 >> 5);
  const int sh = (int)(bit0 & 31);
  if (word << sh) atomicOr(w, word << sh);
  if (sh && (word >> (32 - sh))) atomicOr(w + 1, word >> (32 - sh));
}

__device__ __forceinline__ void emit_fixed(const PqChunk& ch, const PqColOut& out, const uint8_t* src, int64_t row) {
  uint8_t* dst = reinterpret_cast<uint8_t*>(out.values) + row * ch.out_width;
  switch (ch.conv) {
    case CONV_COPY: {
      const uint64_t v = load_le(src, ch.out_width);
      switch (ch.out_width) {
        case 1: *dst = (uint8_t)v; break;
        case 2: *reinterpret_cast<uint16_t*>(dst) = (uint16_t)v; break;
        case 4: *reinterpret_cast<uint32_t*>(dst) = (uint32_t)v; break;
        default: *reinterpret_cast<uint64_t*>(dst) = v; break;
      }
    } break;
    case CONV_DEC_INT: {
      const int n = ch.phys == PQ_INT32 ? 4 : 8;
      uint64_t lo = load_le(src, n);
      if (n == 4) lo = (uint64_t)(int64_t)(int32_t)(uint32_t)lo;
      const uint64_t hi = ((int64_t)lo < 0) ? ~0ull : 0ull;
      reinterpret_cast<uint64_t*>(dst)[0] = lo;
      reinterpret_cast<uint64_t*>(dst)[1] = hi;
    } break;
    case CONV_DEC_FLBA: {
      const int n = ch.type_length;  // big endian two's complement, n <= 16
      const uint64_t fill = (src[0] & 0x80) ? ~0ull : 0ull;
      uint64_t lo = fill, hi = fill;
      for (int i = 0; i < n; ++i) {
        hi = (hi << 8) | (lo >> 56);
        lo = (lo << 8) | src[i];
      }
      reinterpret_cast<uint64_t*>(dst)[0] = lo;
      reinterpret_cast<uint64_t*>(dst)[1] = hi;
    } break;
    default: {  // CONV_FLBA_RAW
      for (int i = 0; i < ch.out_width; ++i) dst[i] = src[i];
    } break;
  }
}

__device__ __forceinline__ int phys_width(const PqChunk& ch) {
  switch (ch.phys) {
    case PQ_INT32: case PQ_FLOAT: return 4;
    case PQ_INT64: case PQ_DOUBLE: return 8;
    case PQ_INT96: return 12;
    case PQ_FLBA: return ch.type_length;
    default: return 0;
  }
}

// n bytes src -> dst, any alignment, one lane. Word loop when the two are congruent mod 4.
__device__ __forceinline__ void copy_bytes(uint8_t* dst, const uint8_t* src, int n) {
  int i = 0;
  if (((reinterpret_cast<uintptr_t>(dst) ^ reinterpret_cast<uintptr_t>(src)) & 3) == 0) {
    for (; i < n && (reinterpret_cast<uintptr_t>(dst + i) & 3); ++i) dst[i] = src[i];
    for (; i + 4 <= n; i += 4) *reinterpret_cast<uint32_t*>(dst + i) = *reinterpret_cast<const uint32_t*>(src + i);
  }
  for (; i < n; ++i) dst[i] = src[i];
}

}  // namespace

// start of every BYTE_ARRAY dictionary entry's payload, one warp per chunk: the length-prefix chain of the
// dictionary page is walked by lane 0 inside a shared-memory window refilled by the whole warp (same scheme as
// the PLAIN BYTE_ARRAY walk of the index pass)
__global__ void __launch_bounds__(128) dict_index_kernel(DecodeArgs A) {
  __shared__ __align__(16) uint8_t s_win_all[4][kWalkWindow];
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (c >= A.nchunks) return;
  const PqChunk ch = A.chunks[c];
  if (ch.dict_nvals <= 0 || ch.phys != PQ_BYTE_ARRAY) return;
  uint8_t* s_win = s_win_all[threadIdx.x >> 5];
  const uint8_t* base = A.raw + ch.dict_off;
  const int64_t avail = ch.dict_size;
  const uintptr_t vaddr = reinterpret_cast<uintptr_t>(base);
  int32_t* pos = A.dict_pos + ch.dict_index;
  const int n = ch.dict_nvals;
  int64_t cursor = 0;
  int j = 0, bad = 0;
  while (j < n && !bad) {
    const uintptr_t wbeg = (vaddr + (uintptr_t)cursor) & ~(uintptr_t)15;
    const uintptr_t vend = vaddr + (uintptr_t)avail;
#pragma unroll
    for (int i = 0; i < kWalkWindow / 512; ++i) {
      const uintptr_t a = wbeg + (uintptr_t)(lane + 32 * i) * 16;
      uint4 q = make_uint4(0, 0, 0, 0);
      if (a < vend) q = *reinterpret_cast<const uint4*>(a);
      reinterpret_cast<uint4*>(s_win)[lane + 32 * i] = q;
    }
    __syncwarp();
    if (lane == 0) {
      const int64_t wofs = (int64_t)(wbeg - vaddr);
      const int64_t lim = min(wofs + (int64_t)kWalkWindow, avail) - 4;
      if (cursor > lim) bad = 1;  // a fresh window always holds the prefix at `cursor` unless the page ends
      while (j < n && cursor <= lim) {
        const uint8_t* q = s_win + (cursor - wofs);
        const uint32_t len = (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16) | ((uint32_t)q[3] << 24);
        pos[j] = (int32_t)(cursor + 4);
        cursor += 4 + (int64_t)len;
        ++j;
      }
    }
    cursor = __shfl_sync(0xffffffffu, cursor, 0);
    j = __shfl_sync(0xffffffffu, j, 0);
    bad = __shfl_sync(0xffffffffu, bad, 0);
    __syncwarp();
  }
  if (lane == 0) {
    if (bad || cursor > avail) {
      atomicExch(A.err, kErrBadPage);
      for (int t = 0; t <= n; ++t) pos[t] = 4;  // every entry reads as empty
    } else {
      pos[n] = (int32_t)(cursor + 4);
    }
  }
}

// ---- index pass: one warp per page ------------------------------------------------------------
__global__ void __launch_bounds__(128) page_index_kernel(DecodeArgs A) {
  const int pgi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (pgi >= A.npages) return;
  PqPage& pg = A.pages[pgi];
  const PqChunk ch = A.chunks[pg.chunk];
  const uint8_t* page = A.raw + pg.src_off;
  const int n = pg.num_values;
  const int nblk = (n + kDecodeBlockRows - 1) / kDecodeBlockRows;
  PqBlock* blocks = A.blocks + pg.first_block;
  PageView v;
  if (!open_page(A, pg, ch, v)) {
    if (lane == 0) {
      atomicExch(A.err, kErrBadPage);
      pg.nonnull = 0;
      pg.values_off = 0;
      for (int k = 0; k < nblk; ++k) { blocks[k].page = -1; blocks[k].str_bytes = 0; }
    }
    return;
  }
  // 1. definition levels: non-null count before every block
  int vb = 0;
  for (int k = 0; k < nblk; ++k) {
    if (lane == 0) {
      PqBlock& b = blocks[k];
      b.page = pgi;
      b.row0 = k * kDecodeBlockRows;
      b.vbase = vb;
      b.pad = 0;
      b.str_bytes = 0;
      b.str_base = 0;
      if (v.has_levels) hybrid_save(v.levels, page, b.levels);
    }
    const int cnt = min(kDecodeBlockRows, n - k * kDecodeBlockRows);
    vb += v.has_levels ? hybrid_count_ones(v.levels, cnt, lane) : cnt;
  }
  const int nonnull = vb;
  if (lane == 0) {
    pg.nonnull = nonnull;
    pg.values_off = (int32_t)(v.values - page);
  }
  __syncwarp();
  // 2. index stream (dictionary indices / RLE booleans): state at the first value of every block
  Hybrid idx;
  if (open_index_stream(pg, ch, v, idx)) {
    int consumed = 0;
    for (int k = 0; k < nblk; ++k) {
      const int target = blocks[k].vbase;
      hybrid_skip(idx, target - consumed);
      consumed = target;
      if (lane == 0) hybrid_save(idx, page, blocks[k].index);
    }
  }
  // 3. PLAIN BYTE_ARRAY: the length-prefix chain can only be walked value by value. The warp stages
  //    the page through a shared-memory window (coalesced 16-byte loads, next window prefetched into
  //    L2) and lane 0 hops from prefix to prefix inside it: ~30 cycles per value instead of a DRAM
  //    round trip per 128-byte line (measured 300 cycles per value when walking global memory).
  if (ch.conv == CONV_VAR && pg.encoding == ENC_PLAIN) {
    __shared__ __align__(16) uint8_t s_win_all[4][kWalkWindow];
    uint8_t* s_win = s_win_all[threadIdx.x >> 5];
    int32_t* vpos = A.vpos + pg.vpos_base;
    const int64_t avail = v.values_end - v.values;
    const uintptr_t vaddr = reinterpret_cast<uintptr_t>(v.values);
    int64_t cursor = 0;  // of the next length prefix, relative to v.values
    int j = 0, k = 0;
    int bad = 0;
    while (true) {
      // window = [wbeg, wbeg + kWalkWindow) in absolute 16-byte aligned addresses, covering `cursor`
      const uintptr_t wbeg = (vaddr + (uintptr_t)cursor) & ~(uintptr_t)15;
      const uintptr_t vend = vaddr + (uintptr_t)avail;
#pragma unroll
      for (int i = 0; i < kWalkWindow / 512; ++i) {
        const uintptr_t a = wbeg + (uintptr_t)(lane + 32 * i) * 16;
        uint4 q = make_uint4(0, 0, 0, 0);
        if (a < vend) q = *reinterpret_cast<const uint4*>(a);  // the raw image is padded by 64 bytes
        reinterpret_cast<uint4*>(s_win)[lane + 32 * i] = q;
      }
      {
        const uintptr_t nxt = wbeg + kWalkWindow + (uintptr_t)lane * 128;
        if (nxt < vend) asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt));
      }
      __syncwarp();
      if (lane == 0) {
        const int64_t wofs = (int64_t)(wbeg - vaddr);          // window start relative to v.values (may be < 0)
        const int64_t wlim = min(wofs + (int64_t)kWalkWindow, avail);  // prefixes must end at or before this
        // next block boundary (value index) kept in a register: the hop loop must not wait on global loads
        int target = k < nblk ? blocks[k].vbase : nonnull;
        int64_t prev_sb = k > 0 ? blocks[k - 1].str_base : 0;
        while (true) {
          while (k <= nblk && j == target) {  // block boundaries reached at value j
            const int64_t sb = cursor - 4ll * j;
            if (k > 0) blocks[k - 1].str_bytes = sb - prev_sb;
            if (k < nblk) blocks[k].str_base = sb;  // page-relative for now (rebased by the scan)
            prev_sb = sb;
            ++k;
            target = k < nblk ? blocks[k].vbase : nonnull;
          }
          if (j >= nonnull) break;
          if (cursor + 4 > avail) { bad = 1; break; }
          // hop inside the window up to the next boundary
          const int64_t lim = wlim - 4;
          while (j < target && cursor <= lim) {
            const uint8_t* q = s_win + (cursor - wofs);
            const uint32_t len = (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16) | ((uint32_t)q[3] << 24);
            vpos[j] = (int32_t)(cursor - 4ll * j);
            cursor += 4 + (int64_t)len;
            ++j;
          }
          if (j < target) break;  // window exhausted (or the page ends: checked after the refill)
        }
      }
      cursor = __shfl_sync(0xffffffffu, cursor, 0);
      j = __shfl_sync(0xffffffffu, j, 0);
      k = __shfl_sync(0xffffffffu, k, 0);
      bad = __shfl_sync(0xffffffffu, bad, 0);
      if (bad || j >= nonnull) break;
      __syncwarp();
    }
    if (lane == 0) {
      if (bad || cursor > avail) {
        atomicExch(A.err, kErrBadPage);
        for (int t = 0; t <= nonnull; ++t) vpos[t] = 0;
        for (int t = 0; t < nblk; ++t) blocks[t].str_bytes = 0;
      } else {
        vpos[nonnull] = (int32_t)(cursor - 4ll * nonnull);
      }
    }
  }
}

// ---- dictionary-coded strings: bytes produced by every block (one warp per block) -----------
__global__ void __launch_bounds__(128) block_count_kernel(DecodeArgs A) {
  const int bi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (bi >= A.nblocks) return;
  PqBlock& blk = A.blocks[bi];
  if (blk.page < 0) return;
  const PqPage pg = A.pages[blk.page];
  const PqChunk ch = A.chunks[pg.chunk];
  const bool dict = pg.encoding == ENC_PLAIN_DICT || pg.encoding == ENC_RLE_DICT;
  if (ch.conv != CONV_VAR || !dict) return;
  const uint8_t* page = A.raw + pg.src_off;
  const uint8_t* values = page + pg.values_off;
  const int n_blk = min(kDecodeBlockRows, pg.num_values - blk.row0);
  const int nv_blk = ((blk.row0 + n_blk >= pg.num_values) ? pg.nonnull : (&blk)[1].vbase) - blk.vbase;
  Hybrid idx;
  hybrid_init(idx, values + 1, page + pg.size, values < page + pg.size ? values[0] : 0);
  hybrid_restore(idx, page, blk.index);
  const int32_t* dpos = A.dict_pos + ch.dict_index;
  int64_t bytes = 0;
  for (int j0 = 0; j0 < nv_blk; j0 += 32) {
    const int cnt = min(32, nv_blk - j0);
    const uint32_t di = hybrid_get(idx, cnt, lane);
    int len = 0;
    if (lane < cnt && (int)di < ch.dict_nvals) len = dpos[di + 1] - 4 - dpos[di];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) len += __shfl_xor_sync(0xffffffffu, len, d);
    bytes += len;
  }
  if (lane == 0) blk.str_bytes = bytes;
}

// exclusive scan of str_bytes over the blocks of each (column, bucket) run of pages: pages are listed
// in row order and so are their blocks; one warp per run
__global__ void __launch_bounds__(128) block_scan_kernel(DecodeArgs A, const int32_t* seg_bounds, const int64_t* seg_base,
                                                         int nseg, int64_t* seg_total) {
  const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (s >= nseg) return;
  const int p0 = seg_bounds[2 * s], p1 = seg_bounds[2 * s + 1];
  if (p0 >= p1) {
    if (lane == 0) seg_total[s] = 0;
    return;
  }
  const int b0 = A.pages[p0].first_block;
  const PqPage& last = A.pages[p1 - 1];
  const int b1 = last.first_block + (last.num_values + kDecodeBlockRows - 1) / kDecodeBlockRows;
  int64_t run = seg_base[s];  // bytes of the bucket's host-decoded files come first
  for (int i0 = b0; i0 < b1; i0 += 32) {
    const int i = i0 + lane;
    const int64_t mine = i < b1 ? A.blocks[i].str_bytes : 0;
    int64_t incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int64_t y = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += y;
    }
    if (i < b1) A.blocks[i].str_base = run + incl - mine;
    run += __shfl_sync(0xffffffffu, incl, 31);
  }
  if (lane == 0) seg_total[s] = run - seg_base[s];
}

// ---- decode pass: one warp per block ------------------------------------------------------------
__global__ void __launch_bounds__(128) block_decode_kernel(DecodeArgs A) {
  const int bi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (bi >= A.nblocks) return;
  const PqBlock blk = A.blocks[bi];
  if (blk.page < 0) return;
  const PqPage pg = A.pages[blk.page];
  const PqChunk ch = A.chunks[pg.chunk];
  const PqColOut out = A.cols[ch.col];
  const uint8_t* page = A.raw + pg.src_off;
  const uint8_t* page_end = page + pg.size;
  const uint8_t* values = page + pg.values_off;
  const int n_blk = min(kDecodeBlockRows, pg.num_values - blk.row0);
  const bool last_block = blk.row0 + n_blk >= pg.num_values;
  const int64_t row_base = pg.row_start + blk.row0;
  const bool dict = pg.encoding == ENC_PLAIN_DICT || pg.encoding == ENC_RLE_DICT;
  const bool rle_bool = pg.encoding == ENC_RLE && ch.conv == CONV_BOOL;
  if (!dict && !rle_bool && pg.encoding != ENC_PLAIN) {
    if (lane == 0) atomicExch(A.err, kErrBadPage);
    return;
  }
  const bool all_valid = pg.nonnull == pg.num_values;
  const int pw = phys_width(ch);

  // ---- fast path: dense PLAIN values of a 4/8-byte type are a (mis)aligned memcpy -------------
  if (all_valid && !dict && ch.conv == CONV_COPY && ch.out_width == pw && (pw == 4 || pw == 8)) {
    const uint8_t* src = values + (int64_t)blk.vbase * pw;
    if (src + (int64_t)n_blk * pw > page_end) {
      if (lane == 0) atomicExch(A.err, kErrBadPage);
      return;
    }
    uint32_t* dst = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(out.values) + row_base * pw);
    const int nwords = n_blk * (pw >> 2);
    const unsigned sh = (unsigned)(reinterpret_cast<uintptr_t>(src) & 3) * 8;
    const uint32_t* sw = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(src) & ~(uintptr_t)3);
    if (sh == 0) {
      for (int i = lane; i < nwords; i += 32) dst[i] = sw[i];
    } else {
      for (int i = lane; i < nwords; i += 32) dst[i] = __funnelshift_r(sw[i], sw[i + 1], sh);
    }
    if (out.validity) {
      for (int r0 = lane * 32; r0 < n_blk; r0 += 32 * 32) {
        const int c = min(32, n_blk - r0);
        or_bits(out.validity, row_base + r0, c == 32 ? 0xffffffffu : ((1u << c) - 1));
      }
    }
    return;
  }

  // ---- general path: 32 rows per step, streams restarted from the block's checkpoint ----------
  Hybrid levels, idx;
  const bool has_levels = ch.max_def > 0 && !all_valid;
  if (has_levels) {
    hybrid_init(levels, page, values, 1);
    if (pg.def_encoding == ENC_BIT_PACKED && pg.v2_def_len < 0) legacy_init(levels, page, pg.num_values, 1);
    else if (pg.v2_def_len >= 0) hybrid_init(levels, page, page + pg.v2_def_len, 1);
    hybrid_restore(levels, page, blk.levels);
  }
  if (dict) {
    hybrid_init(idx, values + 1, page_end, values < page_end ? values[0] : 0);
    hybrid_restore(idx, page, blk.index);
  } else if (rle_bool) {
    const uint32_t len = values + 4 <= page_end ? load_le32(values) : 0;
    hybrid_init(idx, values + 4, (values + 4 + len < page_end) ? (values + 4 + len) : page_end, 1);
    hybrid_restore(idx, page, blk.index);
  }
  const uint8_t* dict_base = A.raw + ch.dict_off;
  const int32_t* dpos = A.dict_pos + ch.dict_index;
  const int32_t* vpos = (ch.conv == CONV_VAR && !dict) ? A.vpos + pg.vpos_base : nullptr;
  // PLAIN BYTE_ARRAY: blk.str_base is the bucket position of the block's first byte and
  // vpos[j] - vpos[blk.vbase] the distance of value j from it
  const int64_t vpos0 = vpos ? vpos[blk.vbase] : 0;
  int64_t vbase = blk.vbase;  // non-null values consumed so far
  int64_t sbytes = 0;         // dictionary strings: bytes produced so far by this block
  int32_t* offs = reinterpret_cast<int32_t*>(out.values);
  uint8_t* sdata = ch.conv == CONV_VAR ? out.data + out.bucket_data[ch.bucket] : nullptr;
  for (int r0 = 0; r0 < n_blk; r0 += 32) {
    const int cnt = min(32, n_blk - r0);
    const int64_t row = row_base + r0 + lane;
    bool valid = lane < cnt;
    if (has_levels) valid = hybrid_get(levels, cnt, lane) != 0 && lane < cnt;
    const uint32_t mask = __ballot_sync(0xffffffffu, valid);
    const int nv = __popc(mask);
    const int rank = __popc(mask & ((1u << lane) - 1));
    if (out.validity && lane == 0) or_bits(out.validity, row_base + r0, mask);
    uint32_t di = 0;
    if (dict || rle_bool) {
      const uint32_t mine = hybrid_get(idx, nv, lane);
      di = __shfl_sync(0xffffffffu, mine, rank);
      if (dict && valid && (int)di >= ch.dict_nvals) { valid = false; atomicExch(A.err, kErrBadPage); }
    }
    if (ch.conv == CONV_VAR) {
      if (dict) {
        const uint8_t* src = nullptr;
        int len = 0;
        if (valid) { src = dict_base + dpos[di]; len = dpos[di + 1] - 4 - dpos[di]; }
        int incl = len;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int y = __shfl_up_sync(0xffffffffu, incl, d);
          if (lane >= d) incl += y;
        }
        const int64_t my_off = blk.str_base + sbytes + (incl - len);
        if (lane < cnt) offs[row] = (int32_t)my_off;
        if (len > 0) copy_bytes(sdata + my_off, src, len);
        sbytes += __shfl_sync(0xffffffffu, incl, 31);
      } else {
        // value j = vbase + rank; a null row takes the offset of the next value (length 0)
        const int64_t j = vbase + rank;
        if (lane < cnt) {
          const int64_t rel = (int64_t)vpos[j] - vpos0;
          const int64_t my_off = blk.str_base + rel;
          offs[row] = (int32_t)my_off;
          if (valid) {
            const int len = vpos[j + 1] - vpos[j];
            const uint8_t* src = values + vpos[j] + 4 * (j + 1);
            if (len > 0 && src + len <= page_end) copy_bytes(sdata + my_off, src, len);
          }
        }
      }
    } else if (ch.conv == CONV_BOOL) {
      bool bit = false;
      if (valid) {
        if (rle_bool) bit = di != 0;
        else if (dict) bit = false;  // booleans are never dictionary encoded
        else {
          const int64_t b = vbase + rank;
          bit = (values[b >> 3] >> (b & 7)) & 1;
        }
      }
      const uint32_t word = __ballot_sync(0xffffffffu, bit);
      if (lane == 0) or_bits(reinterpret_cast<uint8_t*>(out.values), row_base + r0, word);
    } else if (valid) {
      const uint8_t* src = dict ? dict_base + (int64_t)di * pw : values + (vbase + rank) * pw;
      if (dict || src + pw <= page_end) emit_fixed(ch, out, src, row);
    }
    vbase += nv;
  }
  if (ch.conv == CONV_VAR && pg.closes_file && last_block && lane == 0) {
    const int64_t end_off = dict ? blk.str_base + sbytes : blk.str_base + ((int64_t)vpos[pg.nonnull] - vpos0);
    offs[pg.row_start + pg.num_values] = (int32_t)end_off;
  }
}


// ---------------------------------------------------------------------------------------------------
// ZSTD page payloads (the LakeSoul writer compresses every page with ZSTD, multipart_writer.rs:128): one warp
// per page. The format is sequential inside a frame (entropy-coded bit streams read backwards, matches that
// reference earlier output), so one lane parses while the warp does what is parallel: the four Huffman
// streams of a block's literals are decoded by four lanes, literal and match copies by all 32. The tables
// live in shared memory; zstd_decode.h holds the format logic (shared with the host unit test).
// ---------------------------------------------------------------------------------------------------
namespace {

// backward bit reader for the hot loops: a 64-bit register window over the stream, refilled with byte loads only
// when fewer than 32 unread bits are left in it. The stream sits inside the (padded) file image; going below its
// start is caught through bits_left < 0 (the missing low bits read as zeros).
struct FastBack {
  const uint8_t* p;
  int64_t bits_left;   // unread bits of the stream
  uint64_t win;        // bits [wlo, wlo + 64) of the stream, little-endian bit numbering
  int64_t wlo;
  __device__ __forceinline__ void refill() {
    // window ending at the unread position: its top bit is bit (bits_left - 1) when possible
    int64_t lo = ((bits_left + 7) & ~(int64_t)7) - 64;  // the window's top byte is the last byte with unread bits
    if (lo < 0) lo = 0;
    const uint8_t* q = p + (lo >> 3);
    uint64_t v = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) v |= (uint64_t)q[i] << (8 * i);  // at most 7 bytes past the stream end: image padding
    win = v;
    wlo = lo;
  }
  __device__ __forceinline__ bool init(const uint8_t* src, int64_t size) {
    p = src;
    if (size <= 0 || src[size - 1] == 0) return false;
    bits_left = size * 8 - (8 - (31 - __clz((unsigned)src[size - 1])));
    refill();
    return true;
  }
  __device__ __forceinline__ uint32_t peek(int n) {  // n <= 24
    if (n == 0) return 0u;
    int64_t lo = bits_left - n;
    if (lo < 0) {  // over-read: zeros below the start
      const int drop = (int)-lo;
      if (drop >= n) return 0u;
      if (wlo != 0) refill();
      return ((uint32_t)win & ((1u << (n - drop)) - 1)) << drop;
    }
    if (lo < wlo) refill();
    return (uint32_t)(win >> (lo - wlo)) & ((1u << n) - 1);
  }
  __device__ __forceinline__ uint32_t read(int n) {
    const uint32_t v = peek(n);
    bits_left -= n;
    return v;
  }
};

// n bytes src -> dst by one lane, ranges do not overlap: eight loads are in flight before the first store (a
// plain byte loop waits for every load, the compiler cannot move loads above possibly aliasing stores)
__device__ __forceinline__ void copy_lane(uint8_t* __restrict__ d, const uint8_t* __restrict__ s, uint32_t n) {
  uint32_t i = 0;
  for (; i + 8 <= n; i += 8) {
    uint8_t t[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) t[k] = s[i + k];
#pragma unroll
    for (int k = 0; k < 8; ++k) d[i + k] = t[k];
  }
  const uint32_t r = n - i;
  uint8_t t[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) if ((uint32_t)k < r) t[k] = s[i + k];
#pragma unroll
  for (int k = 0; k < 8; ++k) if ((uint32_t)k < r) d[i + k] = t[k];
}

// FSE bit stream of one block -> packed sequence records (device twin of zstd_decode_sequences: register-window reader)
__device__ int64_t sequences_device(const uint8_t* bits, int64_t size, const ZstdBlockPlan& P, const ZstdTables& T, uint64_t* out) {
  using namespace zstd_detail;
  FastBack bb;
  bb.win = 0; bb.wlo = 0;
  if (!bb.init(bits, size)) return ZSTD_E_CORRUPT;
  uint32_t sl = bb.read(P.ll_log), so = bb.read(P.of_log), sm = bb.read(P.ml_log);
  for (int32_t q = 0; q < P.nseq; ++q) {
    const FseEntry eo = T.of[so], em = T.ml[sm], el = T.ll[sl];
    const int oc = eo.symbol, mc = em.symbol, lc = el.symbol;
    if (oc > 29 || mc > 52 || lc > 35) return ZSTD_E_CORRUPT;
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
      sl = el.baseline + bb.read(el.nbits);
      sm = em.baseline + bb.read(em.nbits);
      so = eo.baseline + bb.read(eo.nbits);
    }
    if (bb.bits_left < 0) return ZSTD_E_CORRUPT;
    out[q] = zstd_pack_seq(ll, ml, ov);
  }
  return bb.bits_left == 0 ? 0 : (int64_t)ZSTD_E_CORRUPT;
}

__device__ __forceinline__ int64_t huf_stream_fast(const uint16_t* huf, int log, const uint8_t* src, int64_t size,
                                                   uint8_t* dst, int64_t n) {
  FastBack bb;
  bb.win = 0; bb.wlo = 0;
  if (!bb.init(src, size)) return ZSTD_E_CORRUPT;
  for (int64_t i = 0; i < n; ++i) {
    const uint16_t e = huf[bb.peek(log)];
    dst[i] = (uint8_t)e;
    bb.bits_left -= e >> 8;
  }
  return bb.bits_left == 0 ? n : (int64_t)ZSTD_E_CORRUPT;
}

}  // namespace

// One thread block per page, three steps separated by block barriers (the phased pipeline of zstd_decode.h):
//   plan    thread 0: section headers and entropy tables of the page's blocks, in order (tables can be inherited)
//   entropy 8 threads per block, all blocks side by side: 4 Huffman streams, the FSE sequence stream, raw / RLE literals
//   execute warp 0, in order, 32 sequences per round: repeat offsets (lane 0), output positions (scan), literal runs
//           in parallel, matches whose source lies before the round's output in parallel, the rest one by one
constexpr int kZstdThreads = 64;
// development profile (LAKESOUL_B200_ZSTD_PROFILE=1): cycles of the three phases summed over the pages >= 256 KiB
__device__ unsigned long long g_zstd_prof[16];
__global__ void __launch_bounds__(kZstdThreads) zstd_pages_kernel(const PqZstdPage* pages, const ZstdBlockDesc* blocks_all,
                                                                  ZstdBlockPlan* plans_all, ZstdTables* tables_all, uint8_t* lit_all,
                                                                  uint64_t* seq_all, int32_t* err, const int32_t* order = nullptr) {
  __shared__ ZstdPlanScratch s_plan;
  __shared__ int s_fail;
  __shared__ uint32_t s_round[96];  // ll | ml | resolved offset of the 32 sequences of a round
  const PqZstdPage pg = pages[order ? order[blockIdx.x] : blockIdx.x];
  const uint8_t* src = reinterpret_cast<const uint8_t*>(pg.src_off);
  uint8_t* dst = reinterpret_cast<uint8_t*>(pg.dst_off);
  const ZstdBlockDesc* blocks = blocks_all + pg.first_block;
  ZstdBlockPlan* plans = plans_all + pg.first_block;
  ZstdTables* tables = tables_all + pg.first_block;
  uint8_t* lit = lit_all + pg.lit_off;
  uint64_t* seq = seq_all + pg.seq_off;
  const int nb = pg.nblocks;
  const int tid = threadIdx.x;
  // plan: the tables of a block are built in shared memory (thread 0, serial: a block may inherit the tables of the
  // one before it) and copied to the page's table scratch by the whole thread block
  __shared__ ZstdTables s_tab;
  __shared__ ZstdBlockPlan s_prev_plan, s_cur_plan;
  __shared__ int s_have_prev;
  if (tid == 0) { s_fail = 0; s_have_prev = 0; }
  __syncthreads();
  const bool prof = pg.raw_size >= (256 << 10);
  const long long t_start = clock64();
  {
    int64_t lit_total = 0, seq_total = 0;  // meaningful in thread 0
    for (int k = 0; k < nb; ++k) {
      if (blocks[k].type != 2) continue;  // uniform: block descriptors are the same for every thread
      if (tid == 0 && !s_fail) {
        // s_tab still holds the previous compressed block's tables: inherited entries stay where they are
        const int64_t r = zstd_plan_block(src + blocks[k].src_off, blocks[k].size, s_cur_plan, s_tab, s_have_prev ? &s_prev_plan : nullptr,
                                          s_have_prev ? &s_tab : nullptr, s_plan);
        if (r < 0) s_fail = 1;
        s_cur_plan.lit_base = (int32_t)lit_total;
        s_cur_plan.seq_base = (int32_t)seq_total;
        lit_total += s_cur_plan.regen;
        seq_total += s_cur_plan.nseq;
        if (lit_total > pg.raw_size || seq_total > pg.raw_size / 3 + 1) s_fail = 1;
        s_prev_plan = s_cur_plan;
        s_have_prev = 1;
        plans[k] = s_cur_plan;
      }
      __syncthreads();
      if (!s_fail) {
        const uint32_t* from = reinterpret_cast<const uint32_t*>(&s_tab);
        uint32_t* to = reinterpret_cast<uint32_t*>(&tables[k]);
        for (int i = tid; i < (int)(sizeof(ZstdTables) / 4); i += kZstdThreads) to[i] = from[i];
      }
      __syncthreads();
    }
  }
  __syncthreads();
  if (s_fail) {
    if (tid == 0) atomicExch(err, kErrBadPage);
    return;
  }
  const long long t_plan = clock64();
  // ---- entropy: 8 roles per block --------------------------------------------------------------
  for (int k = tid >> 3; k < nb; k += kZstdThreads >> 3) {
    if (blocks[k].type != 2) continue;
    const ZstdBlockPlan& P = plans[k];
    const uint8_t* b = src + blocks[k].src_off;
    uint8_t* L = lit + P.lit_base;
    const int role = tid & 7;
    int64_t r = 0;
    if (role < 4) {
      if (P.lit_type == 2 && role < P.streams) {
        const int64_t per = P.streams == 4 ? (P.regen + 3) / 4 : P.regen;
        int64_t so = 0;
        for (int st = 0; st < role; ++st) so += P.stream_size[st];
        const int64_t n = role < 3 ? per : P.regen - 3 * per;
        r = huf_stream_fast(tables[k].huf, P.huf_log, b + P.lit_src + so, P.stream_size[role], L + role * per, P.streams == 4 ? n : P.regen);
      }
    } else if (role == 4) {
      if (P.nseq > 0) r = sequences_device(b + P.seq_src, P.seq_size, P, tables[k], seq + P.seq_base);
    } else if (role == 5) {
      if (P.lit_type == 0) for (int i = 0; i < P.regen; ++i) L[i] = b[P.lit_src + i];
      else if (P.lit_type == 1) { const uint8_t v = b[P.lit_src]; for (int i = 0; i < P.regen; ++i) L[i] = v; }
    }
    if (r < 0) s_fail = 1;
  }
  __syncthreads();
  if (s_fail) {
    if (tid == 0) atomicExch(err, kErrBadPage);
    return;
  }
  const long long t_entropy = clock64();
  if (tid >= 32) return;
  // ---- execute: warp 0 ----------------------------------------------------------------------------
  const int lane = tid;
  uint32_t rep0 = 1, rep1 = 4, rep2 = 8;  // kept by lane 0
  int64_t op = 0;
  const int64_t cap = pg.raw_size;
  for (int k = 0; k < nb; ++k) {
    const uint8_t* b = src + blocks[k].src_off;
    if (blocks[k].type != 2) {
      const int64_t n = blocks[k].out_size;
      if (op + n > cap) { if (lane == 0) atomicExch(err, kErrBadPage); return; }
      if (blocks[k].type == 0) for (int64_t i = lane; i < n; i += 32) dst[op + i] = b[i];
      else { const uint8_t v = b[0]; for (int64_t i = lane; i < n; i += 32) dst[op + i] = v; }
      op += n;
      __syncwarp();
      continue;
    }
    const ZstdBlockPlan P = plans[k];
    const uint8_t* lits = lit + P.lit_base;
    const uint64_t* sq = seq + P.seq_base;
    int64_t lpos = 0;
    for (int q0 = 0; q0 < P.nseq; q0 += 32) {
      const int cnt = min(32, P.nseq - q0);
      const uint64_t v = lane < cnt ? sq[q0 + lane] : 0ull;
      const uint32_t ll = (uint32_t)(v & 0x1ffff), ml = (uint32_t)((v >> 17) & 0x1ffff);
      s_round[lane] = ll;
      s_round[64 + lane] = (uint32_t)(v >> 34);
      __syncwarp();
      int bad = 0;
      if (lane == 0) {  // repeat-offset history is a chain through the sequences
        uint32_t rep[3] = {rep0, rep1, rep2};
        for (int i = 0; i < cnt; ++i) {
          const uint32_t o = zstd_resolve_offset(s_round[64 + i], s_round[i], rep);
          if (o == 0) bad = 1;
          s_round[64 + i] = o;
        }
        rep0 = rep[0]; rep1 = rep[1]; rep2 = rep[2];
      }
      __syncwarp();
      const uint32_t offset = lane < cnt ? s_round[64 + lane] : 1u;
      __syncwarp();
      uint32_t io = ll + ml, il = ll;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t yo = __shfl_up_sync(0xffffffffu, io, d), yl = __shfl_up_sync(0xffffffffu, il, d);
        if (lane >= d) { io += yo; il += yl; }
      }
      const uint32_t tot_out = __shfl_sync(0xffffffffu, io, 31), tot_lit = __shfl_sync(0xffffffffu, il, 31);
      const int64_t my_out = op + (io - ll - ml);
      const int64_t my_lit = lpos + (il - ll);
      const int64_t my_match = my_out + ll;
      bad = __shfl_sync(0xffffffffu, bad, 0) || lpos + tot_lit > P.regen || op + tot_out > cap;
      if (__any_sync(0xffffffffu, bad || (lane < cnt && (int64_t)offset > my_match))) {
        if (lane == 0) atomicExch(err, kErrBadPage);
        return;
      }
      if (ll <= 64) copy_lane(dst + my_out, lits + my_lit, ll);
      unsigned big = __ballot_sync(0xffffffffu, ll > 64);
      while (big) {
        const int j = __ffs(big) - 1;
        big &= big - 1;
        const int64_t o = __shfl_sync(0xffffffffu, my_out, j), l0 = __shfl_sync(0xffffffffu, my_lit, j);
        const uint32_t n = __shfl_sync(0xffffffffu, ll, j);
        for (uint32_t i = lane; i < n; i += 32) dst[o + i] = lits[l0 + i];
      }
      __syncwarp();
      const bool indep = lane < cnt && ml <= 64 && (my_match - (int64_t)offset + ml <= op);
      if (indep) copy_lane(dst + my_match, dst + my_match - offset, ml);
      __syncwarp();
      unsigned serial = __ballot_sync(0xffffffffu, lane < cnt && !indep && ml > 0);
      while (serial) {
        const int j = __ffs(serial) - 1;
        serial &= serial - 1;
        const int64_t md = __shfl_sync(0xffffffffu, my_match, j);
        const uint32_t off = __shfl_sync(0xffffffffu, offset, j), n = __shfl_sync(0xffffffffu, ml, j);
        const uint8_t* m = dst + md - off;
        if (off >= n) {
          for (uint32_t i = lane; i < n; i += 32) dst[md + i] = m[i];
        } else {  // overlapping match: only bytes that existed before this sequence are read
          for (uint32_t i = lane; i < n; i += 32) dst[md + i] = m[i % off];
        }
        __syncwarp();
      }
      op += tot_out;
      lpos += tot_lit;
    }
    const int64_t rest = P.regen - lpos;
    if (rest < 0 || op + rest > cap) { if (lane == 0) atomicExch(err, kErrBadPage); return; }
    for (int64_t i = lane; i < rest; i += 32) dst[op + i] = lits[lpos + i];
    op += rest;
    __syncwarp();
  }
  if (op != cap && lane == 0) atomicExch(err, kErrBadPage);
  if (prof && lane == 0) {
    const long long t_end = clock64();
    atomicAdd(&g_zstd_prof[0], (unsigned long long)(t_plan - t_start));
    atomicAdd(&g_zstd_prof[1], (unsigned long long)(t_entropy - t_plan));
    atomicAdd(&g_zstd_prof[2], (unsigned long long)(t_end - t_entropy));
    atomicAdd(&g_zstd_prof[3], 1ull);
    atomicMax(&g_zstd_prof[4], (unsigned long long)(t_plan - t_start));
    atomicMax(&g_zstd_prof[5], (unsigned long long)(t_entropy - t_plan));
    atomicMax(&g_zstd_prof[6], (unsigned long long)(t_end - t_entropy));
  }
}

// ---------------------------------------------------------------------------------------------------
// ZSTD pages, second generation (the default): the same three phases, but
//   * entropy tables live in shared memory (one slot per block of a group), the bit readers keep a 64-bit
//     container in registers and refill it with two aligned 8-byte loads (zstd's BIT_DStream scheme): a symbol
//     costs one shared-memory lookup instead of a chain of global loads (r1 profile: ~470 cycles per symbol);
//   * Huffman streams emit 4 symbols per refill and store them as one word;
//   * the sequence decoder resolves repeat offsets on the fly, relative to the block's INITIAL history
//     (symbolically): the only state that crosses blocks becomes three words per block, chained by one thread;
//   * every block is executed by its own warp; a match that reaches into an earlier block waits for that block's
//     progress counter. Within a round (32 sequences) the positions come from a scan, literal runs and matches
//     whose source lies before the round are copied in parallel, the rest in order.
// ---------------------------------------------------------------------------------------------------
namespace {

// backward bit stream (zstd's BIT_DStream scheme): `c` = the 8 bytes at `ptr`, `consumed` bits of it already used
// from the top. The stream is walked from its end to its start; the aligned 8-byte words it will need next sit in a
// register queue that is filled three words ahead, so a refill is a funnel shift, not a memory round trip (with the
// entropy tables in shared memory the L1 is small: r2 profile, ~800 cycles per sequence on un-queued refills).
struct BitD {
  const uint8_t* start;
  const uint8_t* ptr;
  uint64_t c;
  uint32_t consumed;
  uint64_t q0, q1, q2, q3, q4;  // words at base - 24, base - 16, base - 8, base, base + 8 with base = ptr & ~7
  uintptr_t floor_word;         // lowest word address worth loading
  static __device__ __forceinline__ uint64_t word(uintptr_t a) { return __ldg(reinterpret_cast<const uint64_t*>(a)); }
  __device__ __forceinline__ uint64_t word_clamped(uintptr_t a) const { return word(a < floor_word ? floor_word : a); }
  __device__ __forceinline__ void compose() {
    const unsigned sh = (unsigned)(reinterpret_cast<uintptr_t>(ptr) & 7) * 8;
    c = sh == 0 ? q3 : ((q3 >> sh) | (q4 << (64 - sh)));
  }
  __device__ __forceinline__ void fill() {
    const uintptr_t base = reinterpret_cast<uintptr_t>(ptr) & ~(uintptr_t)7;
    q4 = word(base + 8);  // <= 15 bytes past the stream: inside the padded image
    q3 = word(base);
    q2 = word_clamped(base - 8);
    q1 = word_clamped(base - 16);
    q0 = word_clamped(base - 24);
    compose();
  }
  __device__ __forceinline__ bool init(const uint8_t* src, int64_t size) {
    start = src;
    floor_word = reinterpret_cast<uintptr_t>(src) & ~(uintptr_t)7;
    if (size <= 0) return false;
    const uint32_t last = src[size - 1];
    if (last == 0) return false;
    const uint32_t hb = 31 - __clz(last);
    if (size >= 8) {
      ptr = src + size - 8;
      fill();
      consumed = 8 - hb;
    } else {
      ptr = src;
      fill();
      uint64_t v = 0;
      for (int i = 0; i < (int)size; ++i) v |= (uint64_t)src[i] << (8 * i);
      c = v;
      consumed = 8 - hb + (uint32_t)(8 - size) * 8;
    }
    return true;
  }
  __device__ __forceinline__ uint32_t peek(int n) const {  // 1 <= n <= 32
    return (uint32_t)((c << (consumed & 63)) >> (64 - n));
  }
  __device__ __forceinline__ uint32_t read(int n) {
    if (n == 0) return 0u;
    const uint32_t v = peek(n);
    consumed += n;
    return v;
  }
  __device__ __forceinline__ void move_to(const uint8_t* np) {  // np <= ptr, at most 8 bytes lower
    const uintptr_t ob = reinterpret_cast<uintptr_t>(ptr) & ~(uintptr_t)7, nbase = reinterpret_cast<uintptr_t>(np) & ~(uintptr_t)7;
    ptr = np;
    if (nbase != ob) {  // one word down: shift the queue, fetch the word three below
      q4 = q3; q3 = q2; q2 = q1; q1 = q0;
      q0 = word_clamped(nbase - 24);
    }
    compose();
  }
  __device__ __forceinline__ void reload() {
    if (ptr >= start + 8) {
      const uint32_t nb = consumed >> 3;
      consumed &= 7;
      move_to(ptr - nb);
    } else if (ptr != start) {
      uint32_t nb = consumed >> 3;
      if (ptr - nb < start) nb = (uint32_t)(ptr - start);
      consumed -= nb * 8;
      move_to(ptr - nb);
    }
  }
  __device__ __forceinline__ bool finished_exactly() const { return ptr == start && consumed == 64; }
  __device__ __forceinline__ bool overrun() const { return consumed > 64; }
};

// one Huffman stream -> n literal bytes; `huf` in shared memory
__device__ int64_t huf_stream_v2(const uint16_t* huf, int log, const uint8_t* src, int64_t size, uint8_t* dst, int64_t n) {
  BitD d;
  if (!d.init(src, size)) return ZSTD_E_CORRUPT;
  int64_t i = 0;
  while (i < n && (reinterpret_cast<uintptr_t>(dst + i) & 3)) {
    d.reload();
    const uint16_t e = huf[d.peek(log)];
    dst[i++] = (uint8_t)e;
    d.consumed += e >> 8;
  }
  for (; i + 4 <= n; i += 4) {  // 4 x <= 11 bits after a refill that leaves >= 57
    d.reload();
    uint32_t w = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint16_t e = huf[d.peek(log)];
      w |= (uint32_t)(e & 0xff) << (8 * k);
      d.consumed += e >> 8;
    }
    *reinterpret_cast<uint32_t*>(dst + i) = w;
  }
  for (; i < n; ++i) {
    d.reload();
    const uint16_t e = huf[d.peek(log)];
    dst[i] = (uint8_t)e;
    d.consumed += e >> 8;
  }
  d.reload();
  return (!d.overrun() && d.finished_exactly()) ? n : (int64_t)ZSTD_E_CORRUPT;
}

// sequence record of the v2 path: ll (17) | ml (17) << 17 | tag (2) << 34 | value (28) << 36
//   tag 0: value = the offset;  tag t > 0: offset = rep_init[t - 1] + sign-extended value
constexpr uint32_t kSeqValMask = (1u << 28) - 1;
struct RepSym { int32_t src; int32_t val; };  // src 3 = constant
__device__ __forceinline__ uint64_t pack_seq_v2(uint32_t ll, uint32_t ml, const RepSym& o) {
  const uint64_t tag = o.src == 3 ? 0ull : (uint64_t)(o.src + 1);
  return (uint64_t)ll | ((uint64_t)ml << 17) | (tag << 34) | ((uint64_t)((uint32_t)o.val & kSeqValMask) << 36);
}

// literal-length codes 0..35 then match-length codes 0..52: extra bits << 24 | base value (RFC 8878 3.1.1.3.2.1.1;
// the same numbers as ll_base / ll_bits / ml_base / ml_bits of zstd_decode.h). The kernels copy the table to shared
// memory: one LDS per code instead of two global loads with 64-bit address arithmetic on the bit-consumption chain.
constexpr int kSeqCodeTabSize = 89;
__constant__ uint32_t kSeqCodeTab[kSeqCodeTabSize] = {0x0u, 0x1u, 0x2u, 0x3u, 0x4u, 0x5u, 0x6u, 0x7u, 0x8u, 0x9u, 0xau, 0xbu, 0xcu, 0xdu, 0xeu, 0xfu, 0x1000010u, 0x1000012u, 0x1000014u, 0x1000016u, 0x2000018u, 0x200001cu, 0x3000020u, 0x3000028u, 0x4000030u, 0x6000040u, 0x7000080u, 0x8000100u, 0x9000200u, 0xa000400u, 0xb000800u, 0xc001000u, 0xd002000u, 0xe004000u, 0xf008000u, 0x10010000u, 0x3u, 0x4u, 0x5u, 0x6u, 0x7u, 0x8u, 0x9u, 0xau, 0xbu, 0xcu, 0xdu, 0xeu, 0xfu, 0x10u, 0x11u, 0x12u, 0x13u, 0x14u, 0x15u, 0x16u, 0x17u, 0x18u, 0x19u, 0x1au, 0x1bu, 0x1cu, 0x1du, 0x1eu, 0x1fu, 0x20u, 0x21u, 0x22u, 0x1000023u, 0x1000025u, 0x1000027u, 0x1000029u, 0x200002bu, 0x200002fu, 0x3000033u, 0x300003bu, 0x4000043u, 0x4000053u, 0x5000063u, 0x7000083u, 0x8000103u, 0x9000203u, 0xa000403u, 0xb000803u, 0xc001003u, 0xd002003u, 0xe004003u, 0xf008003u, 0x10010003u};

// FSE bit stream of one block -> records with offsets resolved against the block's initial repeat history.
// Returns 0 or < 0; *ml_total = sum of the match lengths; fin[3] = the final history (symbolic).
__device__ int64_t sequences_v2(const uint8_t* bits, int64_t size, const ZstdBlockPlan& P, const FseEntry* tll, const FseEntry* tml,
                                const FseEntry* tof, uint64_t* out, int64_t* ml_total, RepSym* fin, const uint32_t* codetab) {
  using namespace zstd_detail;
  BitD d;
  if (!d.init(bits, size)) return ZSTD_E_CORRUPT;
  uint32_t sl = d.read(P.ll_log), so = d.read(P.of_log), sm = d.read(P.ml_log);
  RepSym r0{0, 0}, r1{1, 0}, r2{2, 0};
  int64_t mls = 0;
  for (int32_t q = 0; q < P.nseq; ++q) {
    const FseEntry eo = tof[so], em = tml[sm], el = tll[sl];
    const int oc = eo.symbol, mc = em.symbol, lc = el.symbol;
    if (oc > 27 || mc > 52 || lc > 35) return ZSTD_E_CORRUPT;  // offsets >= 256 MiB cannot occur inside a page
    const uint32_t cm = codetab[36 + mc], cl = codetab[lc];
    // The container is refilled only when the fields of this sequence do not fit what is left of it (a refill is ~40 of
    // the ~200 instructions of a sequence; a typical sequence reads 20-30 bits, so every other one skips it), and a
    // second time in the middle only for the rare sequence of more than 57 bits.
    const uint32_t n1 = (uint32_t)oc + (cm >> 24), n2 = (cl >> 24) + el.nbits + em.nbits + eo.nbits;
    if (d.consumed + n1 + n2 > 64u) d.reload();
    const uint32_t ov = (1u << oc) + d.read(oc);                        // <= 27 bits
    const uint32_t ml = (cm & 0xffffffu) + d.read((int)(cm >> 24));     // <= 16 bits
    if (d.consumed + n2 > 64u) d.reload();
    const uint32_t ll = (cl & 0xffffffu) + d.read((int)(cl >> 24));     // <= 16 bits
    if (q + 1 < P.nseq) {                                      // <= 9 + 9 + 8 bits
      sl = el.baseline + d.read(el.nbits);
      sm = em.baseline + d.read(em.nbits);
      so = eo.baseline + d.read(eo.nbits);
    }
    if (d.overrun()) return ZSTD_E_CORRUPT;
    RepSym o;
    if (ov > 3) {
      o = RepSym{3, (int32_t)(ov - 3)};
      r2 = r1; r1 = r0; r0 = o;
    } else {
      const uint32_t idx = ov - 1 + (ll == 0 ? 1u : 0u);
      if (idx == 0) {
        o = r0;
      } else {
        o = idx == 1 ? r1 : (idx == 2 ? r2 : RepSym{r0.src, r0.val - 1});
        if (idx > 1) r2 = r1;
        r1 = r0;
        r0 = o;
      }
    }
    if (o.src != 3 && (o.val < -(1 << 20) || o.val > 0)) return ZSTD_E_UNSUPPORTED;  // a chain of 2^20 "rep0 - 1": not a page
    mls += ml;
    out[q] = pack_seq_v2(ll, ml, o);
  }
  d.reload();
  if (d.overrun() || !d.finished_exactly()) return ZSTD_E_CORRUPT;
  *ml_total = mls;
  fin[0] = r0; fin[1] = r1; fin[2] = r2;
  return 0;
}

constexpr int kZstdMaxTracked = 64;  // blocks per page with a progress counter; larger pages run on one warp

// n bytes, one lane: loads of 8 bytes are in flight before the first store (ranges do not overlap)
__device__ __forceinline__ void copy_lane_cg(uint8_t* d, const uint8_t* s, uint32_t n) {
  uint32_t i = 0;
  for (; i + 8 <= n; i += 8) {
    uint8_t t[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) t[k] = s[i + k];
#pragma unroll
    for (int k = 0; k < 8; ++k) d[i + k] = t[k];
  }
  const uint32_t r = n - i;
  uint8_t t[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) if ((uint32_t)k < r) t[k] = s[i + k];
#pragma unroll
  for (int k = 0; k < 8; ++k) if ((uint32_t)k < r) d[i + k] = t[k];
}
// n bytes, the whole warp, 4 x 32 bytes in flight
__device__ __forceinline__ void copy_warp_cg(uint8_t* d, const uint8_t* s, uint32_t n, int lane) {
  uint32_t i = lane;
  for (; i + 96 < n; i += 128) {
    const uint8_t a = s[i], b = s[i + 32], c = s[i + 64], e = s[i + 96];
    d[i] = a; d[i + 32] = b; d[i + 64] = c; d[i + 96] = e;
  }
  for (; i < n; i += 32) d[i] = s[i];
}

// Matches of one round (<= 32 sequences, one per lane) in dependency LEVELS instead of one at a time: a match whose
// source overlaps the output of an earlier match of the same round (columnar data repeats values a few bytes back: most
// matches do) must wait for it -- but only for it. level = 1 + the highest level among the earlier matches it reads
// from (one pass over the lanes: lane i's level is final when it is broadcast); every level is then copied by all its
// lanes at once. A round of 32 sequences costs (levels) memory round trips instead of (dependent matches) round
// trips (r3: the three ZSTD launches of a config-5 bucket are 9-13 ms chains of such round trips).
__device__ __forceinline__ void zstd_copy_matches(uint8_t* dst, int64_t op, int64_t my_match, uint32_t offset, uint32_t ml, int cnt, int lane) {
  const bool has = lane < cnt && ml > 0;
  const int64_t src0 = my_match - (int64_t)offset;
  const int64_t src1 = src0 + (int64_t)(ml < offset ? ml : offset);  // an overlapping match reads [src0, my_match) only
  int level = 0;
  if (__any_sync(0xffffffffu, has && src1 > op)) {  // some source reaches into this round
    for (int i = 0; i + 1 < cnt; ++i) {
      const int64_t o0 = __shfl_sync(0xffffffffu, my_match, i);
      const uint32_t oml = __shfl_sync(0xffffffffu, ml, i);
      const int li = __shfl_sync(0xffffffffu, level, i);
      if (has && lane > i && oml > 0 && src0 < o0 + (int64_t)oml && src1 > o0) level = max(level, li + 1);
    }
  }
  const int maxl = __reduce_max_sync(0xffffffffu, has ? level : 0);
  for (int L = 0; L <= maxl; ++L) {
    const bool mine = has && level == L;
    const bool small = mine && ml <= 64 && offset >= ml;
    if (small) copy_lane_cg(dst + my_match, dst + src0, ml);
    unsigned big = __ballot_sync(0xffffffffu, mine && !small);
    while (big) {
      const int j = __ffs(big) - 1;
      big &= big - 1;
      const int64_t md = __shfl_sync(0xffffffffu, my_match, j);
      const uint32_t off = __shfl_sync(0xffffffffu, offset, j), n = __shfl_sync(0xffffffffu, ml, j);
      const uint8_t* m = dst + md - off;
      if (off >= n) {
        copy_warp_cg(dst + md, m, n, lane);
      } else {  // overlapping match: only bytes that existed before this sequence are read
        for (uint32_t i = lane; i < n; i += 32) dst[md + i] = m[i % off];
      }
    }
    __syncwarp();
  }
}

}  // namespace

template <int kWarps>
__global__ void __launch_bounds__(kWarps * 32) zstd_pages_kernel_v2(const PqZstdPage* pages, const int32_t* order,
                                                                    const ZstdBlockDesc* blocks_all, ZstdBlockPlan* plans_all,
                                                                    uint8_t* lit_all, uint64_t* seq_all, int32_t* err) {
  constexpr int kThreads = kWarps * 32;
  constexpr int kGroup = kWarps;  // blocks whose tables are resident at a time
  extern __shared__ __align__(16) unsigned char zsm_raw[];
  ZstdTables* s_tabs = reinterpret_cast<ZstdTables*>(zsm_raw);  // kGroup slots
  __shared__ ZstdPlanScratch s_plan;
  __shared__ ZstdBlockPlan s_cur_plan, s_prev_plan;
  __shared__ int s_fail, s_have_prev, s_prev_slot;
  __shared__ uint32_t s_codetab[kSeqCodeTabSize];
  for (int i = threadIdx.x; i < kSeqCodeTabSize; i += kWarps * 32) s_codetab[i] = kSeqCodeTab[i];  // read after the first barrier
  __shared__ int s_done[kZstdMaxTracked];  // absolute output position every block has completed
  __shared__ int s_abort;                  // a block met corrupt data: nobody waits for anybody any more
  const PqZstdPage pg = pages[order[blockIdx.x]];
  const uint8_t* src = reinterpret_cast<const uint8_t*>(pg.src_off);
  uint8_t* dst = reinterpret_cast<uint8_t*>(pg.dst_off);
  const ZstdBlockDesc* blocks = blocks_all + pg.first_block;
  ZstdBlockPlan* plans = plans_all + pg.first_block;
  uint8_t* lit = lit_all + pg.lit_off;
  uint64_t* seq = seq_all + pg.seq_off;
  const int nb = pg.nblocks;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) { s_fail = 0; s_have_prev = 0; s_prev_slot = 0; s_abort = 0; }
  __syncthreads();
  int64_t lit_total = 0, seq_total = 0;  // meaningful in thread 0
  const bool prof = pg.raw_size >= (64 << 10);
  long long c_plan = 0, c_entropy = 0;
  for (int g0 = 0; g0 < nb; g0 += kGroup) {
    const int g1 = min(nb, g0 + kGroup);
    const long long tp0 = clock64();
    // ---- plan (thread 0, in order: tables may be inherited from the previous compressed block) ----
    if (tid == 0) {
      for (int k = g0; k < g1 && !s_fail; ++k) {
        if (blocks[k].type != 2) continue;
        const int slot = k - g0;
        const int64_t r = zstd_plan_block(src + blocks[k].src_off, blocks[k].size, s_cur_plan, s_tabs[slot],
                                          s_have_prev ? &s_prev_plan : nullptr, s_have_prev ? &s_tabs[s_prev_slot] : nullptr, s_plan);
        if (r < 0) { s_fail = 1; break; }
        s_cur_plan.lit_base = (int32_t)lit_total;
        s_cur_plan.seq_base = (int32_t)seq_total;
        lit_total += s_cur_plan.regen;
        seq_total += s_cur_plan.nseq;
        if (lit_total > pg.raw_size || seq_total > pg.raw_size / 3 + 1) { s_fail = 1; break; }
        s_prev_plan = s_cur_plan;
        s_have_prev = 1;
        s_prev_slot = slot;
        plans[k] = s_cur_plan;
      }
    }
    __syncthreads();
    if (s_fail) break;
    const long long tp1 = clock64();
    c_plan += tp1 - tp0;
    // ---- entropy, all blocks of the group side by side. Roles are split by WARP so that lanes of a warp run the
    // same loop: warp 0 = the Huffman streams (lane = block * 4 + stream), warp 1 = the FSE sequence streams and the
    // RLE literal fills (lane = block). (Mixed roles inside one warp serialise: r2 profile, entropy = sum of both.)
    if (warp == 0 && lane < 4 * (g1 - g0)) {
      const int k = g0 + (lane >> 2), role = lane & 3;
      if (blocks[k].type == 2) {
        const ZstdBlockPlan& P = plans[k];
        if (P.lit_type == 2 && role < P.streams) {
          const uint8_t* b = src + blocks[k].src_off;
          uint8_t* L = lit + P.lit_base;
          const int64_t per = P.streams == 4 ? (P.regen + 3) / 4 : P.regen;
          int64_t so = 0;
          for (int st = 0; st < role; ++st) so += P.stream_size[st];
          const int64_t n = P.streams == 4 ? (role < 3 ? per : P.regen - 3 * per) : P.regen;
          const long long th0 = clock64();
          if (huf_stream_v2(s_tabs[k - g0].huf, P.huf_log, b + P.lit_src + so, P.stream_size[role], L + role * per, n) < 0) s_fail = 1;
          if (prof) {
            const unsigned long long c = (unsigned long long)(clock64() - th0);
            if (atomicMax(&g_zstd_prof[10], c) < c) g_zstd_prof[11] = (unsigned long long)n;
          }
        }
      }
    } else if (warp == 1 && lane < (g1 - g0)) {
      const int k = g0 + lane;
      if (blocks[k].type == 2) {
        const ZstdBlockPlan P = plans[k];
        const ZstdTables& T = s_tabs[k - g0];
        const uint8_t* b = src + blocks[k].src_off;
        int64_t mls = 0;
        RepSym fin[3] = {{0, 0}, {1, 0}, {2, 0}};
        int64_t r = 0;
        const long long ts0 = clock64();
        if (P.nseq > 0) r = sequences_v2(b + P.seq_src, P.seq_size, P, T.ll, T.ml, T.of, seq + P.seq_base, &mls, fin, s_codetab);
        if (prof) {
          const unsigned long long c = (unsigned long long)(clock64() - ts0);
          if (atomicMax(&g_zstd_prof[8], c) < c) g_zstd_prof[9] = (unsigned long long)P.nseq;
        }
        if (r < 0) s_fail = 1;
        plans[k].out_size = (int32_t)min((int64_t)INT32_MAX, (int64_t)P.regen + mls);
        for (int j = 0; j < 3; ++j) { plans[k].rep_src[j] = fin[j].src; plans[k].rep_val[j] = fin[j].val; }
        if (P.lit_type == 1) {  // RLE literals (raw literals are read in place by the execute phase)
          uint8_t* L = lit + P.lit_base;
          const uint8_t v = b[P.lit_src];
          for (int i = 0; i < P.regen; ++i) L[i] = v;
        }
      }
    }
    __syncthreads();
    if (s_fail) break;
    c_entropy += clock64() - tp1;
    // tables of the group's last compressed block must survive in a slot the next group's first plan does not
    // overwrite before reading: move them to the last slot and plan the next group from slot 0 upwards
    if (g1 < nb && s_have_prev && s_prev_slot != kGroup - 1) {
      const uint32_t* from = reinterpret_cast<const uint32_t*>(&s_tabs[s_prev_slot]);
      uint32_t* to = reinterpret_cast<uint32_t*>(&s_tabs[kGroup - 1]);
      for (int i = tid; i < (int)(sizeof(ZstdTables) / 4); i += kThreads) to[i] = from[i];
      __syncthreads();
      if (tid == 0) s_prev_slot = kGroup - 1;
      __syncthreads();
    }
  }
  __syncthreads();
  if (s_fail) {
    if (tid == 0) atomicExch(err, kErrBadPage);
    return;
  }
  // ---- chain: output bases and repeat histories of the blocks (thread 0) ----
  if (tid == 0) {
    int64_t op = 0;
    uint32_t rep[3] = {1, 4, 8};
    for (int k = 0; k < nb; ++k) {
      if (blocks[k].type != 2) {
        op += blocks[k].out_size;
        if (k < kZstdMaxTracked) s_done[k] = 0;
        continue;
      }
      ZstdBlockPlan& P = plans[k];
      P.out_base = (int32_t)op;
      for (int j = 0; j < 3; ++j) P.rep_init[j] = rep[j];
      uint32_t nx[3];
      for (int j = 0; j < 3; ++j) nx[j] = P.rep_src[j] == 3 ? (uint32_t)P.rep_val[j] : rep[P.rep_src[j]] + (uint32_t)P.rep_val[j];
      for (int j = 0; j < 3; ++j) rep[j] = nx[j];
      op += P.out_size;
      if (k < kZstdMaxTracked) s_done[k] = 0;
      if (op > pg.raw_size) { s_fail = 1; break; }
    }
    if (op != pg.raw_size) s_fail = 1;
  }
  __syncthreads();
  if (s_fail) {
    if (tid == 0) atomicExch(err, kErrBadPage);
    return;
  }
  // ---- execute: one warp per block (all blocks on warp 0 when the page has more blocks than counters) ----
  const bool tracked = nb <= kZstdMaxTracked;
  if (!tracked && warp != 0) return;
  const long long t_exec0 = clock64();
  long long c_wait = 0;
  if (prof && tid == 0) {
    atomicAdd(&g_zstd_prof[0], (unsigned long long)c_plan);
    atomicAdd(&g_zstd_prof[1], (unsigned long long)c_entropy);
    atomicAdd(&g_zstd_prof[3], 1ull);
    atomicMax(&g_zstd_prof[4], (unsigned long long)c_plan);
    atomicMax(&g_zstd_prof[5], (unsigned long long)c_entropy);
  }
  const int64_t cap = pg.raw_size;
  volatile int* done = s_done;
  volatile int* abort_flag = &s_abort;
  int64_t base_k = 0;  // output base of block k (raw / RLE blocks carry no plan: recomputed while walking)
  for (int k = 0; k < nb; ++k) {
    const int64_t out_k = blocks[k].type == 2 ? plans[k].out_size : blocks[k].out_size;
    const int64_t my_base = base_k;
    base_k += out_k;
    if (tracked && (k % kWarps) != warp) continue;
    const uint8_t* b = src + blocks[k].src_off;
    if (blocks[k].type != 2) {
      if (blocks[k].type == 0) for (int64_t i = lane; i < out_k; i += 32) dst[my_base + i] = b[i];
      else { const uint8_t v = b[0]; for (int64_t i = lane; i < out_k; i += 32) dst[my_base + i] = v; }
      __syncwarp();
      __threadfence_block();
      if (lane == 0 && tracked) done[k] = (int)(my_base + out_k);
      continue;
    }
    const ZstdBlockPlan P = plans[k];
    const uint8_t* lits = P.lit_type == 0 ? b + P.lit_src : lit + P.lit_base;  // raw literals stay where they are
    const uint64_t* sq = seq + P.seq_base;
    const uint32_t rep_i0 = P.rep_init[0], rep_i1 = P.rep_init[1], rep_i2 = P.rep_init[2];
    int64_t op = my_base, lpos = 0;
    uint64_t v_next = (lane < P.nseq) ? sq[lane] : 0ull;
    for (int q0 = 0; q0 < P.nseq; q0 += 32) {
      const int cnt = min(32, P.nseq - q0);
      const uint64_t v = v_next;
      v_next = (q0 + 32 + lane < P.nseq) ? sq[q0 + 32 + lane] : 0ull;  // the next round's records travel meanwhile
      const uint32_t ll = (uint32_t)(v & 0x1ffff), ml = (uint32_t)((v >> 17) & 0x1ffff);
      const uint32_t tag = (uint32_t)(v >> 34) & 3u;
      const int32_t val = (int32_t)((uint32_t)(v >> 36) << 4) >> 4;  // sign-extend 28 bits
      const uint32_t rep_t = tag == 1 ? rep_i0 : (tag == 2 ? rep_i1 : rep_i2);
      const uint32_t offset = lane < cnt ? (tag == 0 ? (uint32_t)val : rep_t + (uint32_t)val) : 1u;
      uint32_t io = ll + ml, il = ll;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t yo = __shfl_up_sync(0xffffffffu, io, d), yl = __shfl_up_sync(0xffffffffu, il, d);
        if (lane >= d) { io += yo; il += yl; }
      }
      const uint32_t tot_out = __shfl_sync(0xffffffffu, io, 31), tot_lit = __shfl_sync(0xffffffffu, il, 31);
      const int64_t my_out = op + (io - ll - ml);
      const int64_t my_lit = lpos + (il - ll);
      const int64_t my_match = my_out + ll;
      const bool bad = lpos + tot_lit > P.regen || op + tot_out > my_base + out_k ||
                       (lane < cnt && (offset == 0 || (int64_t)offset > my_match));
      if (__any_sync(0xffffffffu, bad)) {
        if (lane == 0) { atomicExch(err, kErrBadPage); *abort_flag = 1; }
        return;
      }
      // matches that reach below this block wait for the blocks that own those bytes
      if (tracked && k > 0) {
        const int64_t s0 = (lane < cnt && ml > 0) ? my_match - (int64_t)offset : INT64_MAX;
        int64_t lo = s0;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, d));
        if (lo < my_base) {
          const long long tw0 = clock64();
          int64_t bb = my_base;  // walk down the blocks that end above `lo`
          for (int kk = k - 1; kk >= 0 && bb > lo; --kk) {
            const int64_t need = bb;  // the whole block below us up to its end (conservative, rarely more than one)
            while (done[kk] < (int)need && !*abort_flag) __nanosleep(64);
            bb -= blocks[kk].type == 2 ? plans[kk].out_size : blocks[kk].out_size;
          }
          if (*abort_flag) return;
          __threadfence_block();
          c_wait += clock64() - tw0;
        }
      }
      if (ll <= 64) copy_lane_cg(dst + my_out, lits + my_lit, ll);
      unsigned big = __ballot_sync(0xffffffffu, ll > 64);
      while (big) {
        const int j = __ffs(big) - 1;
        big &= big - 1;
        const int64_t o = __shfl_sync(0xffffffffu, my_out, j), l0 = __shfl_sync(0xffffffffu, my_lit, j);
        const uint32_t n = __shfl_sync(0xffffffffu, ll, j);
        copy_warp_cg(dst + o, lits + l0, n, lane);
      }
      __syncwarp();
      zstd_copy_matches(dst, op, my_match, offset, ml, cnt, lane);
      op += tot_out;
      lpos += tot_lit;
      __threadfence_block();
      if (lane == 0 && tracked) done[k] = (int)op;
    }
    const int64_t rest = P.regen - lpos;
    if (rest < 0 || op + rest != my_base + out_k) {
      if (lane == 0) { atomicExch(err, kErrBadPage); *abort_flag = 1; }
      return;
    }
    copy_warp_cg(dst + op, lits + lpos, (uint32_t)rest, lane);
    op += rest;
    __syncwarp();
    __threadfence_block();
    if (lane == 0 && tracked) done[k] = (int)op;
  }
  (void)cap;
  if (prof && lane == 0) {
    const unsigned long long c = (unsigned long long)(clock64() - t_exec0);
    atomicAdd(&g_zstd_prof[2], c);
    atomicMax(&g_zstd_prof[6], c);
    atomicMax(&g_zstd_prof[7], (unsigned long long)c_wait);
  }
}

// ---------------------------------------------------------------------------------------------------
// One-block pages, eight to a thread block. The delta files of a hash bucket are small (config 5: 63 files of
// 1 162 rows -> a thousand pages of a few KiB, one ZSTD block each); given a thread block of their own they keep
// one lane of one warp busy. Here warp w plans and executes page w of the batch, and the two entropy roles run
// across the batch exactly as they run across the blocks of a large page: warp 0 = the Huffman streams of all eight
// pages (lane = page * 4 + stream), warp 1 = their eight FSE sequence streams (8 x fewer warp instructions there).
// ---------------------------------------------------------------------------------------------------
constexpr int kZstdBatch = 8;

// the sequences of one block -> output (a whole page here: no history before it, no other block to wait for)
__device__ bool zstd_execute_single_block(const ZstdBlockPlan& P, const uint8_t* lits, const uint64_t* sq, uint8_t* dst,
                                          int64_t out_k, int lane) {
  const uint32_t rep_i0 = 1, rep_i1 = 4, rep_i2 = 8;
  int64_t op = 0, lpos = 0;
  uint64_t v_next = (lane < P.nseq) ? sq[lane] : 0ull;
  for (int q0 = 0; q0 < P.nseq; q0 += 32) {
    const int cnt = min(32, P.nseq - q0);
    const uint64_t v = v_next;
    v_next = (q0 + 32 + lane < P.nseq) ? sq[q0 + 32 + lane] : 0ull;
    const uint32_t ll = (uint32_t)(v & 0x1ffff), ml = (uint32_t)((v >> 17) & 0x1ffff);
    const uint32_t tag = (uint32_t)(v >> 34) & 3u;
    const int32_t val = (int32_t)((uint32_t)(v >> 36) << 4) >> 4;
    const uint32_t rep_t = tag == 1 ? rep_i0 : (tag == 2 ? rep_i1 : rep_i2);
    const uint32_t offset = lane < cnt ? (tag == 0 ? (uint32_t)val : rep_t + (uint32_t)val) : 1u;
    uint32_t io = ll + ml, il = ll;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t yo = __shfl_up_sync(0xffffffffu, io, d), yl = __shfl_up_sync(0xffffffffu, il, d);
      if (lane >= d) { io += yo; il += yl; }
    }
    const uint32_t tot_out = __shfl_sync(0xffffffffu, io, 31), tot_lit = __shfl_sync(0xffffffffu, il, 31);
    const int64_t my_out = op + (io - ll - ml);
    const int64_t my_lit = lpos + (il - ll);
    const int64_t my_match = my_out + ll;
    const bool bad = lpos + tot_lit > P.regen || op + tot_out > out_k ||
                     (lane < cnt && (offset == 0 || (int64_t)offset > my_match));
    if (__any_sync(0xffffffffu, bad)) return false;
    if (ll <= 64) copy_lane_cg(dst + my_out, lits + my_lit, ll);
    unsigned big = __ballot_sync(0xffffffffu, ll > 64);
    while (big) {
      const int j = __ffs(big) - 1;
      big &= big - 1;
      const int64_t o = __shfl_sync(0xffffffffu, my_out, j), l0 = __shfl_sync(0xffffffffu, my_lit, j);
      const uint32_t n = __shfl_sync(0xffffffffu, ll, j);
      copy_warp_cg(dst + o, lits + l0, n, lane);
    }
    __syncwarp();
    zstd_copy_matches(dst, op, my_match, offset, ml, cnt, lane);
    op += tot_out;
    lpos += tot_lit;
  }
  const int64_t rest = P.regen - lpos;
  if (rest < 0 || op + rest != out_k) return false;
  copy_warp_cg(dst + op, lits + lpos, (uint32_t)rest, lane);
  return true;
}

__global__ void __launch_bounds__(kZstdBatch * 32) zstd_batch_kernel(const PqZstdPage* pages, const int32_t* order, int npages,
                                                                     const ZstdBlockDesc* blocks_all, ZstdBlockPlan* plans_all,
                                                                     uint8_t* lit_all, uint64_t* seq_all, int32_t* err, int prof) {
  extern __shared__ __align__(16) unsigned char zsm_raw[];
  ZstdTables* s_tabs = reinterpret_cast<ZstdTables*>(zsm_raw);  // one slot per page of the batch
  const long long t_start = prof ? clock64() : 0;
  __shared__ ZstdPlanScratch s_scr[kZstdBatch];
  __shared__ ZstdBlockPlan s_plans[kZstdBatch];
  __shared__ int s_state[kZstdBatch];  // 0 = no page, 1 = compressed block planned, 2 = raw / RLE block, -1 = corrupt
  __shared__ uint32_t s_codetab[kSeqCodeTabSize];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < kSeqCodeTabSize; i += kZstdBatch * 32) s_codetab[i] = kSeqCodeTab[i];  // read after the plan barrier
  const int first = blockIdx.x * kZstdBatch;
  auto page_of = [&](int u) -> const PqZstdPage& { return pages[order[first + u]]; };
  // ---- plan: ONE warp, lane u plans page u (section headers + entropy tables into the page's shared-memory slot). The
  // plans are single-lane chains of a few thousand instructions: eight lanes of one warp issue an eighth of the warp
  // instructions eight warps with one active lane would (measured: no end-to-end difference either way -- 2.5 ms per
  // config-5 bucket at 24 task slots: the decompression is bound by its serial chains, not by issue slots) ----
  if (warp == 0 && lane < kZstdBatch) {
    const int u = lane;
    int st = 0;
    if (first + u < npages) {
      const PqZstdPage& pg = page_of(u);
      const ZstdBlockDesc& blk = blocks_all[pg.first_block];
      const uint8_t* src = reinterpret_cast<const uint8_t*>(pg.src_off);
      if (blk.type != 2) {
        st = blk.out_size == pg.raw_size ? 2 : -1;
      } else {
        ZstdBlockPlan& P = s_plans[u];
        const int64_t r = zstd_plan_block(src + blk.src_off, blk.size, P, s_tabs[u], nullptr, nullptr, s_scr[u]);
        P.lit_base = 0;
        P.seq_base = 0;
        st = (r < 0 || P.regen > pg.raw_size || P.nseq > pg.raw_size / 3 + 1) ? -1 : 1;
      }
    }
    s_state[u] = st;
  }
  __syncthreads();
  const long long t_plan = prof ? clock64() : 0;
  // ---- entropy across the batch ----
  if (warp == 0) {
    const int u = lane >> 2, role = lane & 3;
    if (s_state[u] == 1) {
      const ZstdBlockPlan& P = s_plans[u];
      if (P.lit_type == 2 && role < P.streams) {
        const PqZstdPage& pg = page_of(u);
        const uint8_t* b = reinterpret_cast<const uint8_t*>(pg.src_off) + blocks_all[pg.first_block].src_off;
        uint8_t* L = lit_all + pg.lit_off;
        const int64_t per = P.streams == 4 ? (P.regen + 3) / 4 : P.regen;
        int64_t so = 0;
        for (int st = 0; st < role; ++st) so += P.stream_size[st];
        const int64_t n = P.streams == 4 ? (role < 3 ? per : P.regen - 3 * per) : P.regen;
        if (huf_stream_v2(s_tabs[u].huf, P.huf_log, b + P.lit_src + so, P.stream_size[role], L + role * per, n) < 0) s_state[u] = -1;
      }
    }
  } else if (warp == 1 && lane < kZstdBatch) {
    const int u = lane;
    if (s_state[u] == 1) {
      ZstdBlockPlan& P = s_plans[u];
      const PqZstdPage& pg = page_of(u);
      const uint8_t* b = reinterpret_cast<const uint8_t*>(pg.src_off) + blocks_all[pg.first_block].src_off;
      int64_t mls = 0;
      RepSym fin[3] = {{0, 0}, {1, 0}, {2, 0}};
      int64_t r = 0;
      if (P.nseq > 0) r = sequences_v2(b + P.seq_src, P.seq_size, P, s_tabs[u].ll, s_tabs[u].ml, s_tabs[u].of, seq_all + pg.seq_off, &mls, fin, s_codetab);
      P.out_size = (int32_t)min((int64_t)INT32_MAX, (int64_t)P.regen + mls);
      if (r < 0 || P.out_size != pg.raw_size) s_state[u] = -1;
      if (P.lit_type == 1) {
        uint8_t* L = lit_all + pg.lit_off;
        const uint8_t v = b[P.lit_src];
        for (int i = 0; i < P.regen; ++i) L[i] = v;
      }
    }
  }
  __syncthreads();
  const long long t_entropy = prof ? clock64() : 0;
  if (prof && threadIdx.x == 32) {  // warp 1: how long did the FSE lanes take vs the barrier (Huffman warp)
  }
  // ---- execute: warp w, page w ----
  const int st = s_state[warp];
  if (st == 0) return;
  bool ok = st > 0;
  if (ok) {
    const PqZstdPage& pg = page_of(warp);
    const ZstdBlockDesc& blk = blocks_all[pg.first_block];
    const uint8_t* b = reinterpret_cast<const uint8_t*>(pg.src_off) + blk.src_off;
    uint8_t* dst = reinterpret_cast<uint8_t*>(pg.dst_off);
    if (st == 2) {
      if (blk.type == 0) for (int64_t i = lane; i < blk.out_size; i += 32) dst[i] = b[i];
      else { const uint8_t v = b[0]; for (int64_t i = lane; i < blk.out_size; i += 32) dst[i] = v; }
    } else {
      const ZstdBlockPlan& P = s_plans[warp];
      const uint8_t* lits = P.lit_type == 0 ? b + P.lit_src : lit_all + pg.lit_off;
      ok = zstd_execute_single_block(P, lits, seq_all + pg.seq_off, dst, pg.raw_size, lane);
    }
  }
  if (!ok && lane == 0) atomicExch(err, kErrBadPage);
  (void)plans_all;
  if (prof && lane == 0) {
    const long long t_end = clock64();
    if (warp == 0) {
      atomicAdd(&g_zstd_prof[12], (unsigned long long)(t_plan - t_start));
      atomicAdd(&g_zstd_prof[13], (unsigned long long)(t_entropy - t_plan));
      atomicAdd(&g_zstd_prof[15], 1ull);
    }
    atomicMax(&g_zstd_prof[14], (unsigned long long)(t_end - t_entropy));
  }
}

void launch_zstd_batch(const PqZstdPage* pages, const int32_t* order_one, int none, const ZstdBlockDesc* blocks, ZstdBlockPlan* plans,
                       uint8_t* lit, uint64_t* seq, int32_t* err, cudaStream_t s) {
  if (none <= 0) return;
  static thread_local int attr_device = -1;
  int cur = 0;
  cudaGetDevice(&cur);
  if (attr_device != cur) {
    attr_device = cur;
    cudaFuncSetAttribute(zstd_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kZstdBatch * sizeof(ZstdTables)));
  }
  static const bool prof = [] { const char* v = getenv("LAKESOUL_B200_ZSTD_PROFILE"); return v && *v == '1'; }();
  zstd_batch_kernel<<<(none + kZstdBatch - 1) / kZstdBatch, kZstdBatch * 32, kZstdBatch * sizeof(ZstdTables), s>>>(
      pages, order_one, none, blocks, plans, lit, seq, err, prof ? 1 : 0);
  if (prof) {
    unsigned long long h[16] = {0};
    cudaStreamSynchronize(s);
    cudaMemcpyFromSymbol(h, g_zstd_prof, sizeof(h));
    if (h[15]) fprintf(stderr, "ZSTDPROF batch: %llu thread blocks, mean kcycles plan %.0f entropy %.0f; slowest execute %.0f\n", h[15],
                       h[12] / 1e3 / h[15], h[13] / 1e3 / h[15], h[14] / 1e3);
    unsigned long long z[16] = {0};
    cudaMemcpyToSymbol(g_zstd_prof, z, sizeof(z));
  }
}

void launch_zstd_pages_v2(const PqZstdPage* pages, const int32_t* order_small, int nsmall, const int32_t* order_big, int nbig,
                          const ZstdBlockDesc* blocks, ZstdBlockPlan* plans, ZstdTables* tables_v1, uint8_t* lit, uint64_t* seq, int32_t* err,
                          cudaStream_t s) {
  static thread_local int attr_device = -1;
  int cur = 0;
  cudaGetDevice(&cur);
  if (attr_device != cur) {
    attr_device = cur;
    cudaFuncSetAttribute(zstd_pages_kernel_v2<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(8 * sizeof(ZstdTables)));
    cudaFuncSetAttribute(zstd_pages_kernel_v2<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * sizeof(ZstdTables)));
  }
  if (nbig > 0)
    zstd_pages_kernel_v2<8><<<nbig, 8 * 32, 8 * sizeof(ZstdTables), s>>>(pages, order_big, blocks, plans, lit, seq, err);
  static const bool prof = [] { const char* v = getenv("LAKESOUL_B200_ZSTD_PROFILE"); return v && *v == '1'; }();
  auto dump = [&](const char* what) {
    unsigned long long h[16] = {0}, z[16] = {0};
    cudaStreamSynchronize(s);
    cudaMemcpyFromSymbol(h, g_zstd_prof, sizeof(h));
    cudaMemcpyToSymbol(g_zstd_prof, z, sizeof(z));
    if (h[3]) fprintf(stderr, "ZSTDPROF2b %s slowest sequence stream %.0f kcycles for %llu sequences; slowest huffman stream %.0f kcycles for %llu symbols\n",
                      what, h[8] / 1e3, h[9], h[10] / 1e3, h[11]);
    if (h[3]) fprintf(stderr, "ZSTDPROF2 %s pages>=64K %llu  mean kcycles plan %.0f entropy %.0f execute(per warp) %.0f  max plan %.0f entropy %.0f execute %.0f (max waiting %.0f)\n",
                      what, h[3], h[0] / 1e3 / h[3], h[1] / 1e3 / h[3], h[2] / 1e3 / h[3], h[4] / 1e3, h[5] / 1e3, h[6] / 1e3, h[7] / 1e3);
  };
  if (prof) dump("big");
  if (nsmall > 0 && tables_v1)  // experiment: pages of one or two blocks on the first-generation kernel
    zstd_pages_kernel<<<nsmall, kZstdThreads, 0, s>>>(pages, blocks, plans, tables_v1, lit, seq, err, order_small);
  else if (nsmall > 0)
    zstd_pages_kernel_v2<2><<<nsmall, 2 * 32, 2 * sizeof(ZstdTables), s>>>(pages, order_small, blocks, plans, lit, seq, err);
  if (prof) dump("small");
}

// ---------------------------------------------------------------------------------------------------
// SNAPPY page payloads (pyarrow / Spark default files): one warp per payload. The element stream (tag byte, length,
// literal bytes or copy offset) can only be walked in order, so lane 0 frames 32 elements at a time into shared
// memory; the warp then places them like the ZSTD execute phase does: output positions by a scan, literal runs in
// parallel straight from the compressed stream, copies whose source lies before the round in parallel, the rest in
// order (overlapping copies replicate their period). Format: github.com/google/snappy format_description.txt.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) snappy_pages_kernel(const PqZstdPage* pages, int32_t* err) {
  __shared__ uint32_t s_len[32], s_off[32], s_src[32];
  const PqZstdPage pg = pages[blockIdx.x];
  const uint8_t* src = reinterpret_cast<const uint8_t*>(pg.src_off);
  uint8_t* dst = reinterpret_cast<uint8_t*>(pg.dst_off);
  const int64_t n = pg.comp_size, cap = pg.raw_size;
  const int lane = threadIdx.x;
  // preamble: uncompressed length, base-128 varint
  int64_t ip = 0;
  {
    uint64_t v = 0;
    int sh = 0;
    bool ok = false;
    for (int i = 0; i < 5 && ip < n; ++i, sh += 7) {
      const uint8_t b = src[ip++];
      v |= (uint64_t)(b & 0x7f) << sh;
      if (!(b & 0x80)) { ok = true; break; }
    }
    if (!ok || (int64_t)v != cap) {
      if (lane == 0) atomicExch(err, kErrBadPage);
      return;
    }
  }
  int64_t op = 0;
  while (ip < n) {
    int cnt = 0;
    int bad = 0;
    if (lane == 0) {
      int64_t p = ip;
      while (cnt < 32 && p < n) {
        const uint32_t tag = src[p];
        uint32_t len, off = 0, lsrc = 0;
        if ((tag & 3) == 0) {
          len = (tag >> 2) + 1;
          int extra = 0;
          if (len > 60) {
            extra = (int)len - 60;
            if (p + 1 + extra > n) { bad = 1; break; }
            uint32_t v = 0;
            for (int i = 0; i < extra; ++i) v |= (uint32_t)src[p + 1 + i] << (8 * i);
            len = v + 1;
          }
          p += 1 + extra;
          if (len > (uint32_t)(n - p)) { bad = 1; break; }
          lsrc = (uint32_t)p;
          p += len;
        } else if ((tag & 3) == 1) {
          if (p + 2 > n) { bad = 1; break; }
          len = ((tag >> 2) & 7) + 4;
          off = ((tag >> 5) << 8) | src[p + 1];
          p += 2;
          if (off == 0) { bad = 1; break; }
        } else if ((tag & 3) == 2) {
          if (p + 3 > n) { bad = 1; break; }
          len = (tag >> 2) + 1;
          off = (uint32_t)src[p + 1] | ((uint32_t)src[p + 2] << 8);
          p += 3;
          if (off == 0) { bad = 1; break; }
        } else {
          if (p + 5 > n) { bad = 1; break; }
          len = (tag >> 2) + 1;
          off = (uint32_t)src[p + 1] | ((uint32_t)src[p + 2] << 8) | ((uint32_t)src[p + 3] << 16) | ((uint32_t)src[p + 4] << 24);
          p += 5;
          if (off == 0) { bad = 1; break; }
        }
        s_len[cnt] = len;
        s_off[cnt] = off;
        s_src[cnt] = lsrc;
        ++cnt;
      }
      ip = p;
    }
    cnt = __shfl_sync(0xffffffffu, cnt, 0);
    bad = __shfl_sync(0xffffffffu, bad, 0);
    ip = __shfl_sync(0xffffffffu, ip, 0);
    __syncwarp();
    const uint32_t len = lane < cnt ? s_len[lane] : 0u;
    const uint32_t off = lane < cnt ? s_off[lane] : 0u;
    const uint32_t lsrc = lane < cnt ? s_src[lane] : 0u;
    __syncwarp();
    uint32_t incl = len;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += y;
    }
    const uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
    const int64_t my = op + (incl - len);
    if (bad || op + tot > cap || __any_sync(0xffffffffu, lane < cnt && off != 0 && (int64_t)off > my)) {
      if (lane == 0) atomicExch(err, kErrBadPage);
      return;
    }
    // literals: independent of everything
    const bool lit = lane < cnt && off == 0;
    if (lit && len <= 64) copy_lane_cg(dst + my, src + lsrc, len);
    unsigned big = __ballot_sync(0xffffffffu, lit && len > 64);
    while (big) {
      const int j = __ffs(big) - 1;
      big &= big - 1;
      copy_warp_cg(dst + __shfl_sync(0xffffffffu, my, j), src + __shfl_sync(0xffffffffu, lsrc, j), __shfl_sync(0xffffffffu, len, j), lane);
    }
    __syncwarp();
    // copies whose whole source was produced before this round
    const bool cp = lane < cnt && off != 0;
    const bool indep = cp && len <= 64 && (my - (int64_t)off + len <= op);
    if (indep) copy_lane_cg(dst + my, dst + my - off, len);
    __syncwarp();
    unsigned serial = __ballot_sync(0xffffffffu, cp && !indep);
    while (serial) {
      const int j = __ffs(serial) - 1;
      serial &= serial - 1;
      const int64_t md = __shfl_sync(0xffffffffu, my, j);
      const uint32_t o = __shfl_sync(0xffffffffu, off, j), l = __shfl_sync(0xffffffffu, len, j);
      const uint8_t* m = dst + md - o;
      if (o >= l) copy_warp_cg(dst + md, m, l, lane);
      else for (uint32_t i = lane; i < l; i += 32) dst[md + i] = m[i % o];
      __syncwarp();
    }
    op += tot;
  }
  if (op != cap && lane == 0) atomicExch(err, kErrBadPage);
}

void launch_snappy_pages(const PqZstdPage* pages, int npages, int32_t* err, cudaStream_t s) {
  if (npages > 0) snappy_pages_kernel<<<npages, 32, 0, s>>>(pages, err);
}

void launch_zstd_pages(const PqZstdPage* pages, int npages, const ZstdBlockDesc* blocks, ZstdBlockPlan* plans, ZstdTables* tables,
                       uint8_t* lit, uint64_t* seq, int32_t* err, cudaStream_t s) {
  if (npages > 0) zstd_pages_kernel<<<npages, kZstdThreads, 0, s>>>(pages, blocks, plans, tables, lit, seq, err);
  static const bool prof = [] { const char* v = getenv("LAKESOUL_B200_ZSTD_PROFILE"); return v && *v == '1'; }();
  if (prof && npages > 0) {
    unsigned long long h[8] = {0}, z[8] = {0};
    cudaStreamSynchronize(s);
    cudaMemcpyFromSymbol(h, g_zstd_prof, sizeof(h));
    cudaMemcpyToSymbol(g_zstd_prof, z, sizeof(z));
    if (h[3]) fprintf(stderr, "ZSTDPROF pages>=256K %llu  mean kcycles plan %.0f entropy %.0f execute %.0f  max plan %.0f entropy %.0f execute %.0f\n",
                      h[3], h[0] / 1e3 / h[3], h[1] / 1e3 / h[3], h[2] / 1e3 / h[3], h[4] / 1e3, h[5] / 1e3, h[6] / 1e3);
  }
}

// ---- launchers (called from device_scan.cc) -----------------------------------------------------
void launch_dict_index(const DecodeArgs& A, cudaStream_t s) {
  if (A.nchunks > 0) dict_index_kernel<<<(A.nchunks + 3) / 4, 128, 0, s>>>(A);
}
void launch_page_index(const DecodeArgs& A, cudaStream_t s) {
  if (A.npages > 0) page_index_kernel<<<(A.npages + 3) / 4, 128, 0, s>>>(A);
}
void launch_block_count(const DecodeArgs& A, cudaStream_t s) {
  if (A.nblocks > 0) block_count_kernel<<<(A.nblocks + 3) / 4, 128, 0, s>>>(A);
}
void launch_block_scan(const DecodeArgs& A, const int32_t* seg_bounds, const int64_t* seg_base, int nseg,
                       int64_t* seg_total, cudaStream_t s) {
  if (nseg > 0) block_scan_kernel<<<(nseg + 3) / 4, 128, 0, s>>>(A, seg_bounds, seg_base, nseg, seg_total);
}
void launch_block_decode(const DecodeArgs& A, cudaStream_t s) {
  if (A.nblocks > 0) block_decode_kernel<<<(A.nblocks + 3) / 4, 128, 0, s>>>(A);
}

}  // namespace lsb
