// SPDX-License-Identifier: Apache-2.0
//
// decode_kernels.cu -- K1: Parquet data-page decode on the GPU, one warp per page.
//   * definition levels (RLE / bit-packed hybrid, bit width 1; legacy BIT_PACKED) -> Arrow validity
//   * PLAIN fixed width (INT32/INT64/FLOAT/DOUBLE/FLBA, decimals re-encoded to 16-byte LE), PLAIN BOOLEAN
//   * PLAIN BYTE_ARRAY (length-prefixed) -> int32 offsets + bytes
//   * PLAIN_DICTIONARY / RLE_DICTIONARY (hybrid-coded indices + PLAIN dictionary page)
// Host side (page_plan.cc) frames the pages (thrift headers, decompression) and never touches values.
// Every lane of a warp holds the same reader state; run headers are parsed redundantly (uniform
// loads), values inside a run are addressed directly by lane.
#include <cuda_runtime.h>

#include "decode_types.h"

namespace lsb {

namespace {

constexpr int kErrBadPage = 1;

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

}  // namespace

// start of every BYTE_ARRAY dictionary entry's payload, one thread per chunk
__global__ void dict_index_kernel(DecodeArgs A) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= A.nchunks) return;
  const PqChunk& ch = A.chunks[c];
  if (ch.dict_nvals <= 0 || ch.phys != PQ_BYTE_ARRAY) return;
  const uint8_t* base = A.raw + ch.dict_off;
  const uint8_t* p = base;
  int32_t* pos = A.dict_pos + ch.dict_index;
  for (int i = 0; i < ch.dict_nvals; ++i) {
    const uint32_t len = load_le32(p);
    pos[i] = (int32_t)(p + 4 - base);
    p += 4 + len;
  }
  pos[ch.dict_nvals] = (int32_t)(p + 4 - base);
}

// bytes produced by every page of a variable-width column (needed to place the pages' data)
__global__ void __launch_bounds__(128) page_count_kernel(DecodeArgs A) {
  const int pgi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (pgi >= A.npages) return;
  PqPage& pg = A.pages[pgi];
  const PqChunk& ch = A.chunks[pg.chunk];
  if (ch.conv != CONV_VAR) return;
  PageView v;
  if (!open_page(A, pg, ch, v)) {
    if (lane == 0) atomicExch(A.err, kErrBadPage);
    return;
  }
  const int n = pg.num_values;
  const bool dict = pg.encoding == ENC_PLAIN_DICT || pg.encoding == ENC_RLE_DICT;
  Hybrid idx;
  const int32_t* dpos = A.dict_pos + ch.dict_index;
  if (dict) hybrid_init(idx, v.values + 1, v.values_end, v.values < v.values_end ? v.values[0] : 0);
  int64_t nonnull = 0, bytes = 0;
  for (int r0 = 0; r0 < n; r0 += 32) {
    const int cnt = min(32, n - r0);
    bool valid = lane < cnt;
    if (v.has_levels) valid = hybrid_get(v.levels, cnt, lane) != 0 && lane < cnt;
    const uint32_t mask = __ballot_sync(0xffffffffu, valid);
    const int nv = __popc(mask);
    nonnull += nv;
    if (dict) {
      const int rank = __popc(mask & ((1u << lane) - 1));
      // indices are stored for non-null values only: lane `rank` of the index batch is ours
      const uint32_t mine = hybrid_get(idx, nv, lane);
      const uint32_t di = __shfl_sync(0xffffffffu, mine, rank);
      int len = 0;
      if (valid && (int)di < ch.dict_nvals) len = dpos[di + 1] - 4 - dpos[di];
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) len += __shfl_xor_sync(0xffffffffu, len, d);
      bytes += len;
    }
  }
  if (!dict) bytes = (int64_t)(v.values_end - v.values) - 4 * nonnull;
  if (lane == 0) pg.str_bytes = bytes < 0 ? 0 : bytes;
}

// exclusive scan of str_bytes over the pages of each (column, bucket) run: pages are listed in
// row order, `seg_start[s]..seg_start[s+1]` delimits one run; one thread per run (runs are short)
__global__ void page_scan_kernel(DecodeArgs A, const int32_t* seg_bounds, const int64_t* seg_base, int nseg,
                                 int64_t* seg_total) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nseg) return;
  int64_t run = seg_base[s];  // bytes of the bucket's host-decoded files come first
  for (int i = seg_bounds[2 * s]; i < seg_bounds[2 * s + 1]; ++i) {
    A.pages[i].str_base = run;
    run += A.pages[i].str_bytes;
  }
  seg_total[s] = run - seg_base[s];
}

__global__ void __launch_bounds__(128) page_decode_kernel(DecodeArgs A) {
  const int pgi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (pgi >= A.npages) return;
  const PqPage pg = A.pages[pgi];
  const PqChunk ch = A.chunks[pg.chunk];
  const PqColOut out = A.cols[ch.col];
  PageView v;
  if (!open_page(A, pg, ch, v)) {
    if (lane == 0) atomicExch(A.err, kErrBadPage);
    return;
  }
  const int n = pg.num_values;
  const bool dict = pg.encoding == ENC_PLAIN_DICT || pg.encoding == ENC_RLE_DICT;
  const bool rle_bool = pg.encoding == ENC_RLE && ch.conv == CONV_BOOL;  // data page v2 booleans
  if (!dict && !rle_bool && pg.encoding != ENC_PLAIN) {
    if (lane == 0) atomicExch(A.err, kErrBadPage);
    return;
  }
  Hybrid idx;
  if (dict) hybrid_init(idx, v.values + 1, v.values_end, v.values < v.values_end ? v.values[0] : 0);
  if (rle_bool) {  // <4-byte length> <hybrid runs, bit width 1>
    const uint32_t len = v.values + 4 <= v.values_end ? load_le32(v.values) : 0;
    hybrid_init(idx, v.values + 4, (v.values + 4 + len < v.values_end) ? (v.values + 4 + len) : v.values_end, 1);
  }
  const int pw = phys_width(ch);
  const uint8_t* dict_base = A.raw + ch.dict_off;
  const int32_t* dpos = A.dict_pos + ch.dict_index;
  const uint8_t* cursor = v.values;  // PLAIN BYTE_ARRAY walk
  int64_t vbase = 0;                  // non-null values consumed so far
  int64_t sbytes = 0;                 // string bytes produced so far
  int32_t* offs = reinterpret_cast<int32_t*>(out.values);
  uint8_t* sdata = ch.conv == CONV_VAR ? out.data + out.bucket_data[ch.bucket] : nullptr;
  for (int r0 = 0; r0 < n; r0 += 32) {
    const int cnt = min(32, n - r0);
    const int64_t row = pg.row_start + r0 + lane;
    bool valid = lane < cnt;
    if (v.has_levels) valid = hybrid_get(v.levels, cnt, lane) != 0 && lane < cnt;
    const uint32_t mask = __ballot_sync(0xffffffffu, valid);
    const int nv = __popc(mask);
    const int rank = __popc(mask & ((1u << lane) - 1));
    if (out.validity && lane == 0) or_bits(out.validity, pg.row_start + r0, mask);
    uint32_t di = 0;
    if (dict || rle_bool) {
      const uint32_t mine = hybrid_get(idx, nv, lane);
      di = __shfl_sync(0xffffffffu, mine, rank);
      if (dict && valid && (int)di >= ch.dict_nvals) { valid = false; atomicExch(A.err, kErrBadPage); }
    }
    if (ch.conv == CONV_VAR) {
      const uint8_t* src = nullptr;
      int len = 0;
      if (dict) {
        if (valid) { src = dict_base + dpos[di]; len = dpos[di + 1] - 4 - dpos[di]; }
      } else {
        // length-prefixed chain: the offsets of a batch can only be found by walking the prefixes
        int my_pos = 0, my_len = 0;
        // simple and branch-uniform: every lane walks to its own rank (L1-resident, <= 32 steps)
        {
          const uint8_t* p = cursor;
          for (int t = 0; t < rank && valid; ++t) p += 4 + (int)load_le32(p);
          if (valid) { my_len = (int)load_le32(p); my_pos = (int)(p + 4 - cursor); }
        }
        if (valid) { src = cursor + my_pos; len = my_len; }
        // advance the shared cursor past the whole batch: last valid lane knows the end
        int endpos = valid ? my_pos + my_len : 0;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) endpos = max(endpos, __shfl_xor_sync(0xffffffffu, endpos, d));
        cursor += endpos;
      }
      // exclusive prefix of the lengths inside the batch
      int incl = len;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += y;
      }
      const int64_t my_off = pg.str_base + sbytes + (incl - len);
      if (lane < cnt) offs[row] = (int32_t)my_off;
      if (len > 0) {
        uint8_t* dst = sdata + my_off;
        for (int i = 0; i < len; ++i) dst[i] = src[i];
      }
      sbytes += __shfl_sync(0xffffffffu, incl, 31);
    } else if (ch.conv == CONV_BOOL) {
      bool bit = false;
      if (valid) {
        if (rle_bool) bit = di != 0;
        else if (dict) bit = false;  // booleans are never dictionary encoded
        else {
          const int64_t b = vbase + rank;
          bit = (v.values[b >> 3] >> (b & 7)) & 1;
        }
      }
      const uint32_t word = __ballot_sync(0xffffffffu, bit);
      if (lane == 0) or_bits(reinterpret_cast<uint8_t*>(out.values), pg.row_start + r0, word);
    } else if (valid) {
      const uint8_t* src = dict ? dict_base + (int64_t)di * pw : v.values + (vbase + rank) * pw;
      emit_fixed(ch, out, src, row);
    }
    vbase += nv;
  }
  if (ch.conv == CONV_VAR && pg.closes_file && lane == 0) offs[pg.row_start + n] = (int32_t)(pg.str_base + sbytes);
}

// ---- launchers (called from device_scan.cc) -----------------------------------------------------
void launch_dict_index(const DecodeArgs& A, cudaStream_t s) {
  if (A.nchunks > 0) dict_index_kernel<<<(A.nchunks + 127) / 128, 128, 0, s>>>(A);
}
void launch_page_count(const DecodeArgs& A, cudaStream_t s) {
  if (A.npages > 0) page_count_kernel<<<(A.npages + 3) / 4, 128, 0, s>>>(A);
}
void launch_page_scan(const DecodeArgs& A, const int32_t* seg_bounds, const int64_t* seg_base, int nseg,
                      int64_t* seg_total, cudaStream_t s) {
  if (nseg > 0) page_scan_kernel<<<(nseg + 127) / 128, 128, 0, s>>>(A, seg_bounds, seg_base, nseg, seg_total);
}
void launch_page_decode(const DecodeArgs& A, cudaStream_t s) {
  if (A.npages > 0) page_decode_kernel<<<(A.npages + 3) / 4, 128, 0, s>>>(A);
}

}  // namespace lsb
