// SPDX-License-Identifier: Apache-2.0
//
// mor_engine.cu -- sm_100a kernels + host planning for the per-hash-bucket k-way
// primary-key merge with last-writer-wins selection, and the winner gather.
//
// Algorithm (see DESIGN.md "PK merge"):
//   * Total order of input rows = (pk, file_idx, row_idx)  -- the order in which the
//     reference's loser tree pops ranges (sort_key_range.rs:182-187) and rows inside a range.
//   * The last element of every equal-pk group in that order is exactly the row the
//     reference keeps (sort_key_range.rs:402-412: last row of the run of the highest stream).
//   * Work is cut into tiles of <= C elements by *regular sampling*: level l+1 = every S-th
//     element of every run of level l. Splitters = every m-th element of the merged level
//     l+1 list. With S*m + k*(S-1) <= C no tile can overflow, whatever the key distribution
//     or duplicate structure (bound is on positions in the total order, not on key values).
//   * One CTA merges one tile in shared memory (pairwise rank-merge tree over the k
//     sub-runs, stable: lower file first), then marks group ends and emits winners.
//   Every input key is read from HBM once at level 0 (+1/S, 1/S^2.. for sample levels).
#include "mor_engine.h"
#include "general_merge.h"

#include <algorithm>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <stdexcept>
#include <type_traits>

namespace lsb {

static thread_local std::string g_last_error;
std::string last_error() { return g_last_error; }

cudaError_t stream_wait(cudaStream_t stream) {
  static const bool spin = [] { const char* v = getenv("LAKESOUL_B200_SPIN_SYNC"); return v && *v == '1'; }();
  if (spin) return cudaStreamSynchronize(stream);
  struct Ev {
    cudaEvent_t e = nullptr;
    int dev = -1;
    ~Ev() { if (e) cudaEventDestroy(e); }
  };
  thread_local Ev ev;
  int dev = 0;
  cudaError_t rc = cudaGetDevice(&dev);
  if (rc != cudaSuccess) return rc;
  if (!ev.e || ev.dev != dev) {
    if (ev.e) cudaEventDestroy(ev.e);
    ev.e = nullptr;
    rc = cudaEventCreateWithFlags(&ev.e, cudaEventBlockingSync | cudaEventDisableTiming);
    if (rc != cudaSuccess) return rc;
    ev.dev = dev;
  }
  rc = cudaEventRecord(ev.e, stream);
  if (rc != cudaSuccess) return rc;
  return cudaEventSynchronize(ev.e);
}


#define LSB_CUDA(x)                                                                   \
  do {                                                                                \
    cudaError_t e_ = (x);                                                             \
    if (e_ != cudaSuccess) {                                                          \
      g_last_error = std::string(#x) + " failed: " + cudaGetErrorString(e_);          \
      throw std::runtime_error(g_last_error);                                         \
    }                                                                                 \
  } while (0)

// ---------------------------------------------------------------------------------------
// keys
// ---------------------------------------------------------------------------------------
struct K128 {
  uint64_t hi, lo;
};
__device__ __forceinline__ bool key_lt(uint64_t a, uint64_t b) { return a < b; }
__device__ __forceinline__ bool key_le(uint64_t a, uint64_t b) { return a <= b; }
__device__ __forceinline__ bool key_ne(uint64_t a, uint64_t b) { return a != b; }
__device__ __forceinline__ bool key_lt(const K128& a, const K128& b) {
  return a.hi < b.hi || (a.hi == b.hi && a.lo < b.lo);
}
__device__ __forceinline__ bool key_le(const K128& a, const K128& b) { return !key_lt(b, a); }
__device__ __forceinline__ bool key_ne(const K128& a, const K128& b) {
  return a.hi != b.hi || a.lo != b.lo;
}

// Order-preserving unsigned image of one PK cell (cursor.rs:265-276 compare()/is_eq()):
// signed -> flip sign bit; float -> IEEE total order (arrow total_cmp); null -> sorts first
// and equals null (cursor.rs:444-476) via an extra leading validity bit.
__device__ __forceinline__ uint64_t load_cell(const PkCol& c, int64_t row, bool& valid) {
  valid = true;
  if (c.validity) valid = (c.validity[row >> 3] >> (row & 7)) & 1;
  uint64_t v;
  switch (c.width) {
    case 1: v = reinterpret_cast<const uint8_t*>(c.values)[row]; break;
    case 2: v = reinterpret_cast<const uint16_t*>(c.values)[row]; break;
    case 4: v = reinterpret_cast<const uint32_t*>(c.values)[row]; break;
    default: v = reinterpret_cast<const uint64_t*>(c.values)[row]; break;
  }
  const int bits = c.width * 8;
  const uint64_t top = 1ull << (bits - 1);
  if (c.kind == PK_SIGNED) {
    v ^= top;
  } else if (c.kind == PK_FLOAT) {
    const uint64_t mask = bits == 64 ? ~0ull : ((1ull << bits) - 1);
    v = (v & top) ? (~v & mask) : (v | top);
  }
  return valid ? v : 0ull;
}

template <typename KeyT>
struct KeyLoader;

template <>
struct KeyLoader<uint64_t> {
  static __device__ __forceinline__ uint64_t load(const PkDesc& pk, int64_t row) {
    uint64_t acc = 0;
#pragma unroll
    for (int c = 0; c < kMaxPkCols; ++c) {
      if (c < pk.ncols) {
        bool valid;
        uint64_t v = load_cell(pk.col[c], row, valid);
        const int bits = pk.col[c].width * 8;
        if (pk.col[c].validity) acc = (acc << 1) | (valid ? 1ull : 0ull);
        acc = bits == 64 ? v : ((acc << bits) | v);
      }
    }
    return acc;
  }
};

template <>
struct KeyLoader<K128> {
  static __device__ __forceinline__ void push(K128& a, uint64_t v, int bits) {
    if (bits == 64) {
      a.hi = a.lo;
      a.lo = v;
    } else {
      a.hi = (a.hi << bits) | (a.lo >> (64 - bits));
      a.lo = (a.lo << bits) | v;
    }
  }
  static __device__ __forceinline__ K128 load(const PkDesc& pk, int64_t row) {
    K128 acc{0, 0};
#pragma unroll
    for (int c = 0; c < kMaxPkCols; ++c) {
      if (c < pk.ncols) {
        bool valid;
        uint64_t v = load_cell(pk.col[c], row, valid);
        if (pk.col[c].validity) push(acc, valid ? 1ull : 0ull, 1);
        push(acc, v, pk.col[c].width * 8);
      }
    }
    return acc;
  }
};

// Utf8 / Binary primary key (single column; reference typed fast path sorted_stream_merger.rs:311-315,
// byte-wise unsigned compare cursor.rs:283-324): 8-byte big-endian prefix in shared memory, the rest
// of the string is fetched from HBM only when two prefixes tie.
struct StrKey {
  uint64_t prefix;
  uint32_t row;   // global row, kNoRow for a null key (nulls sort first and are equal, cursor.rs:444-476)
  uint32_t len;
};

template <typename KeyT>
struct KeyOps {
  PkDesc pk;
  __device__ __forceinline__ KeyOps(const PkDesc& p, int) : pk(p) {}
  __device__ __forceinline__ KeyT load(int64_t row) const { return KeyLoader<KeyT>::load(pk, row); }
  __device__ __forceinline__ bool lt(const KeyT& a, const KeyT& b) const { return key_lt(a, b); }
  __device__ __forceinline__ bool le(const KeyT& a, const KeyT& b) const { return key_le(a, b); }
  __device__ __forceinline__ bool ne(const KeyT& a, const KeyT& b) const { return key_ne(a, b); }
};

template <>
struct KeyOps<StrKey> {
  const int32_t* offs;
  const uint8_t* valid;
  const uint8_t* data;  // bucket's bytes
  __device__ __forceinline__ KeyOps(const PkDesc& p, int bucket)
      : offs(reinterpret_cast<const int32_t*>(p.col[0].values)), valid(p.col[0].validity),
        data(p.col[0].data + p.col[0].bucket_data[bucket]) {}
  __device__ __forceinline__ StrKey load(int64_t row) const {
    StrKey k{0, (uint32_t)row, 0};
    if (valid && !((valid[row >> 3] >> (row & 7)) & 1)) { k.row = kNoRow; return k; }
    const int32_t o0 = offs[row], o1 = offs[row + 1];
    k.len = (uint32_t)(o1 - o0);
    const uint8_t* p = data + o0;
    const int n = k.len < 8 ? (int)k.len : 8;
    uint64_t v = 0;
    for (int i = 0; i < n; ++i) v |= (uint64_t)p[i] << (56 - 8 * i);
    k.prefix = v;
    return k;
  }
  // <0, 0, >0
  __device__ __forceinline__ int cmp(const StrKey& a, const StrKey& b) const {
    const bool an = a.row == kNoRow, bn = b.row == kNoRow;
    if (an || bn) return an == bn ? 0 : (an ? -1 : 1);
    if (a.prefix != b.prefix) return a.prefix < b.prefix ? -1 : 1;
    const uint32_t m = a.len < b.len ? a.len : b.len;
    if (m > 8) {
      const uint8_t* pa = data + offs[a.row];
      const uint8_t* pb = data + offs[b.row];
      for (uint32_t i = 8; i < m; ++i) {
        const int d = (int)pa[i] - (int)pb[i];
        if (d) return d;
      }
    }
    return a.len == b.len ? 0 : (a.len < b.len ? -1 : 1);
  }
  __device__ __forceinline__ bool lt(const StrKey& a, const StrKey& b) const { return cmp(a, b) < 0; }
  __device__ __forceinline__ bool le(const StrKey& a, const StrKey& b) const { return cmp(a, b) <= 0; }
  __device__ __forceinline__ bool ne(const StrKey& a, const StrKey& b) const { return cmp(a, b) != 0; }
};

__device__ __forceinline__ uint64_t shfl_xor_key(uint64_t k, int d) { return __shfl_xor_sync(0xffffffffu, k, d); }
__device__ __forceinline__ K128 shfl_xor_key(const K128& k, int d) {
  K128 r;
  r.hi = __shfl_xor_sync(0xffffffffu, k.hi, d);
  r.lo = __shfl_xor_sync(0xffffffffu, k.lo, d);
  return r;
}
__device__ __forceinline__ StrKey shfl_xor_key(const StrKey& k, int d) {
  StrKey r;
  r.prefix = __shfl_xor_sync(0xffffffffu, k.prefix, d);
  r.row = __shfl_xor_sync(0xffffffffu, k.row, d);
  r.len = __shfl_xor_sync(0xffffffffu, k.len, d);
  return r;
}

// ---------------------------------------------------------------------------------------
// device plan view
// ---------------------------------------------------------------------------------------
struct PlanDev {
  const BucketDesc* buckets;
  const int32_t* file_rows;       // F
  const int64_t* file_row_base;   // F+1
  int32_t* splits;                // all levels
  uint2* mlist;                   // merged sample lists (x = file-in-bucket, y = level position)
  const int32_t* tile_bucket;     // all levels, concatenated
  const int32_t* tile_idx;
  int32_t* tile_cnt;              // level-0 tiles
  int64_t* tile_gstart;           // level-0 tiles: global merged position of the tile start
  int64_t* tile_out;              // level-0 tiles + 1
  uint32_t* win_tmp;              // N
  uint32_t* win;                  // N
  uint8_t* wfile_tmp;             // N: file (inside its bucket) of every winner, written next to win_tmp
  uint8_t* wfile;                 // N: dense, parallel to win
  int64_t* out_base;              // nb+1
  int64_t* batch_base;            // nb+1
  const int64_t* dense_base;      // F+1: prefix of file_rows (rows packed without padding)
  uint32_t* R;                    // general path: presence matrix (last row per output key and file)
  const int64_t* rofs;            // general path: offset of each bucket's matrix inside R
  int32_t* err;
  int32_t nb;
  int32_t log_s;
  int32_t tile_cap;
  int32_t batch_size;
};

constexpr int kTileThreads = 512;
constexpr int kErrTileOverflow = 1;

// largest r in [0, n) with arr[r] <= q   (arr non-decreasing, arr[0] <= q)
template <typename T, typename Q>
__device__ __forceinline__ int upper_idx(const T* arr, int n, Q q) {
  int lo = 0, hi = n;  // invariant arr[lo] <= q, answer in [lo, hi)
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (arr[mid] <= q) lo = mid; else hi = mid;
  }
  return lo;
}

// ---------------------------------------------------------------------------------------
// K_part: tile boundaries of level l from the merged list of level l+1
// One warp per tile computes the tile's LOWER boundary; the last tile of a bucket also
// writes the bucket's final boundary.
// ---------------------------------------------------------------------------------------
template <typename KeyT>
__global__ void __launch_bounds__(128) part_kernel(PlanDev P, PkDesc pk, int level, int tile_ofs,
                                                   int ntiles) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= ntiles) return;
  const int t = tile_ofs + warp;
  const int b = P.tile_bucket[t];
  const int i = P.tile_idx[t];
  const BucketDesc& B = P.buckets[b];
  const KeyOps<KeyT> ops(pk, b);
  const int k = B.k;
  const int T = B.ntiles[level];
  int32_t* sp = P.splits + B.split_ofs[level];
  const int sh = level * P.log_s;

  int js = -1;
  int64_t pprime = 0;
  KeyT keyS{};
  if (i > 0) {
    const uint2 s = P.mlist[B.mlist_ofs[level + 1] + (int64_t)i * B.m - 1];
    js = (int)s.x;
    const int64_t ps = s.y;
    const int64_t rowS =
        P.file_row_base[B.file_begin + js] + (((ps + 1) << (sh + P.log_s)) - 1);
    keyS = ops.load(rowS);
    pprime = ((ps + 1) << P.log_s) - 1;
  }
  for (int j = lane; j < k; j += 32) {
    const int f = B.file_begin + j;
    const int32_t nl = (int32_t)((int64_t)P.file_rows[f] >> sh);
    int32_t v = 0;
    if (i > 0) {
      if (j == js) {
        v = (int32_t)(pprime + 1);
      } else {
        const int64_t fb = P.file_row_base[f];
        int lo = 0, hi = nl;
        if (j < js) {  // count of elements with key <= keyS
          while (lo < hi) {
            int mid = (lo + hi) >> 1;
            KeyT km = ops.load(fb + ((((int64_t)mid + 1) << sh) - 1));
            if (ops.le(km, keyS)) lo = mid + 1; else hi = mid;
          }
        } else {  // count of elements with key < keyS
          while (lo < hi) {
            int mid = (lo + hi) >> 1;
            KeyT km = ops.load(fb + ((((int64_t)mid + 1) << sh) - 1));
            if (ops.lt(km, keyS)) lo = mid + 1; else hi = mid;
          }
        }
        v = lo;
      }
    }
    sp[(int64_t)i * k + j] = v;
    if (i == T - 1) sp[(int64_t)T * k + j] = nl;
  }
}

// ---------------------------------------------------------------------------------------
// block scan helper (exclusive), blockDim.x == kTileThreads
// ---------------------------------------------------------------------------------------
template <int kThreads>
__device__ __forceinline__ int block_exclusive_scan(int v, int* warp_sums, int& total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int x = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int y = __shfl_up_sync(0xffffffffu, x, d);
    if (lane >= d) x += y;
  }
  if (lane == 31) warp_sums[w] = x;
  __syncthreads();
  if (w == 0) {
    int s = lane < (kThreads / 32) ? warp_sums[lane] : 0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, s, d);
      if (lane >= d) s += y;
    }
    if (lane < (kThreads / 32)) warp_sums[lane] = s;  // inclusive warp totals
  }
  __syncthreads();
  const int base = w == 0 ? 0 : warp_sums[w - 1];
  total = warp_sums[kThreads / 32 - 1];
  return base + x - v;
}

// ---------------------------------------------------------------------------------------
// K_tile: merge one tile in shared memory.
//   level > 0 : write the merged (file, level-pos) list      -> mlist
//   level == 0: mark group ends, write winners (global rows) -> win_tmp, count -> tile_cnt
// ---------------------------------------------------------------------------------------
template <typename KeyT, bool kGeneral>
__global__ void __launch_bounds__(kTileThreads, 3) tile_kernel(PlanDev P, PkDesc pk, int level,
                                                            int tile_ofs, const uint8_t* __restrict__ tile_flag) {
  // tile_flag != nullptr: only the tiles the fast kernel could not take (their flag is raised)
  if (tile_flag && !tile_flag[tile_ofs + blockIdx.x]) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int C = P.tile_cap;
  KeyT* K = reinterpret_cast<KeyT*>(smem_raw);
  uint16_t* idxA = reinterpret_cast<uint16_t*>(K + C);
  uint16_t* idxB = idxA + C;
  // C is a multiple of 8 (host-checked), so everything below stays 8-byte aligned
  int64_t* frb = reinterpret_cast<int64_t*>(idxB + C);
  int32_t* seg_start = reinterpret_cast<int32_t*>(frb + kMaxFilesPerBucket);
  int32_t* seg_lo = seg_start + kMaxFilesPerBucket + 1;
  int32_t* scan_tmp = seg_lo + kMaxFilesPerBucket;  // 8 + 2 ints
  __shared__ int64_t s_tile_start;
  __shared__ int s_has_next;
  __shared__ KeyT s_next_key;

  const int t = tile_ofs + blockIdx.x;
  const int b = P.tile_bucket[t];
  const int i = P.tile_idx[t];
  const BucketDesc& B = P.buckets[b];
  const KeyOps<KeyT> ops(pk, b);
  const int k = B.k;
  const int32_t* sp_lo = P.splits + B.split_ofs[level] + (int64_t)i * k;
  const int32_t* sp_hi = sp_lo + k;
  const int sh = level * P.log_s;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;

  // ---- segment table -----------------------------------------------------------------
  for (int j = tid; j < k; j += kTileThreads) {
    seg_lo[j] = sp_lo[j];
    seg_start[j + 1] = sp_hi[j] - sp_lo[j];  // lengths, scanned below
    frb[j] = P.file_row_base[B.file_begin + j];
  }
  __syncthreads();
  if (warp == 0) {
    // serial-in-lane scan of <= 8 entries per lane, then warp scan
    const int per = (k + 31) >> 5;
    int sum = 0;
    int64_t lsum = 0;
    for (int x = 0; x < per; ++x) {
      int j = lane * per + x;
      if (j < k) { sum += seg_start[j + 1]; lsum += seg_lo[j]; }
    }
    int incl = sum;
    int64_t lincl = lsum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, incl, d);
      int64_t ly = __shfl_up_sync(0xffffffffu, lincl, d);
      if (lane >= d) { incl += y; lincl += ly; }
    }
    int run = incl - sum;
    for (int x = 0; x < per; ++x) {
      int j = lane * per + x;
      if (j < k) { int len = seg_start[j + 1]; seg_start[j + 1] = run + len; run += len; }
    }
    if (lane == 0) seg_start[0] = 0;
    if (lane == 31) s_tile_start = lincl;
  }
  __syncthreads();
  const int L = seg_start[k];
  if (L > C) {
    if (tid == 0) atomicExch(P.err, kErrTileOverflow);
    return;
  }

  // ---- load keys: elements strided over the whole CTA (independent loads in flight per thread;
  // consecutive elements of a segment are consecutive rows, so warps stay coalesced at level 0)
  for (int e0 = tid; e0 < L; e0 += 4 * kTileThreads) {
    // 4 elements per thread per trip: addresses first, then all global loads, then the stores
    int64_t row[4];
    KeyT kv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + u * kTileThreads;
      row[u] = 0;
      if (e < L) {
        const int j = upper_idx(seg_start, k, e);
        row[u] = frb[j] + ((((int64_t)seg_lo[j] + (e - seg_start[j]) + 1) << sh) - 1);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (e0 + u * kTileThreads < L) kv[u] = ops.load(row[u]);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + u * kTileThreads;
      if (e < L) {
        K[e] = kv[u];
        idxA[e] = (uint16_t)e;
      }
    }
  }
  if (level == 0 && warp == (kTileThreads / 32 - 1)) {
    // peek the first key after this tile (min over the k run heads)
    bool have = false;
    KeyT best{};
    for (int j = lane; j < k; j += 32) {
      const int32_t h = sp_hi[j];
      if (h < P.file_rows[B.file_begin + j]) {
        KeyT kk = ops.load(frb[j] + h);
        if (!have || ops.lt(kk, best)) { best = kk; have = true; }
      }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      bool oh = __shfl_xor_sync(0xffffffffu, (int)have, d);
      const KeyT ob = shfl_xor_key(best, d);
      if (oh && (!have || ops.lt(ob, best))) { best = ob; have = true; }
    }
    if (lane == 0) { s_has_next = have; s_next_key = best; }
  }
  __syncthreads();

  // ---- pairwise merge tree, merge-path co-ranking ----------------------------------------
  // Each pass merges adjacent run pairs (stable: the left run = lower files wins ties). A thread
  // owns E consecutive OUTPUT positions: one diagonal binary search (co-rank) per pair it touches,
  // then a sequential two-pointer merge with both heads cached in registers.
  uint16_t* in = idxA;
  uint16_t* out = idxB;
  int nr = k;
  const int EM = (L + kTileThreads - 1) / kTileThreads;
  const int mo0 = min(tid * EM, L), mo1 = min(mo0 + EM, L);
  for (int pass = 0; nr > 1; ++pass) {
    // run r of this pass spans original segments [r<<pass, (r+1)<<pass)
    auto rb = [&](int r) -> int {
      int s = r << pass;
      return seg_start[s < k ? s : k];
    };
    int q = mo0;
    while (q < mo1) {
      // run containing q: largest r with rb(r) <= q
      int lo = 0, hi = nr;
      while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (rb(mid) <= q) lo = mid; else hi = mid;
      }
      const int base = lo & ~1;
      const int a0 = rb(base);
      if (base + 1 >= nr) {  // unpaired last run: copy through
        for (; q < mo1; ++q) out[q] = in[q];
        break;
      }
      const int m0 = rb(base + 1), e0 = rb(base + 2);
      const int stop = min(mo1, e0);
      const int na = m0 - a0, nbb = e0 - m0;
      const int d = q - a0;
      int i = max(0, d - nbb), ihi = min(d, na);
      while (i < ihi) {  // co-rank: number of left-run elements among the first d outputs
        const int mid = (i + ihi) >> 1;
        const KeyT ka = K[in[a0 + mid]];
        const KeyT kb = K[in[m0 + d - 1 - mid]];
        if (!ops.lt(kb, ka)) i = mid + 1; else ihi = mid;
      }
      int j = d - i;
      uint16_t ea = 0, eb = 0;
      KeyT ka{}, kb{};
      if (i < na) { ea = in[a0 + i]; ka = K[ea]; }
      if (j < nbb) { eb = in[m0 + j]; kb = K[eb]; }
      for (; q < stop; ++q) {
        // predicated two-pointer step: no divergent branches, one index + one key load per output
        const bool take_a = (j >= nbb) || (i < na && !ops.lt(kb, ka));
        out[q] = take_a ? ea : eb;
        i += take_a ? 1 : 0;
        j += take_a ? 0 : 1;
        const int nxt = take_a ? (a0 + i) : (m0 + j);
        const bool more = take_a ? (i < na) : (j < nbb);
        const uint16_t en = in[more ? nxt : a0];
        const KeyT kn = K[en];
        if (take_a) { ea = en; ka = kn; } else { eb = en; kb = kn; }
      }
    }
    __syncthreads();
    uint16_t* tmp = in; in = out; out = tmp;
    nr = (nr + 1) >> 1;
  }
  // `in` now holds the merged order

  if (level > 0) {
    uint2* ml = P.mlist + B.mlist_ofs[level] + s_tile_start;
    for (int q = tid; q < L; q += kTileThreads) {
      const int e = in[q];
      const int j = upper_idx(seg_start, k, e);
      ml[q] = make_uint2((unsigned)j, (unsigned)(seg_lo[j] + (e - seg_start[j])));
    }
    return;
  }

  // ---- level 0: group ends -> winners ---------------------------------------------------
  const int E = (L + kTileThreads - 1) / kTileThreads;
  const int q0 = min(tid * E, L), q1 = min(q0 + E, L);
  const bool has_next = s_has_next != 0;
  const KeyT next_key = s_next_key;
  int cnt = 0;
  for (int q = q0; q < q1; ++q) {
    const KeyT kq = K[in[q]];
    bool end;
    if (q + 1 < L) end = ops.ne(kq, K[in[q + 1]]);
    else end = has_next ? ops.ne(kq, next_key) : true;
    cnt += end ? 1 : 0;
  }
  int total;
  int off = block_exclusive_scan<kTileThreads>(cnt, scan_tmp, total);
  const int t0 = blockIdx.x;  // level-0 tiles start at tile_ofs
  if constexpr (kGeneral) {
    // second pass: R[u][j] = last row of output key u inside file j
    int64_t u_run = P.tile_out[t0] + off - P.out_base[b];
    uint32_t* Rb = P.R + P.rofs[b];
    for (int q = q0; q < q1; ++q) {
      const int e = in[q];
      const KeyT kq = K[e];
      const int j = upper_idx(seg_start, k, e);
      bool end, sub_last;
      if (q + 1 < L) {
        const int e2 = in[q + 1];
        end = ops.ne(kq, K[e2]);
        sub_last = end || upper_idx(seg_start, k, e2) != j;
      } else {
        end = has_next ? ops.ne(kq, next_key) : true;
        const int32_t h = sp_hi[j];
        sub_last = true;
        if (h < P.file_rows[B.file_begin + j]) {
          const KeyT kn = ops.load(frb[j] + h);
          sub_last = ops.ne(kn, kq);
        }
      }
      if (sub_last) Rb[u_run * k + j] = (uint32_t)(frb[j] + seg_lo[j] + (e - seg_start[j]));
      if (end) ++u_run;
    }
    return;
  }
  const int64_t gstart = B.row_base + s_tile_start;
  uint32_t* wt = P.win_tmp + gstart;
  uint8_t* wf = P.wfile_tmp + gstart;
  for (int q = q0; q < q1; ++q) {
    const int e = in[q];
    const KeyT kq = K[e];
    bool end;
    if (q + 1 < L) end = ops.ne(kq, K[in[q + 1]]);
    else end = has_next ? ops.ne(kq, next_key) : true;
    if (end) {
      const int j = upper_idx(seg_start, k, e);
      wt[off] = (uint32_t)(frb[j] + seg_lo[j] + (e - seg_start[j]));
      wf[off] = (uint8_t)j;  // the string gather stages per-file source windows: no search per column there
      ++off;
    }
  }
  if (tid == 0) {
    P.tile_cnt[t0] = total;
    P.tile_gstart[t0] = gstart;
  }
}

// ---------------------------------------------------------------------------------------
// K_tile, fast path for fixed-width keys: the tile's keys are compressed to 64 bits RELATIVE TO THE TILE'S
// SMALLEST KEY (a tile spans a narrow key range; composite 96/128-bit keys fit whenever the span is below 2^64)
// and travel through the merge passes TOGETHER with their 16-bit element index: every step of the two-pointer
// merge reads one 8-byte key and one index by position -- no index -> key double load, no 128-bit compares.
// (r1 profile of tile_kernel<K128>: 590 thread instructions per key, issue-bound.)
// A tile whose span does not fit raises its flag and is merged by tile_kernel afterwards.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ K128 k128_sub(const K128& a, const K128& b) {  // a - b, a >= b
  K128 r;
  r.lo = a.lo - b.lo;
  r.hi = a.hi - b.hi - (a.lo < b.lo ? 1ull : 0ull);
  return r;
}
__device__ __forceinline__ K128 widen(uint64_t v) { return K128{0ull, v}; }
__device__ __forceinline__ K128 widen(const K128& v) { return v; }

static size_t tile_fast_smem_bytes(int C) {
  return (size_t)C * 8 * 2 + (size_t)C * 2 * 2 + kMaxFilesPerBucket * 8 + (kMaxFilesPerBucket + 1) * 4 +
         kMaxFilesPerBucket * 4 + 16 * 4;
}

// Key of one row for tile_fast_kernel. kShape 1 = (8-byte, 4-byte) integer columns without validity bitmaps -- the
// composite key of config 2 and of most fact tables: the same 96-bit image KeyLoader<K128> builds, without its loops
// over columns, widths and kinds (r3 profile: the generic loader carried 46 % of the kernel's stall samples).
template <int kShape, typename KeyT>
__device__ __forceinline__ KeyT load_key_shape(const KeyOps<KeyT>& ops, const PkDesc& pk, int64_t row) {
  if constexpr (kShape == 1 && std::is_same<KeyT, K128>::value) {
    uint64_t a = __ldg(reinterpret_cast<const uint64_t*>(pk.col[0].values) + row);
    uint32_t b = __ldg(reinterpret_cast<const uint32_t*>(pk.col[1].values) + row);
    if (pk.col[0].kind == PK_SIGNED) a ^= 1ull << 63;
    if (pk.col[1].kind == PK_SIGNED) b ^= 1u << 31;
    return K128{a >> 32, (a << 32) | b};
  } else {
    return load_key_shape<kShape>(ops, pk, row);
  }
}

template <typename KeyT, int kThreads, int kShape>
__global__ void __launch_bounds__(kThreads, 3) tile_fast_kernel(PlanDev P, PkDesc pk, int level, int tile_ofs,
                                                                 uint8_t* __restrict__ tile_flag) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int C = P.tile_cap;
  uint64_t* Ka = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* Kb = Ka + C;
  uint16_t* Ia = reinterpret_cast<uint16_t*>(Kb + C);
  uint16_t* Ib = Ia + C;
  int64_t* frb = reinterpret_cast<int64_t*>(Ib + C);  // C is a multiple of 8
  K128* hmin = reinterpret_cast<K128*>(Ka);            // per-file head / tail keys: live only until the keys are loaded
  K128* hmax = hmin + kMaxFilesPerBucket;              // (2 * 256 * 16 bytes <= the key buffers, host-checked)
  int32_t* seg_start = reinterpret_cast<int32_t*>(frb + kMaxFilesPerBucket);
  int32_t* seg_lo = seg_start + kMaxFilesPerBucket + 1;
  int32_t* scan_tmp = seg_lo + kMaxFilesPerBucket;
  __shared__ int64_t s_tile_start;
  __shared__ int s_has_next, s_fits;
  __shared__ K128 s_min;
  __shared__ unsigned long long s_next_rel;  // first key after the tile, relative to s_min (valid when s_has_next == 1)

  const int t = tile_ofs + blockIdx.x;
  const int b = P.tile_bucket[t];
  const int i = P.tile_idx[t];
  const BucketDesc& B = P.buckets[b];
  const KeyOps<KeyT> ops(pk, b);
  const int k = B.k;
  const int32_t* sp_lo = P.splits + B.split_ofs[level] + (int64_t)i * k;
  const int32_t* sp_hi = sp_lo + k;
  const int sh = level * P.log_s;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;

  // ---- segment table + head / tail key of every run ---------------------------------------------
  for (int j = tid; j < k; j += kThreads) {
    const int32_t lo = sp_lo[j], hi = sp_hi[j];
    seg_lo[j] = lo;
    seg_start[j + 1] = hi - lo;  // lengths, scanned below
    const int64_t fb = P.file_row_base[B.file_begin + j];
    frb[j] = fb;
    if (hi > lo) {
      hmin[j] = widen(load_key_shape<kShape>(ops, pk, fb + ((((int64_t)lo + 1) << sh) - 1)));
      hmax[j] = widen(load_key_shape<kShape>(ops, pk, fb + ((((int64_t)hi) << sh) - 1)));
    }
  }
  __syncthreads();
  if (warp == 0) {
    const int per = (k + 31) >> 5;
    int sum = 0;
    int64_t lsum = 0;
    bool have = false;
    K128 mn{0, 0}, mx{0, 0};
    for (int x = 0; x < per; ++x) {
      const int j = lane * per + x;
      if (j < k) {
        const int len = seg_start[j + 1];
        sum += len;
        lsum += seg_lo[j];
        if (len > 0) {
          const K128 a = hmin[j], z = hmax[j];
          if (!have || key_lt(a, mn)) mn = a;
          if (!have || key_lt(mx, z)) mx = z;
          have = true;
        }
      }
    }
    int incl = sum;
    int64_t lincl = lsum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, incl, d);
      const int64_t ly = __shfl_up_sync(0xffffffffu, lincl, d);
      if (lane >= d) { incl += y; lincl += ly; }
    }
    int run = incl - sum;
    for (int x = 0; x < per; ++x) {
      const int j = lane * per + x;
      if (j < k) { const int len = seg_start[j + 1]; seg_start[j + 1] = run + len; run += len; }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const bool oh = __shfl_xor_sync(0xffffffffu, (int)have, d);
      const K128 omn = shfl_xor_key(mn, d), omx = shfl_xor_key(mx, d);
      if (oh) {
        if (!have || key_lt(omn, mn)) mn = omn;
        if (!have || key_lt(mx, omx)) mx = omx;
        have = true;
      }
    }
    if (lane == 0) {
      seg_start[0] = 0;
      s_min = mn;
      s_fits = (!have || k128_sub(mx, mn).hi == 0) ? 1 : 0;
    }
    if (lane == 31) s_tile_start = lincl;
  }
  __syncthreads();
  const int L = seg_start[k];
  if (L > C) {
    if (tid == 0) atomicExch(P.err, kErrTileOverflow);
    return;
  }
  if (!s_fits) {  // span of 2^64 or more: the generic kernel takes this tile
    if (tid == 0) tile_flag[t] = 1;
    return;
  }
  const K128 kmin = s_min;

  // ---- load keys (coalesced inside a run), compress, pair with the element index ---------------
  for (int e0 = tid; e0 < L; e0 += 4 * kThreads) {
    int64_t row[4];
    KeyT kv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + u * kThreads;
      row[u] = 0;
      if (e < L) {
        const int j = upper_idx(seg_start, k, e);
        row[u] = frb[j] + ((((int64_t)seg_lo[j] + (e - seg_start[j]) + 1) << sh) - 1);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (e0 + u * kThreads < L) kv[u] = load_key_shape<kShape>(ops, pk, row[u]);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + u * kThreads;
      if (e < L) {
        Ka[e] = k128_sub(widen(kv[u]), kmin).lo;
        Ia[e] = (uint16_t)e;
      }
    }
  }
  if (level == 0 && warp == (kThreads / 32 - 1)) {
    // peek the first key after this tile (min over the k run heads)
    bool have = false;
    K128 best{0, 0};
    for (int j = lane; j < k; j += 32) {
      const int32_t h = sp_hi[j];
      if (h < P.file_rows[B.file_begin + j]) {
        const K128 kk = widen(load_key_shape<kShape>(ops, pk, frb[j] + h));
        if (!have || key_lt(kk, best)) { best = kk; have = true; }
      }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const bool oh = __shfl_xor_sync(0xffffffffu, (int)have, d);
      const K128 ob = shfl_xor_key(best, d);
      if (oh && (!have || key_lt(ob, best))) { best = ob; have = true; }
    }
    if (lane == 0) {
      // every later key is >= the tile's keys; one that is 2^64 or more above the tile's minimum differs from all of them
      const K128 rel = have ? k128_sub(best, kmin) : K128{0, 0};
      s_has_next = have ? (rel.hi == 0 ? 1 : 2) : 0;
      s_next_rel = rel.lo;
    }
  }
  __syncthreads();

  // ---- pairwise merge tree over (key, index) pairs, merge-path co-ranking ------------------------
  uint64_t* Kin = Ka; uint64_t* Kout = Kb;
  uint16_t* Iin = Ia; uint16_t* Iout = Ib;
  int nr = k;
  const int EM = (L + kThreads - 1) / kThreads;
  const int mo0 = min(tid * EM, L), mo1 = min(mo0 + EM, L);
  for (int pass = 0; nr > 1; ++pass) {
    auto rb = [&](int r) -> int {
      const int s2 = r << pass;
      return seg_start[s2 < k ? s2 : k];
    };
    int q = mo0;
    while (q < mo1) {
      int lo = 0, hi = nr;
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (rb(mid) <= q) lo = mid; else hi = mid;
      }
      const int base = lo & ~1;
      const int a0 = rb(base);
      if (base + 1 >= nr) {  // unpaired last run: copy through
        for (; q < mo1; ++q) { Kout[q] = Kin[q]; Iout[q] = Iin[q]; }
        break;
      }
      const int m0 = rb(base + 1), e0 = rb(base + 2);
      const int stop = min(mo1, e0);
      const int na = m0 - a0, nbb = e0 - m0;
      const int d = q - a0;
      int ia = max(0, d - nbb), ihi = min(d, na);
      while (ia < ihi) {  // co-rank: number of left-run elements among the first d outputs
        const int mid = (ia + ihi) >> 1;
        if (Kin[a0 + mid] <= Kin[m0 + d - 1 - mid]) ia = mid + 1; else ihi = mid;
      }
      int ja = d - ia;
      uint64_t ka = ia < na ? Kin[a0 + ia] : ~0ull;
      uint64_t kb = ja < nbb ? Kin[m0 + ja] : ~0ull;
      uint16_t ea = ia < na ? Iin[a0 + ia] : (uint16_t)0;
      uint16_t eb = ja < nbb ? Iin[m0 + ja] : (uint16_t)0;
      for (; q < stop; ++q) {
        const bool take_a = (ja >= nbb) || (ia < na && ka <= kb);  // stable: the left run (lower files) wins ties
        Kout[q] = take_a ? ka : kb;
        Iout[q] = take_a ? ea : eb;
        ia += take_a ? 1 : 0;
        ja += take_a ? 0 : 1;
        const int nxt = take_a ? (a0 + ia) : (m0 + ja);
        const bool more = take_a ? (ia < na) : (ja < nbb);
        const uint64_t kn = Kin[more ? nxt : a0];
        const uint16_t en = Iin[more ? nxt : a0];
        if (take_a) { ka = kn; ea = en; } else { kb = kn; eb = en; }
      }
    }
    __syncthreads();
    uint64_t* tk = Kin; Kin = Kout; Kout = tk;
    uint16_t* ti = Iin; Iin = Iout; Iout = ti;
    nr = (nr + 1) >> 1;
  }
  // (Kin, Iin) now hold the merged order

  if (level > 0) {
    uint2* ml = P.mlist + B.mlist_ofs[level] + s_tile_start;
    for (int q = tid; q < L; q += kThreads) {
      const int e = Iin[q];
      const int j = upper_idx(seg_start, k, e);
      ml[q] = make_uint2((unsigned)j, (unsigned)(seg_lo[j] + (e - seg_start[j])));
    }
    return;
  }

  // ---- level 0: group ends -> winners ---------------------------------------------------------
  const int E = (L + kThreads - 1) / kThreads;
  const int q0 = min(tid * E, L), q1 = min(q0 + E, L);
  const int has_next = s_has_next;
  const unsigned long long next_rel = s_next_rel;
  unsigned endmask = 0;  // E <= 11 for the tile capacities in use (host-checked: C <= 16 * kThreads)
  int cnt = 0;
  for (int q = q0; q < q1; ++q) {
    const uint64_t kq = Kin[q];
    bool end;
    if (q + 1 < L) end = kq != Kin[q + 1];
    else end = has_next == 0 ? true : (has_next == 2 ? true : kq != next_rel);
    endmask |= end ? (1u << (q - q0)) : 0u;
    cnt += end ? 1 : 0;
  }
  int total;
  int off = block_exclusive_scan<kThreads>(cnt, scan_tmp, total);
  const int t0 = blockIdx.x;
  const int64_t gstart = B.row_base + s_tile_start;
  uint32_t* wt = P.win_tmp + gstart;
  uint8_t* wf = P.wfile_tmp + gstart;
  for (int q = q0; q < q1; ++q) {
    if (endmask & (1u << (q - q0))) {
      const int e = Iin[q];
      const int j = upper_idx(seg_start, k, e);
      wt[off] = (uint32_t)(frb[j] + seg_lo[j] + (e - seg_start[j]));
      wf[off] = (uint8_t)j;
      ++off;
    }
  }
  if (tid == 0) {
    P.tile_cnt[t0] = total;
    P.tile_gstart[t0] = gstart;
  }
}

// ---------------------------------------------------------------------------------------
// K_scan: exclusive scan of level-0 tile counts -> tile_out; bucket output bases; batch bases
// single block of 1024 threads
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) scan_kernel(PlanDev P, int ntiles0) {
  __shared__ int64_t wsum[32];
  __shared__ int64_t carry_s;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  auto scan_chunk = [&](int64_t v) -> int64_t {  // returns exclusive prefix incl. carry
    int64_t x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int64_t y = __shfl_up_sync(0xffffffffu, x, d);
      if (lane >= d) x += y;
    }
    if (lane == 31) wsum[w] = x;
    __syncthreads();
    if (w == 0) {
      int64_t s = wsum[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        int64_t y = __shfl_up_sync(0xffffffffu, s, d);
        if (lane >= d) s += y;
      }
      wsum[lane] = s;
    }
    __syncthreads();
    const int64_t carry = carry_s;
    const int64_t res = carry + (w == 0 ? 0 : wsum[w - 1]) + x - v;
    __syncthreads();
    if (tid == 1023) carry_s = carry + wsum[31];
    __syncthreads();
    return res;
  };
  for (int base = 0; base < ntiles0; base += 1024) {
    const int t = base + tid;
    const int64_t v = t < ntiles0 ? P.tile_cnt[t] : 0;
    const int64_t ex = scan_chunk(v);
    if (t < ntiles0) P.tile_out[t] = ex;
  }
  if (tid == 0) P.tile_out[ntiles0] = carry_s;
  __syncthreads();
  for (int b = tid; b <= P.nb; b += 1024) {
    P.out_base[b] = b < P.nb ? P.tile_out[P.buckets[b].tile_base[0]] : P.tile_out[ntiles0];
  }
}

// batch_base[b] = number of output batches before bucket b (batches never span buckets)
__global__ void __launch_bounds__(1024) batch_base_kernel(PlanDev P) {
  __shared__ int64_t wsum[32];
  __shared__ int64_t carry_s;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < P.nb; base += 1024) {
    const int b = base + tid;
    int64_t v = 0;
    if (b < P.nb) {
      const int64_t u = P.out_base[b + 1] - P.out_base[b];
      v = (u + P.batch_size - 1) / P.batch_size;
    }
    int64_t x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int64_t y = __shfl_up_sync(0xffffffffu, x, d);
      if (lane >= d) x += y;
    }
    if (lane == 31) wsum[w] = x;
    __syncthreads();
    if (w == 0) {
      int64_t s2 = wsum[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        int64_t y = __shfl_up_sync(0xffffffffu, s2, d);
        if (lane >= d) s2 += y;
      }
      wsum[lane] = s2;
    }
    __syncthreads();
    const int64_t carry = carry_s;
    if (b < P.nb) P.batch_base[b] = carry + (w == 0 ? 0 : wsum[w - 1]) + x - v;
    __syncthreads();
    if (tid == 1023) carry_s = carry + wsum[31];
    __syncthreads();
  }
  if (tid == 0) P.batch_base[P.nb] = carry_s;
}

// winners of the "no merge" path: every input row, file order
__global__ void __launch_bounds__(256) identity_kernel(PlanDev P, int F, int64_t n_in) {
  const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n_in) return;
  const int f = upper_idx(P.dense_base, F + 1, u);
  P.win[u] = (uint32_t)(P.file_row_base[f] + (u - P.dense_base[f]));
}
__global__ void __launch_bounds__(256) identity_bases_kernel(PlanDev P, int F) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > P.nb) return;
  P.out_base[b] = b < P.nb ? P.dense_base[P.buckets[b].file_begin] : P.dense_base[F];
}

__global__ void __launch_bounds__(256) compact_kernel(PlanDev P) {
  const int t = blockIdx.x;
  const int cnt = P.tile_cnt[t];
  const uint32_t* src = P.win_tmp + P.tile_gstart[t];
  uint32_t* dst = P.win + P.tile_out[t];
  const uint8_t* fsrc = P.wfile_tmp + P.tile_gstart[t];
  uint8_t* fdst = P.wfile + P.tile_out[t];
  for (int x = threadIdx.x; x < cnt; x += blockDim.x) { dst[x] = src[x]; fdst[x] = fsrc[x]; }
}

// ---------------------------------------------------------------------------------------
// gather: fixed width
// ---------------------------------------------------------------------------------------
constexpr int kGatherThreads = 256;
constexpr int kGatherFixedBlocksDefault = 3;  // 80 registers: 5.92 ms gather vs 6.11 (2 blocks, 114 regs) and 6.03 (4 blocks, spills)
constexpr int kRowsPerChunk = 2048;

struct GatherGeom {
  const uint32_t* win;
  const int64_t* out_base;
  const int64_t* batch_base;
  const uint8_t* win_file;        // file (inside its bucket) of every plan winner, or nullptr
  const BucketDesc* buckets;      // file range of every bucket (windowed string gather)
  const int64_t* file_row_base;   // first global row of every file
  int32_t nb;
  int32_t batch_size;
  int32_t chunks_per_batch;
  int64_t valid_stride;  // bytes per batch
};

__device__ __forceinline__ bool chunk_geometry(const GatherGeom& G, int64_t chunk, int& b,
                                               int64_t& gb, int64_t& u0, int& r_begin, int& r_end,
                                               int& rows_in_batch) {
  gb = chunk / G.chunks_per_batch;
  const int c_in = (int)(chunk % G.chunks_per_batch);
  b = upper_idx(G.batch_base, G.nb + 1, gb);
  // skip buckets with zero batches that share the same base
  const int64_t i = gb - G.batch_base[b];
  const int64_t U = G.out_base[b + 1] - G.out_base[b];
  const int64_t first = i * G.batch_size;
  rows_in_batch = (int)min((int64_t)G.batch_size, U - first);
  u0 = G.out_base[b] + first;
  r_begin = c_in * kRowsPerChunk;
  r_end = min(r_begin + kRowsPerChunk, (rows_in_batch + 31) & ~31);
  return r_begin < r_end;
}

// same as upper_idx over a global array, by one warp: 32-ary search, one round trip per level instead of one per bit
// (r2c profile of var_write_kernel: 9 % of the stall samples sat behind thread 0's bit-by-bit search of batch_base)
__device__ __forceinline__ int warp_upper_idx(const int64_t* __restrict__ arr, int n, int64_t q) {
  const int lane = threadIdx.x & 31;
  int lo = 0, hi = n;  // invariant arr[lo] <= q, answer in [lo, hi)
  while (hi - lo > 1) {
    const int span = hi - lo;
    const int step = (span + 31) >> 5;
    const int idx = lo + lane * step;
    const bool le = idx < hi && __ldg(arr + idx) <= q;
    const unsigned m = __ballot_sync(0xffffffffu, le);  // lanes are monotone: a prefix of ones (lane 0 always)
    const int last = 31 - __clz((int)m);
    lo = lo + last * step;
    hi = min(hi, lo + step);
  }
  return lo;
}

// chunk geometry by one warp (lanes agree on the result)
__device__ __forceinline__ bool chunk_geometry_warp(const GatherGeom& G, int64_t chunk, int& b, int64_t& gb, int64_t& u0, int& r_begin,
                                                    int& r_end, int& rows_in_batch) {
  gb = chunk / G.chunks_per_batch;
  const int c_in = (int)(chunk % G.chunks_per_batch);
  b = warp_upper_idx(G.batch_base, G.nb + 1, gb);
  const int64_t ob0 = __ldg(G.out_base + b), ob1 = __ldg(G.out_base + b + 1);
  const int64_t i = gb - __ldg(G.batch_base + b);
  const int64_t first = i * G.batch_size;
  rows_in_batch = (int)min((int64_t)G.batch_size, ob1 - ob0 - first);
  u0 = ob0 + first;
  r_begin = c_in * kRowsPerChunk;
  r_end = min(r_begin + kRowsPerChunk, (rows_in_batch + 31) & ~31);
  return r_begin < r_end;
}

constexpr int kRowsPerThread = kRowsPerChunk / kGatherThreads;  // 8

constexpr int kMaxVarCols = 16;
struct VarColSet {
  int32_t n;
  VarCol col[kMaxVarCols];
};

__device__ __forceinline__ int var_row_len(const VarCol& col, uint32_t g, bool& valid) {
  valid = g != kNoRow;
  if (valid && col.src_valid) valid = (col.src_valid[g >> 3] >> (g & 7)) & 1;
  return valid ? (col.src_offsets[(int64_t)g + 1] - col.src_offsets[g]) : 0;
}

// One block = 2048 output rows of one batch x ALL fixed-width columns.
// (Fusing pass 1 of the string gather in here was measured slower: the latency-bound string sizing
//  then runs at this kernel's 2 blocks/SM instead of 8 -- gather 7.99 ms vs 7.2 ms per bench step.)
template <int kMinBlocks>
__global__ void __launch_bounds__(kGatherThreads, kMinBlocks) gather_fixed_kernel(GatherGeom G,
                                                                                  const FixedCol* __restrict__ cols,
                                                                                  int ncols) {
  int b, r_begin, r_end, rows;
  int64_t gb, u0;
  if (!chunk_geometry(G, blockIdx.x, b, gb, u0, r_begin, r_end, rows)) return;
  // winners of this thread's rows (shared by every column unless a column brings its own list)
  uint32_t g0[kRowsPerThread];
#pragma unroll
  for (int x = 0; x < kRowsPerThread; ++x) {
    const int r = r_begin + threadIdx.x + x * kGatherThreads;
    g0[x] = (r < rows) ? G.win[u0 + r] : kNoRow;
  }
  for (int c = 0; c < ncols; ++c) {
    const FixedCol col = cols[c];
    uint32_t g[kRowsPerThread];
#pragma unroll
    for (int x = 0; x < kRowsPerThread; ++x) {
      const int r = r_begin + threadIdx.x + x * kGatherThreads;
      g[x] = col.win ? ((r < rows) ? col.win[u0 + r] : kNoRow) : g0[x];
    }
    uint32_t* vwords =
        col.dst_valid ? reinterpret_cast<uint32_t*>(col.dst_valid + gb * G.valid_stride) : nullptr;
    if (col.width == 0) {
      uint32_t* bwords = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(col.dst) + gb * G.valid_stride);
#pragma unroll
      for (int x = 0; x < kRowsPerThread; ++x) {
        const int r = r_begin + threadIdx.x + x * kGatherThreads;
        if (r - (int)(threadIdx.x & 31) >= r_end) continue;  // warp-uniform
        bool v = false;
        if (g[x] != kNoRow) v = (reinterpret_cast<const uint8_t*>(col.src)[g[x] >> 3] >> (g[x] & 7)) & 1;
        const uint32_t word = __ballot_sync(0xffffffffu, v);
        if ((threadIdx.x & 31) == 0) bwords[r >> 5] = word;
      }
    } else {
#define LSB_GATHER_CASE(T, ZERO)                                                             \
  {                                                                                          \
    T v[kRowsPerThread];                                                                     \
    _Pragma("unroll") for (int x = 0; x < kRowsPerThread; ++x)                               \
        v[x] = g[x] != kNoRow ? reinterpret_cast<const T*>(col.src)[g[x]] : ZERO;           \
    _Pragma("unroll") for (int x = 0; x < kRowsPerThread; ++x) {                             \
      const int r = r_begin + threadIdx.x + x * kGatherThreads;                              \
      if (r < rows) reinterpret_cast<T*>(col.dst)[u0 + r] = v[x];                            \
    }                                                                                        \
  }
      switch (col.width) {
        case 1: LSB_GATHER_CASE(uint8_t, (uint8_t)0) break;
        case 2: LSB_GATHER_CASE(uint16_t, (uint16_t)0) break;
        case 4: LSB_GATHER_CASE(uint32_t, 0u) break;
        case 8: LSB_GATHER_CASE(uint64_t, 0ull) break;
        case 16: LSB_GATHER_CASE(uint4, make_uint4(0, 0, 0, 0)) break;
        default: {
          const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
          for (int x = 0; x < kRowsPerThread; ++x) {
            const int r = r_begin + threadIdx.x + x * kGatherThreads;
            if (r < rows) {
              const bool have = g[x] != kNoRow;
              reinterpret_cast<uint4*>(col.dst)[2 * (u0 + r)] = have ? reinterpret_cast<const uint4*>(col.src)[2 * (int64_t)g[x]] : z;
              reinterpret_cast<uint4*>(col.dst)[2 * (u0 + r) + 1] = have ? reinterpret_cast<const uint4*>(col.src)[2 * (int64_t)g[x] + 1] : z;
            }
          }
        } break;
      }
#undef LSB_GATHER_CASE
    }
    if (vwords) {
#pragma unroll
      for (int x = 0; x < kRowsPerThread; ++x) {
        const int r = r_begin + threadIdx.x + x * kGatherThreads;
        if (r - (int)(threadIdx.x & 31) >= r_end) continue;  // warp-uniform
        bool bit = g[x] != kNoRow;
        if (bit && col.src_valid) bit = (col.src_valid[g[x] >> 3] >> (g[x] & 7)) & 1;
        const uint32_t word = __ballot_sync(0xffffffffu, bit);
        if ((threadIdx.x & 31) == 0) vwords[r >> 5] = word;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// gather: variable width (Utf8 / Binary): lengths -> chunk sums -> scan -> offsets + bytes
// ---------------------------------------------------------------------------------------

// exclusive scan in place; block b scans the segment sums[b*stride .. +nblocks] (+ total at [nblocks])
__global__ void __launch_bounds__(1024) scan_sums_inplace(int64_t* sums_all, int nblocks, int64_t stride) {
  __shared__ int64_t wsum[32];
  __shared__ int64_t carry_s;
  int64_t* sums = sums_all + blockIdx.x * stride;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < nblocks; base += 1024) {
    const int i = base + tid;
    const int64_t v = i < nblocks ? sums[i] : 0;
    int64_t x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int64_t y = __shfl_up_sync(0xffffffffu, x, d);
      if (lane >= d) x += y;
    }
    if (lane == 31) wsum[w] = x;
    __syncthreads();
    if (w == 0) {
      int64_t s = wsum[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        int64_t y = __shfl_up_sync(0xffffffffu, s, d);
        if (lane >= d) s += y;
      }
      wsum[lane] = s;
    }
    __syncthreads();
    const int64_t carry = carry_s;
    if (i < nblocks) sums[i] = carry + (w == 0 ? 0 : wsum[w - 1]) + x - v;
    __syncthreads();
    if (tid == 1023) carry_s = carry + wsum[31];
    __syncthreads();
  }
  if (tid == 0) sums[nblocks] = carry_s;
}

// Row copy global -> shared, split into a LOAD phase and a STORE phase so that every global load of a
// row (and of the next row of the same thread) is in flight before the first dependent store:
// byte head until dst is word aligned (<= 3), 8 destination words fed by 9 aligned source words through
// a funnel shift, byte tail (<= 3); rows longer than head + 32 bytes continue with further 8-word
// chunks. Reads at most the aligned word that holds the last source byte.
// (Profile r01: with loads and stores interleaved per piece the kernel sat on 3 serialized DRAM
//  latencies per row -- 54 % long-scoreboard stalls; a "G lanes per row" variant was slower still.)
struct RowCopy {
  uint8_t* dst;
  const uint8_t* src;
  int n, head, nw, tail;
  unsigned sh;
  const uint32_t* sw;
  uint8_t h0, h1, h2, t0, t1, t2;
  uint32_t w[9];
};

__device__ __forceinline__ void row_load(RowCopy& rc, uint8_t* dst, const uint8_t* src, int n) {
  rc.dst = dst;
  rc.src = src;
  rc.n = n;
  rc.head = min(n, (int)((4 - (reinterpret_cast<uintptr_t>(dst) & 3)) & 3));
  const int body = n - rc.head;
  rc.nw = body >> 2;
  rc.tail = body & 3;
  const uint8_t* bsrc = src + rc.head;
  rc.sh = (unsigned)(reinterpret_cast<uintptr_t>(bsrc) & 3) * 8;
  rc.sw = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(bsrc) & ~(uintptr_t)3);
  const uint8_t* tsrc = bsrc + rc.nw * 4;
  rc.h0 = rc.h1 = rc.h2 = rc.t0 = rc.t1 = rc.t2 = 0;
  if (rc.head > 0) rc.h0 = src[0];
  if (rc.head > 1) rc.h1 = src[1];
  if (rc.head > 2) rc.h2 = src[2];
  const int cnt = min(8, rc.nw);
#pragma unroll
  for (int i = 0; i < 9; ++i) rc.w[i] = (i < cnt || (i == cnt && rc.sh != 0 && cnt > 0)) ? rc.sw[i] : 0u;
  if (rc.tail > 0) rc.t0 = tsrc[0];
  if (rc.tail > 1) rc.t1 = tsrc[1];
  if (rc.tail > 2) rc.t2 = tsrc[2];
}

__device__ __forceinline__ void row_store(RowCopy& rc) {
  if (rc.n <= 0) return;
  uint8_t* dst = rc.dst;
  if (rc.head > 0) dst[0] = rc.h0;
  if (rc.head > 1) dst[1] = rc.h1;
  if (rc.head > 2) dst[2] = rc.h2;
  uint32_t* dw = reinterpret_cast<uint32_t*>(dst + rc.head);
  const int cnt = min(8, rc.nw);
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (i < cnt) dw[i] = __funnelshift_r(rc.w[i], rc.w[i + 1], rc.sh);
  for (int base = 8; base < rc.nw; base += 8) {  // long rows
    const int c2 = min(8, rc.nw - base);
    uint32_t w[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) w[i] = (i < c2 || (i == c2 && rc.sh != 0)) ? rc.sw[base + i] : 0u;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < c2) dw[base + i] = __funnelshift_r(w[i], w[i + 1], rc.sh);
  }
  uint8_t* td = dst + rc.head + rc.nw * 4;
  if (rc.tail > 0) td[0] = rc.t0;
  if (rc.tail > 1) td[1] = rc.t1;
  if (rc.tail > 2) td[2] = rc.t2;
}

// TMA bulk copies (cp.async.bulk; SASS UBLKCP) and the mbarrier that completes the loads
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_addr(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_addr(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_addr(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_store(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_addr(smem_src)),
               "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }


// pass 2 (grid.y = column): batch-rebased offsets + the bytes of one chunk.
// Thread t owns the 8 consecutive rows [8t, 8t+8) for the offsets; the bytes are staged through a
// shared-memory window that is address-congruent (mod 16) with the destination, so the flush to HBM
// is made of full, aligned 16-byte stores whatever the string lengths are.

// The byte position of a chunk is resolved inside the kernel with a decoupled look-back scan over the
// per-chunk totals of the column (status word = bit 63 "published" | bytes): no separate sizing pass.
// Chunks look back at lower-numbered chunks of their column, which the hardware dispatches first (blocks
// are dispatched in index order -- the assumption CUB's single-pass scan makes); the column is the
// fastest-moving index so that the chunks in flight span few chunk numbers.
struct VarScan {
  unsigned long long* agg;    // [col][chunk]: bytes of the chunk
  unsigned long long* incl;   // [col][chunk]: bytes up to and including the chunk
  int64_t stride;             // chunks per column
  int32_t ncols;
  int32_t* err;               // plan error flag (kErrScanStall when a predecessor never publishes)
};
constexpr unsigned long long kScanFlag = 1ull << 63;
constexpr int kErrScanStall = 2;
constexpr int kScanSpinLimit = 1 << 22;  // x 40+ ns: seconds, far beyond any healthy wait
__device__ __forceinline__ unsigned long long ld_scan(const unsigned long long* p) {
  return *reinterpret_cast<const volatile unsigned long long*>(p);
}
__device__ __forceinline__ void st_scan(unsigned long long* p, unsigned long long v) {
  *reinterpret_cast<volatile unsigned long long*>(p) = v;
}

// Executed by one warp: publishes `tot` for `chunk`, returns the bytes before the chunk (*c0) and the bytes
// of the chunk's batch before it (*rel; the c_in preceding chunks belong to the same batch).
__device__ __forceinline__ void scan_lookback(unsigned long long* agg, unsigned long long* incl, int64_t chunk, int c_in,
                                              int tot, long long* c0, long long* rel, int32_t* err) {
  const int lane = threadIdx.x & 31;
  if (lane == 0) st_scan(agg + chunk, kScanFlag | (unsigned long long)tot);
  unsigned long long excl = 0, relv = 0;
  int64_t idx = chunk - 1;
  int rel_left = c_in;  // totals still to be added to relv, nearest chunk first
  bool have_excl = chunk == 0;
  int spins = 0;
  while (!have_excl || rel_left > 0) {
    // 32 predecessors per round trip: both status words of each, so that one trip usually settles everything
    const int64_t my = idx - lane;
    const unsigned long long vi = my >= 0 ? ld_scan(incl + my) : 0;
    const unsigned long long va = my >= 0 ? ld_scan(agg + my) : 0;
    const unsigned m_incl = __ballot_sync(0xffffffffu, my >= 0 && (vi & kScanFlag));
    const unsigned m_agg = __ballot_sync(0xffffffffu, my < 0 || (va & kScanFlag));
    const int first_incl = (!have_excl && m_incl) ? (__ffs(m_incl) - 1) : 32;
    // totals needed from this group: lanes before the first inclusive prefix (for excl) and the batch's chunks
    const int n_excl = have_excl ? 0 : first_incl;
    const int n_need = max(n_excl, min(rel_left, 32));
    const unsigned need = n_need >= 32 ? 0xffffffffu : ((1u << n_need) - 1);
    if ((m_agg & need) != need) {  // a predecessor has not published yet
      // blocks are dispatched in index order, so the wait is short; should that ever not hold (preemption, MPS
      // time slicing), give up with an error flag instead of spinning forever
      if (++spins > kScanSpinLimit) {
        if (lane == 0) atomicExch(err, kErrScanStall);
        break;
      }
      __nanosleep(40);
      continue;
    }
    unsigned long long c = 0, r = 0;
    if (my >= 0) {
      if (!have_excl) {
        if (lane < first_incl) c = va & ~kScanFlag;
        else if (lane == first_incl) c = vi & ~kScanFlag;
      }
      if (lane < rel_left) r = va & ~kScanFlag;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      c += __shfl_xor_sync(0xffffffffu, c, d);
      r += __shfl_xor_sync(0xffffffffu, r, d);
    }
    excl += c;
    relv += r;
    rel_left = max(rel_left - 32, 0);
    idx -= 32;
    if (first_incl < 32 || idx < 0) have_excl = true;
  }
  if (lane == 0) {
    st_scan(incl + chunk, kScanFlag | (excl + (unsigned long long)tot));
    *c0 = (long long)excl;
    *rel = (long long)relv;
  }
}

// ---------------------------------------------------------------------------------------
// var_write_kernel (r2b): same chunk protocol as the r1 / r2a kernel (look-back positions, 16-byte congruent stage, TMA bulk-store
// flush), rebuilt around the instruction count per (row, column) cell -- the r2 profile of v1 showed ~350 thread
// instructions per cell at 36 % issue utilisation, a fixed cost that a 1-byte flag pays like a 40-byte comment:
//   * row pass with rows STRIDED over the block (row = x * 256 + thread): winner and offset loads of a warp fall into
//     one or two lines instead of ten, eight independent load chains per thread, validity words by ballot;
//   * chunks whose rows are all one byte long (flag columns) skip the row copies: a thread assembles one stage word
//     from four byte loads;
//   * general rows: destination-word grid -- one dst word = funnel shift of two aligned source words; only the first
//     and the last word of a row are partial (byte stores), everything between is a full STS.32; all loads of two
//     rows are in flight before the first store.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void store_word_part(uint8_t* p, uint32_t v, int lo, int hi) {  // bytes [lo, hi) of the word at p
  if (lo == 0 && hi == 4) {
    *reinterpret_cast<uint32_t*>(p) = v;
    return;
  }
  if (lo <= 0 && hi > 0) p[0] = (uint8_t)v;
  if (lo <= 1 && hi > 1) p[1] = (uint8_t)(v >> 8);
  if (lo <= 2 && hi > 2) p[2] = (uint8_t)(v >> 16);
  if (hi > 3) p[3] = (uint8_t)(v >> 24);
}

__device__ __forceinline__ void store_word_mask(uint32_t* p, uint32_t v, uint32_t m) {  // the bytes of v selected by m
  if (m == 0xffffffffu) {
    *p = v;
    return;
  }
  uint8_t* b = reinterpret_cast<uint8_t*>(p);
  if (m & 0xffu) b[0] = (uint8_t)v;
  if (m & 0xff00u) b[1] = (uint8_t)(v >> 8);
  if (m & 0xff0000u) b[2] = (uint8_t)(v >> 16);
  if (m & 0xff000000u) b[3] = (uint8_t)(v >> 24);
}

// One row on the destination-word grid of the stage. W = destination words whose source words are loaded up front
// (W + 1 aligned source words + the two under the row's last word); the first and the last word of a row are stored
// under a byte mask (neighbouring rows complete the word), the words between are plain stores. (Measured and
// dropped: shared-memory atomicOr for the partial words on a zeroed stage, two rows per thread in flight -- it
// spills at 64 registers --, and a cp.async raw-word scratch that keeps every load of a batch in flight: all within
// 2 % of this version; the kernel is bound by its instruction count and barriers, not by the copy's load latency.)
template <int W, bool kLoop>
struct RowWords {
  uint32_t* p0;         // stage word that holds the row's first byte
  int nk;               // destination words - 1 (-1: nothing to copy)
  unsigned sh;
  uint32_t m0, ml;      // byte masks of the first / last destination word
  const uint32_t* wp;   // aligned source word under the first destination word (kLoop only)
  uint32_t w[W + 1], wl0, wl1;
};

template <int W, bool kLoop>
__device__ __forceinline__ void roww_load(RowWords<W, kLoop>& rc, uint8_t* stage, int a, int n, const uint8_t* s) {
  rc.nk = -1;
  if (n <= 0) return;
  const int lo = a & 3;
  rc.p0 = reinterpret_cast<uint32_t*>(stage + (a - lo));
  rc.nk = (lo + n - 1) >> 2;
  const uintptr_t sa0 = reinterpret_cast<uintptr_t>(s) - (uintptr_t)lo;
  rc.sh = (unsigned)(sa0 & 3) * 8;
  const uint32_t* wp = reinterpret_cast<const uint32_t*>(sa0 & ~(uintptr_t)3);
  if (kLoop) rc.wp = wp;
  // source word i covers row bytes [4i - d0, 4i - d0 + 4): it is read only if it holds a byte of the row (never before
  // / beyond the column's byte buffer): i == 0 needs d0 < 4, i >= 1 needs 4i < d0 + n
  const int d0 = lo + (int)(sa0 & 3);  // 0..6
  const int lim = d0 + n;
  rc.w[0] = d0 < 4 ? __ldg(wp) : 0u;
#pragma unroll
  for (int i = 1; i <= W; ++i) rc.w[i] = (4 * i < lim) ? __ldg(wp + i) : 0u;
  rc.wl0 = rc.wl1 = 0u;
  if (rc.nk > 0) {
    rc.wl0 = __ldg(wp + rc.nk);
    if (4 * (rc.nk + 1) < lim) rc.wl1 = __ldg(wp + rc.nk + 1);
  }
  const int hi = min(4, lo + n);
  rc.m0 = (0xffffffffu << (8 * lo)) & (0xffffffffu >> (8 * (4 - hi)));
  rc.ml = 0xffffffffu >> (8 * (3 - ((lo + n - 1) & 3)));
}

template <int W, bool kLoop>
__device__ __forceinline__ void roww_store(const RowWords<W, kLoop>& rc) {
  if (rc.nk < 0) return;
  store_word_mask(rc.p0, __funnelshift_r(rc.w[0], rc.w[1], rc.sh), rc.m0);
  if (rc.nk == 0) return;
#pragma unroll
  for (int i = 1; i < W; ++i)
    if (i < rc.nk) rc.p0[i] = __funnelshift_r(rc.w[i], rc.w[i + 1], rc.sh);
  if (kLoop) {
    for (int base = W; base < rc.nk; base += 4) {  // rows longer than the preload: four full words per trip
      uint32_t w[5];
#pragma unroll
      for (int i = 0; i < 5; ++i) w[i] = (base + i <= rc.nk) ? __ldg(rc.wp + base + i) : 0u;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (base + i < rc.nk) rc.p0[base + i] = __funnelshift_r(w[i], w[i + 1], rc.sh);
    }
  }
  store_word_mask(rc.p0 + rc.nk, __funnelshift_r(rc.wl0, rc.wl1, rc.sh), rc.ml);
}

// rows [r_lo, r_hi] of the chunk, clipped to the window [w0, w1) of its bytes; R rows per thread and trip, all loads of
// a trip in flight before the first store
template <int W, int R, bool kLoop>
__device__ __forceinline__ void copy_rows(const int32_t* s_off, const int32_t* s_src, uint8_t* s_stage, const uint8_t* src_base,
                                          int r_lo, int r_hi, int w0, int w1, int first_pad) {
  for (int r = r_lo + threadIdx.x; r <= r_hi; r += R * kGatherThreads) {
    RowWords<W, kLoop> rc[R];
#pragma unroll
    for (int q = 0; q < R; ++q) {
      const int rq = r + q * kGatherThreads;
      rc[q].nk = -1;
      if (rq <= r_hi) {
        const int o = s_off[rq], a = max(o, w0), e = min(s_off[rq + 1], w1);
        roww_load<W, kLoop>(rc[q], s_stage, first_pad + (a - w0), e - a, src_base + s_src[rq] + (a - o));
      }
    }
#pragma unroll
    for (int q = 0; q < R; ++q) roww_store<W, kLoop>(rc[q]);
  }
}

// Short rows (the chunk's longest row has at most kMax bytes) of a single-window chunk, byte by byte: kMax predicated
// byte loads, then kMax predicated byte stores. For rows of a few bytes this is a third of the instructions of the
// word-grid copy (whose alignment bookkeeping is a fixed cost per row), and the byte loads of neighbouring lanes fall
// into the same sectors. (The same idea for rows of up to 20 bytes -- a 20-register byte array, or a four-byte loop --
// made the kernel slower: at 64 registers any addition to this kernel spills somewhere else.)
template <int kMax>
__device__ __forceinline__ void copy_rows_bytes(const int32_t* s_off, const int32_t* s_src, uint8_t* s_stage, const uint8_t* __restrict__ src_base,
                                                int nloc, int first_pad) {
  for (int r = threadIdx.x; r < nloc; r += kGatherThreads) {
    const int o = s_off[r];
    const int n = s_off[r + 1] - o;
    const uint8_t* sp = src_base + s_src[r];
    uint8_t* dp = s_stage + first_pad + o;
    uint8_t v[kMax];
#pragma unroll
    for (int i = 0; i < kMax; ++i)
      if (i < n) v[i] = __ldg(sp + i);
#pragma unroll
    for (int i = 0; i < kMax; ++i)
      if (i < n) dp[i] = v[i];
  }
}

constexpr int kStage2Bytes = 24 * 1024;
constexpr int kVarWriteDynSmem = kStage2Bytes + 16;

template <int kMinBlocks>
__global__ void __launch_bounds__(kGatherThreads, kMinBlocks) var_write_kernel(const __grid_constant__ GatherGeom G,
                                                                              const __grid_constant__ VarColSet set, const __grid_constant__ VarScan S,
                                                                              int64_t total_batches) {
  __shared__ int s_scan[kGatherThreads / 32 + 2];
  __shared__ __align__(16) int32_t s_off[kRowsPerChunk + 4];   // row lengths, then chunk-relative destination offsets
  __shared__ __align__(16) int32_t s_src[kRowsPerChunk];       // bucket-relative source offset of every row
  extern __shared__ __align__(16) uint8_t s_stage[];   // kStage2Bytes + 16
  __shared__ long long s_c0, s_rel;
  __shared__ int g_b, g_r_begin, g_rows, g_live;   // chunk geometry: one thread searches, everybody reads
  __shared__ long long g_gb, g_u0;
  __shared__ int s_maxlen;
  const int ci = (int)(blockIdx.x % (unsigned)S.ncols);
  const int64_t chunk = blockIdx.x / (unsigned)S.ncols;
  if (threadIdx.x < 32) {
    int b0, rb0, re0, rows0;
    int64_t gb0, u00;
    const bool lv = chunk_geometry_warp(G, chunk, b0, gb0, u00, rb0, re0, rows0);
    if (threadIdx.x == 0) {
      g_live = lv ? 1 : 0;
      g_b = b0; g_r_begin = rb0; g_rows = rows0; g_gb = gb0; g_u0 = u00;
      s_maxlen = 0;
    }
  }
  __syncthreads();
  const bool live = g_live != 0;
  const int b = g_b, r_begin = g_r_begin, rows = g_rows;
  const int64_t gb = g_gb, u0 = g_u0;
  const VarCol& col = set.col[ci];
  unsigned long long* agg = S.agg + (int64_t)ci * S.stride;
  unsigned long long* incl = S.incl + (int64_t)ci * S.stride;
  const uint32_t* win = (col.win ? col.win : G.win) + u0 + r_begin;
  const int c_in = (int)(chunk % G.chunks_per_batch);
  const uint8_t* src_base = col.src_data + col.src_data_base[b];
  const int nloc = live ? min(rows - r_begin, kRowsPerChunk) : 0;  // rows of this chunk
  const int lane = threadIdx.x & 31;
  // ---- row pass: row = x * 256 + thread ----
  uint32_t g[kRowsPerThread];
#pragma unroll
  for (int x = 0; x < kRowsPerThread; ++x) {
    const int r = x * kGatherThreads + threadIdx.x;
    g[x] = r < nloc ? __ldg(win + r) : kNoRow;
  }
  bool ones = true;
  int maxlen = 0;
  uint32_t* vwords = col.dst_valid ? reinterpret_cast<uint32_t*>(col.dst_valid + gb * G.valid_stride) + (r_begin >> 5) : nullptr;
  // three phases, each over all eight rows, so that the eight offset-pair loads are in flight together (r2c profile:
  // one exposed DRAM round trip per row when loads, shared stores and the validity store alternate)
  unsigned vmask = 0;
#pragma unroll
  for (int x = 0; x < kRowsPerThread; ++x) {
    bool valid = g[x] != kNoRow;
    if (valid && col.src_valid) valid = (__ldg(col.src_valid + (g[x] >> 3)) >> (g[x] & 7)) & 1;
    vmask |= valid ? (1u << x) : 0u;
  }
  int o0[kRowsPerThread], o1[kRowsPerThread];
#pragma unroll
  for (int x = 0; x < kRowsPerThread; ++x) {
    o0[x] = o1[x] = 0;
    if (vmask & (1u << x)) {
      o0[x] = __ldg(col.src_offsets + g[x]);
      o1[x] = __ldg(col.src_offsets + (int64_t)g[x] + 1);
    }
  }
#pragma unroll
  for (int x = 0; x < kRowsPerThread; ++x) {
    const int r = x * kGatherThreads + threadIdx.x;
    // the row's bytes are wanted two barriers from now: start them towards L2 (4.54 -> 4.38 ms gather on config 2)
    if (o1[x] > o0[x]) asm volatile("prefetch.global.L2 [%0];" ::"l"(src_base + o0[x]));
    s_src[r] = o0[x];
    s_off[r] = o1[x] - o0[x];
    ones = ones && (r >= nloc || o1[x] - o0[x] == 1);
    maxlen = max(maxlen, o1[x] - o0[x]);
    if (vwords) {
      const unsigned word = __ballot_sync(0xffffffffu, (vmask >> x) & 1u);
      if (lane == 0 && r < ((nloc + 31) & ~31)) vwords[r >> 5] = word;
    }
  }
  maxlen = __reduce_max_sync(0xffffffffu, maxlen);
  if (lane == 0 && maxlen > 0) atomicMax(&s_maxlen, maxlen);
  const bool all_one = __syncthreads_and(ones) != 0;
  const int maxlen_all = s_maxlen;
  // ---- lengths -> chunk-relative offsets (thread t scans rows [8t, 8t+8)) ----
  int tot;
  {
    int4* po = reinterpret_cast<int4*>(s_off) + 2 * threadIdx.x;
    int4 l0 = po[0], l1 = po[1];
    const int mine = l0.x + l0.y + l0.z + l0.w + l1.x + l1.y + l1.z + l1.w;
    int run = block_exclusive_scan<kGatherThreads>(mine, s_scan, tot);
    int4 e0, e1;
    e0.x = run; run += l0.x;
    e0.y = run; run += l0.y;
    e0.z = run; run += l0.z;
    e0.w = run; run += l0.w;
    e1.x = run; run += l1.x;
    e1.y = run; run += l1.y;
    e1.z = run; run += l1.z;
    e1.w = run;
    po[0] = e0;
    po[1] = e1;
    if (threadIdx.x == 0) s_off[kRowsPerChunk] = tot;
  }
  // publish the chunk total, then resolve the chunk's position by looking back
  if (threadIdx.x < 32) scan_lookback(agg, incl, chunk, c_in, tot, &s_c0, &s_rel, S.err);
  __syncthreads();
  if (!live) return;
  const int64_t c0 = s_c0;
  const int32_t rel = (int32_t)s_rel;
  // offsets (every earlier batch contributed one extra entry)
  int32_t* dst_off = col.dst_offsets + u0 + gb + r_begin;
  for (int r = threadIdx.x; r < nloc; r += kGatherThreads) dst_off[r] = rel + s_off[r];
  const bool last_chunk = r_begin + nloc == rows;
  if (threadIdx.x == 0) {
    if (last_chunk) dst_off[nloc] = rel + tot;
    if (c_in == 0) col.batch_data_start[gb] = c0 - rel;
    if (last_chunk && gb == total_batches - 1) col.batch_data_start[total_batches] = c0 + tot;
  }
  // ---- bytes: windows of the chunk's output range, staged in shared memory --------------------
  uint8_t* dst_bytes = col.dst_data + c0;       // chunk-relative byte p lives at dst_bytes[p]
  const int mis = (int)((reinterpret_cast<uintptr_t>(dst_bytes)) & 15);  // stage[mis + (p - w0)] <-> p
  int w0 = 0;
  int r_lo = 0;
  while (w0 < tot) {
    const int first_pad = (w0 == 0) ? mis : 0;   // later windows start on a 16-byte boundary
    const int w1 = min(tot, w0 + kStage2Bytes - first_pad);
    if (all_one) {
      // one byte per row (one window: tot = nloc <= 2048): stage word j = rows 4j - mis .. 4j - mis + 3
      const int jw1 = (mis + nloc + 3) >> 2;
      for (int j = (mis >> 2) + threadIdx.x; j < jw1; j += kGatherThreads) {
        const int rb = 4 * j - mis;
        uint32_t v = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int r = rb + q;
          if (r >= 0 && r < nloc) v |= (uint32_t)__ldg(src_base + s_src[r]) << (8 * q);
        }
        store_word_part(s_stage + 4 * j, v, max(-rb, 0), min(nloc - rb, 4));
      }
    } else {
      // rows intersecting [w0, w1): most chunks are one window (no search at all); the row that straddles a window
      // boundary is the first row of the next window
      const int r_hi = (w1 >= tot) ? nloc - 1 : upper_idx(s_off, nloc + 1, w1 - 1);
      // preload depth by the longest row of the chunk (block-uniform): rows up to 8 / 20 / 44 bytes never loop
      if (maxlen_all <= 8 && w1 >= tot && w0 == 0) copy_rows_bytes<8>(s_off, s_src, s_stage, src_base, nloc, first_pad);
      else if (maxlen_all <= 8) copy_rows<3, 1, false>(s_off, s_src, s_stage, src_base, r_lo, r_hi, w0, w1, first_pad);
      else if (maxlen_all <= 20) copy_rows<6, 1, false>(s_off, s_src, s_stage, src_base, r_lo, r_hi, w0, w1, first_pad);
      else copy_rows<6, 1, true>(s_off, s_src, s_stage, src_base, r_lo, r_hi, w0, w1, first_pad);
      r_lo = r_hi;
    }
    fence_proxy_async();  // the stage was written through the generic proxy; the bulk store reads it asynchronously
    __syncthreads();
    // flush: global address of stage[first_pad] is dst_bytes + w0. The 16-byte aligned middle leaves with ONE TMA
    // bulk store (cp.async.bulk shared -> global) issued by a single thread; the <= 15 bytes that share a vector with
    // the neighbouring chunks are stored by hand.
    uint8_t* gdst = dst_bytes + w0 - first_pad;  // 16-byte aligned
    const int nbytes = first_pad + (w1 - w0);
    const int v_lo = first_pad ? 16 : 0;
    const int v_hi = nbytes & ~15;
    if (threadIdx.x == 0 && v_hi > v_lo) bulk_store(gdst + v_lo, s_stage + v_lo, (uint32_t)(v_hi - v_lo));
    if (first_pad)
      for (int t = first_pad + threadIdx.x; t < min(16, nbytes); t += kGatherThreads) gdst[t] = s_stage[t];
    for (int t = max(v_hi, v_lo) + threadIdx.x; t < nbytes; t += kGatherThreads)
      if (t >= first_pad) gdst[t] = s_stage[t];
    if (threadIdx.x == 0) bulk_store_wait_read();  // the stage is rewritten by the next window / released at exit
    if (w1 >= tot) break;
    __syncthreads();
    w0 = w1;
  }
}

// ---------------------------------------------------------------------------------------
// Windowed string gather (TMA bulk copies).
// The winners of ONE file are monotone in row order, so the bytes a chunk needs from file j are one
// contiguous window [offs[first winner], offs[last winner + 1]) of the column's byte buffer -- losers in
// between included (about 1/density - 1 extra bytes). Per piece of output (<= kWinOutBytes):
//   1. every row of the piece folds its source range into its file's window (shared-memory atomics);
//   2. warp 0 lays the windows out in shared memory (16-byte congruent with their global address) and
//      issues one cp.async.bulk (global -> shared, mbarrier-completed) per window: fully coalesced DRAM
//      reads and no thread is busy with the load; windows that do not fit are read directly as before;
//   3. rows are moved shared -> shared into the output stage (address-congruent with the destination);
//   4. the stage leaves with one cp.async.bulk shared -> global (+ <= 15 head / tail bytes by hand).
// Replaces the sector-granular random DRAM reads of var_write_kernel (r01: 38 % long-scoreboard stalls).
// ---------------------------------------------------------------------------------------
constexpr int kWinDirect = INT32_MIN;    // s_wbase marker: this file's rows are read straight from HBM

constexpr int kWinMaxPieces = 64;
constexpr bool kWindowGatherDefault = false;  // flipped once the windowed kernel beats the direct one

template <int kOutLog2, int kInBytes>
struct WinSmem {
  static constexpr int kOutBytes = 1 << kOutLog2;
  int32_t off[kRowsPerChunk + 1];     // chunk-relative destination offset of every row
  int32_t src[kRowsPerChunk];         // bucket-relative source offset of every row
  uint8_t file[kRowsPerChunk];        // file (inside the bucket) of every row
  uint32_t frb[kMaxFilesPerBucket];   // first global row of the bucket's files
  int32_t wlo[kMaxFilesPerBucket];    // piece: lowest source offset per file
  int32_t whi[kMaxFilesPerBucket];    // piece: highest source end per file
  int32_t wbase[kMaxFilesPerBucket];  // source offset o of file j lives at in[wbase[j] + o]; kWinDirect = not staged
  int32_t pstart[kWinMaxPieces + 1];  // first row of every output piece
  alignas(16) uint8_t in[kInBytes];
  alignas(16) uint8_t out[kOutBytes + 16];
  alignas(8) uint64_t bar;
  int scan[kGatherThreads / 32 + 2];
  long long c0, rel;
  // chunk geometry, computed once by thread 0
  int g_b, g_r_begin, g_rows, g_live;
  long long g_gb, g_u0;
};

// n bytes src (shared window or HBM) -> dst (shared stage): byte head until dst is word aligned, words fed by
// aligned source words through a funnel shift, byte tail. Loops, not predication: most rows are a few bytes.
__device__ __forceinline__ void copy_row(uint8_t* dst, const uint8_t* src, int n) {
  int head = (int)((4 - (reinterpret_cast<uintptr_t>(dst) & 3)) & 3);
  head = min(head, n);
  for (int i = 0; i < head; ++i) dst[i] = src[i];
  dst += head;
  src += head;
  n -= head;
  const int nw = n >> 2;
  if (nw > 0) {
    const unsigned sh = (unsigned)(reinterpret_cast<uintptr_t>(src) & 3) * 8;
    const uint32_t* sw = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(src) & ~(uintptr_t)3);
    uint32_t* dw = reinterpret_cast<uint32_t*>(dst);
    if (sh == 0) {
      for (int i = 0; i < nw; ++i) dw[i] = sw[i];
    } else {
      uint32_t lo = sw[0];
      for (int i = 0; i < nw; ++i) {
        const uint32_t hi = sw[i + 1];  // holds source bytes of this row (sh != 0): always readable
        dw[i] = __funnelshift_r(lo, hi, sh);
        lo = hi;
      }
    }
  }
  for (int i = nw * 4; i < n; ++i) dst[i] = src[i];
}

template <int kOutLog2, int kInBytes>
__global__ void __launch_bounds__(kGatherThreads) var_window_kernel(GatherGeom G, const VarColSet* __restrict__ set,
                                                                    VarScan S, int64_t total_batches) {
  using Smem = WinSmem<kOutLog2, kInBytes>;
  constexpr int kOutBytes = Smem::kOutBytes;
  extern __shared__ __align__(16) unsigned char win_smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(win_smem_raw);
  const int tid = threadIdx.x;
  const int ci = (int)(blockIdx.x % (unsigned)S.ncols);
  const int64_t chunk = blockIdx.x / (unsigned)S.ncols;
  if (tid == 0) {
    int b, r_begin, r_end, rows;
    int64_t gb, u0;
    sm.g_live = chunk_geometry(G, chunk, b, gb, u0, r_begin, r_end, rows) ? 1 : 0;
    sm.g_b = b; sm.g_r_begin = r_begin; sm.g_rows = rows; sm.g_gb = gb; sm.g_u0 = u0;
    mbar_init(&sm.bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const bool live = sm.g_live != 0;
  const int b = sm.g_b, r_begin = sm.g_r_begin, rows = sm.g_rows;
  const int64_t gb = sm.g_gb, u0 = sm.g_u0;
  const VarCol col = set->col[ci];
  unsigned long long* agg = S.agg + (int64_t)ci * S.stride;
  unsigned long long* incl = S.incl + (int64_t)ci * S.stride;
  const uint32_t* win = col.win ? col.win : G.win;
  const uint8_t* win_file = (col.win == nullptr) ? G.win_file : nullptr;
  const int c_in = (int)(chunk % G.chunks_per_batch);
  const BucketDesc& B = G.buckets[b];
  const int k = B.k;
  const uint8_t* src_base = col.src_data + col.src_data_base[b];
  const int nloc = live ? min(rows - r_begin, kRowsPerChunk) : 0;  // rows of this chunk
  if (!win_file) {
    for (int j = tid; j < k; j += kGatherThreads) sm.frb[j] = (uint32_t)G.file_row_base[B.file_begin + j];
    __syncthreads();
  }
  const int rl0 = tid * kRowsPerThread;
  int len[kRowsPerThread];
  int mine = 0;
  unsigned vbits = 0;
#pragma unroll
  for (int x = 0; x < kRowsPerThread; ++x) {
    len[x] = 0;
    if (rl0 + x < nloc) {
      const int64_t u = u0 + r_begin + rl0 + x;
      const uint32_t g = win[u];
      bool valid;
      len[x] = var_row_len(col, g, valid);
      sm.src[rl0 + x] = valid ? col.src_offsets[g] : 0;
      sm.file[rl0 + x] = win_file ? win_file[u] : (valid ? (uint8_t)upper_idx(sm.frb, k, g) : (uint8_t)0);
      vbits |= valid ? (1u << x) : 0u;
    }
    mine += len[x];
  }
  if (col.dst_valid && rl0 < ((nloc + 31) & ~31))
    (col.dst_valid + gb * G.valid_stride)[(r_begin + rl0) >> 3] = (uint8_t)vbits;
  int tot;
  const int run0 = block_exclusive_scan<kGatherThreads>(mine, sm.scan, tot);
  {
    int run = run0;
#pragma unroll
    for (int x = 0; x < kRowsPerThread; ++x) {
      if (rl0 + x < nloc) {
        sm.off[rl0 + x] = run;
        run += len[x];
      }
    }
  }
  if (tid == 0) sm.off[nloc] = tot;
  if (tid < 32) scan_lookback(agg, incl, chunk, c_in, tot, &sm.c0, &sm.rel, S.err);
  __syncthreads();
  if (!live) return;
  const int64_t c0 = sm.c0;
  const int32_t rel = (int32_t)sm.rel;
  int32_t* dst_off = col.dst_offsets + u0 + gb + r_begin;
  for (int r = tid; r < nloc; r += kGatherThreads) dst_off[r] = rel + sm.off[r];
  const bool last_chunk = r_begin + nloc == rows;
  if (tid == 0) {
    if (last_chunk) dst_off[nloc] = rel + tot;
    if (c_in == 0) col.batch_data_start[gb] = c0 - rel;
    if (last_chunk && gb == total_batches - 1) col.batch_data_start[total_batches] = c0 + tot;
  }
  // ---- output pieces: piece p covers stage-aligned bytes [p << kOutLog2, (p + 1) << kOutLog2) of mis + offset ----
  uint8_t* dst_bytes = col.dst_data + c0;  // chunk-relative byte q lives at dst_bytes[q]
  const int mis = (int)((reinterpret_cast<uintptr_t>(dst_bytes)) & 15);  // out[(mis + q) & (kOutBytes - 1)] <-> q
  const int npieces = tot > 0 ? ((mis + tot - 1) >> kOutLog2) + 1 : 0;
  const bool have_table = npieces <= kWinMaxPieces;
  if (have_table) {
    // first row of every piece: the row that owns the piece's first byte (rows are blocked per thread)
    if (tid == 0) sm.pstart[0] = 0;
    int run = run0;
#pragma unroll
    for (int x = 0; x < kRowsPerThread; ++x) {
      if (len[x] > 0) {
        const int p0 = (mis + run) >> kOutLog2, p1 = (mis + run + len[x] - 1) >> kOutLog2;
        for (int pp = p0 + 1; pp <= p1; ++pp) sm.pstart[pp] = rl0 + x;
        if (((mis + run) & (kOutBytes - 1)) == 0 && p0 > 0) sm.pstart[p0] = rl0 + x;
      }
      run += len[x];
    }
  }
  uint32_t parity = 0;
  for (int p = 0; p < npieces; ++p) {
    const int w0 = p == 0 ? 0 : (p << kOutLog2) - mis;
    const int w1 = min(tot, ((p + 1) << kOutLog2) - mis);
    const int first_pad = (p == 0) ? mis : 0;  // later pieces start on a 16-byte boundary
    // 1. per-file source windows of the piece
    for (int j = tid; j < k; j += kGatherThreads) { sm.wlo[j] = INT32_MAX; sm.whi[j] = 0; }
    __syncthreads();  // (also orders pstart / the previous piece's reads of `in`)
    int r_lo, r_hi;
    if (have_table) {
      r_lo = sm.pstart[p];
      r_hi = (p + 1 < npieces) ? sm.pstart[p + 1] : nloc - 1;
    } else {
      r_lo = upper_idx(sm.off, nloc + 1, w0);
      r_hi = upper_idx(sm.off, nloc + 1, w1 - 1);
    }
    for (int r = r_lo + tid; r <= r_hi; r += kGatherThreads) {
      const int a = max(sm.off[r], w0), e = min(sm.off[r + 1], w1);
      if (e > a) {
        const int so = sm.src[r] + (a - sm.off[r]);
        const int j = sm.file[r];
        atomicMin(&sm.wlo[j], so);
        atomicMax(&sm.whi[j], so + (e - a));
      }
    }
    __syncthreads();
    // 2. layout + bulk loads (warp 0)
    if (tid < 32) {
      const int lane = tid;
      const int per = (k + 31) >> 5;
      int sizes[kMaxFilesPerBucket / 32];
      int sum = 0;
#pragma unroll
      for (int x = 0; x < kMaxFilesPerBucket / 32; ++x) {
        sizes[x] = 0;
        const int j = lane * per + x;
        if (x < per && j < k && sm.whi[j] > sm.wlo[j]) {
          const uintptr_t a0 = reinterpret_cast<uintptr_t>(src_base + sm.wlo[j]);
          const uintptr_t a1 = reinterpret_cast<uintptr_t>(src_base + sm.whi[j]);
          sizes[x] = (int)(((a1 + 15) & ~(uintptr_t)15) - (a0 & ~(uintptr_t)15));
          if (sizes[x] > kInBytes / 2) sizes[x] = -1;  // sparse winners of a big file: read its rows directly
        }
        sum += max(sizes[x], 0);
      }
      int inclv = sum;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, inclv, d);
        if (lane >= d) inclv += y;
      }
      int pos = inclv - sum;
      uint32_t my_bytes = 0;
#pragma unroll
      for (int x = 0; x < kMaxFilesPerBucket / 32; ++x) {
        const int j = lane * per + x;
        if (x < per && j < k) {
          if (sizes[x] > 0 && pos + sizes[x] <= kInBytes) {
            const int lo_al = sm.wlo[j] - (int)(reinterpret_cast<uintptr_t>(src_base + sm.wlo[j]) & 15);
            sm.wbase[j] = pos - lo_al;
            my_bytes += (uint32_t)sizes[x];
          } else {
            sm.wbase[j] = kWinDirect;
          }
          pos += max(sizes[x], 0);
        }
      }
      uint32_t total = my_bytes;
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) total += __shfl_xor_sync(0xffffffffu, total, d);
      if (lane == 0) {
        bulk_store_wait_read();  // the previous piece's bulk store has finished reading `out`
        fence_proxy_async();
        mbar_arrive_expect_tx(&sm.bar, total);
      }
      __syncwarp();
#pragma unroll
      for (int x = 0; x < kMaxFilesPerBucket / 32; ++x) {
        const int j = lane * per + x;
        if (x < per && j < k && sizes[x] > 0 && sm.wbase[j] != kWinDirect) {
          const int lo_al = sm.wlo[j] - (int)(reinterpret_cast<uintptr_t>(src_base + sm.wlo[j]) & 15);
          bulk_load(sm.in + sm.wbase[j] + lo_al, src_base + lo_al, (uint32_t)sizes[x], &sm.bar);
        }
      }
    }
    __syncthreads();
    while (!mbar_try_wait(&sm.bar, parity)) {}
    parity ^= 1;
    // 3. rows: staged window (shared) or HBM -> output stage
    for (int r = r_lo + tid; r <= r_hi; r += kGatherThreads) {
      const int a = max(sm.off[r], w0), e = min(sm.off[r + 1], w1);
      if (e > a) {
        const int so = sm.src[r] + (a - sm.off[r]);
        const int wb = sm.wbase[sm.file[r]];
        const uint8_t* sp = (wb != kWinDirect) ? (sm.in + wb + so) : (src_base + so);
        copy_row(sm.out + first_pad + (a - w0), sp, e - a);
      }
    }
    fence_proxy_async();  // the stage was written through the generic proxy, the bulk store reads it asynchronously
    __syncthreads();
    // 4. flush: global address of out[first_pad] is dst_bytes + w0
    uint8_t* gdst = dst_bytes + w0 - first_pad;  // 16-byte aligned
    const int nbytes = first_pad + (w1 - w0);
    const int v_lo = first_pad ? 16 : 0;  // bytes [0, first_pad) belong to the chunk before
    const int v_hi = nbytes & ~15;
    if (tid == 0 && v_hi > v_lo) bulk_store(gdst + v_lo, sm.out + v_lo, (uint32_t)(v_hi - v_lo));
    if (first_pad) {
      for (int t = first_pad + tid; t < min(16, nbytes); t += kGatherThreads) gdst[t] = sm.out[t];
    }
    for (int t = max(v_hi, v_lo) + tid; t < nbytes; t += kGatherThreads)
      if (t >= first_pad) gdst[t] = sm.out[t];
  }
  if (tid == 0) bulk_store_wait_read();  // shared memory must outlive the last bulk store's reads
}

// device-wide exclusive scan int32 -> int64 (used by the winner compaction), chunk of 4096 per block
constexpr int kScanChunk = 4096;
__global__ void __launch_bounds__(256) scan_block_sums(const int32_t* in, int64_t n, int64_t* sums) {
  const int64_t base = (int64_t)blockIdx.x * kScanChunk;
  int64_t s = 0;
  for (int x = threadIdx.x; x < kScanChunk; x += 256) {
    const int64_t i = base + x;
    if (i < n) s += in[i];
  }
  __shared__ int64_t ws[8];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    int64_t t = 0;
    for (int i = 0; i < 8; ++i) t += ws[i];
    sums[blockIdx.x] = t;
  }
}
__global__ void __launch_bounds__(256) scan_apply(const int32_t* in, int64_t n, const int64_t* sums,
                                                  int64_t* out) {
  __shared__ int64_t ws[8];
  const int64_t base = (int64_t)blockIdx.x * kScanChunk;
  constexpr int per = kScanChunk / 256;
  int32_t v[per];
  int64_t s = 0;
#pragma unroll
  for (int x = 0; x < per; ++x) {
    const int64_t i = base + (int64_t)threadIdx.x * per + x;
    v[x] = i < n ? in[i] : 0;
    s += v[x];
  }
  int64_t incl = s;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int64_t y = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += y;
  }
  if (lane == 31) ws[w] = incl;
  __syncthreads();
  int64_t wbase = 0;
  for (int i = 0; i < w; ++i) wbase += ws[i];
  int64_t run = sums[blockIdx.x] + wbase + incl - s;
#pragma unroll
  for (int x = 0; x < per; ++x) {
    const int64_t i = base + (int64_t)threadIdx.x * per + x;
    if (i < n) out[i] = run;
    run += v[x];
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 255) out[n] = sums[gridDim.x];
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
struct MergePlan {
  int nb = 0, F = 0;
  int batch_size = 8192;
  bool key128 = false;
  bool key_bytes = false;  // last merge used a Utf8/Binary key
  int log_s = 6, tile_cap = 0, max_level = 0;
  int64_t n_total = 0;  // global row span (>= input rows when files are padded apart)
  int64_t n_in = 0;     // input rows
  std::vector<BucketDesc> buckets;
  std::vector<int32_t> level_tile_ofs, level_tile_cnt;  // per level
  int32_t total_tiles = 0;
  // device
  BucketDesc* d_buckets = nullptr;
  int32_t* d_file_rows = nullptr;
  int64_t* d_file_row_base = nullptr;
  int32_t* d_splits = nullptr;
  uint2* d_mlist = nullptr;
  int32_t *d_tile_bucket = nullptr, *d_tile_idx = nullptr, *d_tile_cnt = nullptr;
  int64_t *d_tile_gstart = nullptr, *d_tile_out = nullptr;
  uint32_t *d_win_tmp = nullptr, *d_win = nullptr;
  uint8_t *d_wfile_tmp = nullptr, *d_wfile = nullptr;
  uint8_t* d_tile_flag = nullptr;  // per tile (all levels): 1 = the fast tile kernel left it to the generic one
  bool use_fast_tiles = false;
  bool has_wfile = false;  // d_wfile matches d_win (merge path; the concat path leaves it unset)
  int64_t *d_out_base = nullptr, *d_batch_base = nullptr;
  int64_t* d_dense_base = nullptr;
  uint32_t* d_R = nullptr;     // general path (owned by GeneralMerge, borrowed here while it runs)
  int64_t* d_rofs = nullptr;
  uint32_t* d_ctmp = nullptr;  // compaction scratch
  PkDesc last_pk{};
  int32_t* d_err = nullptr;
  // pinned host mirrors
  int64_t* h_counts = nullptr;  // out_base (nb+1) + batch_base (nb+1) + err (thread-local pinned block)
  cudaStream_t alloc_stream = nullptr;
  std::vector<int64_t> h_frb, h_dense;
  std::vector<int32_t> h_fr32, h_tb, h_ti;
  std::vector<int64_t> out_counts, out_base, batch_base;
  bool merged = false, synced = false;
  MergeStats stats;
  // gather scratch
  int32_t* d_len = nullptr;
  int64_t* d_doff = nullptr;
  int64_t* d_sums = nullptr;
  FixedCol* d_cols = nullptr;
  int d_cols_cap = 0;
  std::vector<void*> var_scratch;        // look-back state + descriptor blocks of the string gathers
  std::deque<VarColSet> h_varsets;       // host copies outlive the asynchronous descriptor uploads
  size_t smem_bytes = 0;
};

static size_t tile_smem_bytes(int C, bool key128) {
  size_t key = key128 ? 16 : 8;
  return (size_t)C * key + (size_t)C * 2 * 2 + kMaxFilesPerBucket * 8 +
         (kMaxFilesPerBucket + 1) * 4 + kMaxFilesPerBucket * 4 + 16 * 4;
}

// stream-ordered pool allocations: plans are created per scan batch, cudaMalloc/cudaFree would
// serialise the device every time
static thread_local DeviceArena* t_arena = nullptr;
void arena_install(DeviceArena* a) { t_arena = a; }
DeviceArena* arena_current() { return t_arena; }

DeviceArena::~DeviceArena() {
  for (auto& b : blocks_) cudaFree(b.p);
}
size_t DeviceArena::capacity() const {
  size_t c = 0;
  for (auto& b : blocks_) c += b.cap;
  return c;
}
void* DeviceArena::alloc(size_t bytes) {
  bytes = (std::max<size_t>(bytes, 1) + 255) & ~(size_t)255;
  while (true) {
    if (cur_ < blocks_.size() && off_ + bytes <= blocks_[cur_].cap) {
      void* p = blocks_[cur_].p + off_;
      off_ += bytes;
      used_ += bytes;
      return p;
    }
    if (cur_ + 1 < blocks_.size()) {  // later block left over from an earlier reader
      ++cur_;
      off_ = 0;
      continue;
    }
    size_t cap = std::max<size_t>(bytes, (size_t)64 << 20);
    if (!blocks_.empty()) cap = std::max(cap, blocks_.back().cap * 2);
    uint8_t* p = nullptr;
    LSB_CUDA(cudaMalloc(reinterpret_cast<void**>(&p), cap));
    blocks_.push_back({p, cap});
    cur_ = blocks_.size() - 1;
    off_ = 0;
  }
}
void DeviceArena::reset() {
  high_ = std::max(high_, used_);
  const size_t cap = capacity();
  // several blocks = the arena had to grow: fold them into one block that fits the largest reader so far
  bool rebuild = blocks_.size() > 1;
  size_t want = high_ + high_ / 4;
  // one huge bucket must not pin its footprint forever: after 8 readers in a row that used less than a quarter
  // of a large arena, shrink it to twice the recent need
  if (cap > ((size_t)256 << 20) && used_ * 4 < cap) {
    recent_ = std::max(recent_, used_);
    if (++idle_ >= 8) {
      rebuild = true;
      want = 2 * recent_;
      high_ = recent_;
    }
  } else {
    idle_ = 0;
    recent_ = 0;
  }
  if (rebuild) {
    for (auto& b : blocks_) cudaFree(b.p);
    blocks_.clear();
    uint8_t* p = nullptr;
    const size_t bytes = std::max<size_t>((want + ((size_t)1 << 20)) & ~(((size_t)1 << 20) - 1), (size_t)64 << 20);
    if (cudaMalloc(reinterpret_cast<void**>(&p), bytes) == cudaSuccess) blocks_.push_back({p, bytes});
    else cudaGetLastError();
    idle_ = 0;
    recent_ = 0;
  }
  cur_ = 0;
  off_ = 0;
  used_ = 0;
}

static thread_local HostStage* t_stage = nullptr;
void host_stage_install(HostStage* s) { t_stage = s; }
HostStage* host_stage_current() { return t_stage; }
HostStage::~HostStage() {
  for (auto& b : blocks_) cudaFreeHost(b.p);
}
void* HostStage::alloc(size_t bytes) {
  bytes = (std::max<size_t>(bytes, 1) + 63) & ~(size_t)63;
  while (true) {
    if (cur_ < blocks_.size() && off_ + bytes <= blocks_[cur_].cap) {
      void* p = blocks_[cur_].p + off_;
      off_ += bytes;
      return p;
    }
    if (cur_ + 1 < blocks_.size()) { ++cur_; off_ = 0; continue; }
    size_t cap = std::max<size_t>(bytes, (size_t)1 << 20);
    if (!blocks_.empty()) cap = std::max(cap, blocks_.back().cap * 2);
    uint8_t* p = nullptr;
    LSB_CUDA(cudaMallocHost(reinterpret_cast<void**>(&p), cap));
    blocks_.push_back({p, cap});
    cur_ = blocks_.size() - 1;
    off_ = 0;
  }
}
void HostStage::reset() { cur_ = 0; off_ = 0; }
cudaError_t h2d_staged(void* dst, const void* src, size_t bytes, cudaStream_t stream) {
  if (bytes == 0) return cudaSuccess;
  if (t_stage) {
    void* p = t_stage->alloc(bytes);
    std::memcpy(p, src, bytes);
    src = p;
  }
  return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream);
}

void* device_alloc_bytes(size_t bytes, cudaStream_t stream) {
  if (t_arena) return t_arena->alloc(bytes);
  void* p = nullptr;
  LSB_CUDA(cudaMallocAsync(&p, std::max<size_t>(bytes, 16), stream));
  return p;
}
void device_free_bytes(void* p, cudaStream_t stream) {
  if (p && !t_arena) cudaFreeAsync(p, stream);
}

template <typename T>
static T* dev_alloc(size_t n, cudaStream_t stream) {
  return static_cast<T*>(device_alloc_bytes(std::max<size_t>(n, 1) * sizeof(T), stream));
}
static void dev_free(void* p, cudaStream_t stream) { device_free_bytes(p, stream); }

// small pinned staging blocks for the per-plan count read-back, recycled per thread
struct PinnedCounts {
  int64_t* p = nullptr;
  size_t cap = 0;
  ~PinnedCounts() { if (p) cudaFreeHost(p); }
};
static int64_t* pinned_counts(size_t n) {
  static thread_local PinnedCounts pc;
  if (pc.cap < n) {
    if (pc.p) cudaFreeHost(pc.p);
    pc.cap = std::max<size_t>(n * 2, 1024);
    LSB_CUDA(cudaMallocHost(&pc.p, pc.cap * sizeof(int64_t)));
  }
  return pc.p;
}

MergePlan* plan_create(int nb, const int32_t* bucket_file_begin, const int64_t* file_rows,
                       const int64_t* file_row_base, int64_t row_span, int key_bits, int batch_size,
                       cudaStream_t stream) {
  if (key_bits <= 0 || key_bits > 128) {
    g_last_error = "primary key wider than 128 bits is not supported";
    throw std::runtime_error(g_last_error);
  }
  if (batch_size <= 0) {
    g_last_error = "batch_size must be positive";
    throw std::runtime_error(g_last_error);
  }
  auto* P = new MergePlan();
  try {
    P->alloc_stream = stream;
    P->nb = nb;
    P->F = bucket_file_begin[nb];
    P->batch_size = batch_size;
    P->key128 = key_bits > 64;
    // 3 CTAs per SM: 3 * smem <= 227 KiB
    P->tile_cap = P->key128 ? 3328 : 5632;
    const int C = P->tile_cap;
    int kmax = 1;
    for (int b = 0; b < nb; ++b) kmax = std::max(kmax, bucket_file_begin[b + 1] - bucket_file_begin[b]);
    if (kmax > kMaxFilesPerBucket) {
      g_last_error = "more than 256 delta files in one hash bucket: compact the table first";
      throw std::runtime_error(g_last_error);
    }
    // S = largest power of two <= 64 that keeps the per-tile slack k*(S-1) under C/5
    int log_s = 6;
    while (log_s > 3 && (int64_t)kmax * ((1 << log_s) - 1) > C / 5) --log_s;
    P->log_s = log_s;
    const int S = 1 << log_s;
    if ((int64_t)kmax * (S - 1) + S > C) {
      g_last_error = "too many delta files per bucket for the tile capacity";
      throw std::runtime_error(g_last_error);
    }

    std::vector<int64_t>& frb = P->h_frb;
    std::vector<int32_t>& fr32 = P->h_fr32;
    frb.assign(P->F + 1, 0);
    fr32.assign(P->F, 0);
    for (int f = 0; f < P->F; ++f) {
      if (file_rows[f] < 0 || file_rows[f] > 0x7fffffff) {
        g_last_error = "file with more than 2^31-1 rows";
        throw std::runtime_error(g_last_error);
      }
      fr32[f] = (int32_t)file_rows[f];
      if (file_row_base) {
        frb[f] = file_row_base[f];
        if (f > 0 && frb[f] < frb[f - 1] + file_rows[f - 1]) {
          g_last_error = "file_row_base overlaps the previous file";
          throw std::runtime_error(g_last_error);
        }
      } else {
        frb[f + 1] = frb[f] + file_rows[f];
      }
      P->n_in += file_rows[f];
    }
    if (file_row_base) frb[P->F] = row_span;
    P->n_total = frb[P->F];
    if (P->F > 0 && P->n_total < frb[P->F - 1] + file_rows[P->F - 1]) {
      g_last_error = "row_span smaller than the last file";
      throw std::runtime_error(g_last_error);
    }
    if (P->n_total >= 0xfffffff0ll) {
      g_last_error = "scan batch exceeds 2^32 rows: split the bucket list";
      throw std::runtime_error(g_last_error);
    }

    P->buckets.resize(nb);
    int64_t split_total = 0, mlist_total = 0;
    int max_level = 0;
    for (int b = 0; b < nb; ++b) {
      BucketDesc& B = P->buckets[b];
      std::memset(&B, 0, sizeof(B));
      B.file_begin = bucket_file_begin[b];
      B.k = bucket_file_begin[b + 1] - bucket_file_begin[b];
      B.row_base = B.k > 0 ? frb[B.file_begin] : (B.file_begin < P->F ? frb[B.file_begin] : P->n_total);
      B.n_rows = 0;
      for (int j = 0; j < B.k; ++j) B.n_rows += file_rows[B.file_begin + j];
      B.m = (C - B.k * (S - 1)) / S;
      // level sizes
      int64_t M[kMaxLevels + 2];
      int L = 0;
      for (int l = 0; l <= kMaxLevels + 1; ++l) {
        int64_t s = 0;
        for (int j = 0; j < B.k; ++j) s += file_rows[B.file_begin + j] >> (l * log_s);
        M[l] = s;
      }
      while (M[L] > C) {
        ++L;
        if (L > kMaxLevels) {
          g_last_error = "bucket too large for the sampling depth";
          throw std::runtime_error(g_last_error);
        }
      }
      B.levels = L;
      max_level = std::max(max_level, L);
      for (int l = 0; l <= L; ++l) {
        B.ntiles[l] = l == L ? 1 : (int32_t)(M[l + 1] / B.m + 1);
        B.split_ofs[l] = split_total;
        split_total += (int64_t)(B.ntiles[l] + 1) * std::max(B.k, 1);
        if (l >= 1) {
          B.mlist_ofs[l] = mlist_total;
          mlist_total += M[l];
        }
      }
    }
    P->max_level = max_level;
    // tiles per level, ordered by bucket then tile index
    P->level_tile_ofs.assign(max_level + 1, 0);
    P->level_tile_cnt.assign(max_level + 1, 0);
    std::vector<int32_t>& tb = P->h_tb;
    std::vector<int32_t>& ti = P->h_ti;
    for (int l = 0; l <= max_level; ++l) {
      P->level_tile_ofs[l] = (int32_t)tb.size();
      for (int b = 0; b < nb; ++b) {
        BucketDesc& B = P->buckets[b];
        if (B.levels < l) continue;
        B.tile_base[l] = (int32_t)tb.size() - P->level_tile_ofs[l];
        for (int i = 0; i < B.ntiles[l]; ++i) {
          tb.push_back(b);
          ti.push_back(i);
        }
      }
      P->level_tile_cnt[l] = (int32_t)tb.size() - P->level_tile_ofs[l];
    }
    P->total_tiles = (int32_t)tb.size();
    const int T0 = P->level_tile_cnt[0];

    P->d_buckets = dev_alloc<BucketDesc>(nb, stream);
    P->d_file_rows = dev_alloc<int32_t>(P->F, stream);
    P->d_file_row_base = dev_alloc<int64_t>(P->F + 1, stream);
    P->d_splits = dev_alloc<int32_t>(split_total, stream);
    P->d_mlist = dev_alloc<uint2>(mlist_total, stream);
    P->d_tile_bucket = dev_alloc<int32_t>(tb.size(), stream);
    P->d_tile_idx = dev_alloc<int32_t>(ti.size(), stream);
    P->d_tile_cnt = dev_alloc<int32_t>(T0, stream);
    P->d_tile_gstart = dev_alloc<int64_t>(T0, stream);
    P->d_tile_out = dev_alloc<int64_t>(T0 + 1, stream);
    P->d_win_tmp = dev_alloc<uint32_t>(P->n_total, stream);
    P->d_win = dev_alloc<uint32_t>(P->n_total, stream);
    P->d_wfile_tmp = dev_alloc<uint8_t>(P->n_total, stream);
    P->d_wfile = dev_alloc<uint8_t>(P->n_total, stream);
    P->d_tile_flag = dev_alloc<uint8_t>(tb.size(), stream);
    LSB_CUDA(cudaMemsetAsync(P->d_tile_flag, 0, std::max<size_t>(tb.size(), 1), stream));
    {
      static const bool off = [] { const char* v = getenv("LAKESOUL_B200_FAST_TILES"); return v && *v == '0'; }();
      P->use_fast_tiles = P->key128 && !off;
    }
    P->d_out_base = dev_alloc<int64_t>(nb + 1, stream);
    P->d_batch_base = dev_alloc<int64_t>(nb + 1, stream);
    P->d_err = dev_alloc<int32_t>(1, stream);
    P->d_dense_base = dev_alloc<int64_t>(P->F + 1, stream);
    std::vector<int64_t>& dense = P->h_dense;
    dense.assign(P->F + 1, 0);
    for (int f = 0; f < P->F; ++f) dense[f + 1] = dense[f] + file_rows[f];
    LSB_CUDA(h2d_staged(P->d_dense_base, dense.data(), sizeof(int64_t) * (P->F + 1), stream));
    LSB_CUDA(h2d_staged(P->d_buckets, P->buckets.data(), sizeof(BucketDesc) * nb, stream));
    LSB_CUDA(h2d_staged(P->d_file_rows, fr32.data(), sizeof(int32_t) * P->F, stream));
    LSB_CUDA(h2d_staged(P->d_file_row_base, frb.data(), sizeof(int64_t) * (P->F + 1), stream));
    LSB_CUDA(h2d_staged(P->d_tile_bucket, tb.data(), sizeof(int32_t) * tb.size(), stream));
    LSB_CUDA(h2d_staged(P->d_tile_idx, ti.data(), sizeof(int32_t) * ti.size(), stream));
    LSB_CUDA(cudaMemsetAsync(P->d_err, 0, sizeof(int32_t), stream));
    // the host vectors are plan members: they outlive the asynchronous copies

    P->smem_bytes = tile_smem_bytes(C, P->key128);
    static thread_local int attr_device = -1;
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);
    if (attr_device != cur_dev) {
    attr_device = cur_dev;
    LSB_CUDA(cudaFuncSetAttribute(tile_kernel<uint64_t, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)tile_smem_bytes(5632, false)));
    LSB_CUDA(cudaFuncSetAttribute(tile_kernel<K128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)tile_smem_bytes(3328, true)));
    LSB_CUDA(cudaFuncSetAttribute(tile_kernel<uint64_t, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)tile_smem_bytes(5632, false)));
    LSB_CUDA(cudaFuncSetAttribute(tile_kernel<K128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)tile_smem_bytes(3328, true)));
    LSB_CUDA(cudaFuncSetAttribute(tile_kernel<StrKey, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)tile_smem_bytes(3328, true)));
    LSB_CUDA(cudaFuncSetAttribute(tile_kernel<StrKey, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)tile_smem_bytes(3328, true)));
    LSB_CUDA(cudaFuncSetAttribute(tile_fast_kernel<K128, kTileThreads, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)tile_fast_smem_bytes(3328)));
    LSB_CUDA(cudaFuncSetAttribute(tile_fast_kernel<K128, kTileThreads, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)tile_fast_smem_bytes(3328)));
    }
    P->stats.n_rows = P->n_in;
    P->stats.tiles_l0 = T0;
    P->stats.levels = max_level;
    P->stats.log_s = log_s;
    P->stats.tile_cap = C;
  } catch (...) {
    plan_destroy(P);
    throw;
  }
  return P;
}

void plan_destroy(MergePlan* P) {
  if (!P) return;
  cudaStream_t st = P->alloc_stream;
  void* ptrs[] = {P->d_buckets, P->d_file_rows, P->d_file_row_base, P->d_splits, P->d_mlist, P->d_tile_bucket,
                  P->d_tile_idx, P->d_tile_cnt, P->d_tile_gstart, P->d_tile_out, P->d_win_tmp, P->d_win, P->d_wfile_tmp, P->d_wfile, P->d_tile_flag,
                  P->d_out_base, P->d_batch_base, P->d_dense_base, P->d_ctmp, P->d_err, P->d_len, P->d_doff,
                  P->d_sums, P->d_cols};
  for (void* p : ptrs) dev_free(p, st);
  for (void* p : P->var_scratch) dev_free(p, st);
  delete P;
}

static PlanDev plan_dev(const MergePlan* P) {
  PlanDev D{};
  D.buckets = P->d_buckets;
  D.file_rows = P->d_file_rows;
  D.file_row_base = P->d_file_row_base;
  D.splits = P->d_splits;
  D.mlist = P->d_mlist;
  D.tile_bucket = P->d_tile_bucket;
  D.tile_idx = P->d_tile_idx;
  D.tile_cnt = P->d_tile_cnt;
  D.tile_gstart = P->d_tile_gstart;
  D.tile_out = P->d_tile_out;
  D.win_tmp = P->d_win_tmp;
  D.win = P->d_win;
  D.wfile_tmp = P->d_wfile_tmp;
  D.wfile = P->d_wfile;
  D.out_base = P->d_out_base;
  D.batch_base = P->d_batch_base;
  D.dense_base = P->d_dense_base;
  D.R = P->d_R;
  D.rofs = P->d_rofs;
  D.err = P->d_err;
  D.nb = P->nb;
  D.log_s = P->log_s;
  D.tile_cap = P->tile_cap;
  D.batch_size = P->batch_size;
  return D;
}

template <typename KeyT>
static int run_levels(MergePlan* P, const PkDesc& pk, cudaStream_t stream) {
  PlanDev D = plan_dev(P);
  int launches = 0;
  for (int l = P->max_level; l >= 0; --l) {
    const int nt = P->level_tile_cnt[l];
    if (nt == 0) continue;
    const int ofs = P->level_tile_ofs[l];
    const int warps_per_block = 4;
    part_kernel<KeyT><<<(nt + warps_per_block - 1) / warps_per_block, 128, 0, stream>>>(D, pk, l, ofs, nt);
    if constexpr (std::is_same<KeyT, K128>::value) {
      if (P->use_fast_tiles) {
        // (256 threads per tile -- twice the elements per thread, half the co-rank searches -- measured 2.38 vs 2.09 ms
        //  merge: the kernel needs the 48 warps per SM more than it needs the amortisation)
        const bool shape1 = pk.ncols == 2 && pk.col[0].width == 8 && pk.col[1].width == 4 && !pk.col[0].validity && !pk.col[1].validity &&
                            pk.col[0].kind != PK_FLOAT && pk.col[1].kind != PK_FLOAT;
        if (shape1)
          tile_fast_kernel<KeyT, kTileThreads, 1><<<nt, kTileThreads, tile_fast_smem_bytes(P->tile_cap), stream>>>(D, pk, l, ofs, P->d_tile_flag);
        else
          tile_fast_kernel<KeyT, kTileThreads, 0><<<nt, kTileThreads, tile_fast_smem_bytes(P->tile_cap), stream>>>(D, pk, l, ofs, P->d_tile_flag);
        tile_kernel<KeyT, false><<<nt, kTileThreads, P->smem_bytes, stream>>>(D, pk, l, ofs, P->d_tile_flag);
        launches += 3;
        continue;
      }
    }
    tile_kernel<KeyT, false><<<nt, kTileThreads, P->smem_bytes, stream>>>(D, pk, l, ofs, nullptr);
    launches += 2;
  }
  return launches;
}

void plan_merge(MergePlan* P, const PkDesc& pk, cudaStream_t stream) {
  const bool want128 = pk.key_bits > 64;
  if (want128 != P->key128) {
    g_last_error = "plan was created for a different key width";
    throw std::runtime_error(g_last_error);
  }
  int launches = 0;
  if (P->n_total > 0) {
    P->key_bytes = pk.ncols == 1 && pk.col[0].kind == PK_BYTES;
    if (P->key_bytes) launches += run_levels<StrKey>(P, pk, stream);
    else launches += P->key128 ? run_levels<K128>(P, pk, stream) : run_levels<uint64_t>(P, pk, stream);
  }
  PlanDev D = plan_dev(P);
  const int T0 = P->level_tile_cnt[0];
  scan_kernel<<<1, 1024, 0, stream>>>(D, T0);
  batch_base_kernel<<<1, 1024, 0, stream>>>(D);
  launches += 2;
  if (T0 > 0) {
    compact_kernel<<<T0, 256, 0, stream>>>(D);
    ++launches;
  }
  LSB_CUDA(cudaGetLastError());
  P->stats.launches = launches;
  P->last_pk = pk;
  P->merged = true;
  P->synced = false;
  P->has_wfile = true;
}

void plan_identity(MergePlan* P, cudaStream_t stream) {
  PlanDev D = plan_dev(P);
  int launches = 0;
  if (P->n_in > 0) {
    identity_kernel<<<(unsigned)((P->n_in + 255) / 256), 256, 0, stream>>>(D, P->F, P->n_in);
    ++launches;
  }
  identity_bases_kernel<<<(P->nb + 1 + 255) / 256, 256, 0, stream>>>(D, P->F);
  batch_base_kernel<<<1, 1024, 0, stream>>>(D);
  launches += 2;
  LSB_CUDA(cudaGetLastError());
  P->stats.launches = launches;
  P->merged = true;
  P->synced = false;
  P->has_wfile = false;
}

const std::vector<int64_t>& plan_sync_counts(MergePlan* P, cudaStream_t stream) {
  if (!P->merged) {
    g_last_error = "plan_sync_counts before plan_merge";
    throw std::runtime_error(g_last_error);
  }
  if (!P->synced) {
    const int nb = P->nb;
    P->h_counts = pinned_counts(2 * (size_t)(nb + 1) + 1);
    LSB_CUDA(cudaMemcpyAsync(P->h_counts, P->d_out_base, sizeof(int64_t) * (nb + 1),
                             cudaMemcpyDeviceToHost, stream));
    LSB_CUDA(cudaMemcpyAsync(P->h_counts + nb + 1, P->d_batch_base, sizeof(int64_t) * (nb + 1),
                             cudaMemcpyDeviceToHost, stream));
    LSB_CUDA(cudaMemcpyAsync(P->h_counts + 2 * (nb + 1), P->d_err, sizeof(int32_t),
                             cudaMemcpyDeviceToHost, stream));
    LSB_CUDA(stream_wait(stream));
    const int32_t err = *reinterpret_cast<int32_t*>(P->h_counts + 2 * (nb + 1));
    if (err != 0) {
      g_last_error = "merge kernel reported error code " + std::to_string(err);
      throw std::runtime_error(g_last_error);
    }
    P->out_base.assign(P->h_counts, P->h_counts + nb + 1);
    P->batch_base.assign(P->h_counts + nb + 1, P->h_counts + 2 * (nb + 1));
    P->out_counts.resize(nb);
    for (int b = 0; b < nb; ++b) P->out_counts[b] = P->out_base[b + 1] - P->out_base[b];
    P->stats.n_out = P->out_base[nb];
    P->synced = true;
  }
  return P->out_counts;
}

const MergeStats& plan_stats(const MergePlan* P) { return P->stats; }
int64_t plan_total_rows(const MergePlan* P) { return P->n_in; }
int plan_num_buckets(const MergePlan* P) { return P->nb; }
int plan_batch_size(const MergePlan* P) { return P->batch_size; }
const uint32_t* plan_winners(const MergePlan* P) { return P->d_win; }
const int64_t* plan_out_base_dev(const MergePlan* P) { return P->d_out_base; }
const int32_t* plan_error_flag_dev(const MergePlan* P) { return P->d_err; }
int64_t plan_valid_stride_bytes(const MergePlan* P) {
  return (((int64_t)P->batch_size + 7) / 8 + 63) / 64 * 64;
}
const std::vector<int64_t>& plan_out_base(const MergePlan* P) { return P->out_base; }
const std::vector<int64_t>& plan_batch_base(const MergePlan* P) { return P->batch_base; }

static GatherGeom make_geom(const MergePlan* P) {
  GatherGeom G{};
  G.win = P->d_win;
  G.out_base = P->d_out_base;
  G.batch_base = P->d_batch_base;
  G.win_file = P->has_wfile ? P->d_wfile : nullptr;
  G.buckets = P->d_buckets;
  G.file_row_base = P->d_file_row_base;
  G.nb = P->nb;
  G.batch_size = P->batch_size;
  G.chunks_per_batch = (P->batch_size + kRowsPerChunk - 1) / kRowsPerChunk;
  G.valid_stride = plan_valid_stride_bytes(P);
  return G;
}

static void require_synced(const MergePlan* P) {
  if (!P->synced) {
    g_last_error = "gather before plan_sync_counts";
    throw std::runtime_error(g_last_error);
  }
}

static void gather_fixed_on(MergePlan* P, const FixedCol* cols, int ncols, cudaStream_t stream, cudaStream_t launch_stream,
                            cudaEvent_t fork);
void gather_fixed(MergePlan* P, const FixedCol* cols, int ncols, cudaStream_t stream) {
  gather_fixed_on(P, cols, ncols, stream, stream, nullptr);
}

// Descriptor upload on `stream`; the kernel on `launch_stream` (behind `fork`, recorded on `stream`, when they differ).
static void gather_fixed_on(MergePlan* P, const FixedCol* cols, int ncols, cudaStream_t stream, cudaStream_t launch_stream,
                            cudaEvent_t fork) {
  require_synced(P);
  if (ncols <= 0) return;
  const int64_t total_batches = P->batch_base[P->nb];
  if (total_batches == 0) return;
  if (P->d_cols_cap < ncols) {
    dev_free(P->d_cols, stream);
    P->d_cols = dev_alloc<FixedCol>(ncols, stream);
    P->d_cols_cap = ncols;
  }
  LSB_CUDA(h2d_staged(P->d_cols, cols, sizeof(FixedCol) * ncols, stream));
  GatherGeom G = make_geom(P);
  const int64_t chunks = total_batches * G.chunks_per_batch;
  // occupancy variants: r1 profile 90 registers -> 2 blocks per SM, latency-bound (24.8 % warps active)
  static const int min_blocks = [] { const char* v = getenv("LAKESOUL_B200_GATHER_BLOCKS"); return v ? atoi(v) : kGatherFixedBlocksDefault; }();
  if (launch_stream != stream) {
    LSB_CUDA(cudaEventRecord(fork, stream));
    LSB_CUDA(cudaStreamWaitEvent(launch_stream, fork, 0));
  }
  if (min_blocks >= 4) gather_fixed_kernel<4><<<(unsigned)chunks, kGatherThreads, 0, launch_stream>>>(G, P->d_cols, ncols);
  else if (min_blocks == 3) gather_fixed_kernel<3><<<(unsigned)chunks, kGatherThreads, 0, launch_stream>>>(G, P->d_cols, ncols);
  else gather_fixed_kernel<2><<<(unsigned)chunks, kGatherThreads, 0, launch_stream>>>(G, P->d_cols, ncols);
  LSB_CUDA(cudaGetLastError());
  // the descriptor table is reused by the next call: order the copy behind this launch
}

void gather_var_multi(MergePlan* P, const VarCol* cols, int ncols, cudaStream_t stream);

// Fixed-width and variable-width columns of one scan: 2 launches.
// The fixed-width gather is HBM-bound, the string gather instruction / latency-bound: they run side by side, the
// fixed-width kernel on a side stream of the calling thread (fork / join by events; LAKESOUL_B200_GATHER_OVERLAP=0
// puts them back in sequence).
struct SideStream {
  int device = -1;
  cudaStream_t stream = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
};
static SideStream& side_stream() {
  static thread_local SideStream ss;
  int cur = 0;
  LSB_CUDA(cudaGetDevice(&cur));
  if (ss.device != cur) {  // streams of a previous device are left to the context
    LSB_CUDA(cudaStreamCreateWithFlags(&ss.stream, cudaStreamNonBlocking));
    LSB_CUDA(cudaEventCreateWithFlags(&ss.fork, cudaEventDisableTiming));
    LSB_CUDA(cudaEventCreateWithFlags(&ss.join, cudaEventDisableTiming));
    ss.device = cur;
  }
  return ss;
}

void gather_all(MergePlan* P, const FixedCol* fixed, int nfixed, const VarCol* vars, int nvars, cudaStream_t stream) {
  static const bool overlap = [] { const char* v = getenv("LAKESOUL_B200_GATHER_OVERLAP"); return !v || *v != '0'; }();
  // (small scans -- one bucket of a reader -- stay on one stream: many readers already share the hardware queues)
  const bool both = nfixed > 0 && nvars > 0 && P->synced && P->batch_base[P->nb] > 0 && P->out_base[P->nb] >= (1 << 20);
  if (both && overlap) {
    SideStream& ss = side_stream();
    gather_fixed_on(P, fixed, nfixed, stream, ss.stream, ss.fork);
    gather_var_multi(P, vars, nvars, stream);
    LSB_CUDA(cudaEventRecord(ss.join, ss.stream));
    LSB_CUDA(cudaStreamWaitEvent(stream, ss.join, 0));
    return;
  }
  if (nfixed > 0) gather_fixed(P, fixed, nfixed, stream);
  if (nvars > 0) gather_var_multi(P, vars, nvars, stream);
}

void plan_reserve_scratch(MergePlan* P, cudaStream_t stream) {
  require_synced(P);
  const int64_t U = P->out_base[P->nb];
  if (!P->d_len) {
    P->d_len = dev_alloc<int32_t>(U + 1, stream);
    P->d_doff = dev_alloc<int64_t>(U + 2, stream);
    P->d_sums = dev_alloc<int64_t>((U + kScanChunk - 1) / kScanChunk + 2, stream);
  }
}

void gather_var_multi(MergePlan* P, const VarCol* cols, int ncols, cudaStream_t stream) {
  require_synced(P);
  const int64_t total_batches = P->batch_base[P->nb];
  if (total_batches == 0 || ncols <= 0) return;
  GatherGeom G = make_geom(P);
  const int64_t chunks = total_batches * G.chunks_per_batch;
  // LAKESOUL_B200_WINDOW_GATHER=0/1 overrides the default
  static const bool no_windows = [] { const char* v = getenv("LAKESOUL_B200_WINDOW_GATHER"); return v ? *v != '1' : !kWindowGatherDefault; }();
  static thread_local int attr_device = -1;
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  if (attr_device != cur_dev) {
    attr_device = cur_dev;
    LSB_CUDA(cudaFuncSetAttribute(var_window_kernel<14, 24576>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)sizeof(WinSmem<14, 24576>)));
    LSB_CUDA(cudaFuncSetAttribute(var_window_kernel<13, 16384>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)sizeof(WinSmem<13, 16384>)));
    LSB_CUDA(cudaFuncSetAttribute(var_write_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kVarWriteDynSmem));
    LSB_CUDA(cudaFuncSetAttribute(var_write_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kVarWriteDynSmem));
  }
  // columns whose rows are file rows of the plan take the windowed kernel; computed columns (merge operators:
  // rows of a scan-wide scratch column) keep the direct one. Each group: one launch per 16 columns.
  for (int pass = 0; pass < 2; ++pass) {
    std::vector<VarCol> group;
    for (int i = 0; i < ncols; ++i) {
      const bool windowed = !cols[i].direct && !no_windows;
      if (windowed == (pass == 0)) group.push_back(cols[i]);
    }
    for (size_t c0 = 0; c0 < group.size(); c0 += kMaxVarCols) {
      const int n = (int)std::min<size_t>(kMaxVarCols, group.size() - c0);
      if (chunks * n > 0x7fffffffll) {
        g_last_error = "too many string chunks in one scan";
        throw std::runtime_error(g_last_error);
      }
      // look-back scan state: two status words per (column, chunk), zeroed = "not published"; every launch of a
      // gather has its own state and descriptor block (the launches are asynchronous)
      const int64_t words = 2 * chunks * n;
      int64_t* state = dev_alloc<int64_t>(words, stream);
      VarColSet* d_set = pass == 0 ? dev_alloc<VarColSet>(1, stream) : nullptr;  // var_write takes the set as a kernel parameter
      P->var_scratch.push_back(state);
      if (d_set) P->var_scratch.push_back(d_set);
      LSB_CUDA(cudaMemsetAsync(state, 0, sizeof(int64_t) * words, stream));
      VarScan S{};
      S.agg = reinterpret_cast<unsigned long long*>(state);
      S.incl = S.agg + chunks * n;
      S.stride = chunks;
      S.ncols = n;
      S.err = P->d_err;
      P->h_varsets.emplace_back();
      VarColSet& set = P->h_varsets.back();
      set.n = n;
      for (int i = 0; i < n; ++i) set.col[i] = group[c0 + i];
      if (d_set) LSB_CUDA(h2d_staged(d_set, &set, sizeof(VarColSet), stream));
      static const int var_blocks = [] { const char* v = getenv("LAKESOUL_B200_VAR_BLOCKS"); return v ? atoi(v) : 4; }();
      static const int win_variant = [] { const char* v = getenv("LAKESOUL_B200_WINDOW_VARIANT"); return v ? atoi(v) : 0; }();
      if (pass == 0 && win_variant == 1)
        var_window_kernel<14, 24576><<<(unsigned)(chunks * n), kGatherThreads, sizeof(WinSmem<14, 24576>), stream>>>(G, d_set, S, total_batches);
      else if (pass == 0)
        var_window_kernel<13, 16384><<<(unsigned)(chunks * n), kGatherThreads, sizeof(WinSmem<13, 16384>), stream>>>(G, d_set, S, total_batches);
      else if (var_blocks >= 4)
        var_write_kernel<4><<<(unsigned)(chunks * n), kGatherThreads, kVarWriteDynSmem, stream>>>(G, set, S, total_batches);
      else
        var_write_kernel<3><<<(unsigned)(chunks * n), kGatherThreads, kVarWriteDynSmem, stream>>>(G, set, S, total_batches);
    }
  }
  LSB_CUDA(cudaGetLastError());
}

void gather_var(MergePlan* P, const VarCol& col, cudaStream_t stream) { gather_var_multi(P, &col, 1, stream); }

// ---------------------------------------------------------------------------------------
// row-encoded keys (general primary-key shapes)
// ---------------------------------------------------------------------------------------
constexpr int kMaxEncCols = 16;
struct EncKeySet {
  int32_t n;
  EncKeyCol col[kMaxEncCols];
};

__device__ __forceinline__ bool enc_valid(const EncKeyCol& c, int64_t g) {
  return !c.validity || ((c.validity[g >> 3] >> (g & 7)) & 1);
}

// dense row u (files packed back to back) -> global row, file
__device__ __forceinline__ int64_t dense_to_global(const PlanDev& P, int F, int64_t u, int* f_out) {
  const int f = upper_idx(P.dense_base, F + 1, u);
  *f_out = f;
  return P.file_row_base[f] + (u - P.dense_base[f]);
}

template <bool kWrite>
__global__ void __launch_bounds__(256) encode_keys_kernel(PlanDev P, int F, int64_t n_in, const EncKeySet* __restrict__ set,
                                                          const int32_t* __restrict__ file_bucket, int32_t* __restrict__ len,
                                                          const int64_t* __restrict__ off, int32_t* __restrict__ enc_off,
                                                          uint8_t* __restrict__ enc_data) {
  const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n_in) return;
  int f;
  const int64_t g = dense_to_global(P, F, u, &f);
  uint8_t* out = kWrite ? enc_data + off[u] : nullptr;
  if (kWrite) {
    enc_off[g] = (int32_t)off[u];
    if (u + 1 == P.dense_base[f + 1]) enc_off[g + 1] = (int32_t)off[u + 1];  // closing entry in the file's spare row
  }
  int64_t n = 0;
  for (int c = 0; c < set->n; ++c) {
    const EncKeyCol& col = set->col[c];
    const bool valid = enc_valid(col, g);
    if (kWrite) out[n] = valid ? 1 : 0;
    ++n;
    if (col.kind == ENC_BYTES) {
      if (valid) {
        const int32_t* offs = reinterpret_cast<const int32_t*>(col.values);
        const uint8_t* p = col.data + col.bucket_data[file_bucket[f]] + offs[g];
        const int32_t l = offs[g + 1] - offs[g];
        for (int32_t i = 0; i < l; ++i) {
          const uint8_t b = p[i];
          if (kWrite) out[n] = b;
          ++n;
          if (b == 0) { if (kWrite) out[n] = 0xFF; ++n; }
        }
      }
      if (kWrite) { out[n] = 0; out[n + 1] = 0; }
      n += 2;
    } else if (col.kind == ENC_BOOL_BITS) {
      const uint8_t v = valid ? ((reinterpret_cast<const uint8_t*>(col.values)[g >> 3] >> (g & 7)) & 1) : 0;
      if (kWrite) out[n] = v;
      ++n;
    } else {
      const int w = col.width;
      if (kWrite) {
        const uint8_t* p = reinterpret_cast<const uint8_t*>(col.values) + g * w;
        const bool neg = (p[w - 1] & 0x80) != 0;
        for (int i = 0; i < w; ++i) {  // big endian, most significant byte first
          uint8_t b = valid ? p[w - 1 - i] : 0;
          if (valid) {
            if (col.kind == ENC_SIGNED) { if (i == 0) b ^= 0x80; }
            else if (col.kind == ENC_FLOAT) { b = neg ? (uint8_t)~b : (i == 0 ? (uint8_t)(b ^ 0x80) : b); }
          }
          out[n + i] = b;
        }
      }
      n += w;
    }
  }
  if (!kWrite) len[u] = (int32_t)min(n, (int64_t)INT32_MAX);
}

void plan_encode_keys(MergePlan* P, const EncKeyCol* cols, int ncols, PkDesc* pk, cudaStream_t stream) {
  if (ncols <= 0 || ncols > kMaxEncCols) {
    g_last_error = "primary key with more than 16 columns is not supported";
    throw std::runtime_error(g_last_error);
  }
  const int64_t n = P->n_in;
  EncKeySet set{};
  set.n = ncols;
  for (int i = 0; i < ncols; ++i) set.col[i] = cols[i];
  EncKeySet* d_set = dev_alloc<EncKeySet>(1, stream);
  LSB_CUDA(h2d_staged(d_set, &set, sizeof(set), stream));
  std::vector<int32_t> fb((size_t)std::max(P->F, 1), 0);
  for (int b = 0; b < P->nb; ++b)
    for (int j = 0; j < P->buckets[b].k; ++j) fb[(size_t)P->buckets[b].file_begin + j] = b;
  int32_t* d_fb = dev_alloc<int32_t>(fb.size(), stream);
  LSB_CUDA(h2d_staged(d_fb, fb.data(), sizeof(int32_t) * fb.size(), stream));
  int32_t* d_off32 = dev_alloc<int32_t>((size_t)P->n_total + 2, stream);
  int64_t* d_zero = dev_alloc<int64_t>((size_t)std::max(P->nb, 1), stream);
  LSB_CUDA(cudaMemsetAsync(d_zero, 0, sizeof(int64_t) * (size_t)std::max(P->nb, 1), stream));
  LSB_CUDA(cudaMemsetAsync(d_off32, 0, sizeof(int32_t) * ((size_t)P->n_total + 2), stream));
  uint8_t* d_data = nullptr;
  P->var_scratch.push_back(d_set);
  P->var_scratch.push_back(d_fb);
  P->var_scratch.push_back(d_off32);
  P->var_scratch.push_back(d_zero);
  if (n > 0) {
    PlanDev D = plan_dev(P);
    int32_t* d_len = dev_alloc<int32_t>((size_t)n, stream);
    const int nblocks = (int)((n + kScanChunk - 1) / kScanChunk);
    int64_t* d_sums = dev_alloc<int64_t>((size_t)nblocks + 2, stream);
    int64_t* d_doff = dev_alloc<int64_t>((size_t)n + 1, stream);
    P->var_scratch.push_back(d_len);
    P->var_scratch.push_back(d_sums);
    P->var_scratch.push_back(d_doff);
    const unsigned grid = (unsigned)((n + 255) / 256);
    encode_keys_kernel<false><<<grid, 256, 0, stream>>>(D, P->F, n, d_set, d_fb, d_len, nullptr, nullptr, nullptr);
    scan_block_sums<<<nblocks, 256, 0, stream>>>(d_len, n, d_sums);
    scan_sums_inplace<<<1, 1024, 0, stream>>>(d_sums, nblocks, 0);
    scan_apply<<<nblocks, 256, 0, stream>>>(d_len, n, d_sums, d_doff);
    int64_t* h_total = pinned_counts(4) + 3;  // (the plan's count read-back uses the front of the same block later)
    LSB_CUDA(cudaMemcpyAsync(h_total, d_doff + n, sizeof(int64_t), cudaMemcpyDeviceToHost, stream));
    LSB_CUDA(stream_wait(stream));
    const int64_t total = *h_total;
    if (total > 0x7fffff00ll) {
      g_last_error = "row-encoded primary keys exceed 2 GiB in one scan";
      throw std::runtime_error(g_last_error);
    }
    d_data = dev_alloc<uint8_t>((size_t)total + 64, stream);
    P->var_scratch.push_back(d_data);
    encode_keys_kernel<true><<<grid, 256, 0, stream>>>(D, P->F, n, d_set, d_fb, nullptr, d_doff, d_off32, d_data);
    LSB_CUDA(cudaGetLastError());
  }
  *pk = PkDesc{};
  pk->ncols = 1;
  pk->key_bits = 128;
  pk->col[0].values = d_off32;
  pk->col[0].validity = nullptr;
  pk->col[0].width = 0;
  pk->col[0].kind = PK_BYTES;
  pk->col[0].data = d_data;
  pk->col[0].bucket_data = d_zero;
}

// ---------------------------------------------------------------------------------------
// winner compaction after the post-merge predicate
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) keep_to_i32(const uint8_t* keep, int64_t n, int32_t* out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = keep[i] ? 1 : 0;
}
__global__ void __launch_bounds__(256) scatter_kept(const uint8_t* keep, const int64_t* pos, int64_t n,
                                                   const uint32_t* src, uint32_t* dst) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && keep[i]) dst[pos[i]] = src[i];
}
__global__ void __launch_bounds__(256) scatter_kept_u8(const uint8_t* keep, const int64_t* pos, int64_t n,
                                                      const uint8_t* src, uint8_t* dst) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && keep[i]) dst[pos[i]] = src[i];
}
__global__ void __launch_bounds__(256) rebase_out(PlanDev P, const int64_t* pos, int64_t* new_base) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b <= P.nb) new_base[b] = pos[P.out_base[b]];
}
__global__ void __launch_bounds__(256) copy_i64(const int64_t* src, int64_t* dst, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}

void plan_compact(MergePlan* P, const uint8_t* keep, const std::vector<uint32_t*>& extra_lists,
                  cudaStream_t stream) {
  require_synced(P);
  const int64_t U = P->out_base[P->nb];
  if (U == 0) return;
  plan_reserve_scratch(P, stream);
  if (!P->d_ctmp) P->d_ctmp = dev_alloc<uint32_t>(P->n_total, stream);
  const unsigned g = (unsigned)((U + 255) / 256);
  keep_to_i32<<<g, 256, 0, stream>>>(keep, U, P->d_len);
  const int nblocks = (int)((U + kScanChunk - 1) / kScanChunk);
  scan_block_sums<<<nblocks, 256, 0, stream>>>(P->d_len, U, P->d_sums);
  scan_sums_inplace<<<1, 1024, 0, stream>>>(P->d_sums, nblocks, 0);
  scan_apply<<<nblocks, 256, 0, stream>>>(P->d_len, U, P->d_sums, P->d_doff);
  // plan winners: scatter into win_tmp, then swap roles
  scatter_kept<<<g, 256, 0, stream>>>(keep, P->d_doff, U, P->d_win, P->d_win_tmp);
  std::swap(P->d_win, P->d_win_tmp);
  if (P->has_wfile) {
    scatter_kept_u8<<<g, 256, 0, stream>>>(keep, P->d_doff, U, P->d_wfile, P->d_wfile_tmp);
    std::swap(P->d_wfile, P->d_wfile_tmp);
  }
  for (uint32_t* list : extra_lists) {
    scatter_kept<<<g, 256, 0, stream>>>(keep, P->d_doff, U, list, P->d_ctmp);
    LSB_CUDA(cudaMemcpyAsync(list, P->d_ctmp, sizeof(uint32_t) * U, cudaMemcpyDeviceToDevice, stream));
  }
  PlanDev D = plan_dev(P);
  // new bases: out_base'[b] = pos[out_base[b]]  (pos has U+1 entries); staged through d_sums' tail
  int64_t* nb_tmp = reinterpret_cast<int64_t*>(P->d_len);  // d_len no longer needed (>= (U+1)*4 bytes)
  if ((int64_t)(P->nb + 1) * 8 > (U + 1) * 4) {
    // tiny outputs with many buckets: use a dedicated buffer
    dev_free(P->d_sums, stream);
    P->d_sums = dev_alloc<int64_t>((U + kScanChunk - 1) / kScanChunk + 2 + P->nb + 1, stream);
    nb_tmp = P->d_sums + (U + kScanChunk - 1) / kScanChunk + 2;
  }
  rebase_out<<<(P->nb + 1 + 255) / 256, 256, 0, stream>>>(D, P->d_doff, nb_tmp);
  copy_i64<<<(P->nb + 1 + 255) / 256, 256, 0, stream>>>(nb_tmp, P->d_out_base, P->nb + 1);
  batch_base_kernel<<<1, 1024, 0, stream>>>(D);
  LSB_CUDA(cudaGetLastError());
  P->synced = false;
}

// ---------------------------------------------------------------------------------------
// write side: hash buckets of rows and a device sort of the row list
//   * plan_hash_buckets: spark murmur3_32 chained over the key columns (seed 42 for the first column, the running
//     hash for the next, a null leaves it unchanged; utils/hash/mod.rs:22-178, spark_murmur3.rs:42-87), mod the
//     bucket count as u32 (physical_plan/repartition/mod.rs:260);
//   * plan_sort: the identity row list of the plan sorted by (key, input position) -- what SortExec over
//     (range partitions, primary keys, aux sort columns) gives the reference's writers (sort_writer.rs:59-76,
//     partitioning_writer.rs:139-171). Tiles of 2048 rows are sorted in shared memory by a bitonic network, then
//     log2(runs) passes merge neighbouring runs pairwise: one block per 2048 outputs, merge-path split by two binary
//     searches through the row list, merge of the two segments in shared memory.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mm3_mix(uint32_t state, uint32_t k) {
  k *= 0xcc9e2d51u;
  k = __funnelshift_l(k, k, 15);
  k *= 0x1b873593u;
  state ^= k;
  state = __funnelshift_l(state, state, 13);
  return state * 5u + 0xe6546b64u;
}
__device__ __forceinline__ uint32_t mm3_fin(uint32_t state, uint32_t n) {
  uint32_t h = state ^ n;
  h ^= h >> 16;
  h *= 0x85ebca6bu;
  h ^= h >> 13;
  h *= 0xc2b2ae35u;
  h ^= h >> 16;
  return h;
}

__global__ void __launch_bounds__(256) hash_buckets_kernel(PlanDev P, int F, int64_t n_in, const EncKeySet* __restrict__ set,
                                                           const int32_t* __restrict__ file_bucket, uint32_t nbuckets,
                                                           uint32_t* __restrict__ out) {
  const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n_in) return;
  int f;
  const int64_t g = dense_to_global(P, F, u, &f);
  uint32_t h = 0;  // the hash buffer starts at 0 (repartition/mod.rs:247-248): a null first column leaves 0
  for (int c = 0; c < set->n; ++c) {
    const EncKeyCol& col = set->col[c];
    if (!enc_valid(col, g)) continue;
    const uint32_t seed = c == 0 ? 42u : h;
    if (col.kind == ENC_BYTES) {
      const int32_t* offs = reinterpret_cast<const int32_t*>(col.values);
      const uint8_t* p = col.data + col.bucket_data[file_bucket[f]] + offs[g];
      const int32_t l = offs[g + 1] - offs[g];
      uint32_t st = seed;
      int32_t i = 0;
      for (; i + 4 <= l; i += 4)
        st = mm3_mix(st, (uint32_t)p[i] | ((uint32_t)p[i + 1] << 8) | ((uint32_t)p[i + 2] << 16) | ((uint32_t)p[i + 3] << 24));
      for (; i < l; ++i) st = mm3_mix(st, (uint32_t)p[i]);
      h = mm3_fin(st, (uint32_t)l);
    } else if (col.kind == ENC_BOOL_BITS) {
      const uint32_t v = (reinterpret_cast<const uint8_t*>(col.values)[g >> 3] >> (g & 7)) & 1;
      h = mm3_fin(mm3_mix(seed, v), 4);
    } else {
      const int w = col.width;
      const uint8_t* p = reinterpret_cast<const uint8_t*>(col.values) + g * w;
      if (w <= 4) {
        uint32_t v;
        if (w == 1) v = col.kind == ENC_SIGNED ? (uint32_t)(int32_t)(int8_t)p[0] : (uint32_t)p[0];
        else if (w == 2) {
          const uint16_t x = *reinterpret_cast<const uint16_t*>(p);
          v = col.kind == ENC_SIGNED ? (uint32_t)(int32_t)(int16_t)x : (uint32_t)x;
          if (col.kind == ENC_FLOAT && x == 0x8000u) v = 0;  // f16 -0.0
        } else {
          v = *reinterpret_cast<const uint32_t*>(p);
          if (col.kind == ENC_FLOAT && v == 0x80000000u) v = 0;  // -0.0 hashes like 0.0 (utils/hash/mod.rs:76-88)
        }
        h = mm3_fin(mm3_mix(seed, v), 4);
      } else {
        uint32_t st = seed;
        const uint32_t* q = reinterpret_cast<const uint32_t*>(p);
        if (w == 8 && col.kind == ENC_FLOAT && q[0] == 0u && q[1] == 0x80000000u) {
          st = mm3_mix(mm3_mix(st, 0u), 0u);
        } else {
          for (int i = 0; i < w / 4; ++i) st = mm3_mix(st, q[i]);
        }
        h = mm3_fin(st, (uint32_t)w);
      }
    }
  }
  out[g] = h % nbuckets;
}

void plan_hash_buckets(MergePlan* P, const EncKeyCol* cols, int ncols, uint32_t nbuckets, uint32_t* d_out, cudaStream_t stream) {
  if (ncols < 0 || ncols > kMaxEncCols || nbuckets == 0) {
    g_last_error = "plan_hash_buckets: bad arguments";
    throw std::runtime_error(g_last_error);
  }
  const int64_t n = P->n_in;
  if (n == 0) return;
  EncKeySet set{};
  set.n = ncols;
  for (int i = 0; i < ncols; ++i) set.col[i] = cols[i];
  EncKeySet* d_set = dev_alloc<EncKeySet>(1, stream);
  LSB_CUDA(h2d_staged(d_set, &set, sizeof(set), stream));
  std::vector<int32_t> fb((size_t)std::max(P->F, 1), 0);
  for (int b = 0; b < P->nb; ++b)
    for (int j = 0; j < P->buckets[b].k; ++j) fb[(size_t)P->buckets[b].file_begin + j] = b;
  int32_t* d_fb = dev_alloc<int32_t>(fb.size(), stream);
  LSB_CUDA(h2d_staged(d_fb, fb.data(), sizeof(int32_t) * fb.size(), stream));
  P->var_scratch.push_back(d_set);
  P->var_scratch.push_back(d_fb);
  hash_buckets_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(plan_dev(P), P->F, n, d_set, d_fb, nbuckets, d_out);
  LSB_CUDA(cudaGetLastError());
}

constexpr int kSortTile = 2048;
constexpr int kSortThreads = 256;
constexpr int kSortPer = kSortTile / kSortThreads;  // 8

// (key, row) of the sort: rows are global row ids; kNoRow pads a tile and sorts last. Equal keys keep their input
// order (the row list starts increasing), which also makes the order total -- a bitonic network needs no stability.
template <typename KeyT>
__device__ __forceinline__ bool sort_lt(const KeyOps<KeyT>& ops, const KeyT& ka, uint32_t ra, const KeyT& kb, uint32_t rb) {
  if (ra == kNoRow || rb == kNoRow) return rb == kNoRow && ra != kNoRow;
  if (ops.lt(ka, kb)) return true;
  if (ops.lt(kb, ka)) return false;
  return ra < rb;
}

template <typename KeyT>
__global__ void __launch_bounds__(kSortThreads) sort_tiles_kernel(PkDesc pk, const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                                 int64_t n) {
  __shared__ KeyT s_key[kSortTile];
  __shared__ uint32_t s_row[kSortTile];
  const KeyOps<KeyT> ops(pk, 0);
  const int64_t base = (int64_t)blockIdx.x * kSortTile;
  for (int i = threadIdx.x; i < kSortTile; i += kSortThreads) {
    const int64_t q = base + i;
    const uint32_t r = q < n ? in[q] : kNoRow;
    s_row[i] = r;
    if (r != kNoRow) s_key[i] = ops.load(r);
  }
  __syncthreads();
  for (int size = 2; size <= kSortTile; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < kSortTile / 2; t += kSortThreads) {
        const int lo = 2 * t - (t & (stride - 1));  // element with bit `stride` clear
        const int hi = lo + stride;
        const bool up = (lo & size) == 0;           // ascending block of the bitonic sequence
        const KeyT ka = s_key[lo], kb = s_key[hi];
        const uint32_t ra = s_row[lo], rb = s_row[hi];
        const bool swap = up ? sort_lt(ops, kb, rb, ka, ra) : sort_lt(ops, ka, ra, kb, rb);
        if (swap) {
          s_key[lo] = kb; s_key[hi] = ka;
          s_row[lo] = rb; s_row[hi] = ra;
        }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < kSortTile; i += kSortThreads) {
    const int64_t q = base + i;
    if (q < n) out[q] = s_row[i];
  }
}

// number of elements taken from run A among the first d outputs of merge(A, B): A = list[a0, a0 + na), B = list[b0, b0 + nb)
template <typename KeyT>
__device__ int64_t sort_corank_global(const KeyOps<KeyT>& ops, const uint32_t* list, int64_t a0, int64_t na, int64_t b0, int64_t nb, int64_t d) {
  int64_t lo = d > nb ? d - nb : 0, hi = d < na ? d : na;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;  // take mid + 1 from A?  A[mid] must precede B[d - 1 - mid]
    const uint32_t ra = list[a0 + mid], rb = list[b0 + d - 1 - mid];
    const KeyT ka = ops.load(ra), kb = ops.load(rb);
    if (sort_lt(ops, ka, ra, kb, rb)) lo = mid + 1; else hi = mid;
  }
  return lo;
}

template <typename KeyT>
__global__ void __launch_bounds__(kSortThreads) sort_merge_kernel(PkDesc pk, const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                                 int64_t n, int64_t run) {
  __shared__ KeyT s_key[kSortTile];
  __shared__ uint32_t s_row[kSortTile];
  __shared__ long long s_split[2];
  const KeyOps<KeyT> ops(pk, 0);
  const int64_t o0 = (int64_t)blockIdx.x * kSortTile;   // first output of this block
  const int64_t pair0 = o0 / (2 * run) * (2 * run);     // the pair of runs it belongs to
  const int64_t a0 = pair0, na = min(run, n - pair0);
  const int64_t b0 = pair0 + na, nb = max((int64_t)0, min(run, n - b0));
  const int64_t d0 = o0 - pair0, d1 = min(d0 + kSortTile, na + nb);
  if (threadIdx.x < 2) s_split[threadIdx.x] = sort_corank_global(ops, in, a0, na, b0, nb, threadIdx.x == 0 ? d0 : d1);
  __syncthreads();
  const int64_t ia0 = s_split[0], ia1 = s_split[1];
  const int64_t ib0 = d0 - ia0, ib1 = d1 - ia1;
  const int ca = (int)(ia1 - ia0), cb = (int)(ib1 - ib0);  // ca + cb = d1 - d0 <= kSortTile
  for (int i = threadIdx.x; i < ca + cb; i += kSortThreads) {
    const uint32_t r = i < ca ? in[a0 + ia0 + i] : in[b0 + ib0 + (i - ca)];
    s_row[i] = r;
    s_key[i] = ops.load(r);
  }
  __syncthreads();
  const int total = ca + cb;
  const int q0 = min((int)threadIdx.x * kSortPer, total), q1 = min(q0 + kSortPer, total);
  if (q0 < q1) {
    int lo = max(0, q0 - cb), hi = min(q0, ca);
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (sort_lt(ops, s_key[mid], s_row[mid], s_key[ca + q0 - 1 - mid], s_row[ca + q0 - 1 - mid])) lo = mid + 1; else hi = mid;
    }
    int ia = lo, ib = q0 - lo;
    for (int q = q0; q < q1; ++q) {
      bool take_a;
      if (ia >= ca) take_a = false;
      else if (ib >= cb) take_a = true;
      else take_a = sort_lt(ops, s_key[ia], s_row[ia], s_key[ca + ib], s_row[ca + ib]);
      out[pair0 + d0 + q] = take_a ? s_row[ia] : s_row[ca + ib];
      ia += take_a ? 1 : 0;
      ib += take_a ? 0 : 1;
    }
  }
}

template <typename KeyT>
static int run_sort(MergePlan* P, const PkDesc& pk, cudaStream_t stream) {
  const int64_t n = P->out_base[P->nb];
  if (n <= 1) return 0;
  const unsigned tiles = (unsigned)((n + kSortTile - 1) / kSortTile);
  uint32_t* a = P->d_win;
  uint32_t* b = P->d_win_tmp;
  sort_tiles_kernel<KeyT><<<tiles, kSortThreads, 0, stream>>>(pk, a, b, n);
  std::swap(a, b);  // a = sorted tiles
  int launches = 1;
  for (int64_t run = kSortTile; run < n; run *= 2) {
    sort_merge_kernel<KeyT><<<tiles, kSortThreads, 0, stream>>>(pk, a, b, n, run);
    std::swap(a, b);
    ++launches;
  }
  if (a != P->d_win) LSB_CUDA(cudaMemcpyAsync(P->d_win, a, sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToDevice, stream));
  LSB_CUDA(cudaGetLastError());
  return launches;
}

// Sorts the plan's row list (plan_identity + plan_sync_counts first: every input row, file order) by `pk`.
void plan_sort(MergePlan* P, const PkDesc& pk, cudaStream_t stream) {
  require_synced(P);
  if (P->nb != 1) {
    g_last_error = "plan_sort: one bucket per plan";
    throw std::runtime_error(g_last_error);
  }
  int launches;
  if (pk.ncols == 1 && pk.col[0].kind == PK_BYTES) launches = run_sort<StrKey>(P, pk, stream);
  else if (pk.key_bits > 64) launches = run_sort<K128>(P, pk, stream);
  else launches = run_sort<uint64_t>(P, pk, stream);
  P->stats.launches += launches;
  P->has_wfile = false;
}

// out[u] = src[win[u]] for a 4-byte column indexed by global row (the bucket ids of the sorted rows)
__global__ void __launch_bounds__(256) take_u32_kernel(const uint32_t* __restrict__ win, const uint32_t* __restrict__ src, int64_t n,
                                                       uint32_t* __restrict__ out) {
  const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u < n) out[u] = src[win[u]];
}
void plan_take_u32(MergePlan* P, const uint32_t* d_src, uint32_t* d_out, cudaStream_t stream) {
  require_synced(P);
  const int64_t n = P->out_base[P->nb];
  if (n > 0) take_u32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(P->d_win, d_src, n, d_out);
  LSB_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------
// general path: partial merge + merge operators, resolved from the presence matrix R
// ---------------------------------------------------------------------------------------
struct GenColDev {
  const void* values;
  const uint8_t* validity;
  const char* present;  // F entries (device)
  uint32_t* win_out;    // U entries
  void* opval;          // U * width (operators that compute values)
  uint32_t* opvalid;    // bitmap words over output rows
  int32_t op;
  int32_t width;
  int32_t num_class;
};

template <typename T>
__device__ __forceinline__ T ld(const void* p, int64_t i) { return reinterpret_cast<const T*>(p)[i]; }
__device__ __forceinline__ bool bit_at(const uint8_t* bm, int64_t i) { return !bm || ((bm[i >> 3] >> (i & 7)) & 1); }

template <typename KeyT>
__device__ __forceinline__ int64_t run_first(const KeyOps<KeyT>& ops, int64_t file_first, int64_t last) {
  // in-file run of equal keys ending at `last` (sort_key_range.rs:99-113)
  const KeyT kl = ops.load(last);
  int64_t r = last;
  while (r > file_first && !ops.ne(ops.load(r - 1), kl)) --r;
  return r;
}

template <typename T>
__device__ __forceinline__ void store_val(void* p, int64_t i, T v) { reinterpret_cast<T*>(p)[i] = v; }

template <typename KeyT, typename T>
__device__ void sum_op(const PlanDev& P, const KeyOps<KeyT>& ops, const GenColDev& c, const BucketDesc& B,
                       const uint32_t* Rrow, int64_t u, bool sum_all, bool& valid) {
  T res = T(0);
  valid = true;
  for (int j = 0; j < B.k && valid; ++j) {
    if (!c.present[B.file_begin + j]) continue;
    const uint32_t last = Rrow[j];
    if (last == kNoRow) continue;
    if (sum_all) {
      const int64_t first = run_first<KeyT>(ops, P.file_row_base[B.file_begin + j], last);
      T part = T(0);
      for (int64_t r = first; r <= (int64_t)last; ++r) {
        if (!bit_at(c.validity, r)) { valid = false; break; }
        part += ld<T>(c.values, r);
      }
      res += part;
    } else {
      if (!bit_at(c.validity, last)) { valid = false; break; }
      res += ld<T>(c.values, last);
    }
  }
  store_val<T>(c.opval, u, valid ? res : T(0));
}

__global__ void __launch_bounds__(256) iota_u32(uint32_t* out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (uint32_t)i;
}

template <typename KeyT>
__global__ void __launch_bounds__(256) resolve_kernel(PlanDev P, PkDesc pk, GenColDev c, int64_t U) {
  const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = u < U;
  bool valid = false;
  if (active) {
    const int b = upper_idx(P.out_base, P.nb + 1, u);
    const BucketDesc& B = P.buckets[b];
    const KeyOps<KeyT> ops(pk, b);
    const uint32_t* Rrow = P.R + P.rofs[b] + (u - P.out_base[b]) * B.k;
    const char* present = c.present + B.file_begin;
    if (c.op == (int)MergeOp::UseLast) {
      uint32_t w = kNoRow;
      for (int j = B.k - 1; j >= 0; --j) {
        if (present[j] && Rrow[j] != kNoRow) { w = Rrow[j]; break; }
      }
      c.win_out[u] = w;
    } else if (c.op == (int)MergeOp::UseLastNotNull) {
      uint32_t w = kNoRow;
      for (int j = B.k - 1; j >= 0 && w == kNoRow; --j) {
        if (!present[j] || Rrow[j] == kNoRow) continue;
        const int64_t last = Rrow[j];
        const int64_t first = run_first<KeyT>(ops, P.file_row_base[B.file_begin + j], last);
        for (int64_t r = last; r >= first; --r) {
          if (bit_at(c.validity, r)) { w = (uint32_t)r; break; }
        }
      }
      c.win_out[u] = w;
    } else {
      const bool all = c.op == (int)MergeOp::SumAll;
      switch (c.num_class) {
        case NC_I8: sum_op<KeyT, int8_t>(P, ops, c, B, Rrow, u, all, valid); break;
        case NC_I16: sum_op<KeyT, int16_t>(P, ops, c, B, Rrow, u, all, valid); break;
        case NC_I32: sum_op<KeyT, int32_t>(P, ops, c, B, Rrow, u, all, valid); break;
        case NC_I64: sum_op<KeyT, int64_t>(P, ops, c, B, Rrow, u, all, valid); break;
        case NC_U8: sum_op<KeyT, uint8_t>(P, ops, c, B, Rrow, u, all, valid); break;
        case NC_U16: sum_op<KeyT, uint16_t>(P, ops, c, B, Rrow, u, all, valid); break;
        case NC_U32: sum_op<KeyT, uint32_t>(P, ops, c, B, Rrow, u, all, valid); break;
        case NC_U64: sum_op<KeyT, uint64_t>(P, ops, c, B, Rrow, u, all, valid); break;
        case NC_F32: sum_op<KeyT, float>(P, ops, c, B, Rrow, u, all, valid); break;
        default: sum_op<KeyT, double>(P, ops, c, B, Rrow, u, all, valid); break;
      }
      // a key that no carrying file contains stays null (merge_operator.rs:61-62)
      bool any = false;
      for (int j = 0; j < B.k; ++j) any |= (present[j] && Rrow[j] != kNoRow);
      valid = valid && any;
      c.win_out[u] = (uint32_t)u;  // gather reads opval[u]
    }
  }
  if (c.opvalid) {
    const uint32_t word = __ballot_sync(0xffffffffu, active && valid);
    if ((threadIdx.x & 31) == 0 && (u - (threadIdx.x & 31)) < U) c.opvalid[u >> 5] = word;
  }
}

// Joined* operators (merge_operator.rs:389-487): the result is a new Utf8 cell per output key.
// Pass 1 (kWrite=false): one thread per output key sizes the cell and decides its validity;
// pass 2 (kWrite=true): one warp per output key copies the pieces to their scanned position.
struct JoinedDev {
  const int32_t* src_offsets;   // global-row offsets of the source column
  const uint8_t* src_data;
  const int64_t* bucket_data;   // per-bucket base of src_data
  int32_t* len;                 // U (pass 1 out)
  const int64_t* ooff;          // U+1 scanned (pass 2 in)
  int32_t* ooff32;              // U+1 (pass 2 out): offsets of the computed column
  uint8_t* out;                 // computed column bytes
  char delim;
  bool all;                     // JoinedAll*: every row of every run, and never a delimiter (:402-406)
};

template <typename KeyT, bool kWrite>
__global__ void __launch_bounds__(256) joined_kernel(PlanDev P, PkDesc pk, GenColDev c, JoinedDev J, int64_t U) {
  const int lane = threadIdx.x & 31;
  const int64_t u = kWrite ? ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5))
                           : ((int64_t)blockIdx.x * blockDim.x + threadIdx.x);
  const bool active = u < U;
  bool valid = false;
  if (active) {
    const int b = upper_idx(P.out_base, P.nb + 1, u);
    const BucketDesc& B = P.buckets[b];
    const KeyOps<KeyT> ops(pk, b);
    const uint32_t* Rrow = P.R + P.rofs[b] + (u - P.out_base[b]) * B.k;
    const char* present = c.present + B.file_begin;
    const uint8_t* data = J.src_data + J.bucket_data[b];
    int64_t pos = kWrite ? J.ooff[u] : 0;
    const int64_t cell_len = kWrite ? (J.ooff[u + 1] - pos) : 0;
    if (kWrite && lane == 0) {
      J.ooff32[u] = (int32_t)pos;
      if (u == U - 1) J.ooff32[U] = (int32_t)J.ooff[U];
    }
    int64_t len = 0;
    bool any = false, ok = true, first = true;
    if (!kWrite || cell_len > 0) {
      for (int j = 0; j < B.k && ok; ++j) {
        if (!present[j] || Rrow[j] == kNoRow) continue;
        any = true;
        const int64_t last = Rrow[j];
        const int64_t r0 = J.all ? run_first<KeyT>(ops, P.file_row_base[B.file_begin + j], last) : last;
        for (int64_t r = r0; r <= last; ++r) {
          if (!bit_at(c.validity, r)) { ok = false; break; }
          if (!J.all && !first) {
            if (kWrite) { if (lane == 0) J.out[pos] = (uint8_t)J.delim; ++pos; }
            else ++len;
          }
          first = false;
          const int32_t o0 = J.src_offsets[r], o1 = J.src_offsets[r + 1];
          if (kWrite) {
            for (int32_t x = lane; x < o1 - o0; x += 32) J.out[pos + x] = data[o0 + x];
            pos += o1 - o0;
          } else {
            len += o1 - o0;
          }
        }
      }
    }
    valid = ok && any;
    if (!kWrite) J.len[u] = valid ? (int32_t)min(len, (int64_t)INT32_MAX) : 0;
  }
  if (!kWrite && c.opvalid) {
    const uint32_t word = __ballot_sync(0xffffffffu, active && valid);
    if (lane == 0 && (u - lane) < U) c.opvalid[u >> 5] = word;
  }
}

struct GeneralMerge {
  MergePlan* plan = nullptr;
  int C = 0;
  int launches = 0;
  uint32_t* d_R = nullptr;
  int64_t* d_rofs = nullptr;
  char* d_present = nullptr;
  struct Col {
    bool uses_plan_winners = true;
    bool computed = false;
    bool may_null = false;
    uint32_t* win = nullptr;
    void* opval = nullptr;       // computed values (Sum*) or int32 offsets (Joined*)
    uint32_t* opvalid = nullptr;
    uint8_t* opdata = nullptr;   // Joined*: bytes of the computed column
    int64_t opdata_bytes = 0;
    bool joined = false;
  };
  int64_t* d_zero_base = nullptr;  // nb zeros: the computed string columns are addressed scan-wide
  std::vector<Col> cols;
};

void general_destroy(GeneralMerge* G) {
  if (!G) return;
  cudaStream_t st = G->plan ? G->plan->alloc_stream : nullptr;
  dev_free(G->d_R, st); dev_free(G->d_rofs, st); dev_free(G->d_present, st);
  for (auto& c : G->cols) { dev_free(c.win, st); dev_free(c.opval, st); dev_free(c.opvalid, st); dev_free(c.opdata, st); }
  dev_free(G->d_zero_base, st);
  delete G;
}

GeneralMerge* general_create(MergePlan* P, const PkDesc& pk, const std::vector<GeneralColumnIn>& in,
                             cudaStream_t stream) {
  require_synced(P);
  auto* G = new GeneralMerge();
  try {
    G->plan = P;
    G->C = (int)in.size();
    G->cols.resize(G->C);
    const int nb = P->nb;
    const int64_t U = P->out_base[nb];
    // which columns need their own resolution
    bool need_R = false;
    for (int c = 0; c < G->C; ++c) {
      bool all_present = true;
      for (int f = 0; f < P->F; ++f) all_present &= in[c].present[f] != 0;
      auto& col = G->cols[c];
      col.uses_plan_winners = all_present && in[c].op == MergeOp::UseLast;
      col.computed = in[c].op == MergeOp::SumAll || in[c].op == MergeOp::SumLast;
      col.may_null = !all_present || in[c].validity != nullptr || in[c].op != MergeOp::UseLast;
      col.joined = in[c].op == MergeOp::JoinedLastByComma || in[c].op == MergeOp::JoinedLastBySemicolon ||
                   in[c].op == MergeOp::JoinedAllByComma || in[c].op == MergeOp::JoinedAllBySemicolon;
      if (col.joined && !in[c].kind.var) {
        g_last_error = "column type doesn't support MergeOperator::Joined (Utf8 only)";  // as_string_array, :399
        throw std::runtime_error(g_last_error);
      }
      col.computed |= col.joined;
      if (col.computed && !col.joined && in[c].num_class == NC_NONE) {
        g_last_error = "column type doesn't support MergeOperator::Sum";
        throw std::runtime_error(g_last_error);
      }
      need_R |= !col.uses_plan_winners;
    }
    if (!need_R || U == 0) return G;

    // presence matrix
    std::vector<int64_t> rofs(nb + 1, 0);
    for (int b = 0; b < nb; ++b)
      rofs[b + 1] = rofs[b] + (P->out_base[b + 1] - P->out_base[b]) * P->buckets[b].k;
    G->d_R = dev_alloc<uint32_t>(rofs[nb], stream);
    G->d_rofs = dev_alloc<int64_t>(nb + 1, stream);
    LSB_CUDA(cudaMemsetAsync(G->d_R, 0xff, sizeof(uint32_t) * rofs[nb], stream));
    LSB_CUDA(h2d_staged(G->d_rofs, rofs.data(), sizeof(int64_t) * (nb + 1), stream));
    std::vector<char> present((size_t)G->C * P->F);
    for (int c = 0; c < G->C; ++c) std::memcpy(present.data() + (size_t)c * P->F, in[c].present, P->F);
    G->d_present = dev_alloc<char>(present.size(), stream);
    LSB_CUDA(h2d_staged(G->d_present, present.data(), present.size(), stream));
    LSB_CUDA(stream_wait(stream));  // host vectors go out of scope below

    P->d_R = G->d_R;
    P->d_rofs = G->d_rofs;
    PlanDev D = plan_dev(P);
    const int T0 = P->level_tile_cnt[0];
    const int ofs0 = P->level_tile_ofs[0];
    if (P->key_bytes) tile_kernel<StrKey, true><<<T0, kTileThreads, P->smem_bytes, stream>>>(D, pk, 0, ofs0, nullptr);
    else if (P->key128) tile_kernel<K128, true><<<T0, kTileThreads, P->smem_bytes, stream>>>(D, pk, 0, ofs0, nullptr);
    else tile_kernel<uint64_t, true><<<T0, kTileThreads, P->smem_bytes, stream>>>(D, pk, 0, ofs0, nullptr);
    ++G->launches;

    for (int c = 0; c < G->C; ++c) {
      auto& col = G->cols[c];
      if (col.uses_plan_winners) continue;
      col.win = dev_alloc<uint32_t>(U, stream);
      GenColDev gc{};
      gc.values = in[c].values;
      gc.validity = in[c].validity;
      gc.present = G->d_present + (size_t)c * P->F;
      gc.win_out = col.win;
      gc.op = (int)in[c].op;
      gc.width = in[c].kind.width;
      gc.num_class = in[c].num_class;
      if (col.joined) {
        // size -> scan -> write; the computed column is addressed scan-wide with int32 offsets
        col.opval = dev_alloc<int32_t>(U + 2, stream);
        col.opvalid = dev_alloc<uint32_t>((U + 31) / 32 + 2, stream);
        gc.opvalid = col.opvalid;
        if (!G->d_zero_base) {
          G->d_zero_base = dev_alloc<int64_t>(nb, stream);
          LSB_CUDA(cudaMemsetAsync(G->d_zero_base, 0, sizeof(int64_t) * nb, stream));
        }
        const int nblocks = (int)((U + kScanChunk - 1) / kScanChunk);
        int32_t* d_len = dev_alloc<int32_t>(U, stream);
        int64_t* d_sums = dev_alloc<int64_t>(nblocks + 2, stream);
        int64_t* d_ooff = dev_alloc<int64_t>(U + 1, stream);
        JoinedDev J{};
        J.src_offsets = static_cast<const int32_t*>(in[c].values);
        J.src_data = in[c].data;
        J.bucket_data = in[c].bucket_data;
        J.len = d_len;
        J.ooff = d_ooff;
        J.ooff32 = static_cast<int32_t*>(col.opval);
        J.delim = (in[c].op == MergeOp::JoinedLastByComma || in[c].op == MergeOp::JoinedAllByComma) ? ',' : ';';
        J.all = in[c].op == MergeOp::JoinedAllByComma || in[c].op == MergeOp::JoinedAllBySemicolon;
        const unsigned g1 = (unsigned)((U + 255) / 256), g2 = (unsigned)((U + 7) / 8);
        int64_t total = 0;
        try {
          if (P->key_bytes) joined_kernel<StrKey, false><<<g1, 256, 0, stream>>>(D, pk, gc, J, U);
          else if (P->key128) joined_kernel<K128, false><<<g1, 256, 0, stream>>>(D, pk, gc, J, U);
          else joined_kernel<uint64_t, false><<<g1, 256, 0, stream>>>(D, pk, gc, J, U);
          scan_block_sums<<<nblocks, 256, 0, stream>>>(d_len, U, d_sums);
          scan_sums_inplace<<<1, 1024, 0, stream>>>(d_sums, nblocks, 0);
          scan_apply<<<nblocks, 256, 0, stream>>>(d_len, U, d_sums, d_ooff);
          LSB_CUDA(cudaMemcpyAsync(&total, d_ooff + U, sizeof(int64_t), cudaMemcpyDeviceToHost, stream));
          LSB_CUDA(stream_wait(stream));
          if (total > (int64_t)INT32_MAX) {
            g_last_error = "Joined* merge operator: the joined column exceeds 2 GiB in one scan";
            throw std::runtime_error(g_last_error);
          }
          col.opdata = dev_alloc<uint8_t>((size_t)total + 64, stream);
          col.opdata_bytes = total;
          J.out = col.opdata;
          if (P->key_bytes) joined_kernel<StrKey, true><<<g2, 256, 0, stream>>>(D, pk, gc, J, U);
          else if (P->key128) joined_kernel<K128, true><<<g2, 256, 0, stream>>>(D, pk, gc, J, U);
          else joined_kernel<uint64_t, true><<<g2, 256, 0, stream>>>(D, pk, gc, J, U);
          G->launches += 5;
        } catch (...) {
          dev_free(d_len, stream); dev_free(d_sums, stream); dev_free(d_ooff, stream);
          throw;
        }
        dev_free(d_len, stream); dev_free(d_sums, stream); dev_free(d_ooff, stream);
        iota_u32<<<g1, 256, 0, stream>>>(col.win, U);
        ++G->launches;
        continue;
      }
      if (col.computed) {
        col.opval = dev_alloc<uint8_t>((size_t)U * std::max(in[c].kind.width, 1) + 32, stream);
        col.opvalid = dev_alloc<uint32_t>((U + 31) / 32 + 2, stream);
        gc.opval = col.opval;
        gc.opvalid = col.opvalid;
      }
      const unsigned g = (unsigned)((U + 255) / 256);
      if (P->key_bytes) resolve_kernel<StrKey><<<g, 256, 0, stream>>>(D, pk, gc, U);
      else if (P->key128) resolve_kernel<K128><<<g, 256, 0, stream>>>(D, pk, gc, U);
      else resolve_kernel<uint64_t><<<g, 256, 0, stream>>>(D, pk, gc, U);
      ++G->launches;
    }
    LSB_CUDA(cudaGetLastError());
    P->d_R = nullptr;
    P->d_rofs = nullptr;
  } catch (...) {
    P->d_R = nullptr;
    P->d_rofs = nullptr;
    general_destroy(G);
    throw;
  }
  return G;
}

int general_launches(const GeneralMerge* G) { return G->launches; }
const uint32_t* general_winners(const GeneralMerge* G, int c) { return G->cols[c].win; }
const void* general_values(const GeneralMerge* G, int c, const void* dflt) {
  return G->cols[c].computed && G->cols[c].opval ? G->cols[c].opval : dflt;
}
const uint8_t* general_validity(const GeneralMerge* G, int c, const uint8_t* dflt) {
  return G->cols[c].computed && G->cols[c].opvalid ? reinterpret_cast<const uint8_t*>(G->cols[c].opvalid) : dflt;
}
const uint8_t* general_data(const GeneralMerge* G, int c, const uint8_t* dflt) {
  return G->cols[c].joined && G->cols[c].opdata ? G->cols[c].opdata : dflt;
}
const int64_t* general_bucket_data(const GeneralMerge* G, int c, const int64_t* dflt) {
  return G->cols[c].joined && G->cols[c].opdata ? G->d_zero_base : dflt;
}
int64_t general_data_bytes(const GeneralMerge* G, int c, int64_t dflt) {
  return G->cols[c].joined && G->cols[c].opdata ? G->cols[c].opdata_bytes : dflt;
}
bool general_may_be_null(const GeneralMerge* G, int c) { return G->cols[c].may_null; }
std::vector<uint32_t*> general_winner_lists(GeneralMerge* G) {
  std::vector<uint32_t*> v;
  for (auto& c : G->cols)
    if (c.win) v.push_back(c.win);
  return v;
}

}  // namespace lsb
