// SPDX-License-Identifier: Apache-2.0
// CPU check of lakesoul_b200/csrc/zstd_decode.h (the decoder the GPU runs) against frames produced by Arrow's
// ZSTD codec: exact round trips, then mutated frames that must be rejected or decoded without ever touching
// memory out of range (build with -fsanitize=address,undefined).
// input file: u32 count, then per case u64 raw_len, u64 comp_len, raw bytes, compressed bytes
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "zstd_decode.h"

// the phased pipeline (scan / plan / entropy / execute) run step by step, the way the GPU kernel schedules it
static int64_t decompress_phased(const uint8_t* src, int64_t size, uint8_t* dst, int64_t cap) {
  using namespace lsb;
  std::vector<ZstdBlockDesc> blocks(4096);
  int64_t end = 0;
  const int64_t nb = zstd_scan_frame(src, size, blocks.data(), (int)blocks.size(), &end);
  if (nb < 0) return nb;
  std::vector<ZstdBlockPlan> plans((size_t)nb);
  std::vector<ZstdTables> tables((size_t)nb);
  static ZstdPlanScratch S;
  int prev = -1;
  int64_t lit_total = 0, seq_total = 0;
  for (int64_t k = 0; k < nb; ++k) {
    if (blocks[k].type != 2) continue;
    const int64_t r = zstd_plan_block(src + blocks[k].src_off, blocks[k].size, plans[k], tables[k], prev >= 0 ? &plans[prev] : nullptr,
                                      prev >= 0 ? &tables[prev] : nullptr, S);
    if (r < 0) return r;
    plans[k].lit_base = (int32_t)lit_total;
    plans[k].seq_base = (int32_t)seq_total;
    lit_total += plans[k].regen;
    seq_total += plans[k].nseq;
    prev = (int)k;
  }
  std::vector<uint8_t> lits((size_t)lit_total + 8);
  std::vector<uint64_t> seqs((size_t)seq_total + 1);
  // entropy stage: every (block, stream) and every block's sequence stream is independent
  for (int64_t k = 0; k < nb; ++k) {
    if (blocks[k].type != 2) continue;
    const ZstdBlockPlan& P = plans[k];
    const uint8_t* b = src + blocks[k].src_off;
    uint8_t* L = lits.data() + P.lit_base;
    if (P.lit_type == 0) memcpy(L, b + P.lit_src, (size_t)P.regen);
    else if (P.lit_type == 1) memset(L, b[P.lit_src], (size_t)P.regen);
    else {
      const int64_t per = P.streams == 4 ? (P.regen + 3) / 4 : P.regen;
      int64_t so = 0;
      for (int st = 0; st < P.streams; ++st) {
        const int64_t n = st < 3 ? (P.streams == 4 ? per : P.regen) : P.regen - 3 * per;
        const int64_t r = zstd_huf_stream(tables[k].huf, P.huf_log, b + P.lit_src + so, P.stream_size[st], L + st * per, n);
        if (r < 0) return r;
        so += P.stream_size[st];
      }
    }
    if (P.nseq > 0) {
      const int64_t r = zstd_decode_sequences(b + P.seq_src, P.seq_size, P, tables[k], seqs.data() + P.seq_base);
      if (r < 0) return r;
    }
  }
  // execute stage, in order
  uint32_t rep[3] = {1, 4, 8};
  int64_t op = 0;
  for (int64_t k = 0; k < nb; ++k) {
    const uint8_t* b = src + blocks[k].src_off;
    if (blocks[k].type == 0) {
      if (op + blocks[k].out_size > cap) return ZSTD_E_DST_SMALL;
      memcpy(dst + op, b, (size_t)blocks[k].out_size); op += blocks[k].out_size;
    } else if (blocks[k].type == 1) {
      if (op + blocks[k].out_size > cap) return ZSTD_E_DST_SMALL;
      memset(dst + op, b[0], (size_t)blocks[k].out_size); op += blocks[k].out_size;
    } else {
      const ZstdBlockPlan& P = plans[k];
      const uint8_t* L = lits.data() + P.lit_base;
      int64_t lpos = 0;
      for (int32_t q = 0; q < P.nseq; ++q) {
        const uint64_t v = seqs[(size_t)P.seq_base + q];
        const uint32_t ll = (uint32_t)(v & 0x1ffff), ml = (uint32_t)((v >> 17) & 0x1ffff), ov = (uint32_t)(v >> 34);
        const uint32_t off = zstd_resolve_offset(ov, ll, rep);
        if (off == 0 || lpos + ll > P.regen || op + ll + ml > cap) return ZSTD_E_CORRUPT;
        memcpy(dst + op, L + lpos, ll); op += ll; lpos += ll;
        if ((int64_t)off > op) return ZSTD_E_CORRUPT;
        for (uint32_t i = 0; i < ml; ++i) dst[op + i] = dst[op - off + i];
        op += ml;
      }
      const int64_t rest = P.regen - lpos;
      if (op + rest > cap) return ZSTD_E_DST_SMALL;
      memcpy(dst + op, L + lpos, (size_t)rest); op += rest;
    }
  }
  return op;
}

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 2;
  uint32_t n;
  if (fread(&n, 4, 1, f) != 1) return 2;
  static lsb::ZstdWork W;
  std::vector<uint8_t> lit(lsb::kZstdBlockMax);
  std::mt19937 rng(7);
  int bad = 0;
  long fuzz = 0, rejected = 0;
  for (uint32_t i = 0; i < n; ++i) {
    uint64_t dl, cl;
    if (fread(&dl, 8, 1, f) != 1 || fread(&cl, 8, 1, f) != 1) return 2;
    std::vector<uint8_t> d(dl), c(cl), o(dl ? dl : 1);
    if (dl && fread(d.data(), 1, dl, f) != dl) return 2;
    if (fread(c.data(), 1, cl, f) != cl) return 2;
    const int64_t r = lsb::zstd_decompress_frame(W, c.data(), (int64_t)cl, o.data(), (int64_t)dl, lit.data());
    const int64_t fcs = lsb::zstd_frame_content_size(c.data(), (int64_t)cl);
    std::vector<uint8_t> o2(dl ? dl : 1);
    const int64_t r2 = decompress_phased(c.data(), (int64_t)cl, o2.data(), (int64_t)dl);
    if (r2 != (int64_t)dl || (dl && memcmp(o2.data(), d.data(), dl) != 0)) {
      ++bad;
      printf("case %u: phased decoder -> %lld (want %llu)\n", i, (long long)r2, (unsigned long long)dl);
    }
    if (r != (int64_t)dl || (dl && memcmp(o.data(), d.data(), dl) != 0) || (fcs >= 0 && fcs != (int64_t)dl)) {
      ++bad;
      printf("case %u: raw %llu comp %llu -> %lld (frame says %lld)\n", i, (unsigned long long)dl, (unsigned long long)cl,
             (long long)r, (long long)fcs);
    }
    if (cl > 200000) continue;
    for (int t = 0; t < 12; ++t) {
      std::vector<uint8_t> m = c;
      const int kind = (int)(rng() % 3);
      if (kind == 0 && m.size() > 1) m.resize(rng() % m.size());
      else if (kind == 1) m[rng() % m.size()] ^= (uint8_t)(1u << (rng() % 8));
      else for (int k = 0; k < 3; ++k) m[rng() % m.size()] = (uint8_t)rng();
      std::vector<uint8_t> in(m.begin(), m.end());  // exact-size buffers: the sanitizer sees any stray access
      std::vector<uint8_t> out(dl ? dl : 1);
      const int64_t rr = lsb::zstd_decompress_frame(W, in.data(), (int64_t)in.size(), out.data(), (int64_t)dl, lit.data());
      ++fuzz;
      if (rr < 0) ++rejected;
      (void)decompress_phased(in.data(), (int64_t)in.size(), out.data(), (int64_t)dl);
    }
  }
  printf("%u frames, %d wrong; %ld mutated frames, %ld rejected\n", n, bad, fuzz, rejected);
  return bad ? 1 : 0;
}
