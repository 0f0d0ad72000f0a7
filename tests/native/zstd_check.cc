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
    }
  }
  printf("%u frames, %d wrong; %ld mutated frames, %ld rejected\n", n, bad, fuzz, rejected);
  return bad ? 1 : 0;
}
