// SPDX-License-Identifier: Apache-2.0
// Host framing (page_plan.cc: footer, page-header walk, row-group / page pruning, lazy segment-wise reads) compiled with
// AddressSanitizer + UBSan by tests/test_framing_sanitized.py and driven through its test hooks.
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>
namespace lsb { std::shared_ptr<uint8_t> pinned_alloc(size_t bytes) { return std::shared_ptr<uint8_t>(new uint8_t[bytes], [](uint8_t* p) { delete[] p; }); } }
extern "C" int64_t lsb_test_page_prune(const char*, const char*, const char*, int32_t, char*, int64_t);
extern "C" int64_t lsb_test_plan_digest(const char*, int32_t, int32_t, int32_t, char*, int64_t);
int main(int argc, char** argv) {
  // argv: file pk filter...
  std::vector<char> out(1 << 24);
  const char* file = argv[1];
  const char* pk = argv[2];
  int bad = 0;
  for (int lazy = 0; lazy < 2; ++lazy) {
    int64_t n = lsb_test_plan_digest(file, lazy, 0, 1 << 30, out.data(), out.size());
    if (n < 0) { printf("digest lazy=%d: %s\n", lazy, out.data()); }
    for (int i = 3; i < argc; ++i) {
      n = lsb_test_page_prune(file, argv[i], pk, lazy, out.data(), out.size());
      if (n < 0) { printf("prune lazy=%d %s: %s\n", lazy, argv[i], out.data()); ++bad; }
    }
  }
  printf("done bad=%d\n", bad);
  return 0;
}
