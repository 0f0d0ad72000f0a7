// SPDX-License-Identifier: Apache-2.0
// pread scaling probe: T threads, each reading its own 8 x 4 MiB files from the page cache into a private buffer.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
static double now() { using namespace std::chrono; return duration<double>(steady_clock::now().time_since_epoch()).count(); }
int main(int argc, char** argv) {
  const char* dir = argc > 1 ? argv[1] : "/tmp";
  const size_t N = 4 << 20; const int F = 8; const int maxT = 32;
  std::vector<char> junk(N);
  for (size_t i = 0; i < N; ++i) junk[i] = (char)(i * 7);
  for (int t = 0; t < maxT; ++t) for (int f = 0; f < F; ++f) {
    std::string n = std::string(dir) + "/prb_" + std::to_string(t) + "_" + std::to_string(f);
    int fd = open(n.c_str(), O_CREAT | O_WRONLY | O_TRUNC, 0644); if (write(fd, junk.data(), N) != (ssize_t)N) return 1; close(fd);
  }
  for (int mode = 0; mode < 3; ++mode)
  for (int T : {1, 8, 16}) {
    std::vector<std::thread> th; std::vector<char*> bufs(T);
    for (int t = 0; t < T; ++t) { bufs[t] = (char*)aligned_alloc(4096, N * F); memset(bufs[t], 1, N * F); }
    const int reps = 6; double t0 = now();
    for (int t = 0; t < T; ++t) th.emplace_back([&, t] {
      for (int r = 0; r < reps; ++r) for (int f = 0; f < F; ++f) {
        std::string n = std::string(dir) + "/prb_" + std::to_string(t) + "_" + std::to_string(f);
        int fd = open(n.c_str(), O_RDONLY); size_t got = 0;
        if (mode == 0) {
          while (got < N) { ssize_t k = pread(fd, bufs[t] + f * N + got, N - got, got); if (k <= 0) break; got += k; }
        } else if (mode == 1) {
          void* m = mmap(nullptr, N, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0);
          memcpy(bufs[t] + f * N, m, N); munmap(m, N);
        } else {  // 256 KiB preads
          while (got < N) { size_t want = N - got < (256u << 10) ? N - got : (256u << 10); ssize_t k = pread(fd, bufs[t] + f * N + got, want, got); if (k <= 0) break; got += k; }
        }
        close(fd);
      }});
    for (auto& x : th) x.join();
    double dt = now() - t0;
    printf("mode %d threads %2d: %.1f GB/s aggregate, %.2f GB/s per thread\n", mode, T, (double)T * reps * F * N / dt / 1e9, (double)reps * F * N / dt / 1e9);
    for (auto b : bufs) free(b);
  }
  return 0;
}
