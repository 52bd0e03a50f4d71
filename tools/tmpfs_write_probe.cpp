// tools/tmpfs_write_probe.cpp -- how fast can N threads fill a NEW file on this box's tmpfs?  The ceiling of
// `tbc_encode -u` at the RangeFit levels, where the GPU needs 0.05 s for what the host needs 0.4 s to write
// (DESIGN.md s7).  Development aid.
//   g++ -O2 -pthread -o tools/_bin/tmpfs_write_probe tools/tmpfs_write_probe.cpp
//   tools/_bin/tmpfs_write_probe <mode> <threads> <MiB> <path>      mode 0: memcpy into a shared mapping (what the slab
//   pipeline does)   1: pwrite at disjoint offsets   2: MADV_POPULATE_WRITE then memcpy   3: MADV_HUGEPAGE then memcpy
#include <sys/mman.h>
#include <fcntl.h>
#include <unistd.h>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#ifndef MADV_POPULATE_WRITE
#define MADV_POPULATE_WRITE 23
#endif

int main(int argc, char** argv)
{
  if (argc < 5) { std::fprintf(stderr, "usage: %s mode threads MiB path\n", argv[0]); return 2; }
  const int mode = std::atoi(argv[1]), nt = std::atoi(argv[2]);
  const size_t total = (size_t)std::atol(argv[3]) << 20, chunk = (size_t)4 << 20;
  const char* path = argv[4];
  char* src = static_cast<char*>(std::malloc(chunk));
  std::memset(src, 1, chunk);
  ::unlink(path);
  const int fd = ::open(path, O_RDWR | O_CREAT | O_TRUNC, 0644);
  if (fd < 0) { std::perror("open"); return 1; }
  const auto t0 = std::chrono::steady_clock::now();
  if (::ftruncate(fd, (off_t)total)) { std::perror("ftruncate"); return 1; }
  char* map = nullptr;
  if (mode != 1) {
    map = static_cast<char*>(::mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0));
    if (map == MAP_FAILED) { std::perror("mmap"); return 1; }
    if (mode == 3 && ::madvise(map, total, MADV_HUGEPAGE)) std::perror("madvise(MADV_HUGEPAGE)");
  }
  std::atomic<size_t> next{0};
  std::vector<std::thread> pool;
  for (int t = 0; t < nt; ++t)
    pool.emplace_back([&] {
      for (;;) {
        const size_t off = next.fetch_add(chunk);
        if (off >= total) break;
        const size_t n = std::min(chunk, total - off);
        if (mode == 1) { if (::pwrite(fd, src, n, (off_t)off) != (ssize_t)n) std::perror("pwrite"); continue; }
        if (mode == 2 && ::madvise(map + off, n, MADV_POPULATE_WRITE)) std::perror("madvise(MADV_POPULATE_WRITE)");
        std::memcpy(map + off, src, n);
      }
    });
  for (std::thread& t : pool) t.join();
  const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  std::printf("{\"mode\": %d, \"threads\": %d, \"mib\": %zu, \"gb_per_s\": %.2f}\n", mode, nt, total >> 20, (double)total / s / 1e9);
  if (map) ::munmap(map, total);
  ::close(fd);
  ::unlink(path);
  return 0;
}
