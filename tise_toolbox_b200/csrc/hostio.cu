// Host-side helpers of the file ingest: no device code.  The Python loader used to stat, read and
// probe every file of a decode group from Python threads -- 3 system calls and several GIL hand-offs per file, ~45 us per
// 150 KB file on one thread and poor scaling beyond four (tools measurements in profiles/README.md) -- which bounded the
// command line at ~37 k files/s.  These entry points do a whole group per call with the GIL released: plain POSIX reads
// into the caller's (pinned) buffer on a few std::threads, then the PNG / JPEG container walks.
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <thread>
#include <vector>

#include "common.cuh"

namespace {

template <typename F>
void parallel_chunks(int n, int n_threads, F&& body) {  // body(i) for i in [0, n), dealt out in runs of 16
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 64) n_threads = 64;
  std::atomic<int> next{0};
  auto work = [&]() {
    for (;;) {
      const int i0 = next.fetch_add(16);
      if (i0 >= n) break;
      const int i1 = i0 + 16 < n ? i0 + 16 : n;
      for (int i = i0; i < i1; ++i) body(i);
    }
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < n_threads && t * 16 < n + 16; ++t) pool.emplace_back(work);
  work();
  for (auto& th : pool) th.join();
}

}  // namespace

extern "C" int tise_file_sizes(const char* const* paths, int n, int64_t* sizes, int n_threads) {
  if (n < 0 || (n > 0 && (!paths || !sizes))) return TISE_ERR_INVALID;
  std::atomic<int> bad{-1};
  parallel_chunks(n, n_threads, [&](int i) {
    struct stat st;
    if (!paths[i] || stat(paths[i], &st) != 0 || !S_ISREG(st.st_mode)) {
      sizes[i] = -1;
      bad.store(i);
    } else {
      sizes[i] = (int64_t)st.st_size;
    }
  });
  return bad.load() < 0 ? TISE_OK : TISE_ERR_INVALID;
}

extern "C" int tise_read_files(const char* const* paths, const int64_t* offsets, const int64_t* sizes, int n, uint8_t* dst,
                               int n_threads, int* failed_index) {
  if (failed_index) *failed_index = -1;
  if (n < 0 || (n > 0 && (!paths || !offsets || !sizes || !dst))) return TISE_ERR_INVALID;
  std::atomic<int> bad{-1};
  parallel_chunks(n, n_threads, [&](int i) {
    const int fd = open(paths[i], O_RDONLY | O_CLOEXEC);
    if (fd < 0) {
      bad.store(i);
      return;
    }
    int64_t done = 0;
    while (done < sizes[i]) {
      const ssize_t got = read(fd, dst + offsets[i] + done, (size_t)(sizes[i] - done));
      if (got <= 0) break;
      done += got;
    }
    close(fd);
    if (done != sizes[i]) bad.store(i);  // shorter than its stat size: the file changed under us
  });
  if (bad.load() >= 0) {
    if (failed_index) *failed_index = bad.load();
    return TISE_ERR_INVALID;
  }
  return TISE_OK;
}

extern "C" int tise_probe_many(const uint8_t* h_base, const int64_t* offsets, const int64_t* sizes, int n, int use_jpeg,
                               tise_png_info* png_infos, tise_jpeg_info* jpeg_infos, int8_t* kind, int32_t* hw, int n_threads) {
  if (n < 0 || (n > 0 && (!h_base || !offsets || !sizes || !png_infos || !jpeg_infos || !kind || !hw))) return TISE_ERR_INVALID;
  parallel_chunks(n, n_threads, [&](int i) {
    kind[i] = 0;
    hw[2 * i] = hw[2 * i + 1] = 0;
    const uint8_t* p = h_base + offsets[i];
    if (tise_png_probe(p, sizes[i], &png_infos[i]) == TISE_OK) {
      if (png_infos[i].supported) {
        kind[i] = 1;
        hw[2 * i] = png_infos[i].height;
        hw[2 * i + 1] = png_infos[i].width;
      }
    } else if (use_jpeg && tise_jpeg_probe(p, sizes[i], &jpeg_infos[i]) == TISE_OK && jpeg_infos[i].supported) {
      kind[i] = 2;
      hw[2 * i] = jpeg_infos[i].height;
      hw[2 * i + 1] = jpeg_infos[i].width;
    }
  });
  return TISE_OK;
}
