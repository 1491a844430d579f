// Test infrastructure (not part of the product libraries): the device functions of
// dpgo_b200/csrc/dense_la.cuh and the launch sequences of dense_la_seq.h compiled for the host through
// tests/native/cuda_emu.h.  A "launch" walks its grid one CTA at a time with 256 real threads
// (__syncthreads is a barrier), so the descriptor arithmetic, the k-range modes and the in-place
// aliasing rules are the ones the device runs.
#define DLA_EXPORT __attribute__((visibility("default")))
#include "cuda_emu.h"

#include <cstdlib>
#include <functional>
#include <mutex>
#include <vector>


inline int atomicMax(int *p, int v) {
  static std::mutex m;
  std::lock_guard<std::mutex> g(m);
  const int old = *p;
  if (v > old) *p = v;
  return old;
}

#include "../../dpgo_b200/csrc/dense_la_seq.h"

using namespace dpgo::dla;

namespace {

// run body(bx, by, bz) for every CTA of the grid, each with 256 threads
void launch(int gx, int gy, int gz, const std::function<void(int, int, int)> &body) {
  for (int wv = 0; wv < emu::kWarps; ++wv)
    if (!emu::warp_barrier[wv]) emu::warp_barrier[wv] = new std::barrier<>(32);
  std::vector<std::thread> th;
  for (unsigned t = 0; t < (unsigned)emu::kThreads; ++t)
    th.emplace_back([&, t]() {
      threadIdx.x = t;
      for (int z = 0; z < gz; ++z)
        for (int y = 0; y < gy; ++y)
          for (int x = 0; x < gx; ++x) {
            body(x, y, z);
            __syncthreads();
          }
    });
  for (auto &t : th) t.join();
}

struct EmuBackend {
  void *alloc(size_t bytes) { return std::calloc(bytes ? bytes : 1, 1); }
  void release(void *p) { std::free(p); }
  bool upload(void *dst, const void *src, size_t bytes) { std::memcpy(dst, src, bytes); return true; }
  bool download(void *dst, const void *src, size_t bytes) { std::memcpy(dst, src, bytes); return true; }
  void diag(const SpdDesc *d, int panel, int count, int *info) {
    static double sL[TS * (TS + 1)], sW[TS * (TS + 1)];
    launch(1, 1, count, [&](int, int, int z) { dla_diag_block(d[z], panel, z, sL, sW, info); });
  }
  void gemm(const GemmDesc *g, int tiles_m, int tiles_n, int count, const GemmFlags &f) {
    static double sA[2 * BK * LDS], sB[2 * BK * LDS];
    launch(tiles_m, tiles_n, count, [&](int x, int y, int z) { dla_gemm_tile(g[z], f, x, y, sA, sB); });
  }
  void copy(const SpdDesc *d, int what, int tiles, int count) {
    launch(tiles, what == 0 ? 1 : tiles, count,
           [&](int x, int y, int z) { dla_copy_tile(d[z], what, x, what == 0 ? x : y); });
  }
};

}  // namespace

// count matrices packed back to back in A (matrix b: n[b] x n[b], leading dimension n[b])
extern "C" DLA_EXPORT int dla_emu_spd_inverse(double *A, const int *n, int count, int symmetrize) {
  std::vector<SeqItem> items(count);
  size_t off = 0;
  for (int b = 0; b < count; ++b) {
    items[b] = SeqItem{A + off, n[b], n[b] > 0 ? n[b] : 1};
    off += (size_t)n[b] * n[b];
  }
  EmuBackend be;
  return spd_inverse_seq(be, items.data(), count, symmetrize != 0);
}

extern "C" DLA_EXPORT int dla_emu_gemm(const double *A, const double *B, double *C, int M, int N, int K, int lda, int ldb,
                                       int ldc, int ta, int tb, int lower_only, int kmode, double alpha, double beta) {
  GemmDesc g{A, B, C, M, N, K, lda, ldb, ldc};
  EmuBackend be;
  be.gemm(&g, tiles_of(M), tiles_of(N), 1, GemmFlags{ta, tb, lower_only, kmode, alpha, beta});
  return 0;
}
