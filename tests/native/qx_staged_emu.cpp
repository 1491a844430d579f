// Test infrastructure (not part of the product libraries): the staged Q*X device function of
// dpgo_b200/csrc/qx_staged.cuh compiled for the host through tests/native/cuda_emu.h; the asynchronous copies
// land at once (a memcpy), commit / wait are no-ops, warps are real threads, so the step generator, the
// buffer rotation, the piece dealing and the shuffle tree are the ones the device runs.
#define QX_EXPORT __attribute__((visibility("default")))
#include "cuda_emu.h"

#include <vector>

inline void __syncwarp() { emu::warp_barrier[threadIdx.x >> 5]->arrive_and_wait(); }
template <int BYTES, bool STREAM = false>
inline void cp_async(void *dst, const void *src) { std::memcpy(dst, src, BYTES); }
inline void cp_async_commit() {}
template <int N>
inline void cp_async_wait() {}

#include "../../dpgo_b200/csrc/qx_staged.cuh"

using namespace dpgo;

namespace {
template <int R, int D>
int run(const int *rowptr, const int *colidx, const double *blocks, const double *X, const double *G, double *out,
        int n, int ctas) {
  std::vector<unsigned char> smem(QxGeo<R, D>::CTA_BYTES + 128);
  unsigned char *base = smem.data() + (128 - (reinterpret_cast<uintptr_t>(smem.data()) & 127)) % 128;
  for (int wv = 0; wv < emu::kWarps; ++wv) emu::warp_barrier[wv] = new std::barrier<>(32);
  const BsrView Q{rowptr, colidx, blocks};
  for (int cta = 0; cta < ctas; ++cta) {
    std::vector<std::thread> th;
    for (unsigned t = 0; t < (unsigned)emu::kThreads; ++t)
      th.emplace_back([&, t]() {
        threadIdx.x = t;
        phase_qx_staged<R, D>(Q, X, G, out, n, base, cta * kWarpsPerBlock + (int)(t >> 5), ctas * kWarpsPerBlock);
      });
    for (auto &x : th) x.join();
  }
  for (int wv = 0; wv < emu::kWarps; ++wv) delete emu::warp_barrier[wv];
  return 0;
}
}  // namespace

extern "C" QX_EXPORT int qx_staged_emu(int r, int d, const int *rowptr, const int *colidx, const double *blocks,
                                       const double *X, const double *G, double *out, int n, int ctas) {
  if (r == 5 && d == 3) return run<5, 3>(rowptr, colidx, blocks, X, G, out, n, ctas);
  if (r == 3 && d == 3) return run<3, 3>(rowptr, colidx, blocks, X, G, out, n, ctas);
  if (r == 3 && d == 2) return run<3, 2>(rowptr, colidx, blocks, X, G, out, n, ctas);
  if (r == 4 && d == 2) return run<4, 2>(rowptr, colidx, blocks, X, G, out, n, ctas);
  return -1;
}
