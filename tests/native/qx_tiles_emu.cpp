// Test infrastructure (not part of the product libraries): the Q*X device functions of dpgo_b200/csrc/kernels.cuh
// (phase_qx, its two-blocks-per-step form and the shared-memory tile-staged form) compiled for the host through
// tests/native/cuda_emu.h and run as one CTA of 256 real threads per virtual CTA.
#define QX_EXPORT __attribute__((visibility("default")))
#include "cuda_emu.h"

#include <vector>


#include "../../dpgo_b200/csrc/kernels.cuh"

using namespace dpgo;

namespace {
template <int R, int D>
int run(int variant, const int *rowptr, const int *colidx, const double *blocks, const double *X, const double *G,
        double *out, int n, int ctas) {
  static unsigned char stage[kWarpsPerBlock][QxTiles<R, D>::WARP_BYTES + 16];
  for (int wv = 0; wv < emu::kWarps; ++wv) emu::warp_barrier[wv] = new std::barrier<>(32);
  const BsrView Q{rowptr, colidx, blocks};
  for (int cta = 0; cta < ctas; ++cta) {
    std::vector<std::thread> th;
    for (unsigned t = 0; t < (unsigned)emu::kThreads; ++t)
      th.emplace_back([&, t]() {
        threadIdx.x = t;
        Ctx c;                                   // make_ctx() of virtual CTA `cta` out of `ctas`
        c.tid = cta * 256 + (int)t;
        c.nthreads = ctas * 256;
        c.warp = c.tid >> 5;
        c.nwarps = c.nthreads >> 5;
        c.lane = (int)t & 31;
        if (variant == 2) phase_qx_tiles<R, D>(c, Q, X, G, out, n, stage[t >> 5]);
        else if (variant == 3) phase_qx<R, D, true>(c, Q, X, G, out, n);
        else phase_qx<R, D, false>(c, Q, X, G, out, n);
      });
    for (auto &x : th) x.join();
  }
  for (int wv = 0; wv < emu::kWarps; ++wv) delete emu::warp_barrier[wv];
  return 0;
}
}  // namespace

extern "C" QX_EXPORT int qx_emu(int variant, int r, int d, const int *rowptr, const int *colidx, const double *blocks,
                                const double *X, const double *G, double *out, int n, int ctas) {
  if (r == 5 && d == 3) return run<5, 3>(variant, rowptr, colidx, blocks, X, G, out, n, ctas);
  if (r == 3 && d == 3) return run<3, 3>(variant, rowptr, colidx, blocks, X, G, out, n, ctas);
  if (r == 3 && d == 2) return run<3, 2>(variant, rowptr, colidx, blocks, X, G, out, n, ctas);
  if (r == 4 && d == 2) return run<4, 2>(variant, rowptr, colidx, blocks, X, G, out, n, ctas);
  return -1;
}
