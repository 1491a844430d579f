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

namespace {
// Hessian pass of the tCG direction: variant 0 = direction update as its own pass (phase_axpby) followed by
// phase_hess; variant 1 = phase_hess_dir (direction formed on the fly).  One virtual CTA; acc = per-thread partial
// sums added up in thread order.
template <int R, int D>
int run_hess(int variant, const int *rowptr, const int *colidx, const double *blocks, const double *Y, const double *S,
             const double *Z, const double *Dold, double beta, double *Dnew, double *HV, const int *pcol, double *HVp,
             double *acc_out, int n) {
  for (int wv = 0; wv < emu::kWarps; ++wv) emu::warp_barrier[wv] = new std::barrier<>(32);
  const BsrView Q{rowptr, colidx, blocks};
  const size_t len = (size_t)R * (D + 1) * n;
  static double part[emu::kThreads][2];
  for (int pass = (variant == 0 ? 0 : 1); pass < 2; ++pass) {
    std::vector<std::thread> th;
    for (unsigned t = 0; t < (unsigned)emu::kThreads; ++t)
      th.emplace_back([&, t]() {
        threadIdx.x = t;
        Ctx c;
        c.tid = (int)t; c.nthreads = 256; c.warp = c.tid >> 5; c.nwarps = 8; c.lane = (int)t & 31;
        double acc[2] = {0.0, 0.0};
        if (variant == 0 && pass == 0) {
          for (size_t k = t; k < len; k += 256) Dnew[k] = Dold[k];
          emu::cta_barrier.arrive_and_wait();
          phase_axpby(c, -1.0, Z, beta, Dnew, len);
        } else if (variant == 0) {
          phase_hess<R, D>(c, Q, Y, S, Dnew, HV, nullptr, n, acc, pcol, HVp);
        } else {
          phase_hess_dir<R, D>(c, Q, Y, S, Z, Dold, beta, Dnew, HV, n, acc, pcol, HVp);
        }
        part[t][0] = acc[0]; part[t][1] = acc[1];
      });
    for (auto &x : th) x.join();
  }
  acc_out[0] = acc_out[1] = 0.0;
  for (int t = 0; t < emu::kThreads; ++t) { acc_out[0] += part[t][0]; acc_out[1] += part[t][1]; }
  for (int wv = 0; wv < emu::kWarps; ++wv) delete emu::warp_barrier[wv];
  return 0;
}
}  // namespace

extern "C" QX_EXPORT int hess_emu(int variant, int r, int d, const int *rowptr, const int *colidx, const double *blocks,
                                  const double *Y, const double *S, const double *Z, const double *Dold, double beta,
                                  double *Dnew, double *HV, const int *pcol, double *HVp, double *acc_out, int n) {
  if (r == 5 && d == 3) return run_hess<5, 3>(variant, rowptr, colidx, blocks, Y, S, Z, Dold, beta, Dnew, HV, pcol, HVp, acc_out, n);
  if (r == 3 && d == 2) return run_hess<3, 2>(variant, rowptr, colidx, blocks, Y, S, Z, Dold, beta, Dnew, HV, pcol, HVp, acc_out, n);
  return -1;
}

extern "C" QX_EXPORT int qx_emu(int variant, int r, int d, const int *rowptr, const int *colidx, const double *blocks,
                                const double *X, const double *G, double *out, int n, int ctas) {
  if (r == 5 && d == 3) return run<5, 3>(variant, rowptr, colidx, blocks, X, G, out, n, ctas);
  if (r == 3 && d == 3) return run<3, 3>(variant, rowptr, colidx, blocks, X, G, out, n, ctas);
  if (r == 3 && d == 2) return run<3, 2>(variant, rowptr, colidx, blocks, X, G, out, n, ctas);
  if (r == 4 && d == 2) return run<4, 2>(variant, rowptr, colidx, blocks, X, G, out, n, ctas);
  return -1;
}
