// Test infrastructure (not part of the product libraries): the persistent fused RTR solver kernel
// k_rtr_fused<R, D, MODE> of dpgo_b200/csrc/fused_kernel.cuh -- the same source nvcc compiles -- built for
// the host through tests/native/cuda_emu.h and run as ONE CTA of 256 real threads (grid barriers become CTA
// barriers; every phase walks all tiles / virtual CTAs).  tests/test_fused_solver_emu_cpu.py compares its
// iterates, iteration counts and objective with the oracle's optimize().
// built with -fvisibility=hidden -Wl,-Bsymbolic: the product library exports host stubs with the same mangled
// names as the kernels compiled here, and must not interpose them when both are loaded in one process
#define TP_EXPORT __attribute__((visibility("default")))
#include "cuda_emu.h"

#include <vector>

#include "../../dpgo_b200/csrc/fused_kernel.cuh"

using namespace dpgo;

extern "C" {

struct EmuProblem {
  int mode, R, d, n;
  // block-CSR Q (blocks row-major) and the linear term G (column-major R x N)
  const int *rowptr, *colidx;
  const double *blocks, *G;
  // mode 0: stage-major tiled dense inverse (k_tile_layout format), ld x ldk
  const double *Pinv;
  int ld, KT, nsplit;
  // solver parameters (dpgo_ropt_params)
  double gradnorm_tol, init_radius, theta, kappa, accept_rho, shrink, magnify;
  int max_outer, max_inner;
  // in / out
  const double *x_in;
  double *x_out;
  double *result;   // [16]: f_init, gn_init, f_opt, gn_opt, outer, inner, accepted, rejected, tcg_status,
                    //       returned_initial, n_qx, n_precon, n_sweeps, n_barriers
};

}  // extern "C"

namespace {

alignas(128) unsigned char g_dsm[kGemvDynSmem];

template <int R, int D, int MODE>
void thread_main(unsigned tid, FusedParams fp) {
  threadIdx.x = tid;
  k_rtr_fused<R, D, MODE>(fp);
}

template <int R, int D, int MODE>
int run(const EmuProblem &e) {
  const int dh = D + 1, N = dh * e.n;
  const size_t cols = (size_t)std::max(e.nsplit * e.KT, N) + 64;
  const size_t len = (size_t)R * cols;
  std::vector<std::vector<double>> vec(13, std::vector<double>(len, 0.0));
  std::vector<double> S((size_t)e.n * D * D, 0.0), S2(S), partials(64, 0.0);
  std::vector<double> zpart((size_t)std::max(e.nsplit, 1) * len, 0.0);
  FusedOut out{};
  FusedParams fp{};
  fp.Q = BsrView{e.rowptr, e.colidx, e.blocks};
  fp.G = e.G;
  fp.Pinv = e.Pinv;
  fp.zpart = zpart.data();
  fp.ld = e.ld; fp.KT = e.KT; fp.nsplit = e.nsplit; fp.n = e.n;
  fp.zstride = len;
  fp.precon_mode = MODE;
  fp.x_in = e.x_in; fp.x_out = e.x_out;
  double **slots[] = {&fp.xa, &fp.xb, &fp.EG, &fp.EG2, &fp.grad, &fp.grad2, &fp.eta, &fp.r, &fp.z, &fp.delta, &fp.Hd};
  for (int i = 0; i < 11; ++i) *slots[i] = vec[i].data();
  fp.S = S.data(); fp.S2 = S2.data();
  fp.partials = partials.data();
  fp.out = &out;
  fp.trace = nullptr;
  fp.gradnorm_tol = e.gradnorm_tol; fp.init_radius = e.init_radius; fp.theta = e.theta; fp.kappa = e.kappa;
  fp.accept_rho = e.accept_rho; fp.shrink = e.shrink; fp.magnify = e.magnify;
  fp.max_outer = e.max_outer; fp.max_inner = e.max_inner;
  emu::dsm = g_dsm;
  for (int wv = 0; wv < emu::kWarps; ++wv) emu::warp_barrier[wv] = new std::barrier<>(32);
  std::vector<std::thread> th;
  for (unsigned t = 0; t < (unsigned)emu::kThreads; ++t) th.emplace_back(thread_main<R, D, MODE>, t, fp);
  for (auto &t : th) t.join();
  for (int wv = 0; wv < emu::kWarps; ++wv) delete emu::warp_barrier[wv];
  const double res[14] = {out.f_init, out.gn_init, out.f_opt, out.gn_opt, (double)out.outer, (double)out.inner,
                          (double)out.accepted, (double)out.rejected, (double)out.tcg_status,
                          (double)out.returned_initial, (double)out.n_qx, (double)out.n_precon,
                          (double)out.n_sweeps, (double)out.n_barriers};
  for (int i = 0; i < 14; ++i) e.result[i] = res[i];
  return 0;
}

template <int R, int D>
int by_mode(const EmuProblem &e) {
  if (e.mode == 0) return run<R, D, 0>(e);
  return -1;
}

}  // namespace

extern "C" TP_EXPORT int fused_solve_emu(const EmuProblem *e) {
  if (e->R == 5 && e->d == 3) return by_mode<5, 3>(*e);
  if (e->R == 3 && e->d == 3) return by_mode<3, 3>(*e);
  if (e->R == 3 && e->d == 2) return by_mode<3, 2>(*e);
  return -1;
}
