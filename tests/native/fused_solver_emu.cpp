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
  // modes 3 / 4: three-phase plan (dd_plan.h) and its stage buffers
  int V, nS, nsplit3, sep_col0, pcols, ycols, prefetch;
  const double *M1, *M3, *M5;
  const void *strips1, *strips3, *strips5;
  const int *cta1, *chunks1, *cta3, *chunks3, *cta5, *chunks5;
  const int *gidx, *icol, *tptr, *tcol, *pcol, *srow;
  // mode 2 (five-phase form): sparse couplings A_SI (rows = separator positions) and A_BS (rows = interior
  // poses with a separator neighbour), block-CSR with permuted scalar columns; bcol = permuted column per row
  const int *si_rowptr, *si_colidx, *bs_rowptr, *bs_colidx, *bcol;
  const double *si_blocks, *bs_blocks;
  int nB;
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

alignas(128) unsigned char g_dsm[(kDd3DynSmem > kGemvDynSmem ? kDd3DynSmem : kGemvDynSmem)];

template <int R, int D, int MODE>
void thread_main(unsigned tid, FusedParams fp) {
  threadIdx.x = tid;
  k_rtr_fused<R, D, MODE>(fp);
}

template <int R, int D, int MODE>
int run(const EmuProblem &e) {
  const int dh = D + 1, N = dh * e.n;
  const size_t cols = (size_t)std::max(std::max(e.nsplit * e.KT, N), e.ycols) + 64;
  const size_t len = (size_t)R * cols;
  std::vector<std::vector<double>> vec(13, std::vector<double>(len, 0.0));
  std::vector<double> S((size_t)e.n * D * D, 0.0), S2(S), partials(64, 0.0);
  std::vector<double> zpart((size_t)std::max(e.nsplit, 1) * len, 0.0);
  std::vector<double> y((size_t)R * std::max(e.ycols, 64), 0.0), w((size_t)R * std::max(e.pcols, 64), 0.0);
  std::vector<double> zs((size_t)std::max(e.nsplit3, 1) * R * std::max(e.pcols, 64), 0.0);
  FusedOut out{};
  FusedParams fp{};
  fp.Q = BsrView{e.rowptr, e.colidx, e.blocks};
  fp.G = e.G;
  fp.Pinv = e.Pinv;
  fp.zpart = zpart.data();
  fp.ld = e.ld; fp.KT = e.KT; fp.nsplit = e.nsplit; fp.n = e.n;
  fp.zstride = len;
  fp.precon_mode = MODE;
  std::vector<double> tt((size_t)R * std::max(e.pcols, 64), 0.0), uu(tt);
  if (MODE == 2) {
    DdView &dd = fp.dd;
    dd.P1 = DdStripSet{e.M1, (const DdStrip *)e.strips1, e.cta1, e.chunks1, nullptr};
    dd.P3 = DdStripSet{e.M3, (const DdStrip *)e.strips3, e.cta3, e.chunks3, nullptr};
    dd.V = e.V; dd.nsplit1 = 1; dd.nsplit3 = e.nsplit3; dd.nS = e.nS; dd.nB = e.nB;
    dd.A_SI = BsrView{e.si_rowptr, e.si_colidx, e.si_blocks};
    dd.A_BS = BsrView{e.bs_rowptr, e.bs_colidx, e.bs_blocks};
    dd.pcol = e.pcol; dd.srow = e.srow; dd.bcol = e.bcol; dd.icol = e.icol;
    dd.sep_col0 = e.sep_col0; dd.pcols = e.pcols;
    dd.y = y.data(); dd.t = tt.data(); dd.zs = zs.data(); dd.u = uu.data(); dd.w = w.data();
    dd.prefetch = e.prefetch;
  }
  if (MODE >= 3) {
    DdView &dd = fp.dd;
    dd.P1 = DdStripSet{e.M1, (const DdStrip *)e.strips1, e.cta1, e.chunks1, nullptr};
    dd.P3 = DdStripSet{e.M3, (const DdStrip *)e.strips3, e.cta3, e.chunks3, nullptr};
    dd.P5 = DdStripSet{e.M5, (const DdStrip *)e.strips5, e.cta5, e.chunks5, e.gidx};
    dd.V = e.V; dd.nsplit1 = 1; dd.nsplit3 = e.nsplit3; dd.nS = e.nS;
    dd.pcol = e.pcol; dd.srow = e.srow; dd.icol = e.icol;
    dd.sep_col0 = e.sep_col0; dd.pcols = e.pcols;
    dd.y = y.data(); dd.zs = zs.data(); dd.w = w.data();
    dd.prefetch = e.prefetch;
    dd.tptr = e.tptr; dd.tcol = e.tcol;
  }
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
  if (e.mode == 2) return run<R, D, 2>(e);
  if (e.mode == 3) return run<R, D, 3>(e);
  if constexpr (D == 3) {
    if (e.mode == 4) return run<R, D, 4>(e);
  }
  return -1;
}

}  // namespace

extern "C" TP_EXPORT int fused_solve_emu(const EmuProblem *e) {
  if (e->R == 5 && e->d == 3) return by_mode<5, 3>(*e);
  if (e->R == 3 && e->d == 3) return by_mode<3, 3>(*e);
  if (e->R == 3 && e->d == 2) return by_mode<3, 2>(*e);
  return -1;
}
