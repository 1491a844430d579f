// Test infrastructure (not part of the product libraries): one application of the three-phase
// preconditioner executed by the DEVICE functions of dpgo_b200/csrc/kernels.cuh themselves
// (phase_strip_gemv with its staging modes and its fused-finish epilogue, phase_dd_finish,
// phase_dd_finish_sep, block_reduce_store), compiled for the host through tests/native/cuda_emu.h and
// run as one CTA of 256 real threads that walks every virtual CTA of the plan.
// built with -fvisibility=hidden -Wl,-Bsymbolic: the product library exports host stubs with the same mangled
// names as the kernels compiled here, and must not interpose them when both are loaded in one process
#define TP_EXPORT __attribute__((visibility("default")))
#include "cuda_emu.h"

#include <vector>

#include "../../dpgo_b200/csrc/kernels.cuh"

using namespace dpgo;

namespace {

struct Args {
  DdView dd;
  const double *Y, *rvec;
  double *z, *neg_out, *zr;
  int n, fused_finish;
};

alignas(128) unsigned char g_dsm[kDd3DynSmem];

template <int R, int D>
void cta_thread(unsigned tid, const Args *a) {
  threadIdx.x = tid;
  const DdView &dd = a->dd;
  constexpr int ST = kDd3Stages;
  GemvPipe pipe = gemv_pipe_init<ST, ST>(g_dsm);
  static StripPlanStore s_plan[3];
  strip_plan_fill(&s_plan[0], dd.P1, dd.V);
  strip_plan_fill(&s_plan[1], dd.P3, dd.V);
  strip_plan_fill(&s_plan[2], dd.P5, dd.V);
  const Ctx ctx = make_ctx();
  const size_t zs = (size_t)dd.pcols * R;
  const bool pf = dd.prefetch != 0;
  double acc[1] = {0.0};
  phase_strip_gemv<R, ST, 1>(pipe, dd.P1, dd.V, &s_plan[0], a->rvec, dd.icol, dd.y, 0);
  if (dd.nS > 0) {
    if (pf) strip_prefetch<ST>(pipe, dd.P3, dd.V, &s_plan[1]);
    __syncthreads();   // grid barrier
    const StageAux a3{dd.y, dd.tptr, dd.tcol, dd.sep_col0, 0, 0};
    phase_strip_gemv<R, ST, 2>(pipe, dd.P3, dd.V, &s_plan[1], a->rvec, dd.icol, dd.zs, zs, pf, &a3);
    if (pf) strip_prefetch<ST>(pipe, dd.P5, dd.V, &s_plan[2]);
    __syncthreads();   // grid barrier
    const StageAux a5{nullptr, nullptr, nullptr, 0, dd.nsplit3, zs};
    if constexpr (D == 3) {
      if (a->fused_finish) {
        StripFinish fin{dd.y, dd.icol, a->Y, a->rvec, a->z, a->neg_out, 0.0};
        phase_strip_gemv<R, ST, 3, D>(pipe, dd.P5, dd.V, &s_plan[2], dd.zs, nullptr, dd.w, 0, pf, &a5, &fin);
        acc[0] += fin.acc;
        phase_dd_finish_sep<R, D>(ctx, dd, a->Y, a->rvec, a->z, a->neg_out, acc);
        block_reduce_store<1>(acc, a->zr);
        return;
      }
    }
    phase_strip_gemv<R, ST, 3>(pipe, dd.P5, dd.V, &s_plan[2], dd.zs, nullptr, dd.w, 0, pf, &a5);
  }
  __syncthreads();     // grid barrier
  phase_dd_finish<R, D>(ctx, dd, a->Y, a->rvec, a->z, a->neg_out, a->n, acc);
  block_reduce_store<1>(acc, a->zr);
}

template <int R, int D>
void run_cta(const Args &a) {
  emu::dsm = g_dsm;
  for (int w = 0; w < emu::kWarps; ++w) emu::warp_barrier[w] = new std::barrier<>(32);
  std::vector<std::thread> th;
  for (unsigned t = 0; t < (unsigned)emu::kThreads; ++t) th.emplace_back(cta_thread<R, D>, t, &a);
  for (auto &t : th) t.join();
  for (int w = 0; w < emu::kWarps; ++w) delete emu::warp_barrier[w];
}

}  // namespace

extern "C" {

// strips*: arrays of DdStrip {int cb, kc0, nchunks, slot; long long data_off}; y / zs / w: zero-initialised work
// arrays of R*ycols, nsplit3*R*pcols, R*pcols doubles.  (R, d) in {(5,3), (3,3), (3,2)}.
TP_EXPORT int tp_apply_device_emu(int R, int d, int n, int V, int nS, int nsplit3, int sep_col0, int pcols, int prefetch,
                        const double *M1, const void *strips1, const int *cta1, const int *chunks1,
                        const double *M3, const void *strips3, const int *cta3, const int *chunks3,
                        const double *M5, const void *strips5, const int *cta5, const int *chunks5,
                        const int *gidx, const int *icol, const int *tptr, const int *tcol, const int *pcol,
                        const int *srow, double *y, double *zs, double *w, const double *Y, const double *rvec,
                        double *z, double *neg_out, double *zr, int fused_finish) {
  Args a{};
  DdView &dd = a.dd;
  dd.P1 = DdStripSet{M1, (const DdStrip *)strips1, cta1, chunks1, nullptr};
  dd.P3 = DdStripSet{M3, (const DdStrip *)strips3, cta3, chunks3, nullptr};
  dd.P5 = DdStripSet{M5, (const DdStrip *)strips5, cta5, chunks5, gidx};
  dd.V = V; dd.nsplit1 = 1; dd.nsplit3 = nsplit3;
  dd.nS = nS; dd.nB = 0;
  dd.pcol = pcol; dd.srow = srow; dd.bcol = nullptr; dd.icol = icol;
  dd.sep_col0 = sep_col0; dd.pcols = pcols;
  dd.y = y; dd.t = nullptr; dd.zs = zs; dd.u = nullptr; dd.w = w;
  dd.prefetch = prefetch;
  dd.tptr = tptr; dd.tcol = tcol;
  a.Y = Y; a.rvec = rvec; a.z = z; a.neg_out = neg_out; a.zr = zr; a.n = n; a.fused_finish = fused_finish;
  if (R == 5 && d == 3) run_cta<5, 3>(a);
  else if (R == 3 && d == 3) run_cta<3, 3>(a);
  else if (R == 3 && d == 2) run_cta<3, 2>(a);
  else return -1;
  return 0;
}

// The RGD step's fused kernel (phase_retract_impl<R, D, true>: Xout = Retraction_X(s * Dir)) beside the
// two-pass form it replaces (Eta = s * Dir rounded, then phase_retract): both outputs, to be compared bit for bit.
TP_EXPORT int emu_retract_scaled(int R, int d, int n, const double *X, const double *Dir, double s, double *out_fused,
                                 double *out_two_pass) {
  if (!(R == 5 && d == 3) && !(R == 3 && d == 2)) return -1;
  const size_t len = (size_t)R * (d + 1) * n;
  std::vector<double> eta(len);
  for (size_t k = 0; k < len; ++k) eta[k] = s * Dir[k];
  for (int wv = 0; wv < emu::kWarps; ++wv) emu::warp_barrier[wv] = new std::barrier<>(32);
  std::vector<std::thread> th;
  for (unsigned t = 0; t < (unsigned)emu::kThreads; ++t)
    th.emplace_back([=, &eta]() {
      threadIdx.x = t;
      const Ctx ctx = make_ctx();
      if (R == 5) {
        phase_retract_impl<5, 3, true>(ctx, X, Dir, out_fused, n, s);
        phase_retract<5, 3>(ctx, X, eta.data(), out_two_pass, n);
      } else {
        phase_retract_impl<3, 2, true>(ctx, X, Dir, out_fused, n, s);
        phase_retract<3, 2>(ctx, X, eta.data(), out_two_pass, n);
      }
    });
  for (auto &t : th) t.join();
  for (int wv = 0; wv < emu::kWarps; ++wv) delete emu::warp_barrier[wv];
  return 0;
}

}  // extern "C"
