// The persistent fused RTR solver kernel (see fused_rtr.cu for the host side): one cooperative launch
// runs the whole QuadraticOptimizer::optimize call.  Kept in a header so that the same source is compiled
// by nvcc for the product library and -- with DPGO_CPU_EMU, tests/native/cuda_emu.h -- by g++ for a CPU
// test that runs the kernel as one CTA of real threads against the oracle.
#pragma once
#ifndef DPGO_CPU_EMU
#include <cooperative_groups.h>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;
#define DPGO_DYNAMIC_SMEM(name) extern __shared__ __align__(128) unsigned char name[]
#endif
#include <math.h>

#include "kernels.cuh"
#include "rtr_logic.h"

namespace dpgo {

struct FusedOut {
  double f_init, gn_init, f_opt, gn_opt;
  int outer, inner, accepted, rejected, tcg_status, returned_initial;
  long long n_qx, n_precon, n_sweeps, n_barriers;
  double phase_ms[16];
};

struct FusedParams {
  BsrView Q;
  const double *G;
  const double *Pinv;
  double *zpart;
  int ld, KT, nsplit, n;
  size_t zstride;
  int precon_mode;
  DdView dd;                              // two-level variant (precon_mode == 2)
  const double *x_in;
  double *x_out;
  double *xa, *xb, *EG, *EG2, *grad, *grad2, *S, *S2, *eta, *r, *z, *delta, *Hd;
  double *partials;  // [2][gridDim.x][4]
  FusedOut *out;
  unsigned long long *trace;   // -DDPGO_TRACE builds: [gridDim.x][16] ns each CTA worked in a phase before its barrier
  double gradnorm_tol, init_radius, theta, kappa, accept_rho, shrink, magnify;
  int max_outer, max_inner;
};

#ifndef DPGO_CPU_EMU
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#endif

// per-phase device time seen by CTA 0 (phase body + the barrier that ends it); the accumulators
// live in shared memory so that they do not occupy registers across the phases
struct PhaseClock {
  unsigned long long *acc;  // [17] in shared memory: 16 phase sums + last timestamp
#ifdef DPGO_TRACE
  // measurement builds: per-CTA time from the release of one barrier to the arrival at the next (the
  // CTA's own work in the phase), accumulated per phase id; `pending` is filled by GridReducer
  unsigned long long *busy;            // [16] in shared memory
  const unsigned long long *pending;
#endif
  __device__ __forceinline__ void start(unsigned long long *smem) {
    acc = smem;
    if (threadIdx.x == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = 0;
      acc[16] = gtimer();
    }
  }
  __device__ __forceinline__ void lap(int id) {
    if (threadIdx.x == 0) {
      const unsigned long long t = gtimer();
      acc[id] += t - acc[16];
      acc[16] = t;
#ifdef DPGO_TRACE
      busy[id] += *pending;
#endif
    }
  }
};

// Grid-wide barrier of the persistent solver: cooperative-groups grid.sync() behind a CTA barrier.  Measured on
// B200 (sphere2500, 302 barriers per solve, two A/B runs each on one box, profiles/r02_barrier_ab.json):
//   grid.sync() alone                                   1.993 ms per solve
//   __syncthreads(); grid.sync()                        1.423 ms   <- this form (1.9 us less per barrier)
//   own counter, red.release.gpu + ld.acquire.gpu spin  1.51  ms
//   own counter, fence + relaxed add / polls + fence    1.63  ms
// The hand-rolled counters lost and are gone.
struct GridReducer {
  double *buf[2];
  int flip;
  int barriers;
  __device__ __forceinline__ void sync(cg::grid_group &grid) {
    __syncthreads();
    grid.sync();
  }
#ifdef DPGO_TRACE
  unsigned long long t_rel, pending;   // thread 0: release time of the last barrier, own work before this one
  __device__ __forceinline__ void arrive() {
    __syncthreads();                   // the whole CTA has finished the phase
    if (threadIdx.x == 0) pending = gtimer() - t_rel;
  }
  __device__ __forceinline__ void release() {
    if (threadIdx.x == 0) t_rel = gtimer();
  }
#else
  __device__ __forceinline__ void arrive() {}
  __device__ __forceinline__ void release() {}
#endif
  // block partials -> grid barrier -> every CTA sums all partials in the same order
  template <int K>
  __device__ __forceinline__ void reduce(cg::grid_group &grid, double (&acc)[K], double (&out)[K]) {
    block_reduce_store<K>(acc, buf[flip] + (size_t)blockIdx.x * K);
    arrive();
    sync(grid);
    release();
    sum_partials<K>(buf[flip], gridDim.x, out);
    flip ^= 1;
    barriers++;
  }
  __device__ __forceinline__ void barrier(cg::grid_group &grid) {
    arrive();
    sync(grid);
    release();
    barriers++;
  }
};

template <int R, int D, int MODE>
__global__ void __launch_bounds__(kBlock, MODE == 2 ? 1 : 2) k_rtr_fused(FusedParams p) {
  DPGO_DYNAMIC_SMEM(dsm);
  cg::grid_group grid = cg::this_grid();
  // per-pose phases: deal the warps over every CTA once there is at least one warp of poses per CTA
  // (measured: sphere2500 -3 %), keep them packed in the first CTAs for small problems (1000
  // poses: packed is 8 % faster)
  const Ctx ctx = ((p.n + Geo<R, D>::GPW - 1) / Geo<R, D>::GPW >= (int)gridDim.x) ? make_ctx_spread() : make_ctx();
  const int n = p.n;
  GemvPipe pipe = gemv_pipe_init<(MODE == 2 ? kDdStages : kStages), (MODE == 2 ? kDdVecChunks : kStages)>(dsm);
  const size_t len = (size_t)R * (D + 1) * n;
  GridReducer red;
  red.buf[0] = p.partials;
  red.buf[1] = p.partials + (size_t)gridDim.x * 4;
  red.flip = 0;
  red.barriers = 0;

  // which buffer of each pair is current: one flag per pair instead of swapped pointers (the pointers stay in the
  // constant bank; the solver state that lives across all phases is what spills under 255 registers)
  bool xf = false;   // iterate: X1 / EGc / GRc / Sc are the second buffers
#define X1 (xf ? p.xb : p.xa)
#define X2 (xf ? p.xa : p.xb)
#define EGc (xf ? p.EG2 : p.EG)
#define EGn (xf ? p.EG : p.EG2)
#define GRc (xf ? p.grad2 : p.grad)
#define GRn (xf ? p.grad : p.grad2)
#define Sc (xf ? p.S2 : p.S)
#define Sn (xf ? p.S : p.S2)
  int n_qx = 0, n_precon = 0, n_sweeps = 0;

  __shared__ unsigned long long s_clk[17];
  PhaseClock clk;
#ifdef DPGO_TRACE
  __shared__ unsigned long long s_busy[16];
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < 16; ++i) s_busy[i] = 0;
  }
  red.pending = 0;
  red.t_rel = gtimer();
  clk.busy = s_busy;
  clk.pending = &red.pending;
#endif
  __shared__ StripPlanStore s_plan[MODE == 2 ? 2 : 1];   // two-level variant: this CTA's strips
  if constexpr (MODE == 2) {
    strip_plan_fill(&s_plan[0], p.dd.P1, p.dd.V);
    strip_plan_fill(&s_plan[1], p.dd.P3, p.dd.V);
    if (p.dd.prefetch != 0 && (threadIdx.x >> 5) == kWarpsPerBlock - 1) {   // a warp without poses in the first phases
      strip_l2_prefetch(p.dd.P1, p.dd.V);
      strip_l2_prefetch(p.dd.P3, p.dd.V);
    }
  }
  bool p1_ready = false;   // two-level form: the first wave of this CTA's first interior strip is already in flight
  // the two forms of the exact preconditioner (compile-time: one kernel instantiation each)
  // permuted: v is the tCG residual and its permuted copy dd.rp is current (written by phase_step_perm)
  auto precon_stream = [&](const double *v, bool permuted) {
    if constexpr (MODE == 2) {
      const DdView &dd = p.dd;
      const size_t zs = (size_t)dd.pcols * R;
      const bool pf = dd.prefetch != 0;
      constexpr int ST = kDdStages;
      if (permuted) phase_strip_gemv<R, ST>(pipe, dd.P1, dd.V, &s_plan[0], dd.rp, nullptr, dd.y, zs, p1_ready);
      else phase_strip_gemv<R, ST>(pipe, dd.P1, dd.V, &s_plan[0], v, dd.icol, dd.y, zs, p1_ready);
      p1_ready = false;
      if (dd.nS > 0) {
        if (pf) strip_prefetch<ST>(pipe, dd.P3, dd.V, &s_plan[1]);
        red.barrier(grid);
        clk.lap(8);
        phase_dd_sep_rhs<R, D>(ctx, dd, v);
        red.barrier(grid);
        clk.lap(9);
        phase_strip_gemv<R, ST>(pipe, dd.P3, dd.V, &s_plan[1], dd.t, nullptr, dd.zs, zs, pf);
        if (pf) strip_prefetch<ST>(pipe, dd.P1, dd.V, &s_plan[0]);
        red.barrier(grid);
        clk.lap(10);
        phase_dd_back_rhs<R, D>(ctx, dd);
        red.barrier(grid);
        clk.lap(11);
        phase_strip_gemv<R, ST>(pipe, dd.P1, dd.V, &s_plan[0], dd.u, nullptr, dd.w, zs, pf);
      }
      // no separator (a single domain): z = y, w stays zero
    } else {
      phase_precon_gemv<R>(pipe, p.Pinv, p.ld, v, p.zpart, p.zstride, p.KT, p.nsplit);
    }
  };
  auto precon_finish = [&](const double *Ycur, const double *rvec, double *neg_out, double (&a1)[1]) {
    if constexpr (MODE == 2)
      phase_dd_finish<R, D>(ctx, p.dd, Ycur, rvec, p.z, neg_out, n, a1);
    else
      phase_precon_finish<R, D>(ctx, p.zpart, p.zstride, p.nsplit, Ycur, rvec, p.z, neg_out, n, a1);
  };

  // ---- statistics at the initial point (fInit, gradNormInit) = first f / Grad of the solver
  clk.start(s_clk);
  phase_copy(ctx, p.x_in, X1, len);
  red.barrier(grid);
  clk.lap(6);
  double f1, gn2;
  {
    double acc[2] = {0.0, 0.0}, sc[2];
    phase_fgrad<R, D>(ctx, p.Q, X1, p.G, EGc, GRc, Sc, n, acc);
    red.reduce<2>(grid, acc, sc);
    clk.lap(0);
    f1 = sc[0];
    gn2 = sc[1];
    n_qx++;
  }
  const double f_init = f1, gn_init = sqrt(gn2);
  int outer = 0, inner_total = 0, accepted_cnt = 0, rejected_cnt = 0, last_status = TCG_MAXITER;
  int returned_initial = 0;

  const bool single = (p.max_outer == 1);  // ref: src/QuadraticOptimizer.cpp:80-98
  double radius = p.init_radius;
  double Delta = p.init_radius;
  double max_Delta = single ? p.init_radius : 5.0 * p.init_radius;
  int total_steps = 0;
  bool run = (gn_init >= p.gradnorm_tol) && (p.max_outer > 0);

  while (run) {
    // ------------------------------------------------------------------ truncated CG
    // One call site per phase: the loop starts with the preconditioner (on grad the first time,
    // on the residual r afterwards), then the direction update, then the Hessian product.
    TcgState s;
    int inner = 0;
    bool first = true;
    for (int j = 0;; ++j) {
      const double *pvec = first ? GRc : p.r;
      precon_stream(pvec, !first);
      if (first) {
        phase_copy(ctx, GRc, p.r, len);
        phase_zero(ctx, p.eta, len);
      }
      red.barrier(grid);
      clk.lap(MODE == 2 ? 12 : 1);
      if constexpr (MODE == 2) {
        // every stage has been consumed; the interior matrix does not change: the first wave of the next interior pass
        // (next tCG iteration, or the next tCG run) goes in flight now, issued by the CTA's last warp, which has no
        // poses in the per-pose phases that follow
        if (p.dd.prefetch != 0) {
          strip_prefetch<kDdStages>(pipe, p.dd.P1, p.dd.V, &s_plan[0], kWarpsPerBlock - 1);
          p1_ready = true;
        }
      }
      {
        double acc[1] = {0.0}, sc[1];
        precon_finish(X1, pvec, first ? p.delta : nullptr, acc);   // first: delta = -z
        red.reduce<1>(grid, acc, sc);
        clk.lap(2);
        n_precon++;
        if (first) {
          tcg_begin(s, gn2, sc[0]);
        } else {
          const double beta = tcg_direction(s, sc[0]);
          phase_axpby(ctx, -1.0, p.z, beta, p.delta, len);
          red.barrier(grid);
          clk.lap(5);
        }
      }
      first = false;
      if (j >= p.max_inner) break;
      double d_Hd;
      {
        double acc[2] = {0.0, 0.0}, sc[2];
        phase_hess<R, D>(ctx, p.Q, X1, Sc, p.delta, p.Hd, nullptr, n, acc);
        red.reduce<2>(grid, acc, sc);
        clk.lap(3);
        d_Hd = sc[0];
        n_qx++;
      }
      inner = j + 1;
      double step;
      if (tcg_curvature(s, d_Hd, Delta, &step)) {
        phase_axpby(ctx, step, p.delta, 1.0, p.eta, len);
        red.barrier(grid);
        clk.lap(4);
        break;
      }
      double r_r;
      {
        double acc[1] = {0.0}, sc[1];
        if constexpr (MODE == 2) phase_step_perm<R, D>(ctx, step, p.delta, p.Hd, p.eta, p.r, p.dd.pcol, p.dd.rp, len, acc);
        else phase_step(ctx, step, p.delta, p.Hd, p.eta, p.r, len, acc);
        red.reduce<1>(grid, acc, sc);
        clk.lap(4);
        r_r = sc[0];
      }
      if (tcg_converged(s, r_r, p.theta, p.kappa)) break;
      if (j + 1 >= p.max_inner) break;   // the reference's loop ends without a further direction
    }
    inner_total += inner;
    last_status = s.status;

    // ------------------------------------------------------------------ candidate + ratio test
    phase_retract<R, D>(ctx, X1, p.eta, X2, n);
    n_sweeps++;
    red.barrier(grid);
    clk.lap(6);
    double f2, gn2_2, eHe, eg;
    {
      double acc[4] = {0.0, 0.0, 0.0, 0.0}, sc[4];
      double a01[2] = {0.0, 0.0}, a23[2] = {0.0, 0.0};
      phase_fgrad<R, D>(ctx, p.Q, X2, p.G, EGn, GRn, Sn, n, a01);
      phase_hess<R, D>(ctx, p.Q, X1, Sc, p.eta, p.Hd, GRc, n, a23);
      acc[0] = a01[0]; acc[1] = a01[1]; acc[2] = a23[0]; acc[3] = a23[1];
      red.reduce<4>(grid, acc, sc);
      clk.lap(0);
      f2 = sc[0]; gn2_2 = sc[1]; eHe = sc[2]; eg = sc[3];
      n_qx += 2;
    }
    double rho;
    const bool acc_step = rtr_accept(f1, f2, eg, eHe, s.status, p.accept_rho, p.shrink, p.magnify,
                                     max_Delta, &Delta, &rho);
    if (acc_step) {
      xf = !xf;
      f1 = f2;
      gn2 = gn2_2;
      accepted_cnt++;
    } else {
      rejected_cnt++;
    }
    outer++;
    if (single) {
      if (acc_step) run = false;
      else if (total_steps > 10) { run = false; returned_initial = 1; }
      else { radius *= 0.25; total_steps++; Delta = radius; max_Delta = radius; }
    } else {
      run = (outer < p.max_outer) && !(sqrt(gn2) < p.gradnorm_tol);
    }
  }

  phase_copy(ctx, X1, p.x_out, len);
  if constexpr (MODE == 2) {   // no bulk copy may be in flight when the CTA exits
    if (p1_ready) strip_drain<kDdStages>(pipe, p.dd.P1, p.dd.V, &s_plan[0]);
  }
#ifdef DPGO_TRACE
  if (threadIdx.x == 0 && p.trace) {
#pragma unroll
    for (int i = 0; i < 16; ++i) p.trace[(size_t)blockIdx.x * 16 + i] = s_busy[i];
  }
#endif
  if (ctx.tid == 0) {
    FusedOut o;
    o.f_init = f_init; o.gn_init = gn_init; o.f_opt = f1; o.gn_opt = sqrt(gn2);
    o.outer = outer; o.inner = inner_total; o.accepted = accepted_cnt; o.rejected = rejected_cnt;
    o.tcg_status = last_status; o.returned_initial = returned_initial;
    o.n_qx = n_qx; o.n_precon = n_precon; o.n_sweeps = n_sweeps; o.n_barriers = red.barriers;
#pragma unroll
    for (int i = 0; i < 16; ++i) o.phase_ms[i] = (double)clk.acc[i] * 1e-6;
    if (MODE == 2) {
#pragma unroll
      for (int i = 8; i <= 12; ++i) o.phase_ms[1] += o.phase_ms[i];
    }
    *p.out = o;
  }
#undef X1
#undef X2
#undef EGc
#undef EGn
#undef GRc
#undef GRn
#undef Sc
#undef Sn
}

}  // namespace dpgo
