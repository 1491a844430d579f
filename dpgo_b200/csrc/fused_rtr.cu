// Persistent fused RTR solver: the whole QuadraticOptimizer::optimize call (cost / gradient
// statistics, Steihaug-Toint tCG with the dense preconditioner, QF retraction, ratio test, radius
// update; ref: src/QuadraticOptimizer.cpp:26-108 + ROPTLIB RTRNewton) runs as ONE cooperative
// kernel.  Phases are the same __device__ functions the stand-alone kernels use (kernels.cuh);
// they are separated by grid-wide barriers, every reduction is a per-CTA partial followed by a
// fixed-order sum that every CTA repeats, so all threads hold identical copies of the control
// scalars (alpha, beta, rho, Delta, ...) and take identical branches: no host round trip until
// the result block is read back.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "device_state.h"
#include "kernels.cuh"
#include "rtr_logic.h"

namespace cg = cooperative_groups;

namespace dpgo {

DdView dd_view(const dpgo_dev *h);  // precon_dd.cu

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
  int precon_mode, symT, symNG, nitems;   // symmetric half-storage variant
  const SymItem *items;
  double *zT;
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

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

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

struct GridReducer {
  double *buf[2];
  int flip;
  int barriers;
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
    grid.sync();
    release();
    sum_partials<K>(buf[flip], gridDim.x, out);
    flip ^= 1;
    barriers++;
  }
  __device__ __forceinline__ void barrier(cg::grid_group &grid) {
    arrive();
    grid.sync();
    release();
    barriers++;
  }
};

template <int R, int D, int MODE>
__global__ void __launch_bounds__(kBlock, MODE >= 2 ? 1 : 2) k_rtr_fused(FusedParams p) {
  extern __shared__ __align__(128) unsigned char dsm[];
  cg::grid_group grid = cg::this_grid();
  // per-pose phases: deal the warps over every CTA once there is at least one warp of poses per CTA
  // (measured: sphere2500 -3 %), keep them packed in the first CTAs for small problems (1000
  // poses: packed is 8 % faster)
  const Ctx ctx = ((p.n + Geo<R, D>::GPW - 1) / Geo<R, D>::GPW >= (int)gridDim.x) ? make_ctx_spread() : make_ctx();
  const int n = p.n;
  GemvPipe pipe = gemv_pipe_init<(MODE >= 3 ? kDd3Stages : (MODE == 2 ? kDdStages : kStages)),
                                 (MODE >= 3 ? kDd3Stages : (MODE == 2 ? kDdVecChunks : kStages))>(dsm);
  const size_t len = (size_t)R * (D + 1) * n;
  GridReducer red;
  red.buf[0] = p.partials;
  red.buf[1] = p.partials + (size_t)gridDim.x * 4;
  red.flip = 0;
  red.barriers = 0;

  double *x1 = p.xa, *x2 = p.xb, *EG = p.EG, *EG2 = p.EG2, *grad = p.grad, *grad2 = p.grad2;
  double *S = p.S, *S2 = p.S2;
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
  __shared__ StripPlanStore s_plan[MODE >= 3 ? 3 : (MODE == 2 ? 2 : 1)];   // two-level variants: this CTA's strips
  if constexpr (MODE >= 2) {
    strip_plan_fill(&s_plan[0], p.dd.P1, p.dd.V);
    strip_plan_fill(&s_plan[1], p.dd.P3, p.dd.V);
  }
  if constexpr (MODE >= 3) strip_plan_fill(&s_plan[2], p.dd.P5, p.dd.V);
  // the three variants of the exact preconditioner (compile-time: one per kernel instantiation)
  // MODE 4 = the three-phase form with the finish of the interior poses in the epilogue of the last strip
  // phase and the separator poses right after it: one application = 3 grid phases, the third one ending in
  // the <z, r> reduction (d = 3 only: 64-column strips hold whole poses)
  auto precon_fused = [&](const double *v, const double *Ycur, double *neg_out, double (&a1)[1]) {
    if constexpr (MODE == 4) {
      const DdView &dd = p.dd;
      const size_t zs = (size_t)dd.pcols * R;
      const bool pf = dd.prefetch != 0;
      constexpr int ST = kDd3Stages;
      phase_strip_gemv<R, ST, 1>(pipe, dd.P1, dd.V, &s_plan[0], v, dd.icol, dd.y, 0);
      if (dd.nS > 0) {
        if (pf) strip_prefetch<ST>(pipe, dd.P3, dd.V, &s_plan[1]);
        red.barrier(grid);
        clk.lap(8);
        const StageAux a3{dd.y, dd.tptr, dd.tcol, dd.sep_col0, 0, 0};
        phase_strip_gemv<R, ST, 2>(pipe, dd.P3, dd.V, &s_plan[1], v, dd.icol, dd.zs, zs, pf, &a3);
        if (pf) strip_prefetch<ST>(pipe, dd.P5, dd.V, &s_plan[2]);
        red.barrier(grid);
        clk.lap(10);
        const StageAux a5{nullptr, nullptr, nullptr, 0, dd.nsplit3, zs};
        StripFinish fin{dd.y, dd.icol, Ycur, v, p.z, neg_out, 0.0};
        phase_strip_gemv<R, ST, 3, D>(pipe, dd.P5, dd.V, &s_plan[2], dd.zs, nullptr, dd.w, 0, pf, &a5, &fin);
        a1[0] += fin.acc;
        phase_dd_finish_sep<R, D>(ctx, dd, Ycur, v, p.z, neg_out, a1);
      } else {   // a single domain: z = Proj(y)
        red.barrier(grid);
        clk.lap(8);
        phase_dd_finish<R, D>(ctx, dd, Ycur, v, p.z, neg_out, n, a1);
      }
    }
  };
  auto precon_stream = [&](const double *v) {
    if constexpr (MODE == 3) {
      // three-phase form: [M_k | C_k] strips -> Sigma^-1 strips (t_S formed while staged) -> C_k^T strips
      const DdView &dd = p.dd;
      const size_t zs = (size_t)dd.pcols * R;
      const bool pf = dd.prefetch != 0;
      constexpr int ST = kDd3Stages;
      phase_strip_gemv<R, ST, 1>(pipe, dd.P1, dd.V, &s_plan[0], v, dd.icol, dd.y, 0);
      if (dd.nS > 0) {
        if (pf) strip_prefetch<ST>(pipe, dd.P3, dd.V, &s_plan[1]);
        red.barrier(grid);
        clk.lap(8);
        const StageAux a3{dd.y, dd.tptr, dd.tcol, dd.sep_col0, 0, 0};
        phase_strip_gemv<R, ST, 2>(pipe, dd.P3, dd.V, &s_plan[1], v, dd.icol, dd.zs, zs, pf, &a3);
        if (pf) strip_prefetch<ST>(pipe, dd.P5, dd.V, &s_plan[2]);
        red.barrier(grid);
        clk.lap(10);
        const StageAux a5{nullptr, nullptr, nullptr, 0, dd.nsplit3, zs};
        phase_strip_gemv<R, ST, 3>(pipe, dd.P5, dd.V, &s_plan[2], dd.zs, nullptr, dd.w, 0, pf, &a5);
      }
    } else if constexpr (MODE == 2) {
      const DdView &dd = p.dd;
      const size_t zs = (size_t)dd.pcols * R;
      const bool pf = dd.prefetch != 0;
      constexpr int ST = kDdStages;
      phase_strip_gemv<R, ST>(pipe, dd.P1, dd.V, &s_plan[0], v, dd.icol, dd.y, zs);
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
    } else if constexpr (MODE == 1) {
      phase_precon_symv<R>(pipe, p.Pinv, p.symT, p.items, p.nitems, v, p.zpart, p.zT, p.zstride);
    } else {
      phase_precon_gemv<R>(pipe, p.Pinv, p.ld, v, p.zpart, p.zstride, p.KT, p.nsplit);
    }
  };
  auto precon_finish = [&](const double *Ycur, const double *rvec, double *neg_out, double (&a1)[1]) {
    if constexpr (MODE >= 2)   // (MODE 4 finishes inside precon_fused)
      phase_dd_finish<R, D>(ctx, p.dd, Ycur, rvec, p.z, neg_out, n, a1);
    else if constexpr (MODE == 1)
      phase_precon_finish_sym<R, D>(pipe.scratch, p.zpart, p.zT, p.zstride, p.symNG, Ycur, rvec, p.z,
                                    neg_out, n, a1);
    else
      phase_precon_finish<R, D>(ctx, p.zpart, p.zstride, p.nsplit, Ycur, rvec, p.z, neg_out, n, a1);
  };

  // ---- statistics at the initial point (fInit, gradNormInit) = first f / Grad of the solver
  clk.start(s_clk);
  phase_copy(ctx, p.x_in, x1, len);
  red.barrier(grid);
  clk.lap(6);
  double f1, gn2;
  {
    double acc[2] = {0.0, 0.0}, sc[2];
    phase_fgrad<R, D>(ctx, p.Q, x1, p.G, EG, grad, S, n, acc);
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
      const double *pvec = first ? grad : p.r;
      if constexpr (MODE == 4) {
        double acc[1] = {0.0}, sc[1];
        precon_fused(pvec, x1, first ? p.delta : nullptr, acc);    // first: delta = -z
        if (first) {
          phase_copy(ctx, grad, p.r, len);
          phase_zero(ctx, p.eta, len);
        }
        red.reduce<1>(grid, acc, sc);
        clk.lap(12);
        n_precon++;
        if (first) {
          tcg_begin(s, gn2, sc[0]);
        } else {
          const double beta = tcg_direction(s, sc[0]);
          phase_axpby(ctx, -1.0, p.z, beta, p.delta, len);
          red.barrier(grid);
          clk.lap(5);
        }
      } else {
      precon_stream(pvec);
      if (first) {
        phase_copy(ctx, grad, p.r, len);
        phase_zero(ctx, p.eta, len);
      }
      red.barrier(grid);
      clk.lap(MODE >= 2 ? 12 : 1);
      {
        double acc[1] = {0.0}, sc[1];
        precon_finish(x1, pvec, first ? p.delta : nullptr, acc);   // first: delta = -z
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
      }
      first = false;
      if (j >= p.max_inner) break;
      double d_Hd;
      {
        double acc[2] = {0.0, 0.0}, sc[2];
        phase_hess<R, D>(ctx, p.Q, x1, S, p.delta, p.Hd, nullptr, n, acc);
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
        phase_step(ctx, step, p.delta, p.Hd, p.eta, p.r, len, acc);
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
    phase_retract<R, D>(ctx, x1, p.eta, x2, n);
    n_sweeps++;
    red.barrier(grid);
    clk.lap(6);
    double f2, gn2_2, eHe, eg;
    {
      double acc[4] = {0.0, 0.0, 0.0, 0.0}, sc[4];
      double a01[2] = {0.0, 0.0}, a23[2] = {0.0, 0.0};
      phase_fgrad<R, D>(ctx, p.Q, x2, p.G, EG2, grad2, S2, n, a01);
      phase_hess<R, D>(ctx, p.Q, x1, S, p.eta, p.Hd, grad, n, a23);
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
      double *t;
      t = x1; x1 = x2; x2 = t;
      t = EG; EG = EG2; EG2 = t;
      t = grad; grad = grad2; grad2 = t;
      t = S; S = S2; S2 = t;
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

  phase_copy(ctx, x1, p.x_out, len);
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
    if (MODE >= 2) {
#pragma unroll
      for (int i = 8; i <= 12; ++i) o.phase_ms[1] += o.phase_ms[i];
    }
    *p.out = o;
  }
}

template <int R, int D, int MODE>
static int launch_fused_v(dpgo_dev *h, FusedParams &fp) {
  constexpr int smem = (MODE >= 3) ? kDd3DynSmem : ((MODE == 2) ? kDdDynSmem : kGemvDynSmem);
  // the shared-memory attribute and the occupancy belong to the device (one handle = one device; a
  // process may hold handles on several): cached per device
  static int occ_by_device[64];
  if (h->device < 0 || h->device >= 64) {
    set_error("device ordinal %d not supported by the fused solver", h->device);
    return DPGO_EINVAL;
  }
  int &occ_cache = occ_by_device[h->device];
  if (occ_cache <= 0) {
    int occ = 0;
    if (cudaFuncSetAttribute(k_rtr_fused<R, D, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_rtr_fused<R, D, MODE>, kBlock, smem) != cudaSuccess ||
        occ < 1) {
      set_error("fused kernel does not fit on the device");
      return DPGO_ECUDA;
    }
    occ_cache = occ;
  }
  const long cap = (long)h->num_sms * occ_cache;
  // enough CTAs for the widest phase, never more than can be co-resident
  long tiles = (long)(h->ld / kGemvCols) * h->nsplit;
  if (h->precon_mode == 1) tiles = h->sym_nitems;
  if (h->precon_mode >= 2) tiles = fp.dd.V;
  const int gpw = 32 / (h->d + 1);
  const long pose_blocks = (((long)h->n + gpw - 1) / gpw + kWarpsPerBlock - 1) / kWarpsPerBlock;
  long grid = std::max(tiles, pose_blocks);
  grid = std::max(1L, std::min(grid, cap));
  if (grid > h->partial_blocks) {
    set_error("partials buffer too small");
    return DPGO_EINVAL;
  }
  void *args[] = {&fp};
  cudaError_t e = cudaLaunchCooperativeKernel((void *)k_rtr_fused<R, D, MODE>, dim3((unsigned)grid), dim3(kBlock),
                                              args, smem, h->stream);
  if (e != cudaSuccess) {
    set_error("cudaLaunchCooperativeKernel failed: %s", cudaGetErrorString(e));
    return DPGO_ECUDA;
  }
  h->launches++;
  h->last_grid = (int)grid;
  return DPGO_OK;
}

template <int R, int D>
static int launch_fused(dpgo_dev *h, FusedParams &fp) {
  if constexpr (D == 3) {
    if (h->precon_mode == 4) return launch_fused_v<R, D, 4>(h, fp);
  }
  if (h->precon_mode >= 3) return launch_fused_v<R, D, 3>(h, fp);
  if (h->precon_mode == 2) return launch_fused_v<R, D, 2>(h, fp);
  if (h->precon_mode == 1) return launch_fused_v<R, D, 1>(h, fp);
  return launch_fused_v<R, D, 0>(h, fp);
}

// Launch half of the fused solve: the kernel and the copy of its result block to pinned host memory
// are queued on the handle's stream; nothing waits.  solve_fused_collect() is the other half.
int solve_fused_launch(dpgo_dev *h, const dpgo_ropt_params *P, const double *x_in, double *x_out) {
  if (!h->d_fused) {
    if (cudaMalloc(&h->d_fused, sizeof(FusedOut)) != cudaSuccess ||
        cudaMallocHost(&h->h_fused, sizeof(FusedOut)) != cudaSuccess) {
      set_error("allocation of the fused result block failed");
      return DPGO_ECUDA;
    }
  }
  FusedParams fp;
  fp.Q = BsrView{h->d_rowptr, h->d_colidx, h->d_blocks};
  fp.G = h->d_G;
  fp.Pinv = h->d_Pinv;
  fp.zpart = h->d_zpart;
  fp.ld = h->ld; fp.KT = h->KT; fp.nsplit = h->nsplit; fp.n = h->n;
  fp.zstride = h->vpad;
  fp.precon_mode = h->precon_mode; fp.symT = h->symT; fp.symNG = h->symNG; fp.nitems = h->sym_nitems;
  fp.items = (const SymItem *)h->d_sym_items; fp.zT = h->d_zT;
  if (h->precon_mode >= 2) fp.dd = dd_view(h); else memset(&fp.dd, 0, sizeof(fp.dd));
  fp.x_in = x_in; fp.x_out = x_out;
  fp.xa = h->d_xa; fp.xb = h->d_xb; fp.EG = h->d_EG; fp.EG2 = h->d_EG2;
  fp.grad = h->d_grad; fp.grad2 = h->d_grad2; fp.S = h->d_S; fp.S2 = h->d_S2;
  fp.eta = h->d_eta; fp.r = h->d_r; fp.z = h->d_z; fp.delta = h->d_delta; fp.Hd = h->d_Hd;
  fp.partials = h->d_partials;
  fp.out = (FusedOut *)h->d_fused;
  fp.trace = nullptr;
#ifdef DPGO_TRACE
  if (!h->d_trace) {
    if (cudaMalloc(&h->d_trace, (size_t)h->partial_blocks * 16 * sizeof(unsigned long long)) != cudaSuccess) {
      set_error("allocation of the phase trace failed");
      return DPGO_ECUDA;
    }
  }
  cudaMemsetAsync(h->d_trace, 0, (size_t)h->partial_blocks * 16 * sizeof(unsigned long long), h->stream);
  fp.trace = (unsigned long long *)h->d_trace;
#endif
  fp.gradnorm_tol = P->gradnorm_tol; fp.init_radius = P->RTR_initial_radius;
  fp.theta = P->tcg_theta; fp.kappa = P->tcg_kappa; fp.accept_rho = P->accept_rho;
  fp.shrink = P->shrink; fp.magnify = P->magnify;
  fp.max_outer = P->RTR_iterations; fp.max_inner = P->RTR_tCG_iterations;
  int rc = DPGO_EINVAL;
  switch (h->d * 16 + h->r) {
    case 2 * 16 + 2: rc = launch_fused<2, 2>(h, fp); break;
    case 2 * 16 + 3: rc = launch_fused<3, 2>(h, fp); break;
    case 2 * 16 + 4: rc = launch_fused<4, 2>(h, fp); break;
    case 2 * 16 + 5: rc = launch_fused<5, 2>(h, fp); break;
    case 3 * 16 + 3: rc = launch_fused<3, 3>(h, fp); break;
    case 3 * 16 + 4: rc = launch_fused<4, 3>(h, fp); break;
    case 3 * 16 + 5: rc = launch_fused<5, 3>(h, fp); break;
    case 3 * 16 + 6: rc = launch_fused<6, 3>(h, fp); break;
    default: set_error("unsupported (d=%d, r=%d)", h->d, h->r); return DPGO_EINVAL;
  }
  if (rc != DPGO_OK) return rc;
  if (cudaMemcpyAsync(h->h_fused, h->d_fused, sizeof(FusedOut), cudaMemcpyDeviceToHost, h->stream) != cudaSuccess) {
    set_error("copy of the fused result block failed: %s", cudaGetErrorString(cudaGetLastError()));
    return DPGO_ECUDA;
  }
  return DPGO_OK;
}

// Collect half: wait for the stream, decode the result block of the most recent launch.
int solve_fused_collect(dpgo_dev *h, int verbose, dpgo_ropt_result *res) {
  if (!h->h_fused) {
    set_error("no fused solve has been launched on this handle");
    return DPGO_ESTATE;
  }
  if (cudaStreamSynchronize(h->stream) != cudaSuccess) {
    set_error("fused RTR kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
    return DPGO_ECUDA;
  }
  const FusedOut &o = *(const FusedOut *)h->h_fused;
  res->success = 1;
  res->tcg_status = o.tcg_status;
  res->f_init = o.f_init; res->gradnorm_init = o.gn_init;
  res->f_opt = o.f_opt; res->gradnorm_opt = o.gn_opt;
  res->outer_iters = o.outer; res->inner_iters = o.inner;
  res->accepted = o.accepted; res->rejected = o.rejected;
  res->n_qx = o.n_qx; res->n_precon = o.n_precon; res->n_pose_sweeps = o.n_sweeps;
  res->n_barriers = o.n_barriers;
  for (int i = 0; i < 16; ++i) res->phase_ms[i] = o.phase_ms[i];
  if (verbose)
    printf("[dpgo_b200] fused RTR: f %.10g -> %.10g, |g| %.4g -> %.4g, %d outer, %d tCG, %lld barriers\n",
           o.f_init, o.f_opt, o.gn_init, o.gn_opt, o.outer, o.inner, o.n_barriers);
  return DPGO_OK;
}

// Per-CTA busy time (ms) per phase id of the last fused solve; only in -DDPGO_TRACE builds.
int fused_phase_trace(dpgo_dev *h, double *busy_ms, int cap_ctas, int *num_ctas) {
#ifdef DPGO_TRACE
  if (!h->d_trace || h->last_grid <= 0) {
    set_error("no fused solve has run on this handle");
    return DPGO_ESTATE;
  }
  *num_ctas = h->last_grid;
  if (cap_ctas < h->last_grid) return DPGO_OK;
  std::vector<unsigned long long> tmp((size_t)h->last_grid * 16);
  if (cudaStreamSynchronize(h->stream) != cudaSuccess ||
      cudaMemcpy(tmp.data(), h->d_trace, tmp.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost) != cudaSuccess) {
    set_error("reading the phase trace failed: %s", cudaGetErrorString(cudaGetLastError()));
    return DPGO_ECUDA;
  }
  for (size_t i = 0; i < tmp.size(); ++i) busy_ms[i] = (double)tmp[i] * 1e-6;
  return DPGO_OK;
#else
  (void)h; (void)busy_ms; (void)cap_ctas; (void)num_ctas;
  set_error("the library was built without -DDPGO_TRACE");
  return DPGO_ESTATE;
#endif
}

int solve_fused(dpgo_dev *h, const dpgo_ropt_params *P, const double *x_in, double *x_out,
                dpgo_ropt_result *res) {
  const int rc = solve_fused_launch(h, P, x_in, x_out);
  if (rc != DPGO_OK) return rc;
  return solve_fused_collect(h, P->verbose, res);
}

}  // namespace dpgo
