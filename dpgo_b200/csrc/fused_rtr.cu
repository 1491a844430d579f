// Persistent fused RTR solver: the whole QuadraticOptimizer::optimize call (cost / gradient
// statistics, Steihaug-Toint tCG with the dense preconditioner, QF retraction, ratio test, radius
// update; ref: src/QuadraticOptimizer.cpp:26-108 + ROPTLIB RTRNewton) runs as ONE cooperative
// kernel.  Phases are the same __device__ functions the stand-alone kernels use (kernels.cuh);
// they are separated by grid-wide barriers, every reduction is a per-CTA partial followed by a
// fixed-order sum that every CTA repeats, so all threads hold identical copies of the control
// scalars (alpha, beta, rho, Delta, ...) and take identical branches: no host round trip until
// the result block is read back.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "device_state.h"
#include "fused_kernel.cuh"

namespace dpgo {

DdView dd_view(const dpgo_dev *h);  // precon_dd.cu

template <int R, int D, int MODE>
static int launch_fused_v(dpgo_dev *h, FusedParams &fp) {
  constexpr int smem = (MODE == 2) ? kDdDynSmem : kGemvDynSmem;
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
  if (h->precon_mode == 2) return launch_fused_v<R, D, 2>(h, fp);
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
  fp.precon_mode = h->precon_mode;
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
