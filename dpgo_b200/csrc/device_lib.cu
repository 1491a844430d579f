// C-ABI implementation: handle lifecycle, host-side construction of the block-CSR connection
// Laplacian, dense preconditioner build (in-tree Cholesky / inverse of dense_la.cu, set-up path only), stand-alone
// kernels (one launch per op) and the host-driven RTR/RGD solver built from them.  The
// persistent single-kernel solver lives in fused_rtr.cu.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "dense_la.h"
#include "device_state.h"
#include "dissect.h"
#include "kernels.cuh"
#include "rtr_logic.h"

namespace dpgo {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

#define CUDA_TRY(expr)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      dpgo::set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, #expr,              \
                      cudaGetErrorString(_e));                                           \
      return DPGO_ECUDA;                                                                 \
    }                                                                                    \
  } while (0)

#define CHECK_ARG(cond)                                                                  \
  do {                                                                                   \
    if (!(cond)) {                                                                       \
      dpgo::set_error("%s:%d contract violated: %s", __FILE__, __LINE__, #cond);         \
      return DPGO_EINVAL;                                                                \
    }                                                                                    \
  } while (0)

#define DPGO_TRY(expr)                                                                   \
  do {                                                                                   \
    int _rc = (expr);                                                                    \
    if (_rc != DPGO_OK) return _rc;                                                      \
  } while (0)

// d in {2,3}; r in [d, d+3]
#define DPGO_DISPATCH(h, ...)                                                            \
  switch ((h)->d * 16 + (h)->r) {                                                        \
    case 2 * 16 + 2: { constexpr int D = 2, R = 2; __VA_ARGS__; } break;                 \
    case 2 * 16 + 3: { constexpr int D = 2, R = 3; __VA_ARGS__; } break;                 \
    case 2 * 16 + 4: { constexpr int D = 2, R = 4; __VA_ARGS__; } break;                 \
    case 2 * 16 + 5: { constexpr int D = 2, R = 5; __VA_ARGS__; } break;                 \
    case 3 * 16 + 3: { constexpr int D = 3, R = 3; __VA_ARGS__; } break;                 \
    case 3 * 16 + 4: { constexpr int D = 3, R = 4; __VA_ARGS__; } break;                 \
    case 3 * 16 + 5: { constexpr int D = 3, R = 5; __VA_ARGS__; } break;                 \
    case 3 * 16 + 6: { constexpr int D = 3, R = 6; __VA_ARGS__; } break;                 \
    default: dpgo::set_error("unsupported (d=%d, r=%d)", (h)->d, (h)->r); return DPGO_EINVAL; \
  }

// ---------------------------------------------------------------------------------------------
// stand-alone kernels
// ---------------------------------------------------------------------------------------------
template <int R, int D>
__global__ void __launch_bounds__(kBlock) k_qx(BsrView Q, const double *X, const double *G,
                                               double *out, int n) {
  phase_qx<R, D>(make_ctx(), Q, X, G, out, n);
}

// lane-group kernel, two blocks per step (spmm_col2): 3 CTAs / SM
template <int R, int D>
__global__ void __launch_bounds__(kBlock, 3) k_qx_pipe(BsrView Q, const double *X, const double *G,
                                                       double *out, int n) {
  phase_qx<R, D, true>(make_ctx(), Q, X, G, out, n);
}

// lane-group kernel with the X tiles of a warp step staged in shared memory (phase_qx_tiles)
template <int R, int D>
__global__ void __launch_bounds__(kBlock, 4) k_qx_tiles(BsrView Q, const double *X, const double *G,
                                                        double *out, int n) {
  __shared__ __align__(16) unsigned char stage[kWarpsPerBlock][QxTiles<R, D>::WARP_BYTES];
  phase_qx_tiles<R, D>(make_ctx(), Q, X, G, out, n, stage[threadIdx.x >> 5]);
}

template <int R, int D>
__global__ void __launch_bounds__(kBlock, 5) k_qx_prefetch(BsrView Q, const double *X, const double *G,
                                                        double *out, int n, int dist) {
  phase_qx_prefetch<R, D>(make_ctx(), Q, X, G, out, n, dist);
}

// per-pose kernels: one thread per pose, tiles staged through shared memory (coalesced global traffic), kPoseBlock
// threads per CTA
// (5 CTAs per SM: 96 registers instead of 142; 262 144 / 1 000 000 poses 0.39 / 0.50 -> 0.43 / 0.62 of the HBM peak)
template <int R, int D>
__global__ void __launch_bounds__(kPoseBlock, 5) k_round(const double *X, const double *anchor, double *T, int n) {
  __shared__ double sw[kPoseBlock / 32][PoseStage<R * (D + 1)>::WARP_DOUBLES];
  round_staged<R, D>(X, anchor, T, n, sw[threadIdx.x >> 5]);
}

template <int R, int D>
__global__ void __launch_bounds__(kBlock) k_fgrad(BsrView Q, const double *X, const double *G,
                                                  double *EG, double *grad, double *S, int n,
                                                  double *partials) {
  double acc[2] = {0.0, 0.0};
  phase_fgrad<R, D>(make_ctx(), Q, X, G, EG, grad, S, n, acc);
  block_reduce_store<2>(acc, partials + 2 * blockIdx.x);
}

template <int R, int D>
__global__ void __launch_bounds__(kBlock) k_hess(BsrView Q, const double *Y, const double *S,
                                                 const double *V, double *HV, const double *W,
                                                 int n, double *partials) {
  double acc[2] = {0.0, 0.0};
  phase_hess<R, D>(make_ctx(), Q, Y, S, V, HV, W, n, acc);
  block_reduce_store<2>(acc, partials + 2 * blockIdx.x);
}

template <int R, int D>
__global__ void __launch_bounds__(kBlock) k_tangent(const double *X, const double *V, double *out,
                                                    int n) {
  phase_tangent<R, D>(make_ctx(), X, V, out, n);
}

template <int R>
__global__ void __launch_bounds__(kBlock) k_precon_gemv(const double *Pinv, int ld,
                                                        const double *vec, double *zpart,
                                                        size_t zstride, int KT, int nsplit) {
  extern __shared__ __align__(128) unsigned char dsm[];
  GemvPipe pp = gemv_pipe_init(dsm);
  phase_precon_gemv<R>(pp, Pinv, ld, vec, zpart, zstride, KT, nsplit);
}

template <int R, int D>
__global__ void __launch_bounds__(kBlock) k_precon_finish(const double *zpart, size_t zstride, int nsplit,
                                                          const double *Y, const double *rvec,
                                                          double *z, double *neg_out, int n,
                                                          double *partials) {
  double acc[1] = {0.0};
  phase_precon_finish<R, D>(make_ctx(), zpart, zstride, nsplit, Y, rvec, z, neg_out, n, acc);
  block_reduce_store<1>(acc, partials + blockIdx.x);
}

template <int R, int D>
__global__ void __launch_bounds__(kPoseBlock) k_retract(const double *X, const double *Eta,
                                                        double *Xout, int n) {
  __shared__ double sw[kPoseBlock / 32][PoseStage<R * (D + 1)>::WARP_DOUBLES];
  retract_staged<R, D, false>(X, Eta, Xout, n, 1.0, sw[threadIdx.x >> 5]);
}

// RGD step: Xout = Retraction_X(s * Dir) -- gradient scale and QF retraction in one pass
template <int R, int D>
__global__ void __launch_bounds__(kPoseBlock) k_retract_scaled(const double *X, const double *Dir, double s,
                                                               double *Xout, int n) {
  __shared__ double sw[kPoseBlock / 32][PoseStage<R * (D + 1)>::WARP_DOUBLES];
  retract_staged<R, D, true>(X, Dir, Xout, n, s, sw[threadIdx.x >> 5]);
}

template <int R, int D>
__global__ void __launch_bounds__(kPoseBlock) k_polar(double ca, const double *A, double cb,
                                                      const double *B, double cc, const double *C,
                                                      double *out, int n) {
  __shared__ double sw[kPoseBlock / 32][PoseStage<R * (D + 1)>::WARP_DOUBLES];
  polar_staged<R, D>(ca, A, cb, B, cc, C, out, n, sw[threadIdx.x >> 5]);
}

__global__ void __launch_bounds__(kBlock) k_step(double a, const double *delta, const double *Hd,
                                                 double *eta, double *r, size_t len,
                                                 double *partials) {
  double acc[1] = {0.0};
  phase_step(make_ctx(), a, delta, Hd, eta, r, len, acc);
  block_reduce_store<1>(acc, partials + blockIdx.x);
}

__global__ void __launch_bounds__(kBlock) k_axpby(double a, const double *x, double b, double *y,
                                                  size_t len) {
  phase_axpby(make_ctx(), a, x, b, y, len);
}

// scalars[k] = sum_b partials[b*K + k], fixed order
__global__ void __launch_bounds__(kBlock) k_finalize(const double *partials, int nblocks, int K,
                                                     double *scalars) {
  __shared__ double sm[kBlock];
  for (int k = 0; k < K; ++k) {
    double x = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += kBlock) x += partials[(size_t)b * K + k];
    sm[threadIdx.x] = x;
    __syncthreads();
    for (int o = kBlock / 2; o > 0; o >>= 1) {
      if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) scalars[k] = sm[0];
    __syncthreads();
  }
}

// dense P = Q + shift*I from the block-CSR (one thread per scalar entry)
__global__ void k_scatter_dense(const int *browidx, const int *colidx, const double *blocks,
                                int nnzb, int dh, double shift, double *P, int ld) {
  const size_t total = (size_t)nnzb * dh * dh;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total;
       t += (size_t)gridDim.x * blockDim.x) {
    const int e = (int)(t / (dh * dh));
    const int ab = (int)(t % (dh * dh));
    const int a = ab / dh, b = ab % dh;  // row-major within the block
    const int i = browidx[e], j = colidx[e];
    const size_t row = (size_t)i * dh + a, col = (size_t)j * dh + b;
    double v = blocks[t];
    if (row == col) v += shift;
    P[row + col * (size_t)ld] = v;
  }
}

// Re-lay the inverse (lower triangle of the col-major potri output A, leading dim lda) into the
// stage-major tiled format the apply kernel streams (see phase_precon_gemv): symmetric, zero
// padded to ld rows x ldk columns.
__global__ void k_tile_layout(const double *A, int N, int lda, double *T, int ld, int ldk) {
  const size_t total = (size_t)ld * ldk;
  const int ncb = ld / kGemvCols;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total;
       t += (size_t)gridDim.x * blockDim.x) {
    // t enumerates the destination: ((kc * ncb + cb) * kStageK + kk) * kGemvCols + jj
    const int jj = (int)(t % kGemvCols);
    const size_t u = t / kGemvCols;
    const int kk = (int)(u % kStageK);
    const size_t v = u / kStageK;
    const int cb = (int)(v % ncb);
    const int kc = (int)(v / ncb);
    const int j = cb * kGemvCols + jj, k = kc * kStageK + kk;
    double val = 0.0;
    if (j < N && k < N) val = (j >= k) ? A[(size_t)j + (size_t)k * lda] : A[(size_t)k + (size_t)j * lda];
    T[t] = val;
  }
}

__global__ void k_gather_tiles(const double *X, const int *idx, int num, int tile, double *out) {
  const size_t total = (size_t)num * tile;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total;
       t += (size_t)gridDim.x * blockDim.x) {
    const int p = (int)(t / tile), q = (int)(t % tile);
    out[t] = X[(size_t)idx[p] * tile + q];
  }
}

// max_i || p_i(A) - p_i(B) ||  -> partial max per block
__global__ void __launch_bounds__(kBlock) k_maxdist(const double *A, const double *B, int n, int r,
                                                    int dh, double *partials) {
  __shared__ double sm[kBlock];
  double m = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const size_t o = ((size_t)i * dh + (dh - 1)) * r;
    double s = 0.0;
    for (int q = 0; q < r; ++q) {
      const double df = A[o + q] - B[o + q];
      s = fma(df, df, s);
    }
    m = fmax(m, sqrt(s));
  }
  sm[threadIdx.x] = m;
  __syncthreads();
  for (int o = kBlock / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) sm[threadIdx.x] = fmax(sm[threadIdx.x], sm[threadIdx.x + o]);
    __syncthreads();
  }
  if (threadIdx.x == 0) partials[blockIdx.x] = sm[0];
}
__global__ void k_finalize_max(const double *partials, int nblocks, double *out) {
  double m = 0.0;
  for (int b = 0; b < nblocks; ++b) m = fmax(m, partials[b]);
  *out = m;
}

// reads buf[0 .. len) and keeps the compiler from dropping the loads (the sum is never the sentinel)
__global__ void k_flush_read(const double *buf, size_t len, double *sink) {
  double s = 0.0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < len; i += (size_t)gridDim.x * blockDim.x)
    s += buf[i];
  if (s == -1.2345e300) *sink = s;
}

__global__ void k_flush(double *buf, size_t len, double v) {
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < len;
       t += (size_t)gridDim.x * blockDim.x)
    buf[t] = v;
}

// ---------------------------------------------------------------------------------------------
// launch helpers
// ---------------------------------------------------------------------------------------------
static inline int pose_grid(const dpgo_dev *h, int lanes_per_pose) {
  // lanes_per_pose = D+1 for lane-group phases, 1 for thread-per-pose phases; one pass over
  // the poses (no SM-count cap: the hardware scheduler balances the many small CTAs)
  const int gpw = 32 / lanes_per_pose;
  const long warps = ((long)h->n + gpw - 1) / gpw;
  long blocks = (warps + kWarpsPerBlock - 1) / kWarpsPerBlock;
  const long cap = (long)h->partial_blocks;   // size of the partials buffer
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}
// staged per-pose kernels: one thread per pose, kPoseBlock threads per CTA, one pass over the poses
static inline int staged_grid(const dpgo_dev *h) {
  const long blocks = ((long)h->n + kPoseBlock - 1) / kPoseBlock;
  return (int)std::max(1L, blocks);
}
static inline int elem_grid(const dpgo_dev *h, size_t len) {
  long blocks = (long)((len + kBlock - 1) / kBlock);
  const long cap = (long)h->num_sms * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}
static int gemv_occupancy(dpgo_dev *h);
// persistent: exactly one wave of co-resident CTAs loops over the tiles
static inline int gemv_grid(dpgo_dev *h) {
  long tiles = (long)(h->ld / kGemvCols) * h->nsplit;
  const long cap = (long)h->num_sms * gemv_occupancy(h);
  if (tiles > cap) tiles = cap;
  if (tiles < 1) tiles = 1;
  return (int)tiles;
}
static inline BsrView qview(const dpgo_dev *h) { return BsrView{h->d_rowptr, h->d_colidx, h->d_blocks}; }
static inline BsrView cview(const dpgo_dev *h) { return BsrView{h->d_crowptr, h->d_ccolidx, h->d_cblocks}; }

template <int R>
static int gemv_setup(int *occ) {
  if (cudaFuncSetAttribute(k_precon_gemv<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemvDynSmem) != cudaSuccess)
    return -1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, k_precon_gemv<R>, kBlock, kGemvDynSmem) != cudaSuccess)
    return -1;
  return 0;
}
static int gemv_occupancy(dpgo_dev *h) {
  if (h->gemv_occ > 0) return h->gemv_occ;
  int occ = 0, rc = -1;
  switch (h->r) {
    case 2: rc = gemv_setup<2>(&occ); break;
    case 3: rc = gemv_setup<3>(&occ); break;
    case 4: rc = gemv_setup<4>(&occ); break;
    case 5: rc = gemv_setup<5>(&occ); break;
    case 6: rc = gemv_setup<6>(&occ); break;
  }
  if (rc != 0 || occ < 1) occ = 1;
  h->gemv_occ = occ;
  return occ;
}

#define LAUNCH_CHECK(h)                       \
  do {                                        \
    (h)->launches++;                          \
    CUDA_TRY(cudaPeekAtLastError());          \
  } while (0)

static int read_scalars(dpgo_dev *h, int nblocks, int K, double *out) {
  k_finalize<<<1, kBlock, 0, h->stream>>>(h->d_partials, nblocks, K, h->d_scalars);
  LAUNCH_CHECK(h);
  CUDA_TRY(cudaMemcpyAsync(h->h_scalars, h->d_scalars, K * sizeof(double), cudaMemcpyDeviceToHost,
                           h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  for (int k = 0; k < K; ++k) out[k] = h->h_scalars[k];
  return DPGO_OK;
}

int read_partials(dpgo_dev *h, int nblocks, int K, double *out) { return read_scalars(h, nblocks, K, out); }

int sync_host_blocks(dpgo_dev *h) {
  if (!h->host_blocks_stale) return DPGO_OK;
  CUDA_TRY(cudaMemcpyAsync(h->blocks.data(), h->d_blocks, h->blocks.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  h->host_blocks_stale = false;
  return DPGO_OK;
}

// --- op launchers (device pointers) ---
int op_qx(dpgo_dev *h, const BsrView &Q, const double *X, const double *G, double *out) {
  const int grid = pose_grid(h, h->d + 1);
  DPGO_DISPATCH(h, k_qx<R, D><<<grid, kBlock, 0, h->stream>>>(Q, X, G, out, h->n));
  LAUNCH_CHECK(h);
  return DPGO_OK;
}
// the stand-alone Q*X of dpgo_qx / dpgo_time_qx, in the variant chosen with dpgo_set_qx_variant
template <int R, int D>
static int qx_prefetch_distance(dpgo_dev *h) {
  // poses covered by the CTAs that are resident together: the row a newly scheduled CTA starts with
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_qx_prefetch<R, D>, kBlock, 0) != cudaSuccess || occ < 1) occ = 1;
  return h->num_sms * occ * kWarpsPerBlock * (32 / (D + 1));
}
// Automatic choice of the stand-alone Q*X form (qx_variant -1), from the round-2 measurements on the synthetic grids
// (profiles/r02_summary.md; fraction of the measured HBM peak, L2 flushed; one CTA per 64 poses / one resident wave):
//                      lane-group     + L2 prefetch    tiles in smem    two blocks / step
//   262 144 poses     0.595 / 0.616   0.546 / 0.575    0.547 / 0.580     0.604 / 0.635
//   1 000 000 poses   0.685 / 0.712   0.720 / 0.624    0.593 / 0.639     0.664 / 0.718
// from 100 000 poses the two-blocks-per-step form in one resident wave; below, Q and X are L2 resident and the plain
// kernel is used.
static const int kQxPipeMinPoses = 100000;

// Grid of the stand-alone Q*X.  One CTA per 8 x GPW poses leaves a tail at scale: a CTA lives ~15 us on the
// HBM-resident grids (its lane groups walk their block rows through dependent DRAM round trips), so the last partial
// wave runs at low occupancy.  The lane-group forms therefore run ONE resident wave (SMs x resident CTAs of the
// kernel) whose CTAs stride over the poses: 262 144 poses 0.604 -> 0.635, 1 000 000 poses 0.664 -> 0.716 of the HBM
// peak for the two-blocks-per-step form (profiles/r02_qx_resident_waves.jsonl; 2 and 4 waves are in between).  The
// prefetch form loses with a capped grid (its distance is tuned to the uncapped layout: 0.72 -> 0.62) and keeps one
// CTA per 64 poses.  DPGO_QX_RESIDENT_WAVES = w overrides (0 = uncapped) for measurements.
template <typename K>
static int qx_grid_for(dpgo_dev *h, K kernel, int default_waves) {
  const int need = pose_grid(h, h->d + 1);
  static int env_waves = -2;
  if (env_waves == -2) {
    const char *e = getenv("DPGO_QX_RESIDENT_WAVES");
    env_waves = e ? atoi(e) : -1;
  }
  const int waves = env_waves >= 0 ? env_waves : default_waves;
  if (waves <= 0) return need;
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kBlock, 0) != cudaSuccess || occ < 1) occ = 1;
  return std::min(need, waves * h->num_sms * occ);
}

int op_qx_main(dpgo_dev *h, const double *X, const double *G, double *out) {
  int variant = h->qx_variant;
  if (variant < 0) {
    variant = h->n >= kQxPipeMinPoses ? 3 : 0;
  }
  if (variant == 0) {
    DPGO_DISPATCH(h, k_qx<R, D><<<qx_grid_for(h, k_qx<R, D>, 1), kBlock, 0, h->stream>>>(qview(h), X, G, out, h->n));
    LAUNCH_CHECK(h);
    return DPGO_OK;
  }
  if (variant == 2) {
    DPGO_DISPATCH(h, k_qx_tiles<R, D><<<qx_grid_for(h, k_qx_tiles<R, D>, 1), kBlock, 0, h->stream>>>(qview(h), X, G, out, h->n));
    LAUNCH_CHECK(h);
    return DPGO_OK;
  }
  if (variant == 3) {
    DPGO_DISPATCH(h, k_qx_pipe<R, D><<<qx_grid_for(h, k_qx_pipe<R, D>, 1), kBlock, 0, h->stream>>>(qview(h), X, G, out, h->n));
    LAUNCH_CHECK(h);
    return DPGO_OK;
  }
  DPGO_DISPATCH(h, {
    const int dist = h->qx_prefetch_dist > 0 ? h->qx_prefetch_dist : qx_prefetch_distance<R, D>(h);
    k_qx_prefetch<R, D><<<qx_grid_for(h, k_qx_prefetch<R, D>, 0), kBlock, 0, h->stream>>>(qview(h), X, G, out, h->n, dist);
  });
  LAUNCH_CHECK(h);
  return DPGO_OK;
}
int op_fgrad(dpgo_dev *h, const double *X, double *EG, double *grad, double *S, double *f,
             double *gn2) {
  const int grid = pose_grid(h, h->d + 1);
  DPGO_DISPATCH(h, k_fgrad<R, D><<<grid, kBlock, 0, h->stream>>>(qview(h), X, h->d_G, EG, grad, S,
                                                               h->n, h->d_partials));
  LAUNCH_CHECK(h);
  double sc[2];
  DPGO_TRY(read_scalars(h, grid, 2, sc));
  *f = sc[0];
  *gn2 = sc[1];
  return DPGO_OK;
}
int op_hess(dpgo_dev *h, const double *Y, const double *S, const double *V, double *HV,
            const double *W, double *vhv, double *vw) {
  const int grid = pose_grid(h, h->d + 1);
  DPGO_DISPATCH(h, k_hess<R, D><<<grid, kBlock, 0, h->stream>>>(qview(h), Y, S, V, HV, W, h->n,
                                                              h->d_partials));
  LAUNCH_CHECK(h);
  if (vhv || vw) {
    double sc[2];
    DPGO_TRY(read_scalars(h, grid, 2, sc));
    if (vhv) *vhv = sc[0];
    if (vw) *vw = sc[1];
  }
  return DPGO_OK;
}
int op_tangent(dpgo_dev *h, const double *X, const double *V, double *out) {
  const int grid = pose_grid(h, h->d + 1);
  DPGO_DISPATCH(h, k_tangent<R, D><<<grid, kBlock, 0, h->stream>>>(X, V, out, h->n));
  LAUNCH_CHECK(h);
  return DPGO_OK;
}
int op_precon(dpgo_dev *h, const double *Y, const double *rvec, double *z, double *neg_out,
              double *z_r) {
  if (!h->has_precon) {
    set_error("preconditioner not built (dpgo_finalize(h, 1))");
    return DPGO_ESTATE;
  }
  if (h->precon_mode >= 2) return op_precon_dd(h, Y, rvec, z, neg_out, z_r);
  const int g1 = gemv_grid(h);
  DPGO_DISPATCH(h, k_precon_gemv<R><<<g1, kBlock, kGemvDynSmem, h->stream>>>(
                       h->d_Pinv, h->ld, rvec, h->d_zpart, h->vpad, h->KT, h->nsplit));
  LAUNCH_CHECK(h);
  const int g2 = pose_grid(h, h->d + 1);
  DPGO_DISPATCH(h, k_precon_finish<R, D><<<g2, kBlock, 0, h->stream>>>(
                       h->d_zpart, h->vpad, h->nsplit, Y, rvec, z, neg_out, h->n, h->d_partials));
  LAUNCH_CHECK(h);
  if (z_r) {
    double sc[1];
    DPGO_TRY(read_scalars(h, g2, 1, sc));
    *z_r = sc[0];
  }
  return DPGO_OK;
}
int op_retract(dpgo_dev *h, const double *X, const double *Eta, double *Xout) {
  const int grid = staged_grid(h);
  DPGO_DISPATCH(h, k_retract<R, D><<<grid, kPoseBlock, 0, h->stream>>>(X, Eta, Xout, h->n));
  LAUNCH_CHECK(h);
  return DPGO_OK;
}
int op_retract_scaled(dpgo_dev *h, const double *X, const double *Dir, double s, double *Xout) {
  const int grid = staged_grid(h);
  DPGO_DISPATCH(h, k_retract_scaled<R, D><<<grid, kPoseBlock, 0, h->stream>>>(X, Dir, s, Xout, h->n));
  LAUNCH_CHECK(h);
  return DPGO_OK;
}
int op_polar(dpgo_dev *h, double ca, const double *A, double cb, const double *B, double cc,
             const double *C, double *out) {
  const int grid = staged_grid(h);
  DPGO_DISPATCH(h, k_polar<R, D><<<grid, kPoseBlock, 0, h->stream>>>(ca, A, cb, B, cc, C, out, h->n));
  LAUNCH_CHECK(h);
  return DPGO_OK;
}
int op_step(dpgo_dev *h, double a, const double *delta, const double *Hd, double *eta, double *r,
            double *r_r) {
  const int grid = elem_grid(h, h->vlen);
  k_step<<<grid, kBlock, 0, h->stream>>>(a, delta, Hd, eta, r, h->vlen, h->d_partials);
  LAUNCH_CHECK(h);
  double sc[1];
  DPGO_TRY(read_scalars(h, grid, 1, sc));
  *r_r = sc[0];
  return DPGO_OK;
}
int op_axpby(dpgo_dev *h, double a, const double *x, double b, double *y) {
  const int grid = elem_grid(h, h->vlen);
  k_axpby<<<grid, kBlock, 0, h->stream>>>(a, x, b, y, h->vlen);
  LAUNCH_CHECK(h);
  return DPGO_OK;
}
int op_copy(dpgo_dev *h, const double *src, double *dst) {
  CUDA_TRY(cudaMemcpyAsync(dst, src, h->vlen * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  return DPGO_OK;
}

// ---------------------------------------------------------------------------------------------
// host-driven solver (one launch per op, one host read-back per scalar)
// ---------------------------------------------------------------------------------------------
struct SolveCtx {
  dpgo_dev *h;
  const dpgo_ropt_params *P;
  dpgo_ropt_result *res;
  double *x1, *x2, *EG, *EG2, *grad, *grad2, *S, *S2;
  double f1, gn2;
};

static int tcg_host(SolveCtx &c, double Delta, int *status, int *inner) {
  dpgo_dev *h = c.h;
  const dpgo_ropt_params &P = *c.P;
  TcgState s;
  DPGO_TRY(op_copy(h, c.grad, h->d_r));
  double z_r = 0.0;
  DPGO_TRY(op_precon(h, c.x1, h->d_r, h->d_z, h->d_delta, &z_r));
  c.res->n_precon++;
  tcg_begin(s, c.gn2, z_r);
  CUDA_TRY(cudaMemsetAsync(h->d_eta, 0, h->vlen * sizeof(double), h->stream));
  int j = 0;
  *inner = 0;
  for (j = 0; j < P.RTR_tCG_iterations; ++j) {
    double d_Hd = 0.0;
    DPGO_TRY(op_hess(h, c.x1, c.S, h->d_delta, h->d_Hd, nullptr, &d_Hd, nullptr));
    c.res->n_qx++;
    double step = 0.0;
    *inner = j + 1;
    if (tcg_curvature(s, d_Hd, Delta, &step)) {
      DPGO_TRY(op_axpby(h, step, h->d_delta, 1.0, h->d_eta));
      break;
    }
    double r_r = 0.0;
    DPGO_TRY(op_step(h, step, h->d_delta, h->d_Hd, h->d_eta, h->d_r, &r_r));
    if (tcg_converged(s, r_r, P.tcg_theta, P.tcg_kappa)) break;
    DPGO_TRY(op_precon(h, c.x1, h->d_r, h->d_z, nullptr, &z_r));
    c.res->n_precon++;
    const double beta = tcg_direction(s, z_r);
    DPGO_TRY(op_axpby(h, -1.0, h->d_z, beta, h->d_delta));
  }
  *status = s.status;
  return DPGO_OK;
}

// one outer RTR iteration from (x1, f1, grad); returns accepted flag, updates Delta
static int rtr_outer_host(SolveCtx &c, double *Delta, double max_Delta, bool *accepted) {
  dpgo_dev *h = c.h;
  const dpgo_ropt_params &P = *c.P;
  int status = TCG_MAXITER, inner = 0;
  DPGO_TRY(tcg_host(c, *Delta, &status, &inner));
  c.res->inner_iters += inner;
  c.res->tcg_status = status;
  DPGO_TRY(op_retract(h, c.x1, h->d_eta, c.x2));
  c.res->n_pose_sweeps++;
  double f2 = 0.0, gn2_2 = 0.0;
  DPGO_TRY(op_fgrad(h, c.x2, c.EG2, c.grad2, c.S2, &f2, &gn2_2));
  c.res->n_qx++;
  double eHe = 0.0, eg = 0.0;
  DPGO_TRY(op_hess(h, c.x1, c.S, h->d_eta, h->d_Hd, c.grad, &eHe, &eg));
  c.res->n_qx++;
  double rho = 0.0;
  *accepted = rtr_accept(c.f1, f2, eg, eHe, status, P.accept_rho, P.shrink, P.magnify, max_Delta,
                         Delta, &rho);
  if (P.verbose)
    printf("[dpgo_b200] RTR f=%.10g -> %.10g rho=%.4f Delta=%.4g inner=%d status=%d %s\n", c.f1, f2,
           rho, *Delta, inner, status, *accepted ? "accepted" : "REJECTED");
  if (*accepted) {
    std::swap(c.x1, c.x2);
    std::swap(c.EG, c.EG2);
    std::swap(c.grad, c.grad2);
    std::swap(c.S, c.S2);
    c.f1 = f2;
    c.gn2 = gn2_2;
    c.res->accepted++;
  } else {
    c.res->rejected++;
  }
  c.res->outer_iters++;
  return DPGO_OK;
}

int solve_host(dpgo_dev *h, const dpgo_ropt_params *P, const double *x_in, double *x_out,
               dpgo_ropt_result *res) {
  SolveCtx c;
  c.h = h; c.P = P; c.res = res;
  c.x1 = h->d_xa; c.x2 = h->d_xb;
  c.EG = h->d_EG; c.EG2 = h->d_EG2; c.grad = h->d_grad; c.grad2 = h->d_grad2;
  c.S = h->d_S; c.S2 = h->d_S2;
  DPGO_TRY(op_copy(h, x_in, c.x1));
  DPGO_TRY(op_fgrad(h, c.x1, c.EG, c.grad, c.S, &c.f1, &c.gn2));
  res->n_qx++;
  res->f_init = c.f1;
  res->gradnorm_init = sqrt(c.gn2);
  res->tcg_status = TCG_MAXITER;
  if (P->method == 1) {
    // QuadraticOptimizer::gradientDescent, ref: src/QuadraticOptimizer.cpp:110-137
    const double *dir = c.grad;
    if (P->RGD_use_preconditioner) {
      DPGO_TRY(op_precon(h, c.x1, c.grad, h->d_z, nullptr, nullptr));
      res->n_precon++;
      dir = h->d_z;
    }
    // eta = -stepsize * dir and the retraction in one kernel (the scale is applied inside)
    DPGO_TRY(op_retract_scaled(h, c.x1, dir, -P->RGD_stepsize, c.x2));
    res->n_pose_sweeps++;
    std::swap(c.x1, c.x2);
    DPGO_TRY(op_fgrad(h, c.x1, c.EG, c.grad, c.S, &c.f1, &c.gn2));
    res->n_qx++;
  } else if (sqrt(c.gn2) >= P->gradnorm_tol) {
    // QuadraticOptimizer::trustRegion, ref: src/QuadraticOptimizer.cpp:50-108
    if (P->RTR_iterations == 1) {
      double radius = P->RTR_initial_radius;
      int total_steps = 0;
      while (true) {
        double Delta = radius;
        bool acc = false;
        DPGO_TRY(rtr_outer_host(c, &Delta, radius, &acc));
        if (acc) break;
        if (total_steps > 10) break;  // "Too many RTR rejections. Returning initial guess."
        radius /= 4.0;
        total_steps++;
      }
    } else {
      double Delta = P->RTR_initial_radius;
      const double max_Delta = 5.0 * P->RTR_initial_radius;
      for (int it = 0; it < P->RTR_iterations; ++it) {
        bool acc = false;
        DPGO_TRY(rtr_outer_host(c, &Delta, max_Delta, &acc));
        if (sqrt(c.gn2) < P->gradnorm_tol) break;
      }
    }
  }
  res->f_opt = c.f1;
  res->gradnorm_opt = sqrt(c.gn2);
  res->success = 1;
  DPGO_TRY(op_copy(h, c.x1, x_out));
  return DPGO_OK;
}

// fused_rtr.cu
int solve_fused(dpgo_dev *h, const dpgo_ropt_params *P, const double *x_in, double *x_out,
                dpgo_ropt_result *res);
int solve_fused_launch(dpgo_dev *h, const dpgo_ropt_params *P, const double *x_in, double *x_out);
int solve_fused_collect(dpgo_dev *h, int verbose, dpgo_ropt_result *res);
int fused_phase_trace(dpgo_dev *h, double *busy_ms, int cap_ctas, int *num_ctas);

// ---------------------------------------------------------------------------------------------
// host-side graph -> block-CSR
// ---------------------------------------------------------------------------------------------
static void edge_T_W(const EdgeSet &E, int k, int d, double *T, double *W) {
  const int dh = d + 1;
  for (int a = 0; a < dh * dh; ++a) T[a] = 0.0;
  for (int a = 0; a < d; ++a) {
    for (int b = 0; b < d; ++b) T[a * dh + b] = E.R[(size_t)k * d * d + a * d + b];
    T[a * dh + d] = E.t[(size_t)k * d + a];
  }
  T[d * dh + d] = 1.0;
  for (int a = 0; a < d; ++a) W[a] = E.weight[k] * E.kappa[k];
  W[d] = E.weight[k] * E.tau[k];
}

// out (row-major dh x dh) for the 4 kinds; T row-major, W diagonal
static void block_of_kind(int kind, const double *T, const double *W, int dh, double *out) {
  for (int a = 0; a < dh; ++a)
    for (int b = 0; b < dh; ++b) {
      double v = 0.0;
      switch (kind) {
        case 0: for (int k = 0; k < dh; ++k) v += T[a * dh + k] * W[k] * T[b * dh + k]; break;
        case 1: v = -T[a * dh + b] * W[b]; break;
        case 2: v = -W[a] * T[b * dh + a]; break;
        case 3: v = (a == b) ? W[a] : 0.0; break;
      }
      out[a * dh + b] = v;
    }
}

static int build_Q_host(dpgo_dev *h) {
  const int d = h->d, dh = d + 1, n = h->n, bs = dh * dh;
  // the sorted contribution list is the sparsity pattern: kept across weight-only rebuilds (GNC)
  const bool reuse = h->weights_only_update && !h->q_contribs.empty();
  std::vector<Contribution> &cs = h->q_contribs;
  if (!reuse) {
    cs.clear();
    cs.reserve((size_t)4 * h->priv.m + h->shared.m + h->prior_idx.size() + n);
    for (int i = 0; i < n; ++i) cs.push_back({i, i, 0, 7});
    for (int k = 0; k < h->priv.m; ++k) {
      const int i = h->priv.a[k], j = h->priv.b[k];
      cs.push_back({i, i, k, 0});
      cs.push_back({i, j, k, 1});
      cs.push_back({j, i, k, 2});
      cs.push_back({j, j, k, 3});
    }
    for (int k = 0; k < h->shared.m; ++k) {
      const int i = h->shared.a[k];
      cs.push_back({i, i, k, (int8_t)(h->shared.outgoing[k] ? 4 : 5)});
    }
    for (size_t k = 0; k < h->prior_idx.size(); ++k) cs.push_back({h->prior_idx[k], h->prior_idx[k], (int32_t)k, 6});
    std::stable_sort(cs.begin(), cs.end(), [](const Contribution &x, const Contribution &y) {
      return x.row != y.row ? x.row < y.row : x.col < y.col;
    });
  }
  h->rowptr.assign(n + 1, 0);
  h->colidx.clear();
  h->blocks.clear();
  double T[16], W[4], B[16];
  int cur_row = -1, cur_col = -1;
  for (const Contribution &c : cs) {
    if (c.row != cur_row || c.col != cur_col) {
      h->colidx.push_back(c.col);
      h->blocks.insert(h->blocks.end(), bs, 0.0);
      h->rowptr[c.row + 1]++;
      cur_row = c.row;
      cur_col = c.col;
    }
    double *dst = h->blocks.data() + h->blocks.size() - bs;
    if (c.kind <= 3) {
      edge_T_W(h->priv, c.src, d, T, W);
      block_of_kind(c.kind, T, W, dh, B);
    } else if (c.kind == 4 || c.kind == 5) {
      edge_T_W(h->shared, c.src, d, T, W);
      block_of_kind(c.kind == 4 ? 0 : 3, T, W, dh, B);  // ref: src/PoseGraph.cpp:433-434, :457
    } else if (c.kind == 6) {
      for (int a = 0; a < bs; ++a) B[a] = 0.0;
      for (int a = 0; a < d; ++a) B[a * dh + a] = h->prior_kappa;  // ref: src/PoseGraph.cpp:462-469
      B[d * dh + d] = h->prior_tau;
    } else {
      continue;
    }
    for (int a = 0; a < bs; ++a) dst[a] += B[a];
  }
  for (int i = 0; i < n; ++i) h->rowptr[i + 1] += h->rowptr[i];
  h->nnzb = (int)h->colidx.size();
  return DPGO_OK;
}

// ---------------------------------------------------------------------------------------------
// weight-only refresh of Q and of the cross blocks on the device (GNC, ref: src/PGOAgent.cpp:1062-1142)
// ---------------------------------------------------------------------------------------------
// The kernels restate edge_T_W / block_of_kind / build_cross_host entry by entry with the same operations in the same
// order and WITHOUT contraction (the host code is compiled for baseline x86-64: separate multiply and add), so a
// refreshed Q has the bits of a host-assembled one (tests/test_gpu_b_team.py::test_device_weight_refresh_bitwise).
struct EdgeView { const double *R, *t, *kappa, *tau, *w; };
__device__ __forceinline__ double edge_T(const EdgeView &E, int k, int d, int a, int b) {   // T = [R t; 0 1]
  if (a < d) return (b < d) ? E.R[(size_t)k * d * d + a * d + b] : E.t[(size_t)k * d + a];
  return (b == d) ? 1.0 : 0.0;
}
__device__ __forceinline__ double edge_W(const EdgeView &E, int k, int d, int a) {
  return __dmul_rn(E.w[k], (a < d) ? E.kappa[k] : E.tau[k]);
}
__device__ __forceinline__ double block_entry(int kind, const EdgeView &E, int k, int d, int a, int b) {
  const int dh = d + 1;
  switch (kind) {
    case 0: {
      double v = 0.0;
      for (int c = 0; c < dh; ++c)
        v = __dadd_rn(v, __dmul_rn(__dmul_rn(edge_T(E, k, d, a, c), edge_W(E, k, d, c)), edge_T(E, k, d, b, c)));
      return v;
    }
    case 1: return __dmul_rn(-edge_T(E, k, d, a, b), edge_W(E, k, d, b));
    case 2: return __dmul_rn(-edge_W(E, k, d, a), edge_T(E, k, d, b, a));
    default: return (a == b) ? edge_W(E, k, d, a) : 0.0;   // 3
  }
}
__global__ void k_refresh_q(int nnzb, int d, const int *cptr, const int *csrc, const signed char *ckind, EdgeView P,
                            EdgeView S, double prior_kappa, double prior_tau, double *blocks) {
  const int dh = d + 1, bs = dh * dh;
  const size_t total = (size_t)nnzb * bs;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int e = (int)(idx / bs), ab = (int)(idx - (size_t)e * bs), a = ab / dh, b = ab - a * dh;
    double v = 0.0;
    for (int c = cptr[e]; c < cptr[e + 1]; ++c) {
      const int kind = ckind[c], src = csrc[c];
      double x;
      if (kind <= 3) x = block_entry(kind, P, src, d, a, b);
      else if (kind == 4 || kind == 5) x = block_entry(kind == 4 ? 0 : 3, S, src, d, a, b);
      else if (kind == 6) x = (a == b) ? (a < d ? prior_kappa : prior_tau) : 0.0;
      else continue;
      v = __dadd_rn(v, x);
    }
    blocks[idx] = v;
  }
}
// cross block p (sorted order) of shared edge k = order[p]:  B[a][b] = -M[b][a], M = W T^T (outgoing) or T W
__global__ void k_refresh_cross(int m, int d, const int *order, const unsigned char *outgoing, EdgeView S, double *cblocks) {
  const int dh = d + 1, bs = dh * dh;
  const size_t total = (size_t)m * bs;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int p = (int)(idx / bs), ab = (int)(idx - (size_t)p * bs), a = ab / dh, b = ab - a * dh;
    const int k = order[p];
    const double mv = outgoing[k] ? __dmul_rn(edge_W(S, k, d, b), edge_T(S, k, d, a, b))
                                  : __dmul_rn(edge_T(S, k, d, b, a), edge_W(S, k, d, a));
    cblocks[idx] = -mv;
  }
}

template <typename T>
static int upload(T **dptr, const std::vector<T> &v, size_t min_elems = 1) {
  if (*dptr) {
    CUDA_TRY(cudaFree(*dptr));
    *dptr = nullptr;
  }
  const size_t ne = std::max(v.size(), min_elems);
  CUDA_TRY(cudaMalloc((void **)dptr, ne * sizeof(T)));
  if (!v.empty()) CUDA_TRY(cudaMemcpy(*dptr, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return DPGO_OK;
}

static int upload_Q(dpgo_dev *h) {
  if (h->weights_only_update && h->d_blocks && h->uploaded_nnzb == h->nnzb) {   // same pattern: values only
    CUDA_TRY(cudaMemcpyAsync(h->d_blocks, h->blocks.data(), h->blocks.size() * sizeof(double), cudaMemcpyHostToDevice,
                             h->stream));
    return DPGO_OK;
  }
  h->uploaded_nnzb = h->nnzb;
  DPGO_TRY(upload(&h->d_rowptr, h->rowptr));
  DPGO_TRY(upload(&h->d_colidx, h->colidx));
  DPGO_TRY(upload(&h->d_blocks, h->blocks));
  std::vector<int32_t> browidx(h->nnzb);
  for (int i = 0; i < h->n; ++i)
    for (int e = h->rowptr[i]; e < h->rowptr[i + 1]; ++e) browidx[e] = i;
  DPGO_TRY(upload(&h->d_browidx, browidx));
  return DPGO_OK;
}

// cross blocks:  G = Gconst + Xnbr * C,  block row = my pose, block col = neighbour slot,
// stored block B (row-major) with out += X_slot * B^T  =>  B = -(M)^T,
// M = W T^T (outgoing, ref: src/PoseGraph.cpp:536) or T W (incoming, ref: :561).
static int build_cross_host(dpgo_dev *h) {
  const int d = h->d, dh = d + 1, n = h->n, bs = dh * dh, r = h->r;
  std::vector<int> order(h->shared.m);
  for (int k = 0; k < h->shared.m; ++k) order[k] = k;
  std::stable_sort(order.begin(), order.end(),
                   [&](int x, int y) { return h->shared.a[x] < h->shared.a[y]; });
  std::vector<int32_t> rowptr(n + 1, 0), colidx;
  std::vector<double> blocks;
  double T[16], W[4];
  for (int k : order) {
    const int i = h->shared.a[k];
    rowptr[i + 1]++;
    colidx.push_back(h->shared.b[k]);
    edge_T_W(h->shared, k, d, T, W);
    double B[16];
    for (int a = 0; a < dh; ++a)
      for (int b = 0; b < dh; ++b) {
        // B[a][b] = -M[b][a]
        double m;
        if (h->shared.outgoing[k]) m = W[b] * T[a * dh + b];   // M = W T^T: M[b][a] = W[b] T[a][b]
        else m = T[b * dh + a] * W[a];                          // M = T W  : M[b][a] = T[b][a] W[a]
        B[a * dh + b] = -m;
      }
    blocks.insert(blocks.end(), B, B + bs);
  }
  for (int i = 0; i < n; ++i) rowptr[i + 1] += rowptr[i];
  if (h->weights_only_update && h->d_cblocks && h->cnnzb == (int)colidx.size()) {   // same pattern: values only
    if (!blocks.empty())
      CUDA_TRY(cudaMemcpy(h->d_cblocks, blocks.data(), blocks.size() * sizeof(double), cudaMemcpyHostToDevice));
    return DPGO_OK;                                                                  // (priors do not carry weights)
  }
  h->cnnzb = (int)colidx.size();
  DPGO_TRY(upload(&h->d_crowptr, rowptr));
  DPGO_TRY(upload(&h->d_ccolidx, colidx));
  DPGO_TRY(upload(&h->d_cblocks, blocks));
  // constant part: priors, G_idx += -P W   (ref: src/PoseGraph.cpp:566-575)
  std::vector<double> Gc(h->vpad, 0.0);
  for (size_t p = 0; p < h->prior_idx.size(); ++p) {
    const int idx = h->prior_idx[p];
    const double *pose = h->prior_poses.data() + p * (size_t)r * dh;
    for (int c = 0; c < dh; ++c) {
      const double w = (c < d) ? h->prior_kappa : h->prior_tau;
      for (int q = 0; q < r; ++q) Gc[((size_t)idx * dh + c) * r + q] += -pose[c * r + q] * w;
    }
  }
  CUDA_TRY(cudaMemcpy(h->d_Gconst, Gc.data(), h->vpad * sizeof(double), cudaMemcpyHostToDevice));
  return DPGO_OK;
}

// Measured on B200 (profiles/r01_summary.md, tools/dd_probe.py): the two-level apply has a fixed cost
// of ~25-30 us (5 grid phases), the full dense apply streams N^2*8 bytes; inside the fused solver
// the two-level variant is ahead from N = 4000 (1.03 -> 0.87 ms per solve) and behind at N = 500.
static const int kAutoTwoLevelMinN = 3000;

// Device copy of what the refresh kernels read: uploaded whenever Q is assembled on the host (pattern may have changed).
static int upload_refresh_plan(dpgo_dev *h) {
  h->refresh_plan_valid = false;
  std::vector<int> cptr, csrc;
  std::vector<signed char> ckind;
  cptr.reserve(h->nnzb + 1);
  int cur_row = -1, cur_col = -1;
  for (const Contribution &c : h->q_contribs) {
    if (c.row != cur_row || c.col != cur_col) {
      cptr.push_back((int)csrc.size());
      cur_row = c.row;
      cur_col = c.col;
    }
    csrc.push_back(c.src);
    ckind.push_back(c.kind);
  }
  cptr.push_back((int)csrc.size());
  if ((int)cptr.size() != h->nnzb + 1) {
    set_error("contribution list does not match the pattern of Q");
    return DPGO_ESTATE;
  }
  DPGO_TRY(upload(&h->d_cptr, cptr));
  DPGO_TRY(upload(&h->d_csrc, csrc));
  DPGO_TRY(upload(&h->d_ckind, ckind));
  auto edges = [&](const EdgeSet &E, dpgo_dev::DeviceEdges &D) -> int {
    DPGO_TRY(upload(&D.R, E.R));
    DPGO_TRY(upload(&D.t, E.t));
    DPGO_TRY(upload(&D.kappa, E.kappa));
    DPGO_TRY(upload(&D.tau, E.tau));
    return upload(&D.w, E.weight);
  };
  DPGO_TRY(edges(h->priv, h->de_priv));
  DPGO_TRY(edges(h->shared, h->de_shared));
  std::vector<int> order(h->shared.m);                      // the row order of build_cross_host
  for (int k = 0; k < h->shared.m; ++k) order[k] = k;
  std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return h->shared.a[x] < h->shared.a[y]; });
  DPGO_TRY(upload(&h->d_corder, order));
  DPGO_TRY(upload(&h->d_sout, h->shared.outgoing));
  h->refresh_plan_valid = true;
  return DPGO_OK;
}

// GNC weight update with an unchanged pattern: new weights up, Q blocks and cross blocks re-weighted in place.
static int refresh_weights_device(dpgo_dev *h) {
  auto view = [](const dpgo_dev::DeviceEdges &D) { return EdgeView{D.R, D.t, D.kappa, D.tau, D.w}; };
  if (h->priv.m > 0)
    CUDA_TRY(cudaMemcpyAsync(h->de_priv.w, h->priv.weight.data(), (size_t)h->priv.m * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  if (h->shared.m > 0)
    CUDA_TRY(cudaMemcpyAsync(h->de_shared.w, h->shared.weight.data(), (size_t)h->shared.m * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  const int bs = (h->d + 1) * (h->d + 1);
  {
    const size_t total = (size_t)h->nnzb * bs;
    const int grid = (int)std::max<size_t>(1, std::min<size_t>((total + 255) / 256, (size_t)h->num_sms * 16));
    k_refresh_q<<<grid, 256, 0, h->stream>>>(h->nnzb, h->d, h->d_cptr, h->d_csrc, h->d_ckind, view(h->de_priv),
                                             view(h->de_shared), h->prior_kappa, h->prior_tau, h->d_blocks);
    LAUNCH_CHECK(h);
  }
  if (h->shared.m > 0) {
    const size_t total = (size_t)h->shared.m * bs;
    const int grid = (int)std::max<size_t>(1, std::min<size_t>((total + 255) / 256, (size_t)h->num_sms * 16));
    k_refresh_cross<<<grid, 256, 0, h->stream>>>(h->shared.m, h->d, h->d_corder, h->d_sout, view(h->de_shared), h->d_cblocks);
    LAUNCH_CHECK(h);
  }
  h->host_blocks_stale = true;
  return DPGO_OK;
}

static int build_precon(dpgo_dev *h) {
  h->precon_mode = (h->precon_request >= 0) ? h->precon_request : (h->N >= kAutoTwoLevelMinN ? 2 : 0);
  if (h->precon_mode == 2) {   // two-level exact preconditioner (precon_dd.cu)
    DPGO_TRY(dd_build(h));
    h->has_precon = true;
    return DPGO_OK;
  }
  const int N = h->N, ld = h->ld;
  // scratch: dense P = Q + 0.1 I, column-major N x N (leading dimension ld)
  double *A = nullptr;
  CUDA_TRY(cudaMalloc((void **)&A, (size_t)ld * N * sizeof(double)));
  CUDA_TRY(cudaMemsetAsync(A, 0, (size_t)ld * N * sizeof(double), h->stream));
  const int dh = h->d + 1;
  {
    const size_t total = (size_t)h->nnzb * dh * dh;
    int grid = (int)std::min<size_t>((total + 255) / 256, (size_t)h->num_sms * 16);
    if (grid < 1) grid = 1;
    k_scatter_dense<<<grid, 256, 0, h->stream>>>(h->d_browidx, h->d_colidx, h->d_blocks, h->nnzb, dh,
                                                 0.1 /* ref: src/PoseGraph.cpp:603 */, A, ld);
    LAUNCH_CHECK(h);
  }
  {
    // in-tree blocked Cholesky + triangular inverse + W^T W (dense_la.cu); lower triangle of A <- inverse
    const dla::SpdItem item{A, N, ld};
    const int rc = dla::spd_inverse_batched(h->stream, &item, 1, false);
    if (rc != 0) {
      cudaFree(A);
      if (rc > 0) {
        set_error("Cholesky of Q + 0.1 I failed (matrix not positive definite)");
        return DPGO_ENUMERIC;
      }
      set_error("%s", dla::last_error());
      return DPGO_ECUDA;
    }
  }
  if (h->d_Pinv) { CUDA_TRY(cudaFree(h->d_Pinv)); h->d_Pinv = nullptr; }
  if (h->d_zpart) { CUDA_TRY(cudaFree(h->d_zpart)); h->d_zpart = nullptr; }
  size_t npart;
  {
    const size_t total = (size_t)ld * h->ldk;
    CUDA_TRY(cudaMalloc((void **)&h->d_Pinv, total * sizeof(double)));
    int grid = (int)std::min<size_t>((total + 255) / 256, (size_t)h->num_sms * 32);
    if (grid < 1) grid = 1;
    k_tile_layout<<<grid, 256, 0, h->stream>>>(A, N, ld, h->d_Pinv, ld, h->ldk);
    LAUNCH_CHECK(h);
    npart = (size_t)h->nsplit;
  }
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  CUDA_TRY(cudaFree(A));
  CUDA_TRY(cudaMalloc((void **)&h->d_zpart, npart * h->vpad * sizeof(double)));
  CUDA_TRY(cudaMemsetAsync(h->d_zpart, 0, npart * h->vpad * sizeof(double), h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  h->has_precon = true;
  return DPGO_OK;
}

static int alloc_vec(dpgo_dev *h, double **p, size_t elems) {
  CUDA_TRY(cudaMalloc((void **)p, elems * sizeof(double)));
  CUDA_TRY(cudaMemset(*p, 0, elems * sizeof(double)));
  return DPGO_OK;
}

// Squared measurement errors kappa |Y1 R~ - Y2|_F^2 + tau |p2 - p1 - Y1 t~|^2 of a set of edges
// (ref: computeMeasurementError, src/DPGO_utils.cpp:501-507): one thread per edge.  Private
// edges take both poses from X; a shared edge takes my pose from X and the neighbour's from the
// slot buffer (outgoing: my pose is the tail).
__global__ void k_edge_errors(int m, int r, int d, const int *a, const int *b, const unsigned char *outgoing,
                              const double *Rm, const double *tm, const double *kappa, const double *tau,
                              const double *X, const double *nbr, double *out) {
  const int dh = d + 1, tile = r * dh;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < m; e += gridDim.x * blockDim.x) {
    const double *T1, *T2;
    if (nbr == nullptr) {                 // private edge: a = p1, b = p2
      T1 = X + (size_t)a[e] * tile;
      T2 = X + (size_t)b[e] * tile;
    } else if (outgoing[e]) {             // shared, my pose a[e] is the tail
      T1 = X + (size_t)a[e] * tile;
      T2 = nbr + (size_t)b[e] * tile;
    } else {
      T1 = nbr + (size_t)b[e] * tile;
      T2 = X + (size_t)a[e] * tile;
    }
    const double *R = Rm + (size_t)e * d * d, *t = tm + (size_t)e * d;
    double rot = 0.0, tr = 0.0;
    for (int q = 0; q < r; ++q) {
      double pt = T2[d * r + q] - T1[d * r + q];
      for (int c = 0; c < d; ++c) {
        double v = -T2[c * r + q];
        for (int k = 0; k < d; ++k) v = fma(T1[k * r + q], R[k * d + c], v);
        rot = fma(v, v, rot);
        pt = fma(-T1[c * r + q], t[c], pt);
      }
      tr = fma(pt, pt, tr);
    }
    out[e] = kappa[e] * rot + tau[e] * tr;
  }
}

static int copy_edges(EdgeSet &E, int m, int d, const int32_t *a, const int32_t *b,
                      const uint8_t *outgoing, const double *R, const double *t, const double *kappa,
                      const double *tau, const double *weight) {
  E.m = m;
  E.a.assign(a, a + m);
  E.b.assign(b, b + m);
  if (outgoing) E.outgoing.assign(outgoing, outgoing + m); else E.outgoing.assign(m, 0);
  E.R.assign(R, R + (size_t)m * d * d);
  E.t.assign(t, t + (size_t)m * d);
  E.kappa.assign(kappa, kappa + m);
  E.tau.assign(tau, tau + m);
  if (weight) E.weight.assign(weight, weight + m); else E.weight.assign(m, 1.0);
  return DPGO_OK;
}

}  // namespace dpgo

using namespace dpgo;

// =============================================================================================
// C-ABI
// =============================================================================================
#define H_CHECK(h)                                \
  do {                                            \
    CHECK_ARG((h) != nullptr);                    \
    CUDA_TRY(cudaSetDevice((h)->device));         \
  } while (0)

extern "C" {

const char *dpgo_last_error(void) { return dpgo::g_err; }
const char *dpgo_version(void) { return "dpgo_b200 0.1 (sm_100a)"; }

void dpgo_default_params(dpgo_ropt_params *p) {
  if (!p) return;
  p->method = 0;
  p->verbose = 0;
  p->gradnorm_tol = 1e-2;
  p->RGD_stepsize = 1e-3;
  p->RGD_use_preconditioner = 1;
  p->RTR_iterations = 3;
  p->RTR_tCG_iterations = 50;
  p->fused = 1;
  p->RTR_initial_radius = 100.0;
  p->tcg_theta = 1.0;
  p->tcg_kappa = 0.1;
  p->accept_rho = 0.1;
  p->shrink = 0.25;
  p->magnify = 2.0;
}

int dpgo_create(int device, int n, int d, int r, void *stream, dpgo_handle *out) {
  CHECK_ARG(out != nullptr);
  CHECK_ARG(n > 0);
  CHECK_ARG(d == 2 || d == 3);
  CHECK_ARG(r >= d && r <= d + 3);  // ref: CHECK(r >= d) src/PoseGraph.cpp:19
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0) {
    set_error("no CUDA device available (%s); dpgo_b200 has no CPU fallback",
              e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    return DPGO_ECUDA;
  }
  CHECK_ARG(device >= 0 && device < count);
  CUDA_TRY(cudaSetDevice(device));
  dpgo_dev *h = new dpgo_dev();
  h->device = device; h->n = n; h->d = d; h->r = r;
  h->N = (d + 1) * n;
  h->ld = ((h->N + kSymB - 1) / kSymB) * kSymB;  // multiple of 128 (and of 64): both apply variants
  h->vlen = (size_t)r * h->N;
  {
    int sms = 0;   // needed by the tiling below
    cudaError_t ea = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (ea != cudaSuccess || sms <= 0) {
      delete h;
      set_error("cannot read the SM count of device %d (%s)", device, cudaGetErrorString(ea));
      return DPGO_ECUDA;
    }
    h->num_sms = sms;
  }
  // tiling of the dense preconditioner apply: tiles of 64 output columns x KT inner indices.
  // Pick the number of inner splits (6..24) whose tile count balances best over one wave of
  // 2 CTAs per SM, counting the zero padding to ldk = nsplit * KT columns as waste.
  {
    const int ncb = h->ld / kGemvCols;
    const long wave = (long)h->num_sms * 2;
    double best = -1.0;
    for (int ns = 6; ns <= 24; ++ns) {
      int KT = (((h->ld + ns - 1) / ns + kStageK - 1) / kStageK) * kStageK;
      if (KT < 256) KT = 256;
      const int nsp = (h->ld + KT - 1) / KT;
      const long tiles = (long)ncb * nsp;
      const long rounds = (tiles + wave - 1) / wave;
      const double eff = ((double)tiles / (double)(rounds * wave)) * ((double)h->ld / ((double)nsp * KT));
      const double score = (tiles < wave) ? 0.5 * eff : eff;   // small problems: prefer filling the wave
      if (score > best + 1e-9) { best = score; h->KT = KT; h->nsplit = nsp; }
    }
    h->ldk = h->nsplit * h->KT;
  }
  h->vpad = (size_t)r * h->ldk;
  {
    // stream-ordered scratch of the preconditioner set-up is kept by the device's default pool instead of going
    // back to the driver at every synchronization (GNC rebuilds the preconditioner at every weight update)
    cudaMemPool_t pool = nullptr;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess && pool) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  if (stream) {
    h->stream = (cudaStream_t)stream;
  } else {
    CUDA_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    h->own_stream = true;
  }
  double **vecs[] = {&h->d_slot[0], &h->d_slot[1], &h->d_slot[2], &h->d_slot[3], &h->d_xa, &h->d_xb,
                     &h->d_EG, &h->d_EG2, &h->d_grad, &h->d_grad2, &h->d_eta, &h->d_r, &h->d_z,
                     &h->d_delta, &h->d_Hd, &h->d_t0, &h->d_t1, &h->d_t2, &h->d_G, &h->d_Gconst};
  for (double **p : vecs) DPGO_TRY(alloc_vec(h, p, h->vpad));
  DPGO_TRY(alloc_vec(h, &h->d_S, (size_t)n * d * d));
  DPGO_TRY(alloc_vec(h, &h->d_S2, (size_t)n * d * d));
  h->partial_blocks = std::max(h->num_sms * 32, (n + 63) / 64 + 1);
  DPGO_TRY(alloc_vec(h, &h->d_partials, (size_t)h->partial_blocks * 8));
  DPGO_TRY(alloc_vec(h, &h->d_scalars, 64));
  CUDA_TRY(cudaMallocHost((void **)&h->h_scalars, 64 * sizeof(double)));
  CUDA_TRY(cudaEventCreate(&h->ev0));
  CUDA_TRY(cudaEventCreate(&h->ev1));
  *out = h;
  return DPGO_OK;
}

int dpgo_destroy(dpgo_handle h) {
  if (!h) return DPGO_OK;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  dd_free(h);
  void *ptrs[] = {h->d_rowptr, h->d_colidx, h->d_browidx, h->d_blocks, h->d_crowptr, h->d_ccolidx,
                  h->d_cblocks, h->d_Gconst, h->d_G, h->d_nbr, h->d_Pinv, h->d_zpart, h->d_slot[0],
                  h->d_slot[1], h->d_slot[2], h->d_slot[3], h->d_xa, h->d_xb, h->d_EG, h->d_EG2,
                  h->d_grad, h->d_grad2, h->d_S, h->d_S2, h->d_eta, h->d_r, h->d_z, h->d_delta,
                  h->d_Hd, h->d_t0, h->d_t1, h->d_t2, h->d_partials, h->d_scalars, h->d_fused,
                  h->d_public_idx, h->d_flush, h->d_trace, h->d_nbr_xy[0], h->d_nbr_xy[1], h->d_cptr, h->d_csrc,
                  h->d_ckind, h->d_corder, h->d_sout, h->de_priv.R, h->de_priv.t, h->de_priv.kappa, h->de_priv.tau,
                  h->de_priv.w, h->de_shared.R, h->de_shared.t, h->de_shared.kappa, h->de_shared.tau, h->de_shared.w};
  for (void *p : ptrs)
    if (p) cudaFree(p);
  if (h->h_scalars) cudaFreeHost(h->h_scalars);
  if (h->h_fused) cudaFreeHost(h->h_fused);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->own_stream) cudaStreamDestroy(h->stream);
  delete h;
  return DPGO_OK;
}

int dpgo_dims(dpgo_handle h, int *n, int *d, int *r) {
  CHECK_ARG(h != nullptr);
  if (n) *n = h->n;
  if (d) *d = h->d;
  if (r) *r = h->r;
  return DPGO_OK;
}

int dpgo_launch_count(dpgo_handle h, int64_t *count) {
  CHECK_ARG(h != nullptr && count != nullptr);
  *count = h->launches;
  return DPGO_OK;
}

int dpgo_sync(dpgo_handle h) {
  H_CHECK(h);
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return DPGO_OK;
}

int dpgo_set_private_edges(dpgo_handle h, int m, const int32_t *p1, const int32_t *p2,
                           const double *R, const double *t, const double *kappa,
                           const double *tau, const double *weight) {
  H_CHECK(h);
  CHECK_ARG(m >= 0);
  CHECK_ARG(m == 0 || (p1 && p2 && R && t && kappa && tau));
  for (int k = 0; k < m; ++k) {
    CHECK_ARG(p1[k] >= 0 && p1[k] < h->n);
    CHECK_ARG(p2[k] >= 0 && p2[k] < h->n);
  }
  h->finalized = false;
  return copy_edges(h->priv, m, h->d, p1, p2, nullptr, R, t, kappa, tau, weight);
}

int dpgo_set_shared_edges(dpgo_handle h, int m, int num_nbr_slots, const int32_t *my_idx,
                          const int32_t *nbr_slot, const uint8_t *outgoing, const double *R,
                          const double *t, const double *kappa, const double *tau,
                          const double *weight) {
  H_CHECK(h);
  CHECK_ARG(m >= 0 && num_nbr_slots >= 0);
  CHECK_ARG(m == 0 || (my_idx && nbr_slot && outgoing && R && t && kappa && tau));
  for (int k = 0; k < m; ++k) {
    CHECK_ARG(my_idx[k] >= 0 && my_idx[k] < h->n);
    CHECK_ARG(nbr_slot[k] >= 0 && nbr_slot[k] < num_nbr_slots);
  }
  h->finalized = false;
  h->num_nbr_slots = num_nbr_slots;
  if (h->d_nbr) { CUDA_TRY(cudaFree(h->d_nbr)); h->d_nbr = nullptr; }
  const size_t ne = (size_t)std::max(num_nbr_slots, 1) * h->r * (h->d + 1);
  DPGO_TRY(alloc_vec(h, &h->d_nbr, ne));
  return copy_edges(h->shared, m, h->d, my_idx, nbr_slot, outgoing, R, t, kappa, tau, weight);
}

int dpgo_set_priors(dpgo_handle h, int num, const int32_t *idx, const double *poses,
                    double prior_kappa, double prior_tau) {
  H_CHECK(h);
  CHECK_ARG(num >= 0);
  CHECK_ARG(num == 0 || (idx && poses));
  for (int k = 0; k < num; ++k) CHECK_ARG(idx[k] >= 0 && idx[k] < h->n);  // ref: CHECK_LT(index, n())
  h->prior_idx.assign(idx, idx + num);
  h->prior_poses.assign(poses, poses + (size_t)num * h->r * (h->d + 1));
  h->prior_kappa = prior_kappa;
  h->prior_tau = prior_tau;
  h->finalized = false;
  return DPGO_OK;
}

int dpgo_set_precon_mode(dpgo_handle h, int mode) {
  CHECK_ARG(h != nullptr);
  CHECK_ARG(mode == -1 || mode == 0 || mode == 2);
  if (mode != h->precon_request) {
    h->precon_request = mode;
    h->has_precon = false;
  }
  return DPGO_OK;
}

int dpgo_two_level_partition(int n, const int32_t *rowptr, const int32_t *colidx, int dh,
                             int max_domain_poses, int32_t *group, int *num_domains) {
  CHECK_ARG(n >= 0 && rowptr && (colidx || n == 0 || rowptr[n] == 0) && group && num_domains);
  CHECK_ARG(dh >= 2 && dh <= 4);
  for (int i = 0; i < n; ++i) {
    CHECK_ARG(rowptr[i] <= rowptr[i + 1]);
    for (int e = rowptr[i]; e < rowptr[i + 1]; ++e) CHECK_ARG(colidx[e] >= 0 && colidx[e] < n);
  }
  const std::vector<std::vector<int>> adj = bsr_adjacency(n, rowptr, colidx);
  Dissector ds(adj, max_domain_poses > 0 ? max_domain_poses : two_level_max_domain_poses(dh));
  std::vector<int> all(n);
  for (int i = 0; i < n; ++i) all[i] = i;
  ds.run(std::move(all));
  for (int i = 0; i < n; ++i) group[i] = -1;
  for (size_t k = 0; k < ds.domains.size(); ++k)
    for (int v : ds.domains[k]) group[v] = (int32_t)k;
  *num_domains = (int)ds.domains.size();
  return DPGO_OK;
}

int dpgo_get_precon_mode(dpgo_handle h, int *mode) {
  CHECK_ARG(h != nullptr && mode != nullptr);
  if (!h->has_precon) { set_error("preconditioner not built"); return DPGO_ESTATE; }
  *mode = h->precon_mode;
  return DPGO_OK;
}

int dpgo_set_precon_tuning(dpgo_handle h, int split_interior, int split_schur, int prefetch) {
  CHECK_ARG(h != nullptr);
  CHECK_ARG(split_interior <= 16 && split_schur <= 64);
  h->dd_split1 = split_interior > 0 ? split_interior : 0;
  h->dd_split3 = split_schur > 0 ? split_schur : 0;
  h->dd_prefetch = (prefetch < 0) ? 1 : (prefetch ? 1 : 0);
  h->has_precon = false;
  return DPGO_OK;
}

int dpgo_set_qx_variant(dpgo_handle h, int variant, int prefetch_distance) {
  CHECK_ARG(h != nullptr && variant >= -1 && variant <= 3 && prefetch_distance >= 0);
  h->qx_variant = variant;
  h->qx_prefetch_dist = prefetch_distance;
  return DPGO_OK;
}

int dpgo_set_two_level_domain_size(dpgo_handle h, int max_domain_poses) {
  CHECK_ARG(h != nullptr && max_domain_poses >= 0);
  h->dd_max_domain = max_domain_poses;
  h->has_precon = false;
  return DPGO_OK;
}

int dpgo_finalize(dpgo_handle h, int build_precon_flag) {
  H_CHECK(h);
  if (h->weights_only_update && h->refresh_plan_valid && h->cnnzb == h->shared.m) {
    DPGO_TRY(refresh_weights_device(h));       // GNC: same pattern, values re-weighted on the device
  } else {
    DPGO_TRY(build_Q_host(h));
    h->host_blocks_stale = false;
    DPGO_TRY(upload_Q(h));
    DPGO_TRY(build_cross_host(h));
    DPGO_TRY(upload_refresh_plan(h));
  }
  // G starts as its constant part (no neighbour poses yet)
  CUDA_TRY(cudaMemcpyAsync(h->d_G, h->d_Gconst, h->vpad * sizeof(double), cudaMemcpyDeviceToDevice,
                           h->stream));
  h->has_precon = false;
  if (build_precon_flag) DPGO_TRY(build_precon(h));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  h->finalized = true;
  return DPGO_OK;
}

int dpgo_update_weights(dpgo_handle h, const double *w_private, const double *w_shared,
                        int build_precon_flag) {
  H_CHECK(h);
  if (w_private) h->priv.weight.assign(w_private, w_private + h->priv.m);
  if (w_shared) h->shared.weight.assign(w_shared, w_shared + h->shared.m);
  h->weights_only_update = h->finalized;     // same pattern: the two-level set-up keeps its symbolic part
  const int rc = dpgo_finalize(h, build_precon_flag);
  h->weights_only_update = false;
  return rc;
}

int dpgo_get_Q_bsr(dpgo_handle h, int *nnzb, int32_t *rowptr, int32_t *colidx, double *blocks) {
  CHECK_ARG(h != nullptr);
  if (!h->finalized) { set_error("dpgo_finalize not called"); return DPGO_ESTATE; }
  if (nnzb) *nnzb = h->nnzb;
  if (rowptr) memcpy(rowptr, h->rowptr.data(), h->rowptr.size() * sizeof(int32_t));
  if (colidx) memcpy(colidx, h->colidx.data(), h->colidx.size() * sizeof(int32_t));
  if (blocks) {
    cudaSetDevice(h->device);
    DPGO_TRY(sync_host_blocks(h));   // re-weighted on the device since the host assembled them
    memcpy(blocks, h->blocks.data(), h->blocks.size() * sizeof(double));
  }
  return DPGO_OK;
}

#define NEED_FINAL(h)                                                   \
  do {                                                                  \
    if (!(h)->finalized) {                                              \
      set_error("dpgo_finalize must be called before compute calls");   \
      return DPGO_ESTATE;                                               \
    }                                                                   \
  } while (0)

static int h2d(dpgo_dev *h, double *dst, const double *src) {
  CUDA_TRY(cudaMemcpyAsync(dst, src, h->vlen * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  return DPGO_OK;
}
static int d2h(dpgo_dev *h, double *dst, const double *src) {
  CUDA_TRY(cudaMemcpyAsync(dst, src, h->vlen * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return DPGO_OK;
}

int dpgo_set_G(dpgo_handle h, const double *G_host) {
  H_CHECK(h);
  CHECK_ARG(G_host != nullptr);
  return h2d(h, h->d_G, G_host);
}

static int rebuild_G(dpgo_dev *h) {
  return op_qx(h, cview(h), h->d_nbr, h->d_Gconst, h->d_G);
}

int dpgo_set_neighbor_poses(dpgo_handle h, const double *tiles_host) {
  H_CHECK(h);
  NEED_FINAL(h);
  CHECK_ARG(h->num_nbr_slots == 0 || tiles_host != nullptr);
  if (h->num_nbr_slots > 0)
    CUDA_TRY(cudaMemcpyAsync(h->d_nbr, tiles_host,
                             (size_t)h->num_nbr_slots * h->r * (h->d + 1) * sizeof(double),
                             cudaMemcpyHostToDevice, h->stream));
  return rebuild_G(h);
}

int dpgo_set_neighbor_poses_dev(dpgo_handle h, const double *tiles_dev) {
  H_CHECK(h);
  NEED_FINAL(h);
  CHECK_ARG(h->num_nbr_slots == 0 || tiles_dev != nullptr);
  if (h->num_nbr_slots > 0)
    CUDA_TRY(cudaMemcpyAsync(h->d_nbr, tiles_dev,
                             (size_t)h->num_nbr_slots * h->r * (h->d + 1) * sizeof(double),
                             cudaMemcpyDeviceToDevice, h->stream));
  return rebuild_G(h);
}

int dpgo_get_G(dpgo_handle h, double *G_host) {
  H_CHECK(h);
  CHECK_ARG(G_host != nullptr);
  return d2h(h, G_host, h->d_G);
}

int dpgo_qx(dpgo_handle h, const double *X, double *out) {
  H_CHECK(h); NEED_FINAL(h);
  CHECK_ARG(X && out);
  DPGO_TRY(h2d(h, h->d_t0, X));
  DPGO_TRY(op_qx_main(h, h->d_t0, nullptr, h->d_t1));
  return d2h(h, out, h->d_t1);
}

int dpgo_f(dpgo_handle h, const double *X, double *f) {
  H_CHECK(h); NEED_FINAL(h);
  CHECK_ARG(X && f);
  DPGO_TRY(h2d(h, h->d_t0, X));
  double gn2;
  return op_fgrad(h, h->d_t0, h->d_EG2, h->d_grad2, h->d_S2, f, &gn2);
}

int dpgo_egrad(dpgo_handle h, const double *X, double *out) {
  H_CHECK(h); NEED_FINAL(h);
  CHECK_ARG(X && out);
  DPGO_TRY(h2d(h, h->d_t0, X));
  DPGO_TRY(op_qx_main(h, h->d_t0, h->d_G, h->d_t1));
  return d2h(h, out, h->d_t1);
}

int dpgo_rgrad(dpgo_handle h, const double *X, double *out, double *norm) {
  H_CHECK(h); NEED_FINAL(h);
  CHECK_ARG(X != nullptr);
  DPGO_TRY(h2d(h, h->d_t0, X));
  double f, gn2;
  DPGO_TRY(op_fgrad(h, h->d_t0, h->d_EG2, h->d_grad2, h->d_S2, &f, &gn2));
  if (norm) *norm = sqrt(gn2);
  if (out) return d2h(h, out, h->d_grad2);
  return DPGO_OK;
}

int dpgo_hessvec(dpgo_handle h, const double *X, const double *V, double *out) {
  H_CHECK(h); NEED_FINAL(h);
  CHECK_ARG(X && V && out);
  DPGO_TRY(h2d(h, h->d_t0, X));
  DPGO_TRY(h2d(h, h->d_t1, V));
  double f, gn2;
  DPGO_TRY(op_fgrad(h, h->d_t0, h->d_EG2, h->d_grad2, h->d_S2, &f, &gn2));
  DPGO_TRY(op_hess(h, h->d_t0, h->d_S2, h->d_t1, h->d_t2, nullptr, nullptr, nullptr));
  return d2h(h, out, h->d_t2);
}

int dpgo_precon(dpgo_handle h, const double *X, const double *V, double *out) {
  H_CHECK(h); NEED_FINAL(h);
  CHECK_ARG(X && V && out);
  DPGO_TRY(h2d(h, h->d_t0, X));
  DPGO_TRY(h2d(h, h->d_t1, V));
  DPGO_TRY(op_precon(h, h->d_t0, h->d_t1, h->d_t2, nullptr, nullptr));
  return d2h(h, out, h->d_t2);
}

int dpgo_tangent_project(dpgo_handle h, const double *X, const double *V, double *out) {
  H_CHECK(h);
  CHECK_ARG(X && V && out);
  DPGO_TRY(h2d(h, h->d_t0, X));
  DPGO_TRY(h2d(h, h->d_t1, V));
  DPGO_TRY(op_tangent(h, h->d_t0, h->d_t1, h->d_t2));
  return d2h(h, out, h->d_t2);
}

int dpgo_retract(dpgo_handle h, const double *X, const double *V, double *out) {
  H_CHECK(h);
  CHECK_ARG(X && V && out);
  DPGO_TRY(h2d(h, h->d_t0, X));
  DPGO_TRY(h2d(h, h->d_t1, V));
  DPGO_TRY(op_retract(h, h->d_t0, h->d_t1, h->d_t2));
  return d2h(h, out, h->d_t2);
}

int dpgo_project_manifold(dpgo_handle h, const double *M, double *out) {
  H_CHECK(h);
  CHECK_ARG(M && out);
  DPGO_TRY(h2d(h, h->d_t0, M));
  DPGO_TRY(op_polar(h, 1.0, h->d_t0, 0.0, nullptr, 0.0, nullptr, h->d_t2));
  return d2h(h, out, h->d_t2);
}

static int run_solver(dpgo_dev *h, const dpgo_ropt_params *params, const double *x_in_dev,
                      double *x_out_dev, dpgo_ropt_result *result) {
  dpgo_ropt_params P;
  if (params) P = *params; else dpgo_default_params(&P);
  CHECK_ARG(P.method == 0 || P.method == 1);
  CHECK_ARG(P.RTR_iterations >= 0 && P.RTR_tCG_iterations >= 0);
  if ((P.method == 0 || P.RGD_use_preconditioner) && !h->has_precon) {
    set_error("preconditioner not built (dpgo_finalize(h, 1))");
    return DPGO_ESTATE;
  }
  dpgo_ropt_result res;
  memset(&res, 0, sizeof(res));
  const int64_t l0 = h->launches;
  CUDA_TRY(cudaEventRecord(h->ev0, h->stream));
  int rc;
  if (P.fused && P.method == 0) rc = solve_fused(h, &P, x_in_dev, x_out_dev, &res);
  else rc = solve_host(h, &P, x_in_dev, x_out_dev, &res);
  if (rc != DPGO_OK) return rc;
  CUDA_TRY(cudaEventRecord(h->ev1, h->stream));
  CUDA_TRY(cudaEventSynchronize(h->ev1));
  float ms = 0.f;
  CUDA_TRY(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  res.elapsed_ms = ms;
  res.n_launches = h->launches - l0;
  if (result) *result = res;
  return DPGO_OK;
}

int dpgo_optimize(dpgo_handle h, const dpgo_ropt_params *params, const double *X0, double *Xout,
                  dpgo_ropt_result *result) {
  H_CHECK(h); NEED_FINAL(h);
  if (X0) DPGO_TRY(h2d(h, h->d_slot[DPGO_SLOT_X], X0));
  DPGO_TRY(run_solver(h, params, h->d_slot[DPGO_SLOT_X], h->d_slot[DPGO_SLOT_X], result));
  if (Xout) return d2h(h, Xout, h->d_slot[DPGO_SLOT_X]);
  return DPGO_OK;
}

int dpgo_optimize_slot(dpgo_handle h, const dpgo_ropt_params *params, int from,
                       dpgo_ropt_result *result) {
  H_CHECK(h); NEED_FINAL(h);
  CHECK_ARG(from >= 0 && from < 4);
  return run_solver(h, params, h->d_slot[from], h->d_slot[DPGO_SLOT_X], result);
}

// Stream-ordered form of dpgo_optimize_slot: the fused solve and the copy of its result block are
// queued on the handle's stream and the call returns without waiting, so that a driver can queue a
// whole RBCD round (pack, exchange, G, solve, Nesterov updates) ahead of the device.
int dpgo_optimize_slot_async(dpgo_handle h, const dpgo_ropt_params *params, int from) {
  H_CHECK(h); NEED_FINAL(h);
  CHECK_ARG(from >= 0 && from < 4);
  dpgo_ropt_params P;
  if (params) P = *params; else dpgo_default_params(&P);
  CHECK_ARG(P.method == 0 && P.fused != 0);   // only the single-launch RTR solver has no host round trips
  CHECK_ARG(P.RTR_iterations >= 0 && P.RTR_tCG_iterations >= 0);
  if (!h->has_precon) {
    set_error("preconditioner not built (dpgo_finalize(h, 1))");
    return DPGO_ESTATE;
  }
  h->pending_l0 = h->launches;
  h->pending_verbose = P.verbose;
  CUDA_TRY(cudaEventRecord(h->ev0, h->stream));
  DPGO_TRY(solve_fused_launch(h, &P, h->d_slot[from], h->d_slot[DPGO_SLOT_X]));
  CUDA_TRY(cudaEventRecord(h->ev1, h->stream));
  h->pending_launches = h->launches - h->pending_l0;
  h->pending = true;
  return DPGO_OK;
}

// Waits for the most recent dpgo_optimize_slot_async of this handle and returns its result block.
int dpgo_optimize_result(dpgo_handle h, dpgo_ropt_result *result) {
  H_CHECK(h);
  CHECK_ARG(result != nullptr);
  if (!h->pending) {
    set_error("no asynchronous solve is pending on this handle");
    return DPGO_ESTATE;
  }
  dpgo_ropt_result res;
  memset(&res, 0, sizeof(res));
  DPGO_TRY(solve_fused_collect(h, h->pending_verbose, &res));
  CUDA_TRY(cudaEventSynchronize(h->ev1));
  float ms = 0.f;
  CUDA_TRY(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  res.elapsed_ms = ms;
  res.n_launches = h->pending_launches;
  h->pending = false;
  *result = res;
  return DPGO_OK;
}

int dpgo_slot_set(dpgo_handle h, int slot, const double *host) {
  H_CHECK(h);
  CHECK_ARG(slot >= 0 && slot < 4 && host);
  return h2d(h, h->d_slot[slot], host);
}
int dpgo_slot_get(dpgo_handle h, int slot, double *host) {
  H_CHECK(h);
  CHECK_ARG(slot >= 0 && slot < 4 && host);
  return d2h(h, host, h->d_slot[slot]);
}
int dpgo_slot_copy(dpgo_handle h, int dst, int src) {
  H_CHECK(h);
  CHECK_ARG(dst >= 0 && dst < 4 && src >= 0 && src < 4);
  if (dst == src) return DPGO_OK;
  return op_copy(h, h->d_slot[src], h->d_slot[dst]);
}

int dpgo_nesterov_update_Y(dpgo_handle h, double alpha) {
  H_CHECK(h);
  return op_polar(h, 1.0 - alpha, h->d_slot[DPGO_SLOT_X], alpha, h->d_slot[DPGO_SLOT_V], 0.0, nullptr,
                  h->d_slot[DPGO_SLOT_Y]);
}
int dpgo_nesterov_update_V(dpgo_handle h, double gamma) {
  H_CHECK(h);
  return op_polar(h, 1.0, h->d_slot[DPGO_SLOT_V], gamma, h->d_slot[DPGO_SLOT_X], -gamma,
                  h->d_slot[DPGO_SLOT_Y], h->d_slot[DPGO_SLOT_V]);
}

int dpgo_set_public_indices(dpgo_handle h, int num_public, const int32_t *idx) {
  H_CHECK(h);
  CHECK_ARG(num_public >= 0);
  CHECK_ARG(num_public == 0 || idx);
  for (int k = 0; k < num_public; ++k) CHECK_ARG(idx[k] >= 0 && idx[k] < h->n);
  std::vector<int32_t> v(idx, idx + num_public);
  h->num_public = num_public;
  return upload(&h->d_public_idx, v);
}

int dpgo_pack_public_dev(dpgo_handle h, int slot, double *tiles_dev) {
  H_CHECK(h);
  CHECK_ARG(slot >= 0 && slot < 4);
  if (h->num_public == 0) return DPGO_OK;
  CHECK_ARG(tiles_dev != nullptr);
  const int tile = h->r * (h->d + 1);
  const size_t total = (size_t)h->num_public * tile;
  int grid = (int)std::min<size_t>((total + 255) / 256, (size_t)h->num_sms * 8);
  k_gather_tiles<<<grid, 256, 0, h->stream>>>(h->d_slot[slot], h->d_public_idx, h->num_public, tile,
                                              tiles_dev);
  LAUNCH_CHECK(h);
  return DPGO_OK;
}

int dpgo_gather_tiles_dev(dpgo_handle h, int slot, int num, const int32_t *idx_dev,
                          double *tiles_dev) {
  H_CHECK(h);
  CHECK_ARG(slot >= 0 && slot < 4 && num >= 0);
  if (num == 0) return DPGO_OK;
  CHECK_ARG(idx_dev != nullptr && tiles_dev != nullptr);
  const int tile = h->r * (h->d + 1);
  const size_t total = (size_t)num * tile;
  int grid = (int)std::min<size_t>((total + 255) / 256, (size_t)h->num_sms * 8);
  k_gather_tiles<<<grid, 256, 0, h->stream>>>(h->d_slot[slot], idx_dev, num, tile, tiles_dev);
  LAUNCH_CHECK(h);
  return DPGO_OK;
}

}  // extern "C"

namespace {
template <typename T>
int to_device(const std::vector<T> &v, T **out) {
  *out = nullptr;
  if (v.empty()) return DPGO_OK;
  CUDA_TRY(cudaMalloc((void **)out, v.size() * sizeof(T)));
  CUDA_TRY(cudaMemcpy(*out, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return DPGO_OK;
}

// errors of one edge set: uploads the (small) edge arrays, runs k_edge_errors, copies the result back
int edge_errors(dpgo_dev *h, const EdgeSet &E, const double *X, const double *nbr, double *out_host) {
  if (E.m == 0) return DPGO_OK;
  int *a = nullptr, *b = nullptr;
  unsigned char *og = nullptr;
  double *R = nullptr, *t = nullptr, *ka = nullptr, *ta = nullptr, *out = nullptr;
  int rc = DPGO_OK;
  auto cleanup = [&]() {
    cudaFree(a); cudaFree(b); cudaFree(og); cudaFree(R); cudaFree(t); cudaFree(ka); cudaFree(ta); cudaFree(out);
  };
  std::vector<int> av(E.a.begin(), E.a.end()), bv(E.b.begin(), E.b.end());
  std::vector<unsigned char> ogv(E.outgoing.begin(), E.outgoing.end());
  if ((rc = to_device(av, &a)) || (rc = to_device(bv, &b)) || (rc = to_device(ogv, &og)) ||
      (rc = to_device(E.R, &R)) || (rc = to_device(E.t, &t)) || (rc = to_device(E.kappa, &ka)) ||
      (rc = to_device(E.tau, &ta))) {
    cleanup();
    return rc;
  }
  if (cudaMalloc((void **)&out, (size_t)E.m * sizeof(double)) != cudaSuccess) {
    cleanup();
    set_error("allocation of the edge error buffer failed");
    return DPGO_ECUDA;
  }
  const int grid = std::max(1, std::min((E.m + 255) / 256, h->num_sms * 8));
  k_edge_errors<<<grid, 256, 0, h->stream>>>(E.m, h->r, h->d, a, b, og, R, t, ka, ta, X, nbr, out);
  h->launches++;
  cudaError_t e = cudaPeekAtLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(out_host, out, (size_t)E.m * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cleanup();
  if (e != cudaSuccess) {
    set_error("edge error kernel failed: %s", cudaGetErrorString(e));
    return DPGO_ECUDA;
  }
  return DPGO_OK;
}
}  // namespace

extern "C" {

int dpgo_measurement_errors(dpgo_handle h, int slot, const double *nbr_poses_dev, double *err_private,
                            double *err_shared) {
  H_CHECK(h);
  CHECK_ARG(slot >= 0 && slot < 4);
  CHECK_ARG((h->priv.m == 0 || err_private) && (h->shared.m == 0 || err_shared));
  if (h->priv.m > 0) DPGO_TRY(edge_errors(h, h->priv, h->d_slot[slot], nullptr, err_private));
  if (h->shared.m > 0) {
    const double *nbr = nbr_poses_dev ? nbr_poses_dev : h->d_nbr;
    if (!nbr) { set_error("no neighbour poses"); return DPGO_ESTATE; }
    DPGO_TRY(edge_errors(h, h->shared, h->d_slot[slot], nbr, err_shared));
  }
  return DPGO_OK;
}

int dpgo_round_trajectory(dpgo_handle h, int slot, const double *anchor_tile, double *T_host) {
  H_CHECK(h);
  CHECK_ARG(slot >= 0 && slot < 4 && T_host != nullptr);
  const int tile = h->r * (h->d + 1), out_tile = h->d * (h->d + 1);
  // scratch: d_t0 holds the rounded poses (d(d+1) n doubles <= r(d+1) n), d_t1 the anchor tile
  const double *anchor = h->d_slot[slot];                 // local frame: pose 0 of the slot itself
  if (anchor_tile) {
    CUDA_TRY(cudaMemcpyAsync(h->d_t1, anchor_tile, (size_t)tile * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    anchor = h->d_t1;
  }
  const int grid = staged_grid(h);
  DPGO_DISPATCH(h, k_round<R, D><<<grid, kPoseBlock, 0, h->stream>>>(h->d_slot[slot], anchor, h->d_t0, h->n));
  LAUNCH_CHECK(h);
  CUDA_TRY(cudaMemcpyAsync(T_host, h->d_t0, (size_t)out_tile * h->n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return DPGO_OK;
}

int dpgo_max_translation_distance(dpgo_handle h, int slot_a, int slot_b, double *out) {
  H_CHECK(h);
  CHECK_ARG(slot_a >= 0 && slot_a < 4 && slot_b >= 0 && slot_b < 4 && out);
  const int grid = pose_grid(h, 1);
  k_maxdist<<<grid, kBlock, 0, h->stream>>>(h->d_slot[slot_a], h->d_slot[slot_b], h->n, h->r,
                                            h->d + 1, h->d_partials);
  LAUNCH_CHECK(h);
  k_finalize_max<<<1, 1, 0, h->stream>>>(h->d_partials, grid, h->d_scalars);
  LAUNCH_CHECK(h);
  CUDA_TRY(cudaMemcpyAsync(h->h_scalars, h->d_scalars, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  *out = h->h_scalars[0];
  return DPGO_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// Chordal initialization on the device.
// ref: chordalInitialization src/DPGO_solver.cpp:220-269, recoverTranslations src/DPGO_utils.cpp.
//
// The reference solves two sparse least-squares problems with SPQR:
//   rotations     min sum_e kappa_e || R_j - R_i R_ij ||_F^2   with R_0 = I, then every R_i is projected to SO(d)
//   translations  min sum_e tau_e  || t_j - t_i - R_i t_ij ||^2 with t_0 = 0
// Both are quadratic forms of connection Laplacians the library already builds: with the rotation-only measurements
// (tau = 0, t = 0) the cost of X = [R_1 0 | R_2 0 | ...] is tr(X Q_rot X^T); with the full measurements the cost of
// X = [R_1 t_1 | ...] at fixed rotations is tr(X Q X^T), quadratic in the t_i.  So each stage is a handle at r = d,
// and the normal equations  P Q P x = -P Q x0  (P = keep the unknown columns: rotation columns / translation column
// of the poses >= 1) are solved by conjugate gradients preconditioned with P (Q + 0.1 I)^-1 P -- the exact operator
// the solver uses as its preconditioner -- all on the device: Q*X, the two-level / dense inverse, the update kernels
// of the tCG loop.  The rows of X are independent right-hand sides of the same system; they are iterated as one
// vector (one alpha / beta per iteration).
// ---------------------------------------------------------------------------------------------
namespace dpgo {

// a <- P a  (zero outside the unknown columns: cls 0 = rotation columns, 1 = translation column, poses >= 1);
// partials[block] = <P a, b>
__global__ void k_mask_dot(double *a, const double *b, int r, int dh, int cls, size_t len, double *partials) {
  double acc[1] = {0.0};
  for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < len; k += (size_t)gridDim.x * blockDim.x) {
    const size_t col = k / r;
    const size_t pose = col / dh;
    const int c = (int)(col - pose * dh);
    const bool keep = pose >= 1 && (cls == 0 ? c < dh - 1 : c == dh - 1);
    const double v = keep ? a[k] : 0.0;
    a[k] = v;
    acc[0] = fma(v, b[k], acc[0]);
  }
  block_reduce_store<1>(acc, partials + blockIdx.x);
}

static int mask_dot(dpgo_dev *h, double *a, const double *b, int cls, double *out) {
  const int grid = elem_grid(h, h->vlen);
  k_mask_dot<<<grid, kBlock, 0, h->stream>>>(a, b, h->r, h->d + 1, cls, h->vlen, h->d_partials);
  LAUNCH_CHECK(h);
  return read_scalars(h, grid, 1, out);
}

// Solves P Q P x = -g for the unknown columns of class `cls`; g = h->d_r on entry (masked here), x = h->d_eta on
// exit.  h->d_t2 must be zero (stands in for the base point of the preconditioner's tangent projection, which
// is the identity there).
static int chordal_pcg(dpgo_dev *h, int cls, double tol, int max_iter, int *iters, double *relres) {
  double rr0 = 0.0, z_r = 0.0, r_r = 0.0;
  CUDA_TRY(cudaMemsetAsync(h->d_eta, 0, h->vpad * sizeof(double), h->stream));
  DPGO_TRY(mask_dot(h, h->d_r, h->d_r, cls, &rr0));
  *iters = 0;
  *relres = 0.0;
  if (!(rr0 > 0.0)) return DPGO_OK;
  DPGO_TRY(op_precon(h, h->d_t2, h->d_r, h->d_z, nullptr, nullptr));
  DPGO_TRY(mask_dot(h, h->d_z, h->d_r, cls, &z_r));
  DPGO_TRY(op_axpby(h, -1.0, h->d_z, 0.0, h->d_delta));          // delta = -z
  r_r = rr0;
  for (int j = 0; j < max_iter; ++j) {
    double d_Hd = 0.0;
    DPGO_TRY(op_qx_main(h, h->d_delta, nullptr, h->d_Hd));
    DPGO_TRY(mask_dot(h, h->d_Hd, h->d_delta, cls, &d_Hd));
    if (!(d_Hd > 0.0)) break;                                      // null direction (disconnected part): stop
    const double alpha = z_r / d_Hd;
    DPGO_TRY(op_step(h, alpha, h->d_delta, h->d_Hd, h->d_eta, h->d_r, &r_r));
    *iters = j + 1;
    if (sqrt(r_r) <= tol * sqrt(rr0)) break;
    double z_r_new = 0.0;
    DPGO_TRY(op_precon(h, h->d_t2, h->d_r, h->d_z, nullptr, nullptr));
    DPGO_TRY(mask_dot(h, h->d_z, h->d_r, cls, &z_r_new));
    const double beta = z_r_new / z_r;
    z_r = z_r_new;
    DPGO_TRY(op_axpby(h, -1.0, h->d_z, beta, h->d_delta));
  }
  *relres = sqrt(r_r / rr0);
  return DPGO_OK;
}

// x0 (slot 0) -> g = P (x0 Q) in d_r; solve; slot 0 += x
static int chordal_stage(dpgo_dev *h, int cls, double tol, int max_iter, int *iters, double *relres) {
  CUDA_TRY(cudaMemsetAsync(h->d_t2, 0, h->vpad * sizeof(double), h->stream));
  DPGO_TRY(op_qx_main(h, h->d_slot[0], nullptr, h->d_r));
  DPGO_TRY(chordal_pcg(h, cls, tol, max_iter, iters, relres));
  return op_axpby(h, 1.0, h->d_eta, 1.0, h->d_slot[0]);
}

}  // namespace dpgo

extern "C" int dpgo_chordal_initialization(int device, int n, int d, int m, const int32_t *p1, const int32_t *p2,
                                           const double *R, const double *t, const double *kappa, const double *tau,
                                           double *T_host, dpgo_chordal_info *info) {
  CHECK_ARG(n >= 1 && (d == 2 || d == 3) && m >= 0 && T_host != nullptr);
  CHECK_ARG(m == 0 || (p1 && p2 && R && t && kappa && tau));
  const int dh = d + 1;
  const size_t len = (size_t)d * dh * n;
  dpgo_chordal_info inf;
  memset(&inf, 0, sizeof(inf));
  // identity poses: the answer for n == 1, and the fixed first pose otherwise
  std::vector<double> X(len, 0.0);
  for (int i = 0; i < n; ++i)
    for (int a = 0; a < d; ++a) X[((size_t)i * dh + a) * d + a] = (i == 0 || n == 1) ? 1.0 : 0.0;
  if (n == 1 || m == 0) {
    for (int i = 0; i < n; ++i)
      for (int a = 0; a < d; ++a) X[((size_t)i * dh + a) * d + a] = 1.0;
    memcpy(T_host, X.data(), len * sizeof(double));
    if (info) *info = inf;
    return DPGO_OK;
  }
  if (device < 0) CUDA_TRY(cudaGetDevice(&device));
  const double tol = 1e-13;
  const int max_iter = 5000;
  int rc = DPGO_OK;
  dpgo_handle hr = nullptr, ht = nullptr;
  auto done = [&](int code) {
    if (hr) dpgo_destroy(hr);
    if (ht) dpgo_destroy(ht);
    if (info) *info = inf;
    return code;
  };
  // ---- rotations (ref :224-251)
  {
    std::vector<double> zt((size_t)m * d, 0.0), ztau((size_t)m, 0.0);
    if ((rc = dpgo_create(device, n, d, d, nullptr, &hr)) != DPGO_OK) return done(rc);
    if ((rc = dpgo_set_private_edges(hr, m, p1, p2, R, zt.data(), kappa, ztau.data(), nullptr)) != DPGO_OK) return done(rc);
    if ((rc = dpgo_finalize(hr, 1)) != DPGO_OK) return done(rc);
    if ((rc = h2d(hr, hr->d_slot[0], X.data())) != DPGO_OK) return done(rc);
    if ((rc = chordal_stage(hr, 0, tol, max_iter, &inf.rotation_iterations, &inf.rotation_residual)) != DPGO_OK) return done(rc);
    // projectToRotationGroup of every block (ref :247-250): the rounding kernel in the frame of the identity pose
    // (slot 0's first tile is [I 0]); its d x (d+1) output tiles are the r = d pose tiles [R_i 0]
    const int grid = staged_grid(hr);
    DPGO_DISPATCH(hr, k_round<R, D><<<grid, kPoseBlock, 0, hr->stream>>>(hr->d_slot[0], hr->d_slot[0], hr->d_t0, hr->n));
    LAUNCH_CHECK(hr);
    if ((rc = d2h(hr, X.data(), hr->d_t0)) != DPGO_OK) return done(rc);
    inf.launches += hr->launches;
    dpgo_destroy(hr);
    hr = nullptr;
  }
  // ---- translations (ref :253-256, recoverTranslations)
  {
    if ((rc = dpgo_create(device, n, d, d, nullptr, &ht)) != DPGO_OK) return done(rc);
    if ((rc = dpgo_set_private_edges(ht, m, p1, p2, R, t, kappa, tau, nullptr)) != DPGO_OK) return done(rc);
    if ((rc = dpgo_finalize(ht, 1)) != DPGO_OK) return done(rc);
    if ((rc = h2d(ht, ht->d_slot[0], X.data())) != DPGO_OK) return done(rc);
    if ((rc = chordal_stage(ht, 1, tol, max_iter, &inf.translation_iterations, &inf.translation_residual)) != DPGO_OK) return done(rc);
    if ((rc = d2h(ht, T_host, ht->d_slot[0])) != DPGO_OK) return done(rc);
    inf.launches += ht->launches;
  }
  return done(DPGO_OK);
}

// ---- measurement helpers -------------------------------------------------------------------
static int ensure_flush(dpgo_dev *h) {
  if (!h->d_flush) {
    h->flush_bytes = (size_t)512 << 20;  // > 126 MB L2
    CUDA_TRY(cudaMalloc((void **)&h->d_flush, h->flush_bytes));
  }
  return DPGO_OK;
}
// L2 flush in front of a timed launch: write a buffer four times the L2, then READ half of it.  The write alone
// leaves the L2 full of dirty lines whose write-back the timed kernel would pay for (measured: +7 us on an 81 us
// Q*X); after the read pass the L2 holds clean lines of unrelated data -- cold and clean, as ncu's own cache
// control leaves it.
static int do_flush(dpgo_dev *h) {
  const size_t len = h->flush_bytes / sizeof(double);
  k_flush<<<h->num_sms * 8, 256, 0, h->stream>>>(h->d_flush, len, 1.0);
  k_flush_read<<<h->num_sms * 8, 256, 0, h->stream>>>(h->d_flush, len / 2, h->d_flush + len - 1);
  CUDA_TRY(cudaPeekAtLastError());
  return DPGO_OK;
}

template <typename F>
static int time_launches(dpgo_dev *h, int reps, int flush, F launch, double *usec) {
  CHECK_ARG(reps > 0 && usec);
  if (flush) DPGO_TRY(ensure_flush(h));
  for (int w = 0; w < 3; ++w) DPGO_TRY(launch());
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  double total_ms = 0.0;
  if (flush) {
    for (int k = 0; k < reps; ++k) {
      DPGO_TRY(do_flush(h));
      CUDA_TRY(cudaEventRecord(h->ev0, h->stream));
      DPGO_TRY(launch());
      CUDA_TRY(cudaEventRecord(h->ev1, h->stream));
      CUDA_TRY(cudaEventSynchronize(h->ev1));
      float ms = 0.f;
      CUDA_TRY(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
      total_ms += ms;
    }
  } else {
    CUDA_TRY(cudaEventRecord(h->ev0, h->stream));
    for (int k = 0; k < reps; ++k) DPGO_TRY(launch());
    CUDA_TRY(cudaEventRecord(h->ev1, h->stream));
    CUDA_TRY(cudaEventSynchronize(h->ev1));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    total_ms = ms;
  }
  *usec = total_ms * 1000.0 / reps;
  return DPGO_OK;
}

extern "C" {

int dpgo_time_qx(dpgo_handle h, int reps, int flush_l2, double *usec) {
  H_CHECK(h); NEED_FINAL(h);
  return time_launches(h, reps, flush_l2,
                       [&]() { return op_qx_main(h, h->d_slot[0], nullptr, h->d_t1); }, usec);
}

int dpgo_time_precon(dpgo_handle h, int reps, int flush_l2, double *usec) {
  H_CHECK(h); NEED_FINAL(h);
  if (!h->has_precon) { set_error("preconditioner not built"); return DPGO_ESTATE; }
  if (h->precon_mode >= 2)
    return time_launches(h, reps, flush_l2, [&]() { return dd_time_apply(h, h->d_slot[0]); }, usec);
  return time_launches(h, reps, flush_l2, [&]() {
    const int g1 = gemv_grid(h);
    DPGO_DISPATCH(h, k_precon_gemv<R><<<g1, kBlock, kGemvDynSmem, h->stream>>>(
                         h->d_Pinv, h->ld, h->d_slot[0], h->d_zpart, h->vpad, h->KT, h->nsplit));
    LAUNCH_CHECK(h);
    return DPGO_OK;
  }, usec);
}

int dpgo_time_pose_op(dpgo_handle h, int op, int reps, int flush_l2, double *usec) {
  H_CHECK(h);
  CHECK_ARG(op >= 0 && op <= 2);
  if (op == 0)
    return time_launches(h, reps, flush_l2, [&]() { return op_retract(h, h->d_slot[0], h->d_slot[1], h->d_t2); }, usec);
  if (op == 1)
    return time_launches(h, reps, flush_l2, [&]() {
      return op_polar(h, 0.5, h->d_slot[0], 0.3, h->d_slot[1], 0.2, h->d_slot[2], h->d_t2); }, usec);
  return time_launches(h, reps, flush_l2, [&]() {
    const int grid = staged_grid(h);
    DPGO_DISPATCH(h, k_round<R, D><<<grid, kPoseBlock, 0, h->stream>>>(h->d_slot[0], h->d_slot[0], h->d_t0, h->n));
    LAUNCH_CHECK(h);
    return DPGO_OK;
  }, usec);
}

int dpgo_phase_trace(dpgo_handle h, double *busy_ms, int cap_ctas, int *num_ctas) {
  H_CHECK(h);
  CHECK_ARG(num_ctas != nullptr && cap_ctas >= 0 && (busy_ms != nullptr || cap_ctas == 0));
  return fused_phase_trace(h, busy_ms, cap_ctas, num_ctas);
}

int dpgo_bytes_qx(dpgo_handle h, double *bytes) {
  CHECK_ARG(h && bytes);
  NEED_FINAL(h);
  const double dh = h->d + 1;
  // nnzb*((d+1)^2*8 + 4) + (n+1)*4 + 2*r*(d+1)*n*8
  *bytes = (double)h->nnzb * (dh * dh * 8 + 4) + ((double)h->n + 1) * 4 + 2.0 * h->r * dh * h->n * 8;
  return DPGO_OK;
}

int dpgo_bytes_precon(dpgo_handle h, double *bytes) {
  CHECK_ARG(h && bytes);
  // dense inverse read once (all of it, or its lower triangle incl. diagonal when the symmetric
  // half-storage variant is active) + vector read + result written
  if (h->precon_mode >= 2) {
    *bytes = dd_bytes(h);
    return DPGO_OK;
  }
  const double N = (double)h->N;
  const double mat = N * N * 8;
  *bytes = mat + 2.0 * h->r * N * 8;
  return DPGO_OK;
}

}  // extern "C"
