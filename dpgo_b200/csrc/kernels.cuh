// Device-side building blocks of the RBCD local solve (sm_100a).  Every phase is a __device__
// function over a "grid context" so that the same code runs (a) as a stand-alone kernel (one
// launch per op: parity tests, host-driven solver) and (b) inside the persistent fused RTR
// kernel with grid-wide barriers between phases.
//
// Data layout: every r x N array is column-major FP64; pose i is the contiguous tile
// [i*TILE, (i+1)*TILE) of R*(D+1) doubles (160 B for r=5,d=3).  A pose is processed by a
// "lane group" of D+1 adjacent lanes of a warp, lane c owning column c of the tile (R doubles in
// registers); d x d cross-column quantities are exchanged with warp shuffles.
#pragma once
// DPGO_CPU_EMU: tests/native/cuda_emu.h has defined the CUDA keywords, builtins and the PTX helpers below
// for a host build that runs ONE CTA with real threads (a CPU test of the device functions of the
// three-phase preconditioner; never part of the product libraries)
#ifndef DPGO_CPU_EMU
#include <cuda_runtime.h>
#endif
#include <stdint.h>


namespace dpgo {

constexpr int kBlock = 256;          // threads per CTA for every phase
constexpr int kWarpsPerBlock = kBlock / 32;
constexpr int kGemvCols = 64;        // output columns per dense-precon tile (2 per lane)

struct BsrView {
  const int *rowptr;     // [n+1]
  const int *colidx;     // [nnzb]
  const double *blocks;  // [nnzb][(D+1)*(D+1)]  block (i,j) of Q, row-major
};

template <int R, int D>
struct Geo {
  static constexpr int DH = D + 1;
  static constexpr int TILE = R * DH;
  static constexpr int GPW = 32 / DH;  // pose groups per warp
};

struct Ctx {
  int tid;      // global thread id
  int nthreads; // threads in grid
  int warp;     // global warp id
  int nwarps;   // warps in grid
  int lane;
};

__device__ __forceinline__ Ctx make_ctx() {
  Ctx c;
  c.tid = blockIdx.x * blockDim.x + threadIdx.x;
  c.nthreads = gridDim.x * blockDim.x;
  c.warp = c.tid >> 5;
  c.nwarps = c.nthreads >> 5;
  c.lane = threadIdx.x & 31;
  return c;
}

// Warp numbering for the persistent fused solver: warp w of CTA b is global warp b + grid * w, so
// that a phase with fewer warps of work than the grid holds (2500 poses = 313 warps) is dealt over
// every SM instead of filling the first CTAs (latency-bound phases: more LSUs / L1s in parallel).
__device__ __forceinline__ Ctx make_ctx_spread() {
  Ctx c;
  c.nthreads = gridDim.x * blockDim.x;
  c.nwarps = c.nthreads >> 5;
  c.lane = threadIdx.x & 31;
  c.warp = blockIdx.x + gridDim.x * (threadIdx.x >> 5);
  c.tid = c.warp * 32 + c.lane;
  return c;
}

struct LanePos {
  int grp;        // pose group within the warp
  int c;          // column of the tile owned by this lane
  int base_lane;  // first lane of the group
  bool ok;        // lane participates (false for the 32 % (D+1) spare lanes)
};

template <int D>
__device__ __forceinline__ LanePos lane_pos(int lane) {
  constexpr int DH = D + 1;
  LanePos p;
  p.grp = lane / DH;
  p.c = lane - p.grp * DH;
  p.ok = p.grp < (32 / DH);
  p.base_lane = p.grp * DH;
  if (!p.ok) p.base_lane = 32 - DH;  // keep shuffle sources in range
  return p;
}

// ---------------------------------------------------------------------------------------------
// reductions
// ---------------------------------------------------------------------------------------------
// Deterministic block reduction of K running sums; thread 0 stores them to out[0..K).
template <int K>
__device__ __forceinline__ void block_reduce_store(double (&v)[K], double *out) {
  __shared__ double sred[K][kWarpsPerBlock];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double x = v[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    if (lane == 0) sred[k][w] = x;
  }
  __syncthreads();
  if (w == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      double x = (lane < kWarpsPerBlock) ? sred[k][lane] : 0.0;
#pragma unroll
      for (int o = kWarpsPerBlock / 2; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
      if (lane == 0) out[k] = x;
    }
  }
  __syncthreads();
}

// Sum per-block partials [nblocks][K] in a fixed order; every thread of the CTA gets the result.
// (Plain loads: the partials were written by other CTAs before a grid-wide barrier.)
template <int K>
__device__ __forceinline__ void sum_partials(const double *partials, int nblocks, double (&out)[K]) {
  __shared__ double sres[K];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (w == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      double x = 0.0;
      for (int b = lane; b < nblocks; b += 32) x += partials[(size_t)b * K + k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
      if (lane == 0) sres[k] = x;
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) out[k] = sres[k];
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// tile helpers
// ---------------------------------------------------------------------------------------------
template <int R>
__device__ __forceinline__ void load_col(const double *p, double (&v)[R]) {
#pragma unroll
  for (int k = 0; k < R; ++k) v[k] = p[k];
}
template <int R>
__device__ __forceinline__ void store_col(double *p, const double (&v)[R]) {
#pragma unroll
  for (int k = 0; k < R; ++k) p[k] = v[k];
}
template <int R>
__device__ __forceinline__ double dot_col(const double (&a)[R], const double (&b)[R]) {
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < R; ++k) s = fma(a[k], b[k], s);
  return s;
}

// Loads of the lane-group SpMM.  At scale the kernel is bound by L1 data-pipe wavefronts (ncu, 262 144-pose grid:
// l1tex__data_pipe_lsu_wavefronts 72 % of peak, DRAM 47 %): a lane group reads its Q row and the whole X tile, 16
// bytes per lane and instruction, 8 different lines per warp instruction.  256-bit loads (LDG.E.256, one per
// 32-byte sector) were measured in round 2 and are not used: 92.3 / 298.3 us against 88.2 / 282.1 us with 128-bit
// loads on the 262 144- / 1 000 000-pose grids (profiles/r02_qx_scale.md).
// row c of block e of Q (DH doubles) and the pose tile j of X (TILE doubles), widest loads the shape allows
template <int DH>
__device__ __forceinline__ void load_q_row(const double *m, double (&mk)[DH]) {
  if constexpr (DH % 2 == 0) {            // row c of the block is 16-byte aligned: 128-bit read-only loads
#pragma unroll
    for (int k = 0; k < DH / 2; ++k) {
      const double2 v = __ldg(reinterpret_cast<const double2 *>(m) + k);
      mk[2 * k] = v.x;
      mk[2 * k + 1] = v.y;
    }
  } else {
#pragma unroll
    for (int k = 0; k < DH; ++k) mk[k] = __ldg(m + k);
  }
}
template <int TILE>
__device__ __forceinline__ void load_x_tile(const double *xj, double (&x)[TILE]) {
  if constexpr (TILE % 2 == 0) {          // 16-byte aligned tile: 128-bit loads
#pragma unroll
    for (int k = 0; k < TILE / 2; ++k) {
      const double2 v = *(reinterpret_cast<const double2 *>(xj) + k);
      x[2 * k] = v.x;
      x[2 * k + 1] = v.y;
    }
  } else {
#pragma unroll
    for (int k = 0; k < TILE; ++k) x[k] = xj[k];
  }
}

// One column (c) of (X * Q) for pose i: sum over the block row of Q.
//   out[:, (i,c)] = sum_j sum_k X_j[:, k] * Q_ij[c][k]        (Q symmetric)
// X tiles are gathered with plain (coherent) loads; Q blocks / indices through the read-only path.
template <int R, int D>
__device__ __forceinline__ void spmm_col(const BsrView &Q, const double *X, int i, int c,
                                         double (&acc)[R]) {
  constexpr int DH = D + 1, TILE = R * DH;
  const int e0 = __ldg(Q.rowptr + i), e1 = __ldg(Q.rowptr + i + 1);
  for (int e = e0; e < e1; ++e) {
    const int j = __ldg(Q.colidx + e);
    const double *m = Q.blocks + (size_t)e * (DH * DH) + c * DH;
    const double *xj = X + (size_t)j * TILE;
    double mk[DH], x[TILE];
    load_q_row<DH>(m, mk);
    load_x_tile<TILE>(xj, x);
#pragma unroll
    for (int k = 0; k < DH; ++k) {
#pragma unroll
      for (int q = 0; q < R; ++q) acc[q] = fma(x[k * R + q], mk[k], acc[q]);
    }
  }
}

// Same sum in the same order (bit-identical), two blocks per step: the loads of blocks e and e + 1 (Q rows and
// X tiles) are all issued before the first FMA, and the column indices of the next step are fetched during
// the current one, so a step costs one memory latency instead of the dependent chain colidx -> X tile, and a
// lane group keeps two blocks' loads in flight.  Used where the row walk is latency bound: Q*X streaming from
// HBM (stand-alone kernel at scale) -- more registers than spmm_col (two tiles live).
template <int R, int D>
__device__ __forceinline__ void spmm_col2(const BsrView &Q, const double *X, int i, int c, double (&acc)[R]) {
  constexpr int DH = D + 1, TILE = R * DH;
  const int e0 = __ldg(Q.rowptr + i), e1 = __ldg(Q.rowptr + i + 1);
  auto load_m = [&](int e, double (&mk)[DH]) { load_q_row<DH>(Q.blocks + (size_t)e * (DH * DH) + c * DH, mk); };
  auto load_x = [&](int j, double (&x)[TILE]) { load_x_tile<TILE>(X + (size_t)j * TILE, x); };
  auto fmas = [&](const double (&x)[TILE], const double (&mk)[DH]) {
#pragma unroll
    for (int k = 0; k < DH; ++k) {
#pragma unroll
      for (int q = 0; q < R; ++q) acc[q] = fma(x[k * R + q], mk[k], acc[q]);
    }
  };
  int e = e0;
  int ja = (e < e1) ? __ldg(Q.colidx + e) : 0;
  int jb = (e + 1 < e1) ? __ldg(Q.colidx + e + 1) : 0;
  while (e + 1 < e1) {
    const int j0 = ja, j1 = jb;
    ja = (e + 2 < e1) ? __ldg(Q.colidx + e + 2) : 0;
    jb = (e + 3 < e1) ? __ldg(Q.colidx + e + 3) : 0;
    double m0[DH], m1[DH], x0[TILE], x1[TILE];
    load_m(e, m0);
    load_m(e + 1, m1);
    load_x(j0, x0);
    load_x(j1, x1);
    fmas(x0, m0);
    fmas(x1, m1);
    e += 2;
  }
  if (e < e1) {
    double m0[DH], x0[TILE];
    load_m(e, m0);
    load_x(ja, x0);
    fmas(x0, m0);
  }
}

// Stiefel tangent projection of one pose, column-distributed over the lane group:
//   w_Y <- w_Y - Y sym(Y^T w_Y), translation column untouched.
// Every lane of the warp must call this (shuffles); `valid` guards memory access only.
// sym[a] returns S[a][c] = sym(Y^T W)[a][c] for this lane's column c (< D).
template <int R, int D>
__device__ __forceinline__ void group_tangent(const double *Ytile, double (&w)[R], const LanePos &lp,
                                              bool valid, double (&sym)[D]) {
  double y[D][R];
#pragma unroll
  for (int a = 0; a < D; ++a) {
#pragma unroll
    for (int q = 0; q < R; ++q) y[a][q] = valid ? Ytile[a * R + q] : 0.0;
  }
  double mcol[D];  // M[a][c] = Y_a . w_c
#pragma unroll
  for (int a = 0; a < D; ++a) {
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < R; ++q) s = fma(y[a][q], w[q], s);
    mcol[a] = s;
  }
  double mt[D];  // M[c][a], fetched from lane (base + a)
#pragma unroll
  for (int a = 0; a < D; ++a) {
    mt[a] = 0.0;
#pragma unroll
    for (int k = 0; k < D; ++k) {
      const double v = __shfl_sync(0xffffffffu, mcol[k], lp.base_lane + a);
      if (k == lp.c) mt[a] = v;
    }
  }
#pragma unroll
  for (int a = 0; a < D; ++a) sym[a] = 0.5 * (mcol[a] + mt[a]);
  if (lp.c < D) {
#pragma unroll
    for (int a = 0; a < D; ++a) {
#pragma unroll
      for (int q = 0; q < R; ++q) w[q] = fma(-y[a][q], sym[a], w[q]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// phases
// ---------------------------------------------------------------------------------------------
// out = X * Q (+ G if G != nullptr).  PIPE: walk the block row two blocks per step (spmm_col2, same bits)
template <int R, int D, bool PIPE = false>
__device__ __forceinline__ void phase_qx(const Ctx &ctx, const BsrView &Q, const double *X,
                                         const double *G, double *out, int n) {
  using Gm = Geo<R, D>;
  const LanePos lp = lane_pos<D>(ctx.lane);
  for (int base = ctx.warp * Gm::GPW; base < n; base += ctx.nwarps * Gm::GPW) {
    const int i = base + lp.grp;
    if (lp.ok && i < n) {
      double acc[R];
      const size_t off = ((size_t)i * Gm::DH + lp.c) * R;
      if (G) load_col<R>(G + off, acc);
      else {
#pragma unroll
        for (int q = 0; q < R; ++q) acc[q] = 0.0;
      }
      if constexpr (PIPE) spmm_col2<R, D>(Q, X, i, lp.c, acc);
      else spmm_col<R, D>(Q, X, i, lp.c, acc);
      store_col<R>(out + off, acc);
    }
  }
}

// Q*X with the gathered pose tiles staged in shared memory (measurement variant 2 of the stand-alone kernel).
// Same lane-group mapping and the same sums in the same order as phase_qx; what changes is how an X tile reaches
// the d+1 lanes that need all of it.  In phase_qx every lane loads the whole tile (r(d+1)/2 128-bit loads whose 32
// lanes touch 8 different lines: the L1 data pipe is the limiter at scale).  Here the 8 tiles of a warp step are
// fetched ONCE, 16-byte pieces dealt over the 32 lanes (a tile = 10 consecutive lanes = whole lines), written to
// the warp's staging area and read back by the lane groups as shared-memory broadcasts: about half the L1
// wavefronts per block.  The fetch of step s+1 is issued before the FMAs of step s, the column index one step
// earlier still, so a step costs one memory latency.
template <int R, int D>
struct QxTiles {
  static constexpr int DH = D + 1, GPW = 32 / DH, TILE = R * DH;
  static constexpr int PB = (TILE % 2 == 0) ? 16 : 8;        // bytes per piece
  static constexpr int NP = TILE * 8 / PB;                   // pieces per tile
  static constexpr int NPW = (GPW * NP + 31) / 32;           // pieces per lane and step
  static constexpr int TP = TILE * 8 + 16;                   // tile pitch in the staging area (bank spread)
  static constexpr int WARP_BYTES = ((GPW * TP + 15) / 16) * 16;
};

template <int R, int D>
__device__ __forceinline__ void phase_qx_tiles(const Ctx &ctx, const BsrView &Q, const double *X, const double *G,
                                               double *out, int n, unsigned char *wsm) {
  using T = QxTiles<R, D>;
  constexpr int DH = T::DH, GPW = T::GPW, TILE = T::TILE;
  const LanePos lp = lane_pos<D>(ctx.lane);
  struct Piece { double a, b; };
  for (int base = ctx.warp * GPW; base < n; base += ctx.nwarps * GPW) {      // warp-uniform
    const int i = base + lp.grp;
    const bool active = lp.ok && i < n;
    const int e0 = active ? __ldg(Q.rowptr + i) : 0, e1 = active ? __ldg(Q.rowptr + i + 1) : 0;
    const int len = e1 - e0;
    int maxlen = len;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) maxlen = max(maxlen, (int)__shfl_xor_sync(0xffffffffu, maxlen, o));
    double acc[R];
    const size_t off = ((size_t)i * DH + lp.c) * R;
    if (active && G) load_col<R>(G + off, acc);
    else {
#pragma unroll
      for (int q = 0; q < R; ++q) acc[q] = 0.0;
    }
    Piece t[T::NPW];
    auto fetch = [&](int j) {      // every lane takes part (shuffles); j = this group's column index or -1
#pragma unroll
      for (int m = 0; m < T::NPW; ++m) {
        const int p = ctx.lane + 32 * m;
        const bool valid = p < GPW * T::NP;
        const int tile = valid ? p / T::NP : 0, piece = p - tile * T::NP;
        const int jt = (int)__shfl_sync(0xffffffffu, j, tile * DH);
        t[m].a = 0.0; t[m].b = 0.0;
        if (valid && jt >= 0) {
          const double *src = X + (size_t)jt * TILE + piece * (T::PB / 8);
          if constexpr (T::PB == 16) {
            const double2 v = *reinterpret_cast<const double2 *>(src);
            t[m].a = v.x; t[m].b = v.y;
          } else {
            t[m].a = src[0];
          }
        }
      }
    };
    int j = (len > 0) ? __ldg(Q.colidx + e0) : -1;
    fetch(j);
    for (int s = 0; s < maxlen; ++s) {
      const int jn = (s + 1 < len) ? __ldg(Q.colidx + e0 + s + 1) : -1;
      double mk[DH];
      if (s < len) load_q_row<DH>(Q.blocks + (size_t)(e0 + s) * (DH * DH) + lp.c * DH, mk);
#pragma unroll
      for (int m = 0; m < T::NPW; ++m) {
        const int p = ctx.lane + 32 * m;
        if (p < GPW * T::NP) {
          const int tile = p / T::NP, piece = p - tile * T::NP;
          double *dst = reinterpret_cast<double *>(wsm + tile * T::TP + piece * T::PB);
          dst[0] = t[m].a;
          if constexpr (T::PB == 16) dst[1] = t[m].b;
        }
      }
      __syncwarp();
      fetch(jn);                                   // next step's tiles in flight during the FMAs
      if (s < len) {
        const double *xs = reinterpret_cast<const double *>(wsm + lp.grp * T::TP);
        double x[TILE];
        if constexpr (TILE % 2 == 0) {
#pragma unroll
          for (int k = 0; k < TILE / 2; ++k) {
            const double2 v = reinterpret_cast<const double2 *>(xs)[k];
            x[2 * k] = v.x; x[2 * k + 1] = v.y;
          }
        } else {
#pragma unroll
          for (int k = 0; k < TILE; ++k) x[k] = xs[k];
        }
#pragma unroll
        for (int k = 0; k < DH; ++k) {
#pragma unroll
          for (int q = 0; q < R; ++q) acc[q] = fma(x[k * R + q], mk[k], acc[q]);
        }
      }
      __syncwarp();                                // the staging area may be overwritten
    }
    if (active) store_col<R>(out + off, acc);
  }
}

// TMA bulk prefetch into L2 (no destination in the SM, no registers held): `bytes` from a 16-byte aligned
// address, a multiple of 16
#ifndef DPGO_CPU_EMU
__device__ __forceinline__ void bulk_prefetch_l2(const void *p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
#endif

// Measurement variant of phase_qx for problems that stream from HBM (dpgo_set_qx_variant(h, 1)): the same
// product, plus a software prefetch -- one lane of every warp asks the L2 for the Q blocks, the column
// indices and the own X tiles of the block rows `dist` poses further on (the rows a CTA that becomes
// resident when this one retires will read), one bulk prefetch each.  The dependent chain rowptr -> colidx -> X
// otherwise leaves only the loads of one block row per lane group in flight.
template <int R, int D>
__device__ __forceinline__ void phase_qx_prefetch(const Ctx &ctx, const BsrView &Q, const double *X,
                                                  const double *G, double *out, int n, int dist) {
  using Gm = Geo<R, D>;
  constexpr int BLK = Gm::DH * Gm::DH;
  const LanePos lp = lane_pos<D>(ctx.lane);
  for (int base = ctx.warp * Gm::GPW; base < n; base += ctx.nwarps * Gm::GPW) {
    const int i = base + lp.grp;
    // one lane per warp: the GPW block rows `dist` poses ahead are contiguous in the block-CSR, so their
    // Q blocks, their column indices and their own X tiles are one contiguous span each
    const int ip0 = base + dist;
    if (ctx.lane == 0 && ip0 < n) {
      const int ip1 = min(ip0 + Gm::GPW, n);
      const int f0 = __ldg(Q.rowptr + ip0), f1 = __ldg(Q.rowptr + ip1);
      if constexpr ((BLK * 8) % 16 == 0) {
        if (f1 > f0) bulk_prefetch_l2(Q.blocks + (size_t)f0 * BLK, (uint32_t)(f1 - f0) * (BLK * 8));
      }
      if constexpr ((Gm::TILE * 8) % 16 == 0)
        bulk_prefetch_l2(X + (size_t)ip0 * Gm::TILE, (uint32_t)(ip1 - ip0) * (Gm::TILE * 8));
      // colidx entries are 4 bytes: prefetch the 16-byte aligned span that covers the rows
      const size_t a0 = ((size_t)f0 * 4) & ~(size_t)15, a1 = (((size_t)f1 * 4) + 15) & ~(size_t)15;
      if (a1 > a0) bulk_prefetch_l2(reinterpret_cast<const char *>(Q.colidx) + a0, (uint32_t)(a1 - a0));
    }
    if (lp.ok && i < n) {
      double acc[R];
      const size_t off = ((size_t)i * Gm::DH + lp.c) * R;
      if (G) load_col<R>(G + off, acc);
      else {
#pragma unroll
        for (int q = 0; q < R; ++q) acc[q] = 0.0;
      }
      spmm_col<R, D>(Q, X, i, lp.c, acc);
      store_col<R>(out + off, acc);
    }
  }
}

// Cost, Euclidean gradient, Riemannian gradient and the per-pose S = sym(Y^T EG_Y) in one pass:
//   EG = X Q + G; f = 0.5 <EG + G, X>; grad = Proj_X(EG); acc = {f, <grad,grad>}
// ref: QuadraticProblem::f / EucGrad / RieGrad, src/QuadraticProblem.cpp:29-47,71-83.
template <int R, int D>
__device__ __forceinline__ void phase_fgrad(const Ctx &ctx, const BsrView &Q, const double *X,
                                            const double *G, double *EG, double *grad, double *S,
                                            int n, double (&acc)[2]) {
  using Gm = Geo<R, D>;
  const LanePos lp = lane_pos<D>(ctx.lane);
  for (int base = ctx.warp * Gm::GPW; base < n; base += ctx.nwarps * Gm::GPW) {
    const int i = base + lp.grp;
    const bool valid = lp.ok && i < n;
    const size_t off = valid ? ((size_t)i * Gm::DH + lp.c) * R : 0;
    double eg[R], g[R], x[R];
#pragma unroll
    for (int q = 0; q < R; ++q) { eg[q] = 0.0; g[q] = 0.0; x[q] = 0.0; }
    if (valid) {
      load_col<R>(G + off, g);
      load_col<R>(X + off, x);
#pragma unroll
      for (int q = 0; q < R; ++q) eg[q] = g[q];
      spmm_col<R, D>(Q, X, i, lp.c, eg);
      store_col<R>(EG + off, eg);
      acc[0] += 0.5 * (dot_col<R>(eg, x) + dot_col<R>(g, x));
    }
    double sym[D];
    group_tangent<R, D>(X + (valid ? (size_t)i * Gm::TILE : 0), eg, lp, valid, sym);
    if (valid) {
      store_col<R>(grad + off, eg);
      acc[1] += dot_col<R>(eg, eg);
      if (lp.c < D) {
#pragma unroll
        for (int a = 0; a < D; ++a) S[(size_t)i * (D * D) + lp.c * D + a] = sym[a];
      }
    }
  }
}

// Riemannian Hessian-vector product at Y (S = sym(Y^T EG_Y) cached by phase_fgrad):
//   HV = Proj_Y( V Q - [V_Y S, 0] );  acc = {<V, HV>, <V, W>}  (W optional)
// ref: QuadraticProblem::EucHessianEta src/QuadraticProblem.cpp:49-54 + Stiefel EucHvToHv.
template <int R, int D>
__device__ __forceinline__ void phase_hess(const Ctx &ctx, const BsrView &Q, const double *Y,
                                           const double *S, const double *V, double *HV,
                                           const double *W, int n, double (&acc)[2]) {
  using Gm = Geo<R, D>;
  const LanePos lp = lane_pos<D>(ctx.lane);
  for (int base = ctx.warp * Gm::GPW; base < n; base += ctx.nwarps * Gm::GPW) {
    const int i = base + lp.grp;
    const bool valid = lp.ok && i < n;
    const size_t off = valid ? ((size_t)i * Gm::DH + lp.c) * R : 0;
    double out[R], v[R];
#pragma unroll
    for (int q = 0; q < R; ++q) { out[q] = 0.0; v[q] = 0.0; }
    if (valid) {
      spmm_col<R, D>(Q, V, i, lp.c, out);
      load_col<R>(V + off, v);
      if (lp.c < D) {
        const double *Vi = V + (size_t)i * Gm::TILE;
        const double *Si = S + (size_t)i * (D * D) + lp.c * D;
#pragma unroll
        for (int k = 0; k < D; ++k) {
          const double sk = Si[k];
#pragma unroll
          for (int q = 0; q < R; ++q) out[q] = fma(-Vi[k * R + q], sk, out[q]);
        }
      }
    }
    double sym[D];
    group_tangent<R, D>(Y + (valid ? (size_t)i * Gm::TILE : 0), out, lp, valid, sym);
    if (valid) {
      store_col<R>(HV + off, out);
      acc[0] += dot_col<R>(v, out);
      if (W) {
        double w[R];
        load_col<R>(W + off, w);
        acc[1] += dot_col<R>(v, w);
      }
    }
  }
}

// Stand-alone tangent projection out = Proj_X(V).
template <int R, int D>
__device__ __forceinline__ void phase_tangent(const Ctx &ctx, const double *X, const double *V,
                                              double *out, int n) {
  using Gm = Geo<R, D>;
  const LanePos lp = lane_pos<D>(ctx.lane);
  for (int base = ctx.warp * Gm::GPW; base < n; base += ctx.nwarps * Gm::GPW) {
    const int i = base + lp.grp;
    const bool valid = lp.ok && i < n;
    const size_t off = valid ? ((size_t)i * Gm::DH + lp.c) * R : 0;
    double w[R];
#pragma unroll
    for (int q = 0; q < R; ++q) w[q] = 0.0;
    if (valid) load_col<R>(V + off, w);
    double sym[D];
    group_tangent<R, D>(X + (valid ? (size_t)i * Gm::TILE : 0), w, lp, valid, sym);
    if (valid) store_col<R>(out + off, w);
  }
}

// ---- TMA (bulk async copy) + mbarrier helpers -------------------------------------------------
#ifndef DPGO_CPU_EMU
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
// 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar) : "memory");
}
#endif  // DPGO_CPU_EMU

constexpr int kStageK = 32;                                  // inner indices per pipeline stage
constexpr int kStages = 4;                                   // stages in flight per CTA
constexpr int kStageDoubles = kStageK * kGemvCols;           // 2048 doubles = 16 KB
constexpr int kGemvMaxR = 6;
// symmetric (half-storage) variant: 128 x 128 blocks, 16 inner indices x 128 columns per stage
constexpr int kSymB = 128;
constexpr int kSymStageK = 16;
constexpr int kSymStagesPerTile = kSymB / kSymStageK;        // 8 stages = one 128 KB block
constexpr int kSymS = 2;                                     // blocks per side of a work item
static_assert(kSymStageK * kSymB == kStageDoubles, "both variants stream 16 KB stages");
// scratch shared by the two variants: cross-warp reduction buffers (+ transposed accumulators)
constexpr int kGemvScratch = (4 * kGemvMaxR * kSymB + kSymS * kSymB * kGemvMaxR) * 8;
static_assert(kGemvScratch >= kWarpsPerBlock * kGemvMaxR * kGemvCols * 8, "scratch too small");
constexpr int kGemvDynSmem =
    kStages * kStageDoubles * 8 + kStages * 8 + kStages * kStageK * kGemvMaxR * 8 + kGemvScratch;

struct GemvPipe {
  double *stage;      // [STAGES][kStageK][kGemvCols]
  double *svec;       // [STAGES][kStageK * R] slice of the input array that goes with a stage
  double *scratch;    // kGemvScratch bytes
  uint32_t bar;       // shared address of the first mbarrier
  uint32_t slot;      // ring slot of the next chunk to consume (persists across phases)
  uint32_t parity;    // mbarrier phase parity of that slot
};

// one-time set-up of the pipeline barriers (all threads of the CTA must call); STAGES = ring depth
// (kStages for the HBM-bound dense variants, kDdStages for the L2-resident two-level variant)
template <int STAGES = kStages, int VCH = STAGES>
__device__ __forceinline__ GemvPipe gemv_pipe_init(unsigned char *dsm) {
  GemvPipe pp;
  pp.stage = reinterpret_cast<double *>(dsm);
  pp.bar = smem_u32(dsm + (size_t)STAGES * kStageDoubles * 8);
  pp.svec = reinterpret_cast<double *>(dsm + (size_t)STAGES * kStageDoubles * 8 + STAGES * 8);
  pp.scratch = pp.svec + VCH * kStageK * kGemvMaxR;   // svec holds VCH chunks' worth of the input
  pp.slot = 0;
  pp.parity = 0;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) mbar_init(pp.bar + 8 * s, 1);
    mbar_fence_init();
  }
  __syncthreads();
  return pp;
}

// position of a pipeline chunk: tile (column block cb, inner split s) and chunk c within it.
// Advanced incrementally -- no integer division in the per-chunk path.
struct GemvCursor {
  int t, c, cb, s;
};

// Dense preconditioner, part 1: partial products of the r x N array `vec` with the dense
// symmetric inverse Pinv.  Tile = kGemvCols output columns x KT inner indices.  Pinv is stored
// "stage-major": the kStageK x kGemvCols sub-block (inner chunk kc, column block cb) is one
// contiguous 16 KB run at ((kc * ncb + cb) * kStageDoubles), element (kk, jj) at kk*64 + jj, zero
// padded to ld rows x ldk columns -- so that one TMA bulk copy (cp.async.bulk, SASS UBLKCP)
// moves a whole pipeline stage HBM -> shared memory; kStages stages are in flight per CTA, so
// the bytes in flight do not cost registers; the 8 warps split the inner indices of a stage,
// lanes own 2 adjacent output columns, partial sums meet in shared memory at the end of a tile.
//   zpart[s][:, j] = sum_{k in split s} vec[:, k] * Pinv[j, k]
// ref: QuadraticProblem::PreConditioner, src/QuadraticProblem.cpp:56-69 (the solve).
template <int R>
__device__ __forceinline__ void phase_precon_gemv(GemvPipe &pp, const double *Pinv, int ld,
                                                  const double *vec, double *zpart, size_t zstride,
                                                  int KT, int nsplit) {
  double(*sacc)[R][kGemvCols] = reinterpret_cast<double(*)[R][kGemvCols]>(pp.scratch);  // [8][R][64]
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int ncb = ld / kGemvCols;
  const int ntiles = ncb * nsplit;
  const int cpt = KT / kStageK;  // chunks per tile
  const int my_tiles = ((int)blockIdx.x < ntiles) ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int G = my_tiles * cpt;
  if (G == 0) return;
  const uint32_t stage0 = smem_u32(pp.stage);
  const int cps = KT / kStageK;  // chunks per split (== cpt)

  auto cursor_next = [&](GemvCursor &cu) {
    if (++cu.c == cpt) {
      cu.c = 0;
      cu.t += gridDim.x;
      cu.cb = cu.t % ncb;
      cu.s = cu.t / ncb;
    }
  };
  // producer side: thread 0 streams the 16 KB stage, R*32 threads copy the matching slice of vec
  auto produce = [&](const GemvCursor &cu, uint32_t slot) {
    const int kc = cu.s * cps + cu.c;
    if (threadIdx.x == 0) {
      const uint32_t bar = pp.bar + 8 * slot;
      mbar_expect_tx(bar, kStageDoubles * 8);
      bulk_g2s(stage0 + slot * (kStageDoubles * 8), Pinv + ((size_t)kc * ncb + cu.cb) * kStageDoubles,
               kStageDoubles * 8, bar);
    }
    if (threadIdx.x < kStageK * R)
      pp.svec[slot * (kStageK * R) + threadIdx.x] = vec[(size_t)kc * (kStageK * R) + threadIdx.x];
  };

  GemvCursor ci, cc;  // issue / consume cursors
  ci.t = cc.t = blockIdx.x;
  ci.c = cc.c = 0;
  ci.cb = cc.cb = ci.t % ncb;
  ci.s = cc.s = ci.t / ncb;
  uint32_t slot_i = pp.slot;
  int issued = 0;
  for (; issued < kStages && issued < G; ++issued) {
    produce(ci, slot_i);
    cursor_next(ci);
    slot_i = (slot_i + 1 == kStages) ? 0 : slot_i + 1;
  }
  __syncthreads();
  double a0[R], a1[R];
#pragma unroll
  for (int q = 0; q < R; ++q) { a0[q] = 0.0; a1[q] = 0.0; }
  for (int g = 0; g < G; ++g) {
    const uint32_t slot = pp.slot;
    mbar_wait(pp.bar + 8 * slot, pp.parity);
    const double *st = pp.stage + (size_t)slot * kStageDoubles;
    const double *sv = pp.svec + slot * (kStageK * R);
#pragma unroll
    for (int u = 0; u < kStageK / kWarpsPerBlock; ++u) {
      const int kk = w + u * kWarpsPerBlock;
      const double2 pv = *reinterpret_cast<const double2 *>(st + kk * kGemvCols + 2 * lane);
#pragma unroll
      for (int q = 0; q < R; ++q) {
        const double x = sv[kk * R + q];
        a0[q] = fma(pv.x, x, a0[q]);
        a1[q] = fma(pv.y, x, a1[q]);
      }
    }
    const bool tile_end = (cc.c == cpt - 1);
    if (tile_end) {  // tile finished: combine the 8 warps, emit the partial result
#pragma unroll
      for (int q = 0; q < R; ++q) {
        sacc[w][q][2 * lane] = a0[q];
        sacc[w][q][2 * lane + 1] = a1[q];
        a0[q] = 0.0;
        a1[q] = 0.0;
      }
    }
    __syncthreads();  // every warp is done with this stage (and sacc is complete)
    if (issued < G) {  // refill the slot just freed (its readers meet another barrier first)
      produce(ci, slot);
      cursor_next(ci);
      ++issued;
    }
    if (tile_end) {
      for (int o = threadIdx.x; o < R * kGemvCols; o += kBlock) {
        const int q = o / kGemvCols, jj = o % kGemvCols;
        double x = 0.0;
#pragma unroll
        for (int ww = 0; ww < kWarpsPerBlock; ++ww) x += sacc[ww][q][jj];
        zpart[(size_t)cc.s * zstride + (size_t)(cc.cb * kGemvCols + jj) * R + q] = x;
      }
      __syncthreads();
    }
    cursor_next(cc);
    if (pp.slot + 1 == kStages) { pp.slot = 0; pp.parity ^= 1u; } else { pp.slot += 1; }
  }
}

// ---- two-level (domain decomposition) exact preconditioner ---------------------------------------
// Nested dissection of the pose graph gives interior domains D_1..D_K (no edges between them) and a
// separator S.  With A = Q + 0.1 I permuted to [D_1 .. D_K | S]:
//     A^-1 r :   y_I = A_II^-1 r_I ;  t = r_S - A_SI y_I ;  z_S = Sigma^-1 t ;  z_I = y_I - A_II^-1 (A_IS z_S)
// where A_II^-1 = blockdiag(A_k^-1) and Sigma = A_SS - A_SI A_II^-1 A_IS are stored as dense inverses
// (exact block elimination: same operator as the full dense inverse up to rounding, ~14x fewer
// bytes on sphere2500 and L2-resident).  All work arrays live in the permuted, 64-padded scalar
// column space; a dense block is streamed as "strips" (64 output columns x a range of 32-index
// chunks, stage-major like the full variant).
// The two-level kernels run ONE CTA per SM with a deep ring: the dense blocks are L2 resident, so a
// strip phase is bound by the latency of a bulk copy, not by bandwidth -- a whole strip (<= 12
// stages) is put in flight at once.
constexpr int kDdStages = 10;
constexpr int kDdVecChunks = kDdStages;   // chunks of the input array staged with a wave
constexpr int kDdScratch = kWarpsPerBlock * kGemvMaxR * kGemvCols * 8;
constexpr int kDdDynSmem =
    kDdStages * kStageDoubles * 8 + kDdStages * 8 + kDdVecChunks * kStageK * kGemvMaxR * 8 + kDdScratch;
static_assert(kDdDynSmem <= 227 * 1024, "two-level pipeline does not fit in shared memory");
struct DdStrip {
  int cb, kc0, nchunks, slot;   // output column block, first inner chunk, #chunks, partial slot
  long long data_off;           // first stage of the strip in the matrix buffer (units of stages)
};

// strips of one phase, grouped by "virtual CTA" (balanced on the host, longest strip first):
// virtual CTA v owns strips[cta[v] .. cta[v+1]) = chunks[v] pipeline stages
struct DdStripSet {
  const double *M;
  const DdStrip *strips;
  const int *cta;      // [V + 1]
  const int *chunks;   // [V]
};


struct DdView {
  DdStripSet P1;                // blockdiag(A_k^-1)  (nsplit1 partial slots)
  DdStripSet P3;                // Sigma^-1           (nsplit3 partial slots)
  int V;                        // virtual CTAs (= SMs of the device the plan was built for)
  int nsplit1, nsplit3;
  BsrView A_SI;                 // rows: separator poses; colidx: permuted scalar column of the interior pose
  BsrView A_BS;                 // rows: interior poses WITH a separator neighbour; colidx: permuted column of it
  int nS, nB;
  const int *pcol;              // [n]  permuted scalar column of pose i (tile start)
  const int *srow;              // [nS] original pose id of separator row s (its permuted column is sep_col0 + s*(d+1))
  const int *bcol;              // [nB] permuted scalar column of boundary row b
  const int *icol;              // [pcols] original scalar column of a permuted column (-1: padding)
  int sep_col0, pcols;          // first separator column; padded column count
  double *rp;                   // the tCG residual in permuted order (fused solver, phase_step_perm)
  double *y, *t, *zs, *u, *w;   // permuted work arrays, R x pcols (y, w: nsplit1 partial slots; zs:
                                // nsplit3 partial slots; u is zero outside the boundary rows)
  int prefetch;                 // issue the first matrix stages of the next strip phase before the preceding barrier
};

// Per-CTA cache (shared memory, filled once per kernel) of the head of this CTA's strip list of one
// phase: the schedule is static, and reading it from global memory would put two dependent
// L2 round trips in front of every strip phase.
constexpr int kDdPlanMax = 4;
struct StripPlanStore {
  DdStrip strip[kDdPlanMax];
  int G, n, si0, end0;
};
struct StripPlan {
  const DdStrip *cached;   // this CTA's first `ncached` strips in processing order
  int ncached, G, si0, end0;
};

// all threads of the CTA call
__device__ __forceinline__ void strip_plan_fill(StripPlanStore *st, const DdStripSet &S, int V) {
  if (threadIdx.x == 0) {
    int G = 0;
    for (int v = blockIdx.x; v < V; v += gridDim.x) G += S.chunks[v];
    int v = blockIdx.x, si = 0, end = 0;
    if (v < V) {
      si = S.cta[v];
      end = S.cta[v + 1];
    }
    st->G = G;
    st->si0 = si;
    st->end0 = end;
    int n = 0;
    while (n < kDdPlanMax) {
      while (v < V && si >= end) {
        v += gridDim.x;
        if (v < V) {
          si = S.cta[v];
          end = S.cta[v + 1];
        }
      }
      if (v >= V) break;
      st->strip[n++] = S.strips[si++];
    }
    st->n = n;
  }
  __syncthreads();
}
__device__ __forceinline__ StripPlan strip_plan_load(const StripPlanStore *st) {
  StripPlan p;
  p.cached = st->strip;
  p.ncached = st->n;
  p.G = st->G;
  p.si0 = st->si0;
  p.end0 = st->end0;
  return p;
}

struct StripCursor {
  int v, si, end, c, j;   // virtual CTA, strip index (global table), end of v's strips, chunk, sequence number
  DdStrip d;
};

// position the cursor on the strip at (v, si) or on the first strip of the next non-empty virtual
// CTA this CTA owns
__device__ __forceinline__ void strip_cursor_seek(StripCursor &cu, const DdStripSet &S, int V, const StripPlan &pl) {
  while (cu.v < V && cu.si >= cu.end) {
    cu.v += gridDim.x;
    if (cu.v < V) {
      cu.si = __ldg(S.cta + cu.v);
      cu.end = __ldg(S.cta + cu.v + 1);
    }
  }
  cu.c = 0;
  if (cu.v < V) cu.d = (cu.j < pl.ncached) ? pl.cached[cu.j] : S.strips[cu.si];
}
__device__ __forceinline__ void strip_cursor_init(StripCursor &cu, const DdStripSet &S, int V, const StripPlan &pl) {
  cu.v = blockIdx.x;
  cu.si = pl.si0;
  cu.end = pl.end0;
  cu.j = 0;
  strip_cursor_seek(cu, S, V, pl);
}
__device__ __forceinline__ void strip_cursor_next_strip(StripCursor &cu, const DdStripSet &S, int V,
                                                        const StripPlan &pl) {
  ++cu.si;
  ++cu.j;
  strip_cursor_seek(cu, S, V, pl);
}

// one warp (all its lanes call): put chunks [c0, c0 + nw) of strip d in flight into stage slots 0 .. nw-1, lane i
// issues the copy of stage i (a single thread issuing a 10-stage wave is ~0.4 us on the critical path of a phase)
__device__ __forceinline__ void strip_issue_wave(const GemvPipe &pp, const DdStripSet &S, const DdStrip &d, int c0,
                                                 int nw) {
  const int i = threadIdx.x & 31;
  if (i < nw) {
    const uint32_t stage0 = smem_u32(pp.stage);
    const uint32_t bar = pp.bar + 8 * i;
    mbar_expect_tx(bar, kStageDoubles * 8);
    bulk_g2s(stage0 + i * (kStageDoubles * 8), S.M + (size_t)(d.data_off + c0 + i) * kStageDoubles,
             kStageDoubles * 8, bar);
  }
}

// Issue the TMA copies of the first wave of this CTA's first strip of a strip phase.  The matrix
// does not depend on the phases before it, so this is called BEFORE the grid barrier that makes the
// input array visible; phase_strip_gemv(..., prefetched = true) then skips that issue.
// Precondition: every stage has been consumed (true between strip phases).
// `issuer` = the warp of the CTA that issues (0, or a warp without work in the phase the call sits in).
template <int STAGES>
__device__ __forceinline__ void strip_prefetch(const GemvPipe &pp, const DdStripSet &S, int V,
                                               const StripPlanStore *st, int issuer = 0) {
  static_assert(STAGES <= 32, "one lane per stage");
  if ((int)(threadIdx.x >> 5) != issuer) return;
  const StripPlan pl = strip_plan_load(st);
  StripCursor cu;
  strip_cursor_init(cu, S, V, pl);
  if (cu.v < V) strip_issue_wave(pp, S, cu.d, 0, min(STAGES, cu.d.nchunks));
}

// Ask the L2 for every stage of this CTA's strips of one phase (one warp calls, lane l takes the chunks l, l + 32, ..
// of each strip): the dense blocks are read from DRAM only by the first application after something else used the
// L2 (another agent's solve on the same GPU, the flush of the bench); issued at the start of the solver kernel, the
// fetch runs under the cost / gradient phases instead of in front of the first strip pass.
__device__ __forceinline__ void strip_l2_prefetch(const DdStripSet &S, int V) {
  const int lane = threadIdx.x & 31;
  for (int v = blockIdx.x; v < V; v += gridDim.x) {
    const int s0 = __ldg(S.cta + v), s1 = __ldg(S.cta + v + 1);
    for (int si = s0; si < s1; ++si) {
      const DdStrip d = S.strips[si];
      for (int c = lane; c < d.nchunks; c += 32)
        bulk_prefetch_l2(S.M + (size_t)(d.data_off + c) * kStageDoubles, kStageDoubles * 8);
    }
  }
}

// Wait for a prefetched first wave that no strip pass will consume (end of the kernel) and flip the stage parities.
// All threads of the CTA call.
template <int STAGES>
__device__ __forceinline__ void strip_drain(GemvPipe &pp, const DdStripSet &S, int V, const StripPlanStore *st) {
  const StripPlan pl = strip_plan_load(st);
  StripCursor cu;
  strip_cursor_init(cu, S, V, pl);
  if (cu.v < V) {
    const int nw = min(STAGES, cu.d.nchunks);
    for (int ch = 0; ch < nw; ++ch) mbar_wait(pp.bar + 8 * ch, (pp.parity >> ch) & 1u);
    pp.parity ^= (1u << nw) - 1u;
  }
  __syncthreads();
}

// out[slot][:, 64 cb + jj] = sum over the strip's chunks of vec[:, k] * M(jj, k)
// A strip is processed in waves of <= STAGES chunks: thread 0 puts the whole wave in flight (one
// TMA bulk copy per 16 KB stage, each with its own mbarrier), all threads stage the matching slice
// of `vec` (one round of global-load latency per wave), then the 8 warps split the wave into
// (chunk, 8-row) units and each waits only for the stages it reads -- no CTA-wide barrier per
// stage.  pp.parity is a bit mask here: bit i = phase parity of stage slot i.
template <int R, int STAGES>
__device__ __forceinline__ void phase_strip_gemv(GemvPipe &pp, const DdStripSet &S, int V, const StripPlanStore *st,
                                                 const double *vec, const int *icol, double *out,
                                                 size_t outstride, bool prefetched = false) {
  // icol != nullptr: `vec` is in the ORIGINAL column order and is gathered through icol while it
  // is staged (saves a separate permutation pass + grid barrier).
  static_assert(STAGES <= 32 && kStageK == 32, "wave bookkeeping");
  double(*sacc)[R][kGemvCols] = reinterpret_cast<double(*)[R][kGemvCols]>(pp.scratch);  // [8][R][64]
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const StripPlan pl = strip_plan_load(st);
  if (pl.G == 0) return;
  StripCursor cu;
  strip_cursor_init(cu, S, V, pl);
  bool in_flight = prefetched;   // first wave of the current strip already issued
  while (cu.v < V) {
    const DdStrip d = cu.d;
    double a0[R], a1[R];
#pragma unroll
    for (int q = 0; q < R; ++q) { a0[q] = 0.0; a1[q] = 0.0; }
    for (int c0 = 0; c0 < d.nchunks; c0 += STAGES) {
      const int nw = min(STAGES, d.nchunks - c0);
      if (!in_flight && threadIdx.x < 32) strip_issue_wave(pp, S, d, c0, nw);
      in_flight = false;
      {  // the slice of vec that goes with the wave
        const int k0 = (d.kc0 + c0) * kStageK;
        const int cnt = nw * kStageK * R;
        for (int o = threadIdx.x; o < cnt; o += kBlock) {
          double val;
          if (icol) {
            const int k = o / R, q = o - k * R;
            const int oc = __ldg(icol + k0 + k);
            val = (oc >= 0) ? vec[(size_t)oc * R + q] : 0.0;
          } else {
            val = vec[(size_t)k0 * R + o];
          }
          pp.svec[o] = val;
        }
      }
      __syncthreads();
      for (int u = w; u < 4 * nw; u += kWarpsPerBlock) {
        const int ch = u >> 2, r0 = (u & 3) * 8;
        mbar_wait(pp.bar + 8 * ch, (pp.parity >> ch) & 1u);
        const double *stp = pp.stage + (size_t)ch * kStageDoubles + r0 * kGemvCols + 2 * lane;
        const double *sv = pp.svec + (ch * kStageK + r0) * R;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const double2 pv = *reinterpret_cast<const double2 *>(stp + kk * kGemvCols);
#pragma unroll
          for (int q = 0; q < R; ++q) {
            const double x = sv[kk * R + q];
            a0[q] = fma(pv.x, x, a0[q]);
            a1[q] = fma(pv.y, x, a1[q]);
          }
        }
      }
      pp.parity ^= (1u << nw) - 1u;
      __syncthreads();   // the wave is consumed: stage slots and svec may be reused
    }
    // put the first wave of the next strip in flight, then finish this one
    strip_cursor_next_strip(cu, S, V, pl);
    if (cu.v < V) {
      if (threadIdx.x < 32) strip_issue_wave(pp, S, cu.d, 0, min(STAGES, cu.d.nchunks));
      in_flight = true;
    }
#pragma unroll
    for (int q = 0; q < R; ++q) {
      sacc[w][q][2 * lane] = a0[q];
      sacc[w][q][2 * lane + 1] = a1[q];
    }
    __syncthreads();
    for (int o = threadIdx.x; o < R * kGemvCols; o += kBlock) {
      const int q = o / kGemvCols, jj = o % kGemvCols;
      double x = 0.0;
#pragma unroll
      for (int ww = 0; ww < kWarpsPerBlock; ++ww) x += sacc[ww][q][jj];
      out[(size_t)d.slot * outstride + ((size_t)d.cb * kGemvCols + jj) * R + q] = x;
    }
    __syncthreads();
  }
}

// One warp per block row of a sparse coupling matrix: lane = (el, c) handles the blocks
// e0 + el, e0 + el + 8, ... of the row and column c of the pose tile, so the dependent loads
// (colidx -> tile of X) of up to 8 blocks are in flight together; the 8 partial sums are combined
// by a fixed shuffle tree.  Returns (lanes c < d+1) column c of sum_e X[col_e] * B_e^T, where the
// tile of block e starts at scalar column colidx[e] of the permuted array X = sum of `nsum` arrays
// `xstride` apart.
template <int R, int D>
__device__ __forceinline__ void dd_row_product(const BsrView &B, const double *X, int nsum, size_t xstride,
                                               int row, int lane, double (&acc)[R]) {
  constexpr int DH = D + 1, TILE = R * DH, EL = 8;
  static_assert(EL * DH <= 32, "lane layout");
  const int el = lane / DH, c = lane - el * DH;
  const bool active = lane < EL * DH;
  const int e0 = __ldg(B.rowptr + row), e1 = __ldg(B.rowptr + row + 1);
#pragma unroll
  for (int q = 0; q < R; ++q) acc[q] = 0.0;
  for (int eb = e0; eb < e1; eb += EL) {
    const int e = eb + el;
    if (active && e < e1) {
      const double *xj = X + (size_t)__ldg(B.colidx + e) * R;
      const double *m = B.blocks + (size_t)e * (DH * DH) + c * DH;
      double x[TILE];
#pragma unroll
      for (int k = 0; k < TILE; ++k) x[k] = xj[k];
      // partial slots: 3 at a time, so that their loads are in flight together (fixed order)
      for (int s = 1; s < nsum; s += 3) {
        const double *xa = xj + (size_t)s * xstride;
        const bool hb = s + 1 < nsum, hc = s + 2 < nsum;
        const double *xb = hb ? xa + xstride : xa;
        const double *xc = hc ? xb + xstride : xa;
        double ta[TILE], tb[TILE], tc[TILE];
#pragma unroll
        for (int k = 0; k < TILE; ++k) { ta[k] = xa[k]; tb[k] = xb[k]; tc[k] = xc[k]; }
#pragma unroll
        for (int k = 0; k < TILE; ++k) {
          x[k] += ta[k];
          if (hb) x[k] += tb[k];
          if (hc) x[k] += tc[k];
        }
      }
#pragma unroll
      for (int k = 0; k < DH; ++k) {
        const double mk = __ldg(m + k);
#pragma unroll
        for (int q = 0; q < R; ++q) acc[q] = fma(x[k * R + q], mk, acc[q]);
      }
    }
  }
#pragma unroll
  for (int delta = 4 * DH; delta >= DH; delta >>= 1) {
#pragma unroll
    for (int q = 0; q < R; ++q) acc[q] += __shfl_down_sync(0xffffffffu, acc[q], delta);
  }
}

// P2:  t_S = r_S - A_SI y_I   (r in the original column order, y = sum of its partial slots)
template <int R, int D>
__device__ __forceinline__ void phase_dd_sep_rhs(const Ctx &ctx, const DdView &dd, const double *rvec) {
  constexpr int DH = D + 1;
  const size_t zstride = (size_t)dd.pcols * R;
  // rows are dealt round-robin over the CTAs (row -> CTA row % grid), so that a short phase uses
  // every SM instead of filling the first CTAs' warps
  for (int s = blockIdx.x + gridDim.x * (threadIdx.x >> 5); s < dd.nS; s += gridDim.x * kWarpsPerBlock) {
    double acc[R], rr[R];
    const int cl = (ctx.lane < DH) ? ctx.lane : 0;
    const size_t ooff = ((size_t)__ldg(dd.srow + s) * DH + cl) * R;   // issued before the product: independent
#pragma unroll
    for (int q = 0; q < R; ++q) rr[q] = rvec[ooff + q];
    dd_row_product<R, D>(dd.A_SI, dd.y, dd.nsplit1, zstride, s, ctx.lane, acc);
    if (ctx.lane < DH) {
      const size_t off = ((size_t)dd.sep_col0 + (size_t)s * DH + ctx.lane) * R;
#pragma unroll
      for (int q = 0; q < R; ++q) dd.t[off + q] = rr[q] - acc[q];
    }
  }
}

// P4:  u_B = A_BS z_S  with z_S = sum of the nsplit3 partial results of the Schur GEMV (u stays
// zero at interior poses without a separator neighbour)
template <int R, int D>
__device__ __forceinline__ void phase_dd_back_rhs(const Ctx &ctx, const DdView &dd) {
  constexpr int DH = D + 1;
  const size_t zstride = (size_t)dd.pcols * R;
  for (int b = blockIdx.x + gridDim.x * (threadIdx.x >> 5); b < dd.nB; b += gridDim.x * kWarpsPerBlock) {
    double acc[R];
    const int col0 = __ldg(dd.bcol + b);
    dd_row_product<R, D>(dd.A_BS, dd.zs, dd.nsplit3, zstride, b, ctx.lane, acc);
    if (ctx.lane < DH) {
      const size_t off = ((size_t)col0 + ctx.lane) * R;
#pragma unroll
      for (int q = 0; q < R; ++q) dd.u[off + q] = acc[q];
    }
  }
}

// finish:  z = Proj_Y( unpermute( interior: y - w ; separator: sum of zs partials ) ); acc = {<z, rvec>}
template <int R, int D>
__device__ __forceinline__ void phase_dd_finish(const Ctx &ctx, const DdView &dd, const double *Y,
                                                const double *rvec, double *z, double *neg_out, int n,
                                                double (&acc)[1]) {
  using Gm = Geo<R, D>;
  const LanePos lp = lane_pos<D>(ctx.lane);
  const size_t zstride = (size_t)dd.pcols * R;
  for (int base = ctx.warp * Gm::GPW; base < n; base += ctx.nwarps * Gm::GPW) {
    const int i = base + lp.grp;
    const bool valid = lp.ok && i < n;
    const size_t off = valid ? ((size_t)i * Gm::DH + lp.c) * R : 0;
    double wv[R];
#pragma unroll
    for (int q = 0; q < R; ++q) wv[q] = 0.0;
    if (valid) {
      const int pc0 = dd.pcol[i];
      const size_t poff = ((size_t)pc0 + lp.c) * R;
      if (pc0 >= dd.sep_col0) {
        // partial slots of the Schur product, 4 at a time (loads in flight together, fixed order)
        for (int s = 0; s < dd.nsplit3; s += 4) {
          double tq[4][R];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const size_t so = (size_t)min(s + j, dd.nsplit3 - 1) * zstride + poff;
#pragma unroll
            for (int q = 0; q < R; ++q) tq[j][q] = dd.zs[so + q];
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (s + j < dd.nsplit3) {
#pragma unroll
              for (int q = 0; q < R; ++q) wv[q] += tq[j][q];
            }
          }
        }
      } else {
        for (int s = 0; s < dd.nsplit1; ++s) {
#pragma unroll
          for (int q = 0; q < R; ++q)
            wv[q] += dd.y[(size_t)s * zstride + poff + q] - dd.w[(size_t)s * zstride + poff + q];
        }
      }
    }
    double sym[D];
    group_tangent<R, D>(Y + (valid ? (size_t)i * Gm::TILE : 0), wv, lp, valid, sym);
    if (valid) {
      store_col<R>(z + off, wv);
      double rr[R];
      load_col<R>(rvec + off, rr);
      acc[0] += dot_col<R>(wv, rr);
      if (neg_out) {
#pragma unroll
        for (int q = 0; q < R; ++q) neg_out[off + q] = -wv[q];
      }
    }
  }
}

// Dense preconditioner, part 2: z = Proj_Y( sum_s zpart[s] ); acc = {<z, rvec>};
// optionally also writes delta = -z (first tCG direction).
template <int R, int D>
__device__ __forceinline__ void phase_precon_finish(const Ctx &ctx, const double *zpart,
                                                    size_t zstride, int nsplit, const double *Y, const double *rvec,
                                                    double *z, double *neg_out, int n,
                                                    double (&acc)[1]) {
  using Gm = Geo<R, D>;
  const LanePos lp = lane_pos<D>(ctx.lane);
  for (int base = ctx.warp * Gm::GPW; base < n; base += ctx.nwarps * Gm::GPW) {
    const int i = base + lp.grp;
    const bool valid = lp.ok && i < n;
    const size_t off = valid ? ((size_t)i * Gm::DH + lp.c) * R : 0;
    double w[R];
#pragma unroll
    for (int q = 0; q < R; ++q) w[q] = 0.0;
    if (valid) {
      for (int s = 0; s < nsplit; ++s) {
        const double *zp = zpart + (size_t)s * zstride + off;
#pragma unroll
        for (int q = 0; q < R; ++q) w[q] += zp[q];
      }
    }
    double sym[D];
    group_tangent<R, D>(Y + (valid ? (size_t)i * Gm::TILE : 0), w, lp, valid, sym);
    if (valid) {
      store_col<R>(z + off, w);
      double rr[R];
      load_col<R>(rvec + off, rr);
      acc[0] += dot_col<R>(w, rr);
      if (neg_out) {
#pragma unroll
        for (int q = 0; q < R; ++q) neg_out[off + q] = -w[q];
      }
    }
  }
}

// QF retraction, one thread per pose: Y+ = qf(Y + eta_Y) (Gram-Schmidt, positive diagonal),
// p+ = p + eta_p.   ref: ProductManifold::Retraction (src/QuadraticOptimizer.cpp:134).
// SCALED: the step is s * Eta (the gradient scale of QuadraticOptimizer::gradientDescent,
// src/QuadraticOptimizer.cpp:133-134, applied inside the retraction: round(s * e) then the add, i.e. the
// same bits as scaling into a separate array first).
// one pose: val(k) = entry k of the tile Y + eta (column-major r x (d+1)); o = the retracted tile
template <int R, int D, class F>
__device__ __forceinline__ void retract_pose(F val, double *o) {
  {
    double a[D][R];
#pragma unroll
    for (int k = 0; k < D; ++k) {
#pragma unroll
      for (int q = 0; q < R; ++q) a[k][q] = val(k * R + q);
    }
#pragma unroll
    for (int k = 0; k < D; ++k) {
#pragma unroll
      for (int p = 0; p < k; ++p) {
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < R; ++q) s = fma(a[p][q], a[k][q], s);
#pragma unroll
        for (int q = 0; q < R; ++q) a[k][q] = fma(-s, a[p][q], a[k][q]);
      }
      // second orthogonalization pass (keeps Y^T Y = I to rounding even for large steps)
#pragma unroll
      for (int p = 0; p < k; ++p) {
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < R; ++q) s = fma(a[p][q], a[k][q], s);
#pragma unroll
        for (int q = 0; q < R; ++q) a[k][q] = fma(-s, a[p][q], a[k][q]);
      }
      double nn = 0.0;
#pragma unroll
      for (int q = 0; q < R; ++q) nn = fma(a[k][q], a[k][q], nn);
      const double inv = 1.0 / sqrt(nn);
#pragma unroll
      for (int q = 0; q < R; ++q) a[k][q] *= inv;
    }
#pragma unroll
    for (int k = 0; k < D; ++k) {
#pragma unroll
      for (int q = 0; q < R; ++q) o[k * R + q] = a[k][q];
    }
#pragma unroll
    for (int q = 0; q < R; ++q) o[D * R + q] = val(D * R + q);
  }
}
template <int R, int D, bool SCALED>
__device__ __forceinline__ void phase_retract_impl(const Ctx &ctx, const double *X, const double *Eta,
                                                   double *Xout, int n, double s) {
  constexpr int TILE = R * (D + 1);
  for (int i = ctx.tid; i < n; i += ctx.nthreads) {
    const double *x = X + (size_t)i * TILE;
    const double *e = Eta + (size_t)i * TILE;
    retract_pose<R, D>([&](int k) { return SCALED ? __dadd_rn(x[k], __dmul_rn(s, e[k])) : x[k] + e[k]; },
                       Xout + (size_t)i * TILE);
  }
}
template <int R, int D>
__device__ __forceinline__ void phase_retract(const Ctx &ctx, const double *X, const double *Eta,
                                              double *Xout, int n) {
  phase_retract_impl<R, D, false>(ctx, X, Eta, Xout, n, 1.0);
}

// ---- per-pose kernels at scale: tiles staged through shared memory ------------------------------------------------
// One thread per pose keeps the arithmetic of a pose in one thread's registers, but its global accesses are 32
// tiles apart per warp instruction (measured on 262 144 / 1 000 000 poses, profiles/r02_pose_op_scale.jsonl:
// retraction 0.38 / 0.44, polar projection 0.28 / 0.36, rounding 0.27 / 0.33 of the HBM peak).  The staged forms
// below move the 32 consecutive tiles of a warp step with coalesced loads / stores (consecutive lanes = consecutive
// doubles) through a per-warp shared-memory area with an odd pose stride (bank-conflict free), and run the SAME
// per-pose functions on the staged values: identical bits.  Blocks of kPoseBlock threads.
constexpr int kPoseBlock = 128;
template <int TILE>
struct PoseStage {
  static constexpr int STRIDE = (TILE % 2) ? TILE : TILE + 1;
  static constexpr int WARP_DOUBLES = 32 * STRIDE;
};
// in(g) = value of global element g of the (combined) input array; pose(sw_tile, lane_valid) transforms the staged
// tile in place (OUT_TILE <= TILE doubles per pose are written back to out)
template <int TILE, int OUT_TILE, class In, class Pose>
__device__ __forceinline__ void pose_staged(int n, double *out, double *sw, In in, Pose pose) {
  using PS = PoseStage<TILE>;
  const int lane = threadIdx.x & 31;
  const int warp = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int nwarps = (int)((gridDim.x * blockDim.x) >> 5);
  for (int b0 = warp * 32; b0 < n; b0 += nwarps * 32) {
    const int np = min(32, n - b0);
    const size_t g0 = (size_t)b0 * TILE;
    if (np == 32) {
      double v[TILE];
#pragma unroll
      for (int k = 0; k < TILE; ++k) v[k] = in(g0 + lane + 32 * k);      // all loads in flight before the first use
#pragma unroll
      for (int k = 0; k < TILE; ++k) {
        const int e = lane + 32 * k, pl = e / TILE;
        sw[pl * PS::STRIDE + (e - pl * TILE)] = v[k];
      }
    } else {
      for (int e = lane; e < np * TILE; e += 32) {
        const int pl = e / TILE;
        sw[pl * PS::STRIDE + (e - pl * TILE)] = in(g0 + e);
      }
    }
    __syncwarp();
    if (lane < np) pose(sw + lane * PS::STRIDE);
    __syncwarp();
    const size_t o0 = (size_t)b0 * OUT_TILE;
    for (int e = lane; e < np * OUT_TILE; e += 32) {
      const int pl = e / OUT_TILE;
      out[o0 + e] = sw[pl * PS::STRIDE + (e - pl * OUT_TILE)];
    }
    __syncwarp();
  }
}
template <int R, int D, bool SCALED>
__device__ __forceinline__ void retract_staged(const double *X, const double *Eta, double *Xout, int n, double s,
                                               double *sw) {
  constexpr int TILE = R * (D + 1);
  pose_staged<TILE, TILE>(
      n, Xout, sw, [&](size_t g) { return SCALED ? __dadd_rn(X[g], __dmul_rn(s, Eta[g])) : X[g] + Eta[g]; },
      [&](double *t) { retract_pose<R, D>([&](int k) { return t[k]; }, t); });
}

// Polar projection of the Stiefel block (U V^T of the thin SVD) by one-sided Jacobi, one
// thread per pose (tiles staged through shared memory, polar_staged), of the combination M = ca*A + cb*B + cc*C (B, C optional); translation is
// the same combination, unprojected.
// ref: LiftedSEManifold::project src/manifold/LiftedSEManifold.cpp:34-45,
//      projectToStiefelManifold src/DPGO_utils.cpp:480-486, PGOAgent::updateY/updateV
//      src/PGOAgent.cpp:922-936.
// one pose: val(k) = entry k of the tile M; out = [polar factor of the Stiefel block | translation column of M]
template <int R, int D, class F>
__device__ __forceinline__ void polar_pose(F val, double *out) {
  {
    constexpr size_t o = 0;
    double a[D][R], v[D][D];
#pragma unroll
    for (int k = 0; k < D; ++k) {
#pragma unroll
      for (int q = 0; q < R; ++q) a[k][q] = val(k * R + q);
#pragma unroll
      for (int l = 0; l < D; ++l) v[k][l] = (k == l) ? 1.0 : 0.0;
    }
    for (int sweep = 0; sweep < 30; ++sweep) {
      bool converged = true;
#pragma unroll
      for (int p = 0; p < D - 1; ++p) {
#pragma unroll
        for (int q2 = p + 1; q2 < D; ++q2) {
          double al = 0.0, be = 0.0, ga = 0.0;
#pragma unroll
          for (int q = 0; q < R; ++q) {
            al = fma(a[p][q], a[p][q], al);
            be = fma(a[q2][q], a[q2][q], be);
            ga = fma(a[p][q], a[q2][q], ga);
          }
          // |ga| / sqrt(al be) against the two thresholds, as comparisons of squares: the kernel is bound by the
          // FP64 divide / square-root sequences of this loop, not by memory
          const double g2 = ga * ga, ab = al * be;
          if (g2 > 1e-30 * ab) converged = false;          // relative off-diagonal above 1e-15
          if (g2 > 1e-32 * ab) {                            // above 1e-16: rotate
            const double zeta = (be - al) / (2.0 * ga);
            const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
            const double cs = rsqrt(1.0 + t * t), sn = cs * t;
#pragma unroll
            for (int q = 0; q < R; ++q) {
              const double xp = a[p][q], xq = a[q2][q];
              a[p][q] = cs * xp - sn * xq;
              a[q2][q] = sn * xp + cs * xq;
            }
#pragma unroll
            for (int l = 0; l < D; ++l) {
              const double vp = v[p][l], vq = v[q2][l];
              v[p][l] = cs * vp - sn * vq;
              v[q2][l] = sn * vp + cs * vq;
            }
          }
        }
      }
      if (converged) break;
    }
    // a[k] = sigma_k u_k (column k of A V), v[k][l] = V[l][k];  U V^T = sum_k u_k v_k^T
    double u[D][R];
#pragma unroll
    for (int k = 0; k < D; ++k) {
      double nn = 0.0;
#pragma unroll
      for (int q = 0; q < R; ++q) nn = fma(a[k][q], a[k][q], nn);
      const double inv = (nn > 0.0) ? 1.0 / sqrt(nn) : 0.0;
#pragma unroll
      for (int q = 0; q < R; ++q) u[k][q] = a[k][q] * inv;
    }
#pragma unroll
    for (int l = 0; l < D; ++l) {
#pragma unroll
      for (int q = 0; q < R; ++q) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < D; ++k) s = fma(u[k][q], v[k][l], s);
        out[o + l * R + q] = s;
      }
    }
#pragma unroll
    for (int q = 0; q < R; ++q) out[o + D * R + q] = val(D * R + q);
  }
}
// M = ca*A + cb*B + cc*C, entry g of the arrays (B, C optional)
__device__ __forceinline__ double polar_combination(double ca, const double *A, double cb, const double *B, double cc,
                                                    const double *C, size_t g) {
  double m = ca * A[g];
  if (B) m = fma(cb, B[g], m);
  if (C) m = fma(cc, C[g], m);
  return m;
}
template <int R, int D>
__device__ __forceinline__ void polar_staged(double ca, const double *A, double cb, const double *B, double cc,
                                             const double *C, double *out, int n, double *sw) {
  constexpr int TILE = R * (D + 1);
  pose_staged<TILE, TILE>(
      n, out, sw, [&](size_t g) { return polar_combination(ca, A, cb, B, cc, C, g); },
      [&](double *t) { polar_pose<R, D>([&](int k) { return t[k]; }, t); });
}

// Rounding of the lifted iterate to SE(d) poses in the frame of an anchor pose (one thread per pose):
//   R_i = projectToRotationGroup(Ya^T Y_i),  t_i = Ya^T p_i - Ya^T pa
// ref: PGOAgent::getTrajectoryInLocalFrame / getTrajectoryInGlobalFrame src/PGOAgent.cpp:718-767,
//      projectToRotationGroup src/DPGO_utils.cpp:464-478 (SVD, last column of U negated when det U det V < 0 --
//      the singular values are sorted there, so the negated direction is the one of the smallest singular value).
// `anchor` is a lifted pose tile r x (d+1) (rotation Ya, translation pa); T is d x (d+1)n, column-major.
// one pose: x = the lifted tile, (ya, pa) = the anchor; t = the d x (d+1) rounded pose (may alias x: the tile is
// read into registers first)
template <int R, int D>
__device__ __forceinline__ void round_pose(const double *xin, const double (&ya)[D][R], const double (&pa)[R],
                                           double *t) {
  constexpr int DH = D + 1, TILE = R * DH;
  {
    double x[TILE];
#pragma unroll
    for (int k = 0; k < TILE; ++k) x[k] = xin[k];
    // a[k][l] = (Ya^T Y_i)[l][k]: column k of M as a[k][.]
    double a[D][D], v[D][D];
#pragma unroll
    for (int k = 0; k < D; ++k) {
#pragma unroll
      for (int l = 0; l < D; ++l) {
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < R; ++q) s = fma(ya[l][q], x[k * R + q], s);
        a[k][l] = s;
        v[k][l] = (k == l) ? 1.0 : 0.0;
      }
    }
    for (int sweep = 0; sweep < 30; ++sweep) {      // one-sided Jacobi: columns of M V become orthogonal
      bool converged = true;
#pragma unroll
      for (int p = 0; p < D - 1; ++p) {
#pragma unroll
        for (int q2 = p + 1; q2 < D; ++q2) {
          double al = 0.0, be = 0.0, ga = 0.0;
#pragma unroll
          for (int l = 0; l < D; ++l) {
            al = fma(a[p][l], a[p][l], al);
            be = fma(a[q2][l], a[q2][l], be);
            ga = fma(a[p][l], a[q2][l], ga);
          }
          // |ga| / sqrt(al be) against the two thresholds, as comparisons of squares: the kernel is bound by the
          // FP64 divide / square-root sequences of this loop, not by memory
          const double g2 = ga * ga, ab = al * be;
          if (g2 > 1e-30 * ab) converged = false;          // relative off-diagonal above 1e-15
          if (g2 > 1e-32 * ab) {                            // above 1e-16: rotate
            const double zeta = (be - al) / (2.0 * ga);
            const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
            const double cs = rsqrt(1.0 + t * t), sn = cs * t;
#pragma unroll
            for (int l = 0; l < D; ++l) {
              const double xp = a[p][l], xq = a[q2][l];
              a[p][l] = cs * xp - sn * xq;
              a[q2][l] = sn * xp + cs * xq;
              const double vp = v[p][l], vq = v[q2][l];
              v[p][l] = cs * vp - sn * vq;
              v[q2][l] = sn * vp + cs * vq;
            }
          }
        }
      }
      if (converged) break;
    }
    // a[k] = sigma_k u_k, v[k][l] = V[l][k];  U V^T = sum_k u_k v_k^T
    double sig[D], rot[D][D];   // rot[c][l] = (U V^T)[l][c]
    int kmin = 0;
#pragma unroll
    for (int k = 0; k < D; ++k) {
      double nn = 0.0;
#pragma unroll
      for (int l = 0; l < D; ++l) nn = fma(a[k][l], a[k][l], nn);
      sig[k] = sqrt(nn);
      const double inv = (nn > 0.0) ? 1.0 / sig[k] : 0.0;
#pragma unroll
      for (int l = 0; l < D; ++l) a[k][l] *= inv;
      if (sig[k] < sig[kmin]) kmin = k;
    }
#pragma unroll
    for (int c = 0; c < D; ++c)
#pragma unroll
      for (int l = 0; l < D; ++l) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < D; ++k) s = fma(a[k][l], v[k][c], s);
        rot[c][l] = s;
      }
    double det;
    if constexpr (D == 2) det = rot[0][0] * rot[1][1] - rot[1][0] * rot[0][1];
    else det = rot[0][0] * (rot[1][1] * rot[2][2] - rot[2][1] * rot[1][2]) -
               rot[1][0] * (rot[0][1] * rot[2][2] - rot[2][1] * rot[0][2]) +
               rot[2][0] * (rot[0][1] * rot[1][2] - rot[1][1] * rot[0][2]);
    if (det < 0.0) {   // reflect the direction of the smallest singular value
#pragma unroll
      for (int k = 0; k < D; ++k) {
        if (k == kmin) {
#pragma unroll
          for (int c = 0; c < D; ++c)
#pragma unroll
            for (int l = 0; l < D; ++l) rot[c][l] = fma(-2.0 * a[k][l], v[k][c], rot[c][l]);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < D; ++c)
#pragma unroll
      for (int l = 0; l < D; ++l) t[c * D + l] = rot[c][l];
#pragma unroll
    for (int l = 0; l < D; ++l) {
      double s = 0.0;
#pragma unroll
      for (int q = 0; q < R; ++q) s = fma(ya[l][q], x[D * R + q] - pa[q], s);
      t[D * D + l] = s;
    }
  }
}
template <int R, int D>
__device__ __forceinline__ void load_anchor(const double *anchor, double (&ya)[D][R], double (&pa)[R]) {
#pragma unroll
  for (int k = 0; k < D; ++k)
#pragma unroll
    for (int q = 0; q < R; ++q) ya[k][q] = anchor[k * R + q];
#pragma unroll
  for (int q = 0; q < R; ++q) pa[q] = anchor[D * R + q];
}
template <int R, int D>
__device__ __forceinline__ void round_staged(const double *X, const double *anchor, double *T, int n, double *sw) {
  constexpr int DH = D + 1, TILE = R * DH;
  double ya[D][R], pa[R];
  load_anchor<R, D>(anchor, ya, pa);
  pose_staged<TILE, D * DH>(
      n, T, sw, [&](size_t g) { return X[g]; }, [&](double *t) { round_pose<R, D>(t, ya, pa, t); });
}

// Elementwise phases over the r x N arrays (len = R * N doubles).
//   eta += a*delta; r += a*Hd; acc = {<r,r>}
__device__ __forceinline__ void phase_step(const Ctx &ctx, double a, const double *delta,
                                           const double *Hd, double *eta, double *r, size_t len,
                                           double (&acc)[1]) {
  for (size_t k = ctx.tid; k < len; k += ctx.nthreads) {
    eta[k] = fma(a, delta[k], eta[k]);
    const double rn = fma(a, Hd[k], r[k]);
    r[k] = rn;
    acc[0] = fma(rn, rn, acc[0]);
  }
}
// The same update, and the new residual also written in the PERMUTED column order of the two-level preconditioner
// (rp, R x pcols, padding columns stay zero): the first strip phase of the next application then stages its input
// slice with contiguous loads instead of gathering it through icol (measured: the gathering interior phase took
// 8.9 us per application against 5.5 us for the same strips fed from a permuted array).
template <int R, int D>
__device__ __forceinline__ void phase_step_perm(const Ctx &ctx, double a, const double *delta, const double *Hd,
                                                double *eta, double *r, const int *pcol, double *rp, size_t len,
                                                double (&acc)[1]) {
  constexpr int DH = D + 1;
  for (size_t k = ctx.tid; k < len; k += ctx.nthreads) {
    eta[k] = fma(a, delta[k], eta[k]);
    const double rn = fma(a, Hd[k], r[k]);
    r[k] = rn;
    const int col = (int)(k / R), q = (int)(k - (size_t)col * R);
    const int pose = col / DH, c = col - pose * DH;
    rp[(size_t)(__ldg(pcol + pose) + c) * R + q] = rn;
    acc[0] = fma(rn, rn, acc[0]);
  }
}
//   y = a*x + b*y
__device__ __forceinline__ void phase_axpby(const Ctx &ctx, double a, const double *x, double b,
                                            double *y, size_t len) {
  for (size_t k = ctx.tid; k < len; k += ctx.nthreads) y[k] = fma(a, x[k], b * y[k]);
}
__device__ __forceinline__ void phase_copy(const Ctx &ctx, const double *x, double *y, size_t len) {
  for (size_t k = ctx.tid; k < len; k += ctx.nthreads) y[k] = x[k];
}
__device__ __forceinline__ void phase_zero(const Ctx &ctx, double *y, size_t len) {
  for (size_t k = ctx.tid; k < len; k += ctx.nthreads) y[k] = 0.0;
}

}  // namespace dpgo
