// In-tree FP64 dense linear algebra for the set-up of the exact preconditioner (Q + 0.1 I)^-1
// (ref: src/PoseGraph.cpp:598-613 -- CHOLMOD `compute(Q + 0.1 I)` there).  Device functions only;
// the launch sequences are in dense_la.cu.  Everything is column-major, lower-triangle convention,
// tiles and panels of TS = 64, batched over blockIdx.z through descriptor arrays so that all interior
// domains of the two-level form are factorized in lockstep.
//
//   SPD inverse of A (n x n), in place in the lower triangle:
//     1. blocked right-looking Cholesky  A = L L^T   (dla_diag_block, gemm NT x2 per 64-column panel)
//     2. W = L^-1 by recursive doubling  [[A,0],[B,C]]^-1 = [[A^-1,0],[-C^-1 B A^-1, C^-1]]
//        (the 64 x 64 diagonal inverses come out of step 1; log2(n/64) levels of two batched GEMMs)
//     3. A^-1 = W^T W                    (one GEMM TN over the lower tiles)
//   All three are n^3/3 flops of tile GEMM; the triangular structure is used at tile granularity
//   through the k-range modes of the GEMM device function (strictly-upper tiles are never read).
//
// The bodies take their block coordinates as arguments (no blockIdx inside) so that the CPU
// emulation harness of tests/native can run the same code one CTA at a time.
#pragma once
#include <cstddef>

#ifndef DPGO_CPU_EMU
#include <cuda_runtime.h>
#endif

namespace dpgo {
namespace dla {

constexpr int TS = 64;          // tile edge = Cholesky panel width
constexpr int BK = 16;          // k-chunk of the tile GEMM
constexpr int kThreads = 256;   // 16 x 16 threads, 4 x 4 outputs each
constexpr int LDS = TS + 4;     // shared-memory row pitch (doubles); keeps 32-byte alignment of the 4-vectors

// k-range modes (tile-granular use of triangular operands)
enum KMode : int {
  K_FULL = 0,
  K_FROM_COL_TILE = 1,   // k >= j0            (right operand lower triangular)
  K_UPTO_ROW_TILE = 2,   // k <  i0 + TS       (left operand lower triangular)
  K_FROM_ROW_TILE = 3    // k >= i0            (W^T W on lower tiles: k >= i >= j)
};

// C (M x N, ldc) = alpha * op(A) * op(B) + beta * C ; op(A) is M x K, op(B) is K x N.
// ta = 0: A stored M x K ; ta = 1: A stored K x M (op = transpose).  Same for tb with B stored K x N / N x K.
struct GemmDesc {
  const double *A;
  const double *B;
  double *C;
  int M, N, K;
  int lda, ldb, ldc;
};

struct GemmFlags {
  int ta, tb;
  int lower_only;   // compute tiles with i0 >= j0 only (square outputs)
  int kmode;
  double alpha, beta;
};

// One 64 x 64 output tile.  sA / sB: BK x LDS doubles each, two buffers.
__device__ __forceinline__ void dla_gemm_tile(const GemmDesc &g, const GemmFlags &f, int bi, int bj, double *sA,
                                              double *sB) {
  const int i0 = bi * TS, j0 = bj * TS;
  if (i0 >= g.M || j0 >= g.N) return;
  if (f.lower_only && i0 < j0) return;
  int k_lo = 0, k_hi = g.K;
  if (f.kmode == K_FROM_COL_TILE) k_lo = j0;
  else if (f.kmode == K_UPTO_ROW_TILE) k_hi = min(g.K, i0 + TS);
  else if (f.kmode == K_FROM_ROW_TILE) k_lo = i0;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;   // thread owns rows 4*tx.., columns 4*ty.. of the tile
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;

  // global -> registers for one k-chunk: every thread carries 4 elements of each operand
  double ra[4], rb[4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      // element index e in [0, 1024): operand tile is TS (tile dim) x BK (k)
      const int e = tid + q * kThreads;
      int ia, ka;
      if (f.ta == 0) { ia = e & (TS - 1); ka = e >> 6; }      // contiguous along the tile dimension
      else { ka = e & (BK - 1); ia = e >> 4; }                 // contiguous along k
      const int gi = i0 + ia, gk = k0 + ka;
      double v = 0.0;
      if (gi < g.M && gk < k_hi)
        v = f.ta == 0 ? g.A[(size_t)gi + (size_t)gk * g.lda] : g.A[(size_t)gk + (size_t)gi * g.lda];
      ra[q] = v;
      int jb, kb;
      if (f.tb == 0) { kb = e & (BK - 1); jb = e >> 4; }       // B stored K x N: contiguous along k
      else { jb = e & (TS - 1); kb = e >> 6; }                 // B stored N x K: contiguous along the tile dimension
      const int gj = j0 + jb, gk2 = k0 + kb;
      double w = 0.0;
      if (gj < g.N && gk2 < k_hi)
        w = f.tb == 0 ? g.B[(size_t)gk2 + (size_t)gj * g.ldb] : g.B[(size_t)gj + (size_t)gk2 * g.ldb];
      rb[q] = w;
    }
  };
  auto stash = [&](int buf) {
    double *a = sA + buf * (BK * LDS), *b = sB + buf * (BK * LDS);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int e = tid + q * kThreads;
      int ia, ka, jb, kb;
      if (f.ta == 0) { ia = e & (TS - 1); ka = e >> 6; } else { ka = e & (BK - 1); ia = e >> 4; }
      if (f.tb == 0) { kb = e & (BK - 1); jb = e >> 4; } else { jb = e & (TS - 1); kb = e >> 6; }
      a[ka * LDS + ia] = ra[q];
      b[kb * LDS + jb] = rb[q];
    }
  };

  if (k_lo < k_hi) {
    fetch(k_lo);
    stash(0);
  }
  __syncthreads();
  int buf = 0;
  for (int k0 = k_lo; k0 < k_hi; k0 += BK) {
    const bool more = k0 + BK < k_hi;
    if (more) fetch(k0 + BK);
    const double *a = sA + buf * (BK * LDS) + 4 * tx, *b = sB + buf * (BK * LDS) + 4 * ty;
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      double av[4], bv[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) { av[q] = a[kk * LDS + q]; bv[q] = b[kk * LDS + q]; }
#pragma unroll
      for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) acc[x][y] = fma(av[x], bv[y], acc[x][y]);
    }
    if (more) stash(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }
  // epilogue (after every read of this CTA: C may alias the rows of A this tile was computed from)
#pragma unroll
  for (int y = 0; y < 4; ++y) {
    const int gj = j0 + 4 * ty + y;
    if (gj >= g.N) continue;
#pragma unroll
    for (int x = 0; x < 4; ++x) {
      const int gi = i0 + 4 * tx + x;
      if (gi >= g.M) continue;
      double *c = g.C + (size_t)gi + (size_t)gj * g.ldc;
      *c = f.beta == 0.0 ? f.alpha * acc[x][y] : fma(f.alpha, acc[x][y], f.beta * *c);
    }
  }
}

// ---- diagonal block: unblocked Cholesky + triangular inverse of one (<=) 64 x 64 block -------------------
struct SpdDesc {
  double *A;      // n x n, lower triangle in / out
  int n, lda;
  double *dinv;   // ceil(n / 64) blocks of 64 x 64 (ld 64): inverses of the diagonal blocks of L
  double *tmp;    // n x n scratch (ld ldt)
  int ldt;
};

// sL, sW: TS x (TS + 1) doubles each.  info: set to 1 + batch index when a pivot is not positive.
// 256 threads = 64 rows (columns) x 4 parts: a row's dot product is split over 4 adjacent lanes and combined with
// two shuffles, so a Cholesky column costs two CTA barriers and <= 16 FMAs per thread, and the 64 columns of the
// triangular inverse advance together one row per step.  The block is a chain of 2 x 64 latency-bound steps
// (measured 89-94 us per block with sqrt + divide in every step and a single running sum).
__device__ __forceinline__ void dla_diag_block(const SpdDesc &d, int panel, int batch, double *sL, double *sW,
                                               int *info) {
  constexpr int P = TS + 1;
  const int j0 = panel * TS;
  if (j0 >= d.n) return;
  const int nb = min(TS, d.n - j0);
  const int tid = threadIdx.x;
  const int row = tid >> 2, part = tid & 3;       // 4 adjacent lanes share a row (Cholesky) / a column (inverse)
  double *A = d.A + (size_t)j0 + (size_t)j0 * d.lda;
  for (int e = tid; e < TS * TS; e += kThreads) {
    const int i = e & (TS - 1), j = e >> 6;
    sL[i * P + j] = (i < nb && j < nb && i >= j) ? A[(size_t)i + (size_t)j * d.lda] : 0.0;
    sW[i * P + j] = 0.0;
  }
  __shared__ int s_bad;
  __shared__ double s_piv;
  __shared__ double s_rdiag[TS];                  // 1 / L[i][i]
  if (tid == 0) s_bad = 0;
  __syncthreads();
  // ---- left-looking Cholesky: column j = (A[:, j] - L[:, :j] L[j, :j]^T) / sqrt(pivot).  The step is a latency
  // chain (dot product -> shuffles -> barrier -> reciprocal square root -> barrier): four independent partial sums
  // per lane and rsqrt instead of sqrt + divide keep it short.
  for (int j = 0; j < nb; ++j) {
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    if (row >= j && row < nb) {
      const double *lr = sL + row * P, *lj = sL + j * P;
      int k = part;
      for (; k + 12 < j; k += 16) {
        s0 = fma(lr[k], lj[k], s0);
        s1 = fma(lr[k + 4], lj[k + 4], s1);
        s2 = fma(lr[k + 8], lj[k + 8], s2);
        s3 = fma(lr[k + 12], lj[k + 12], s3);
      }
      for (; k < j; k += 4) s0 = fma(lr[k], lj[k], s0);
    }
    double s = (s0 + s1) + (s2 + s3);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    const double v = sL[row * P + j] - s;         // unscaled entry (row, j); meaningful for j <= row < nb
    if (part == 0 && row == j) s_piv = v;
    __syncthreads();
    const double piv = s_piv;
    if (!(piv > 0.0)) {                           // the same value in every thread: uniform exit
      if (tid == 0) { s_bad = 1; atomicMax(info, 1 + batch); }
      break;
    }
    const double rinv = rsqrt(piv);
    if (part == 0 && row >= j && row < nb) {
      sL[row * P + j] = (row == j) ? piv * rinv : v * rinv;
      if (row == j) s_rdiag[j] = rinv;
    }
    __syncthreads();
  }
  __syncthreads();
  if (s_bad) return;
  // ---- W = L^-1 (lower): column c = thread group `row`; x_i = (delta_ic - sum_{c <= k < i} L[i][k] x_k) / L[i][i]
  {
    const int c = row;
    for (int i = 0; i < nb; ++i) {                // all groups step through the rows together (warp-synchronous)
      double s0 = 0.0, s1 = 0.0;
      if (c < nb && i > c) {
        const double *li = sL + i * P;
        int k = c + part;
        for (; k + 4 < i; k += 8) {
          s0 = fma(li[k], sW[k * P + c], s0);
          s1 = fma(li[k + 4], sW[(k + 4) * P + c], s1);
        }
        for (; k < i; k += 4) s0 = fma(li[k], sW[k * P + c], s0);
      }
      double s = s0 + s1;
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      if (part == 0 && c < nb && i >= c) sW[i * P + c] = (((i == c) ? 1.0 : 0.0) - s) * s_rdiag[i];
      __syncwarp();                               // x_i of this warp's columns is visible to its 4-lane groups
    }
  }
  __syncthreads();
  double *D = d.dinv + (size_t)panel * TS * TS;
  for (int e = tid; e < TS * TS; e += kThreads) {
    const int i = e & (TS - 1), j = e >> 6;
    D[e] = sW[i * P + j];                                             // zero above the diagonal and in the padding
    if (i < nb && j < nb) A[(size_t)i + (size_t)j * d.lda] = sL[i * P + j];   // L, zero above the diagonal
  }
}

// ---- tile copies ---------------------------------------------------------------------------------------------
// what = 0: diagonal blocks of A <- dinv (start of the triangular inverse)
// what = 1: lower tiles of A <- tmp (result of W^T W), diagonal tiles whole
// what = 2: upper triangle of A <- transpose of the lower one (full symmetric matrix for a plain GEMM)
__device__ __forceinline__ void dla_copy_tile(const SpdDesc &d, int what, int bi, int bj) {
  const int i0 = bi * TS, j0 = bj * TS;
  if (i0 >= d.n || j0 >= d.n) return;
  if (what == 0 && bi != bj) return;
  if (what == 1 && bi < bj) return;
  if (what == 2 && bi < bj) return;
  for (int e = threadIdx.x; e < TS * TS; e += kThreads) {
    const int i = e & (TS - 1), j = e >> 6;
    const int gi = i0 + i, gj = j0 + j;
    if (gi >= d.n || gj >= d.n) continue;
    if (what == 0) d.A[(size_t)gi + (size_t)gj * d.lda] = d.dinv[(size_t)bi * TS * TS + e];
    else if (what == 1) d.A[(size_t)gi + (size_t)gj * d.lda] = d.tmp[(size_t)gi + (size_t)gj * d.ldt];
    else if (gi > gj) d.A[(size_t)gj + (size_t)gi * d.lda] = d.A[(size_t)gi + (size_t)gj * d.lda];
  }
}

}  // namespace dla
}  // namespace dpgo
