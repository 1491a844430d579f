// Index arithmetic of the three-phase form that is shared between the CUDA kernels and a host test
// (tests/native/three_phase_host.cpp compiles this header with g++ and compares with the numpy
// replay): how one element of the staged input slice of a strip phase is produced, and which element
// of a coupling block goes where in the stage-major strip image.  No state, no memory traffic of its
// own beyond the loads written here.
#pragma once
#include <stddef.h>

#if defined(__CUDACC__)
#define DPGO_HD __host__ __device__ __forceinline__
#else
#define DPGO_HD inline
#endif
#if defined(__CUDA_ARCH__)
#define DPGO_RO(p) __ldg(p)   // immutable index arrays: read-only path on the device
#else
#define DPGO_RO(p) (*(p))
#endif

namespace dpgo {

// Extra inputs of the staging step of a strip phase in the three-phase form (dd_plan.h):
//   SRC 2: vec is r in the original order; the value of scalar column c of the S segment is
//          r[icol[c]] - sum_e sub[tcol[e]],  e in [tptr[c - col0], tptr[c - col0 + 1])
//   SRC 3: vec is an array of `nslots` partial results `slotstride` apart, read through gidx
struct StageAux {
  const double *sub;
  const int *tptr, *tcol;
  int col0;
  int nslots;
  size_t slotstride;
};

// Row q of element `idx` of the input slice of a strip (idx = 32 * (kc0 + chunk) + kk):
//   SRC 1: idx is a permuted scalar column, vec (original order) is gathered through icol
//   SRC 2: the same minus the scattered coupling terms (forms t_S = r_S - sum_k g_k)
//   SRC 3: idx is a position of the gather list gidx; the partial slots of vec are summed in order
template <int SRC>
DPGO_HD double strip_stage_value(int idx, int q, int R, const double *vec, const int *icol, const int *gidx,
                                 const StageAux *ax) {
  if constexpr (SRC == 3) {
    const int col = DPGO_RO(gidx + idx);
    double val = 0.0;
    if (col >= 0) {
      const double *zp = vec + (size_t)col * R + q;
      for (int sl = 0; sl < ax->nslots; ++sl) val += zp[(size_t)sl * ax->slotstride];
    }
    return val;
  } else {
    const int oc = DPGO_RO(icol + idx);
    double val = (oc >= 0) ? vec[(size_t)oc * R + q] : 0.0;
    if constexpr (SRC == 2) {
      const int j = idx - ax->col0;
      const int e0 = DPGO_RO(ax->tptr + j), e1 = DPGO_RO(ax->tptr + j + 1);
      for (int e = e0; e < e1; ++e) val -= ax->sub[(size_t)DPGO_RO(ax->tcol + e) * R + q];
    }
    return val;
  }
}

// Element t of the stage-major strips of a coupling block C = A_k^-1 A_kS (m rows, col-major, leading
// dimension ldc) restricted to the columns colmap[0 .. ncomp) (nullptr: C holds exactly those columns),
// zero padded; stage (ob, c) is the (ob * nch + c)-th run of 32 x 64 values, element (kk, jj) at kk*64 + jj:
//   form 0 (phase 1, output = compact column):  C(32 c + kk, colmap[64 ob + jj])
//   form 1 (phase 5, output = domain row):      C(64 ob + jj, colmap[32 c + kk])
DPGO_HD double layout_rect_value(const double *C, int m, int ldc, const int *colmap, int ncomp, int form, int nch,
                                 size_t t) {
  const int jj = (int)(t % 64);
  const size_t u = t / 64;
  const int kk = (int)(u % 32);
  const size_t v = u / 32;
  const int c = (int)(v % nch);
  const int ob = (int)(v / nch);
  const int row = form == 0 ? c * 32 + kk : ob * 64 + jj;
  const int comp = form == 0 ? ob * 64 + jj : c * 32 + kk;
  if (row < m && comp < ncomp) return C[(size_t)row + (size_t)(colmap ? colmap[comp] : comp) * ldc];
  return 0.0;
}

}  // namespace dpgo
