// Test infrastructure (not part of the product libraries): the Q*X device functions of dpgo_b200/csrc/kernels.cuh
// (phase_qx, its two-blocks-per-step form and the shared-memory tile-staged form) compiled for the host through
// tests/native/cuda_emu.h and run as one CTA of 256 real threads per virtual CTA.
#define QX_EXPORT __attribute__((visibility("default")))
#include "cuda_emu.h"

#include <vector>


#include "../../dpgo_b200/csrc/kernels.cuh"

using namespace dpgo;

namespace {
template <int R, int D>
int run(int variant, const int *rowptr, const int *colidx, const double *blocks, const double *X, const double *G,
        double *out, int n, int ctas) {
  static unsigned char stage[kWarpsPerBlock][QxTiles<R, D>::WARP_BYTES + 16];
  for (int wv = 0; wv < emu::kWarps; ++wv) emu::warp_barrier[wv] = new std::barrier<>(32);
  const BsrView Q{rowptr, colidx, blocks};
  for (int cta = 0; cta < ctas; ++cta) {
    std::vector<std::thread> th;
    for (unsigned t = 0; t < (unsigned)emu::kThreads; ++t)
      th.emplace_back([&, t]() {
        threadIdx.x = t;
        Ctx c;                                   // make_ctx() of virtual CTA `cta` out of `ctas`
        c.tid = cta * 256 + (int)t;
        c.nthreads = ctas * 256;
        c.warp = c.tid >> 5;
        c.nwarps = c.nthreads >> 5;
        c.lane = (int)t & 31;
        if (variant == 2) phase_qx_tiles<R, D>(c, Q, X, G, out, n, stage[t >> 5]);
        else if (variant == 3) phase_qx<R, D, true>(c, Q, X, G, out, n);
        else phase_qx<R, D, false>(c, Q, X, G, out, n);
      });
    for (auto &x : th) x.join();
  }
  for (int wv = 0; wv < emu::kWarps; ++wv) delete emu::warp_barrier[wv];
  return 0;
}
}  // namespace

extern "C" QX_EXPORT int qx_emu(int variant, int r, int d, const int *rowptr, const int *colidx, const double *blocks,
                                const double *X, const double *G, double *out, int n, int ctas) {
  if (r == 5 && d == 3) return run<5, 3>(variant, rowptr, colidx, blocks, X, G, out, n, ctas);
  if (r == 3 && d == 3) return run<3, 3>(variant, rowptr, colidx, blocks, X, G, out, n, ctas);
  if (r == 3 && d == 2) return run<3, 2>(variant, rowptr, colidx, blocks, X, G, out, n, ctas);
  if (r == 4 && d == 2) return run<4, 2>(variant, rowptr, colidx, blocks, X, G, out, n, ctas);
  return -1;
}

namespace {
// Per-pose kernels: op 0 = QF retraction of X + Eta, 1 = polar projection of 0.5 A + 0.3 B + 0.2 C, 2 = rounding in
// the frame of pose 0.  staged = 1: the stand-alone form (tiles of a warp step through shared memory, pose_staged);
// staged = 0: the per-pose function applied to every tile by one host thread.  Same function, same values: same bits.
template <int R, int D>
int run_pose_op(int op, int staged, const double *A, const double *B, const double *C, double *out, int n) {
  constexpr int DH = D + 1, TILE = R * DH;
  if (!staged) {
    double ya[D][R], pa[R];
    load_anchor<R, D>(A, ya, pa);
    for (int i = 0; i < n; ++i) {
      const size_t o = (size_t)i * TILE;
      if (op == 0) retract_pose<R, D>([&](int k) { return A[o + k] + B[o + k]; }, out + o);
      else if (op == 1) polar_pose<R, D>([&](int k) { return polar_combination(0.5, A, 0.3, B, 0.2, C, o + k); }, out + o);
      else round_pose<R, D>(A + o, ya, pa, out + (size_t)i * (D * DH));
    }
    return 0;
  }
  static double sw[emu::kWarps][PoseStage<TILE>::WARP_DOUBLES];
  for (int wv = 0; wv < emu::kWarps; ++wv) emu::warp_barrier[wv] = new std::barrier<>(32);
  std::vector<std::thread> th;
  for (unsigned t = 0; t < (unsigned)emu::kThreads; ++t)
    th.emplace_back([&, t]() {
      threadIdx.x = t;
      if (op == 0) retract_staged<R, D, false>(A, B, out, n, 1.0, sw[t >> 5]);
      else if (op == 1) polar_staged<R, D>(0.5, A, 0.3, B, 0.2, C, out, n, sw[t >> 5]);
      else round_staged<R, D>(A, A, out, n, sw[t >> 5]);
    });
  for (auto &x : th) x.join();
  for (int wv = 0; wv < emu::kWarps; ++wv) delete emu::warp_barrier[wv];
  return 0;
}
}  // namespace

extern "C" QX_EXPORT int pose_op_emu(int op, int staged, int r, int d, const double *A, const double *B, const double *C,
                                     double *out, int n) {
  if (r == 5 && d == 3) return run_pose_op<5, 3>(op, staged, A, B, C, out, n);
  if (r == 3 && d == 2) return run_pose_op<3, 2>(op, staged, A, B, C, out, n);
  if (r == 4 && d == 2) return run_pose_op<4, 2>(op, staged, A, B, C, out, n);
  return -1;
}
