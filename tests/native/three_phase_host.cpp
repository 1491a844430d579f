// Test infrastructure (not part of the product libraries): the index arithmetic the CUDA kernels of the
// three-phase preconditioner share with the host (dpgo_b200/csrc/dd_stage.h), compiled with g++ so that
// tests/test_three_phase_plan_cpu.py can compare it with the numpy replay without a device.
// built with -fvisibility=hidden -Wl,-Bsymbolic: the product library exports host stubs with the same mangled
// names as the kernels compiled here, and must not interpose them when both are loaded in one process
#define TP_EXPORT __attribute__((visibility("default")))
#include <stdint.h>

#include "../../dpgo_b200/csrc/dd_stage.h"

extern "C" {

// out[i * R + q] = row q of element idx[i] of the staged input slice, staging mode src (1, 2 or 3)
TP_EXPORT int tp_stage_values(int src, int R, int nidx, const int32_t *idx, const double *vec, const int32_t *icol,
                    const int32_t *gidx, const double *sub, const int32_t *tptr, const int32_t *tcol, int col0,
                    int nslots, int64_t slotstride, double *out) {
  const dpgo::StageAux ax{sub, tptr, tcol, col0, nslots, (size_t)slotstride};
  for (int i = 0; i < nidx; ++i)
    for (int q = 0; q < R; ++q) {
      double v;
      if (src == 1) v = dpgo::strip_stage_value<1>(idx[i], q, R, vec, icol, gidx, &ax);
      else if (src == 2) v = dpgo::strip_stage_value<2>(idx[i], q, R, vec, icol, gidx, &ax);
      else if (src == 3) v = dpgo::strip_stage_value<3>(idx[i], q, R, vec, icol, gidx, &ax);
      else return -1;
      out[(size_t)i * R + q] = v;
    }
  return 0;
}

// dst = stage-major strips (nob output blocks x nch chunks) of the coupling block C, as k_dd_layout_rect writes them
TP_EXPORT int tp_layout_rect(const double *C, int m, int ldc, const int32_t *colmap, int ncomp, int form, int nob, int nch,
                   double *dst) {
  const size_t total = (size_t)nob * nch * 32 * 64;
  for (size_t t = 0; t < total; ++t) dst[t] = dpgo::layout_rect_value(C, m, ldc, colmap, ncomp, form, nch, t);
  return 0;
}

}  // extern "C"
