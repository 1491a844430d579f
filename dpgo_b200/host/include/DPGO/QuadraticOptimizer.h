// Local Riemannian solver -- drop-in for the reference's QuadraticOptimizer
// (include/DPGO/QuadraticOptimizer.h:27-81).  optimize() runs the whole trust-region solve
// (or one gradient step) as a single persistent CUDA kernel (dpgo_optimize).
#ifndef DPGO_B200_QUADRATICOPTIMIZER_H
#define DPGO_B200_QUADRATICOPTIMIZER_H

#include <DPGO/DPGO_types.h>
#include <DPGO/QuadraticProblem.h>

namespace DPGO {

class QuadraticOptimizer {
 public:
  QuadraticOptimizer(QuadraticProblem *p, ROptParameters params = ROptParameters());
  ~QuadraticOptimizer() = default;

  /// optimize from Y; returns the new iterate (reference: src/QuadraticOptimizer.cpp:26-48)
  Matrix optimize(const Matrix &Y);

  void setVerbose(bool v) { params_.verbose = v; }
  void setAlgorithm(ROptParameters::ROptMethod alg) { params_.method = alg; }
  void setRGDStepsize(double s) { params_.RGD_stepsize = s; }
  void setRTRIterations(int iter) { params_.RTR_iterations = iter; }
  void setGradientNormTolerance(double tol) { params_.gradnorm_tol = tol; }
  void setRTRInitialRadius(double radius) { params_.RTR_initial_radius = radius; }
  void setRTRtCGIterations(int iter) { params_.RTR_tCG_iterations = iter; }

  ROPTResult getOptResult() const { return result_; }

 private:
  QuadraticProblem *problem_;  // not owned (reference: QuadraticOptimizer.h:85)
  ROptParameters params_;
  ROPTResult result_;
};

}  // namespace DPGO
#endif
