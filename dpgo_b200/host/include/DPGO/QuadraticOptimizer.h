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
  /// result of the last optimize() call
  ROPTResult getOptResult() const { return mLastResult; }

  // configuration (same setters as the reference; they only edit the parameter block that
  // optimize() hands to the device solver)
  void setProblem(QuadraticProblem *problem);
  void setVerbose(bool on);
  void setAlgorithm(ROptParameters::ROptMethod method);
  void setRGDStepsize(double stepsize);
  void setGradientNormTolerance(double tolerance);
  void setRTRIterations(int outer_iterations);
  void setRTRtCGIterations(int inner_iterations);
  void setRTRInitialRadius(double initial_radius);

 private:
  QuadraticProblem *mProblem;  // not owned
  ROptParameters mOptions;
  ROPTResult mLastResult;
};

}  // namespace DPGO
#endif
