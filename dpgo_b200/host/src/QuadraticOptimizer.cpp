#include <DPGO/QuadraticOptimizer.h>

#include "check.h"

namespace DPGO {

QuadraticOptimizer::QuadraticOptimizer(QuadraticProblem *p, ROptParameters params)
    : mProblem(p), mOptions(params), mLastResult(false) {
  DPGO_CHECK(p != nullptr);
}

Matrix QuadraticOptimizer::optimize(const Matrix &Y) {
  DPGO_CHECK(static_cast<unsigned>(Y.rows()) == mProblem->relaxation_rank());
  DPGO_CHECK(static_cast<unsigned>(Y.cols()) == (mProblem->dimension() + 1) * mProblem->num_poses());
  const auto &graph = mProblem->poseGraph();
  const bool rtr = (mOptions.method == ROptParameters::ROptMethod::RTR);
  bool use_precon = rtr || mOptions.RGD_use_preconditioner;
  if (use_precon && !graph->hasPreconditioner()) {
    std::fprintf(stderr, "[QuadraticOptimizer] Failed to compute preconditioner.\n");
    DPGO_CHECK(!rtr);   // RTR needs it (the reference would run unpreconditioned tCG; unsupported here)
    use_precon = false;
  }
  DPGO_CHECK(graph->constructDataMatrices());

  dpgo_ropt_params prm;
  dpgo_default_params(&prm);
  prm.method = rtr ? 0 : 1;
  prm.verbose = mOptions.verbose ? 1 : 0;
  prm.gradnorm_tol = mOptions.gradnorm_tol;
  prm.RGD_stepsize = mOptions.RGD_stepsize;
  prm.RGD_use_preconditioner = use_precon ? 1 : 0;
  prm.RTR_iterations = mOptions.RTR_iterations;
  prm.RTR_tCG_iterations = mOptions.RTR_tCG_iterations;
  prm.RTR_initial_radius = mOptions.RTR_initial_radius;

  dpgo_ropt_result res;
  Matrix out(Y.rows(), Y.cols());
  DPGO_DEVICE_CALL(dpgo_optimize(graph->device(), &prm, Y.data(), out.data(), &res));
  mLastResult.success = res.success != 0;
  mLastResult.fInit = res.f_init;
  mLastResult.gradNormInit = res.gradnorm_init;
  mLastResult.fOpt = res.f_opt;
  mLastResult.gradNormOpt = res.gradnorm_opt;
  mLastResult.elapsedMs = res.elapsed_ms;
  mLastResult.tCGStatus = static_cast<tCGstatusSet>(res.tcg_status);
  return out;
}

void QuadraticOptimizer::setProblem(QuadraticProblem *problem) {
  DPGO_CHECK(problem != nullptr);
  mProblem = problem;
}
void QuadraticOptimizer::setVerbose(bool on) { mOptions.verbose = on; }
void QuadraticOptimizer::setAlgorithm(ROptParameters::ROptMethod method) { mOptions.method = method; }
void QuadraticOptimizer::setRGDStepsize(double stepsize) { mOptions.RGD_stepsize = stepsize; }
void QuadraticOptimizer::setGradientNormTolerance(double tolerance) { mOptions.gradnorm_tol = tolerance; }
void QuadraticOptimizer::setRTRIterations(int outer_iterations) { mOptions.RTR_iterations = outer_iterations; }
void QuadraticOptimizer::setRTRtCGIterations(int inner_iterations) { mOptions.RTR_tCG_iterations = inner_iterations; }
void QuadraticOptimizer::setRTRInitialRadius(double initial_radius) { mOptions.RTR_initial_radius = initial_radius; }

}  // namespace DPGO
