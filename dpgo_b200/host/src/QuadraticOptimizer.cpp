#include <DPGO/QuadraticOptimizer.h>

#include "check.h"

namespace DPGO {

QuadraticOptimizer::QuadraticOptimizer(QuadraticProblem *p, ROptParameters params)
    : problem_(p), params_(params), result_(false) {
  DPGO_CHECK(p != nullptr);
}

Matrix QuadraticOptimizer::optimize(const Matrix &Y) {
  DPGO_CHECK(static_cast<unsigned>(Y.rows()) == problem_->relaxation_rank());
  DPGO_CHECK(static_cast<unsigned>(Y.cols()) == (problem_->dimension() + 1) * problem_->num_poses());
  const auto &graph = problem_->poseGraph();
  const bool rtr = (params_.method == ROptParameters::ROptMethod::RTR);
  bool use_precon = rtr || params_.RGD_use_preconditioner;
  if (use_precon && !graph->hasPreconditioner()) {
    std::fprintf(stderr, "[QuadraticOptimizer] Failed to compute preconditioner.\n");
    DPGO_CHECK(!rtr);   // RTR needs it (the reference would run unpreconditioned tCG; unsupported here)
    use_precon = false;
  }
  DPGO_CHECK(graph->constructDataMatrices());

  dpgo_ropt_params prm;
  dpgo_default_params(&prm);
  prm.method = rtr ? 0 : 1;
  prm.verbose = params_.verbose ? 1 : 0;
  prm.gradnorm_tol = params_.gradnorm_tol;
  prm.RGD_stepsize = params_.RGD_stepsize;
  prm.RGD_use_preconditioner = use_precon ? 1 : 0;
  prm.RTR_iterations = params_.RTR_iterations;
  prm.RTR_tCG_iterations = params_.RTR_tCG_iterations;
  prm.RTR_initial_radius = params_.RTR_initial_radius;

  dpgo_ropt_result res;
  Matrix out(Y.rows(), Y.cols());
  DPGO_DEVICE_CALL(dpgo_optimize(graph->device(), &prm, Y.data(), out.data(), &res));
  result_.success = res.success != 0;
  result_.fInit = res.f_init;
  result_.gradNormInit = res.gradnorm_init;
  result_.fOpt = res.f_opt;
  result_.gradNormOpt = res.gradnorm_opt;
  result_.elapsedMs = res.elapsed_ms;
  result_.tCGStatus = static_cast<tCGstatusSet>(res.tcg_status);
  return out;
}

}  // namespace DPGO
