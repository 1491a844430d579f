// Chordal / odometry initialization and the single-robot solve (host side, cold path).
#include <DPGO/DPGO_solver.h>
#include <DPGO/DPGO_utils.h>
#include <DPGO/PoseGraph.h>
#include <DPGO/QuadraticOptimizer.h>
#include <DPGO/QuadraticProblem.h>

#include <cmath>
#include <cstdio>
#include <vector>

#include "check.h"

namespace DPGO {

// Device path (dpgo_chordal_initialization: both least-squares problems solved on the GPU with the library's
// Q*X and exact-preconditioner kernels; the reference factorizes with SPQR, src/DPGO_solver.cpp:220-269).
PoseArray chordalInitialization(const std::vector<RelativeSEMeasurement> &measurements) {
  size_t d, n;
  get_dimension_and_num_poses(measurements, d, n);
  PoseArray T(static_cast<unsigned>(d), static_cast<unsigned>(n));
  const size_t m = measurements.size();
  std::vector<int32_t> p1(m), p2(m);
  std::vector<double> R(m * d * d), t(m * d), kappa(m), tau(m);
  for (size_t k = 0; k < m; ++k) {
    const RelativeSEMeasurement &e = measurements[k];
    p1[k] = static_cast<int32_t>(e.p1);
    p2[k] = static_cast<int32_t>(e.p2);
    for (size_t a = 0; a < d; ++a) {
      for (size_t b = 0; b < d; ++b) R[(k * d + a) * d + b] = e.R(a, b);
      t[k * d + a] = e.t(a, 0);
    }
    kappa[k] = e.kappa;
    tau[k] = e.tau;
  }
  Matrix Tm(static_cast<std::ptrdiff_t>(d), static_cast<std::ptrdiff_t>(n * (d + 1)));
  dpgo_chordal_info info;
  DPGO_DEVICE_CALL(dpgo_chordal_initialization(defaultDevice(), static_cast<int>(n), static_cast<int>(d),
                                               static_cast<int>(m), p1.data(), p2.data(), R.data(), t.data(),
                                               kappa.data(), tau.data(), Tm.data(), &info));
  if (info.rotation_residual > 1e-8 || info.translation_residual > 1e-8)
    std::fprintf(stderr, "[DPGO] chordalInitialization: linear solves stopped at relative residuals %.3e (rotations, "
                         "%d iterations) / %.3e (translations, %d iterations); the initial guess may be inaccurate\n",
                 info.rotation_residual, info.rotation_iterations, info.translation_residual,
                 info.translation_iterations);
  T.setData(Tm);
  return T;
}

PoseArray odometryInitialization(const std::vector<RelativeSEMeasurement> &odometry, const PoseArray *partial) {
  size_t d, n;
  get_dimension_and_num_poses(odometry, d, n);
  PoseArray T(static_cast<unsigned>(d), static_cast<unsigned>(n));
  unsigned next = 1;
  if (partial && partial->n() > 0) {
    DPGO_CHECK(partial->d() == d);
    for (next = 0; next < partial->n() && next < n; ++next) T.pose(next) = partial->pose(next);
  }
  for (size_t dst = next; dst < n; ++dst) {
    const RelativeSEMeasurement &m = odometry[dst - 1];
    DPGO_CHECK(m.p1 == dst - 1 && m.p2 == dst);
    const Matrix Rs = T.rotation(static_cast<unsigned>(dst - 1));
    const Matrix ts = T.translation(static_cast<unsigned>(dst - 1));
    T.rotation(static_cast<unsigned>(dst)) = Rs * m.R;
    T.translation(static_cast<unsigned>(dst)) = ts + Rs * m.t;
  }
  return T;
}

PoseArray solvePGO(const std::vector<RelativeSEMeasurement> &measurements, const ROptParameters &params,
                   const PoseArray *T0) {
  size_t d, n;
  get_dimension_and_num_poses(measurements, d, n);
  const unsigned robot_id = static_cast<unsigned>(measurements[0].r1);
  PoseArray T = T0 ? *T0 : chordalInitialization(measurements);
  DPGO_CHECK(T.d() == d && T.n() == n);
  auto pose_graph = std::make_shared<PoseGraph>(robot_id, static_cast<unsigned>(d), static_cast<unsigned>(d));
  pose_graph->setMeasurements(measurements);
  QuadraticProblem problem(pose_graph);
  QuadraticOptimizer optimizer(&problem, params);
  T.setData(optimizer.optimize(T.getData()));
  return T;
}

// ---- averaging (reference :23-218) --------------------------------------------------------------------
namespace {
Vector weightsOrOnes(const Vector &w, size_t n, double fill = 1.0) {
  if (static_cast<size_t>(w.rows()) == n) return w;
  Vector o(static_cast<std::ptrdiff_t>(n), 1);
  for (size_t i = 0; i < n; ++i) o(static_cast<std::ptrdiff_t>(i)) = fill;
  return o;
}
Vector hadamard(const Vector &a, const Vector &b) {
  Vector o(a.rows(), 1);
  for (std::ptrdiff_t i = 0; i < a.rows(); ++i) o(i) = a(i) * b(i);
  return o;
}
// GNC-TLS loop shared by the rotation and the pose averaging: solve(weights) refreshes the estimate,
// residualSq(i) evaluates measurement i at it (reference :92-133, :170-217)
template <typename Solve, typename Residual>
void gncAveraging(size_t n, double errorThreshold, unsigned maxIters, Solve solve, Residual residualSq,
                  std::vector<size_t> &inlierIndices) {
  const double w_tol = 1e-8;
  Vector weights = weightsOrOnes(Vector(), n);
  solve(weights);
  double rmax = 0;
  for (size_t i = 0; i < n; ++i) rmax = std::max(rmax, residualSq(i));
  const double barcSq = errorThreshold * errorThreshold;
  const double muInit = std::min(barcSq / (2 * rmax - barcSq), 1e-5);
  if (muInit > 0) {  // a negative value means every residual is already small: no GNC needed
    RobustCostParameters params(RobustCostParameters::Type::GNC_TLS);
    params.GNCBarc = errorThreshold;
    params.GNCMaxNumIters = maxIters;
    params.GNCInitMu = muInit;
    RobustCost cost(params);
    for (unsigned iter = 0; iter < maxIters; ++iter) {
      solve(weights);
      size_t converged = 0;
      for (size_t i = 0; i < n; ++i) {
        const double wi = cost.weight(std::sqrt(residualSq(i)));
        if (wi < w_tol || wi > 1 - w_tol) converged++;
        weights(static_cast<std::ptrdiff_t>(i)) = wi;
      }
      if (converged == n) break;
      cost.update();
    }
  }
  inlierIndices.clear();
  for (size_t i = 0; i < n; ++i)
    if (weights(static_cast<std::ptrdiff_t>(i)) > 1 - w_tol) inlierIndices.push_back(i);
}
}  // namespace

void singleTranslationAveraging(Vector &tOpt, const std::vector<Vector> &tVec, const Vector &tau) {
  DPGO_CHECK(!tVec.empty());
  const Vector w = weightsOrOnes(tau, tVec.size());
  Vector s(tVec[0].rows(), 1);
  double wsum = 0;
  for (size_t i = 0; i < tVec.size(); ++i) {
    s += tVec[i] * w(static_cast<std::ptrdiff_t>(i));
    wsum += w(static_cast<std::ptrdiff_t>(i));
  }
  tOpt = s * (1.0 / wsum);
}

void singleRotationAveraging(Matrix &ROpt, const std::vector<Matrix> &RVec, const Vector &kappa) {
  DPGO_CHECK(!RVec.empty());
  const Vector w = weightsOrOnes(kappa, RVec.size());
  Matrix M(RVec[0].rows(), RVec[0].rows());
  for (size_t i = 0; i < RVec.size(); ++i) M += RVec[i] * w(static_cast<std::ptrdiff_t>(i));
  ROpt = projectToRotationGroup(M);
}

void singlePoseAveraging(Matrix &ROpt, Vector &tOpt, const std::vector<Matrix> &RVec, const std::vector<Vector> &tVec,
                         const Vector &kappa, const Vector &tau) {
  DPGO_CHECK(!RVec.empty() && RVec.size() == tVec.size());
  DPGO_CHECK(RVec[0].rows() == tVec[0].rows());
  singleTranslationAveraging(tOpt, tVec, tau);
  singleRotationAveraging(ROpt, RVec, kappa);
}

void robustSingleRotationAveraging(Matrix &ROpt, std::vector<size_t> &inlierIndices, const std::vector<Matrix> &RVec,
                                   const Vector &kappa, double errorThreshold) {
  const size_t n = RVec.size();
  DPGO_CHECK(n > 0);
  const Vector kappa_ = weightsOrOnes(kappa, n);
  for (const Matrix &Ri : RVec) checkRotationMatrix(Ri);
  gncAveraging(
      n, errorThreshold, 1000, [&](const Vector &w) { singleRotationAveraging(ROpt, RVec, hadamard(kappa_, w)); },
      [&](size_t i) { return kappa_(static_cast<std::ptrdiff_t>(i)) * (ROpt - RVec[i]).squaredNorm(); }, inlierIndices);
}

void robustSinglePoseAveraging(Matrix &ROpt, Vector &tOpt, std::vector<size_t> &inlierIndices,
                               const std::vector<Matrix> &RVec, const std::vector<Vector> &tVec, const Vector &kappa,
                               const Vector &tau, double errorThreshold) {
  const size_t n = RVec.size();
  DPGO_CHECK(n > 0 && tVec.size() == n);
  const Vector kappa_ = weightsOrOnes(kappa, n, 10000.0), tau_ = weightsOrOnes(tau, n, 100.0);
  for (const Matrix &Ri : RVec) checkRotationMatrix(Ri);
  gncAveraging(
      n, errorThreshold, 10000,
      [&](const Vector &w) { singlePoseAveraging(ROpt, tOpt, RVec, tVec, hadamard(kappa_, w), hadamard(tau_, w)); },
      [&](size_t i) {
        const std::ptrdiff_t k = static_cast<std::ptrdiff_t>(i);
        return kappa_(k) * (ROpt - RVec[i]).squaredNorm() + tau_(k) * (tOpt - tVec[i]).squaredNorm();
      },
      inlierIndices);
}

// ---- robust single-robot solve (reference :335-412) -------------------------------------------------------
PoseArray solveRobustPGO(std::vector<RelativeSEMeasurement> &ms, const solveRobustPGOParams &params,
                         const PoseArray *T0) {
  DPGO_CHECK(params.robust_params.costType == RobustCostParameters::Type::GNC_TLS);
  const double w_tol = 1e-8;
  PoseArray T = solvePGO(ms, params.opt_params, T0);
  auto errorSq = [&](const RelativeSEMeasurement &m) {
    return computeMeasurementError(m, T.rotation(static_cast<unsigned>(m.p1)), T.translation(static_cast<unsigned>(m.p1)),
                                   T.rotation(static_cast<unsigned>(m.p2)), T.translation(static_cast<unsigned>(m.p2)));
  };
  double rmax = 0;
  for (RelativeSEMeasurement &m : ms) {
    m.weight = 1.0;
    rmax = std::max(rmax, errorSq(m));
  }
  const double barcSq = params.robust_params.GNCBarc * params.robust_params.GNCBarc;
  RobustCostParameters gnc = params.robust_params;
  gnc.GNCInitMu = barcSq / (2 * rmax - barcSq);
  if (params.verbose) std::printf("[solveRobustPGO] Initial value for mu: %g\n", gnc.GNCInitMu);
  if (gnc.GNCInitMu > 0) {
    RobustCost cost(gnc);
    for (unsigned iter = 0; iter < gnc.GNCMaxNumIters; ++iter) {
      T = solvePGO(ms, params.opt_params, T0);
      int inliers = 0, outliers = 0, undecided = 0;
      for (RelativeSEMeasurement &m : ms) {
        if (m.fixedWeight) continue;
        m.weight = cost.weight(std::sqrt(errorSq(m)));
        if (m.weight < w_tol) outliers++;
        else if (m.weight > 1.0 - w_tol) inliers++;
        else undecided++;
      }
      if (params.verbose)
        std::printf("[solveRobustPGO] Iteration %u: %d inliers, %d outliers, %d undecided.\n", iter, inliers, outliers,
                    undecided);
      if (undecided == 0) break;
      cost.update();
    }
  }
  return solvePGO(ms, params.opt_params, T0);
}

}  // namespace DPGO
