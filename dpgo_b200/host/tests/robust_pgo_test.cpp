// Restatement of the reference's testRobustPGO (tests/testPGO.cpp:193-271) against the drop-in, on a
// GPU box: a 4-pose chain with fixed-weight odometry, one correct and one wrong loop closure;
// solveRobustPGO must give the correct one weight 1 and the wrong one weight 0.  Several random
// instances (the reference draws one).  Exit code 0 = all passed.
#include <DPGO/DPGO_solver.h>
#include <DPGO/PoseGraph.h>

#include <cmath>
#include <cstdio>
#include <memory>
#include <random>

using namespace DPGO;

static Matrix randomRotation(std::mt19937 &rng) {
  std::normal_distribution<double> g(0.0, 1.0);
  double q[4], nrm = 0;
  for (double &v : q) { v = g(rng); nrm += v * v; }
  nrm = std::sqrt(nrm);
  const double x = q[0] / nrm, y = q[1] / nrm, z = q[2] / nrm, w = q[3] / nrm;
  Matrix R(3, 3);
  R(0, 0) = 1 - 2 * (y * y + z * z); R(0, 1) = 2 * (x * y - w * z);     R(0, 2) = 2 * (x * z + w * y);
  R(1, 0) = 2 * (x * y + w * z);     R(1, 1) = 1 - 2 * (x * x + z * z); R(1, 2) = 2 * (y * z - w * x);
  R(2, 0) = 2 * (x * z - w * y);     R(2, 1) = 2 * (y * z + w * x);     R(2, 2) = 1 - 2 * (x * x + y * y);
  return R;
}

int main() {
  const unsigned d = 3, n = 4;
  const double kappa = 10000, tau = 100;
  int failed = 0;
  std::mt19937 rng(12345);
  for (int trial = 0; trial < 5; ++trial) {
    std::vector<Pose> gt;
    for (unsigned i = 0; i < n; ++i) {
      Pose Ti(d);
      Ti.rotation() = randomRotation(rng);
      Matrix t(3, 1);
      t(0, 0) = t(1, 0) = t(2, 0) = static_cast<double>(i);
      Ti.translation() = t;
      gt.push_back(Ti);
    }
    auto relative = [&](unsigned i, unsigned j, bool fixed) {
      const Pose Tij = gt[i].inverse() * gt[j];
      RelativeSEMeasurement m(0, 0, i, j, Tij.rotation(), Tij.translation(), kappa, tau);
      m.fixedWeight = fixed;
      return m;
    };
    std::vector<RelativeSEMeasurement> ms;
    for (unsigned i = 0; i + 1 < n; ++i) ms.push_back(relative(i, i + 1, true));
    ms.push_back(relative(0, 3, false));                       // inlier loop closure
    RelativeSEMeasurement outlier(0, 0, 1, 3, randomRotation(rng), Matrix(3, 1), kappa, tau);
    ms.push_back(outlier);                                     // wrong rotation, zero translation
    auto graph = std::make_shared<PoseGraph>(0, d, d);
    graph->setMeasurements(ms);
    solveRobustPGOParams params;
    params.verbose = false;
    params.opt_params.verbose = false;
    params.opt_params.gradnorm_tol = 1e-1;
    params.opt_params.RTR_iterations = 50;
    params.robust_params.GNCBarc = 7.0;
    PoseArray TOdom = odometryInitialization(graph->odometry());
    std::vector<RelativeSEMeasurement> mutable_ms = ms;
    PoseArray T = solveRobustPGO(mutable_ms, params, &TOdom);
    for (const RelativeSEMeasurement &m : mutable_ms) {
      if (m.fixedWeight) continue;
      const double want = (m.p1 == 0 && m.p2 == 3) ? 1.0 : 0.0;
      if (std::fabs(m.weight - want) > 1e-6) {
        std::printf("trial %d: edge %zu->%zu weight %.9g, expected %.0f\n", trial, m.p1, m.p2, m.weight, want);
        failed++;
      }
    }
    // with the outlier rejected the chain + inlier closure is consistent: relative poses match
    for (unsigned i = 0; i + 1 < n; ++i) {
      const Pose est = Pose(T.pose(i)).inverse() * Pose(T.pose(i + 1));
      const Pose ref = gt[i].inverse() * gt[i + 1];
      if ((est.getData() - ref.getData()).norm() > 1e-2) {
        std::printf("trial %d: relative pose %u->%u off by %.3g\n", trial, i, i + 1, (est.getData() - ref.getData()).norm());
        failed++;
      }
    }
  }
  if (failed) {
    std::printf("%d check(s) FAILED\n", failed);
    return 1;
  }
  std::printf("robust PGO test passed\n");
  return 0;
}
