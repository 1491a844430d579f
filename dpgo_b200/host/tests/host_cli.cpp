// CPU-only command line probe of the host-side (non-CUDA) parts of the drop-in API, used by the
// "not gpu" tests:  host_cli parse <file.g2o>   |   host_cli chordal <file.g2o>  (chordal: needs a GPU)
//                   host_cli logroundtrip <file.g2o> <dir/>  (PGOLogger: write + reload CSV logs)
//                   host_cli averaging <file>  (robust single rotation / pose averaging of the
//                                               candidate alignments listed in <file>)
#include <DPGO/DPGO_solver.h>
#include <DPGO/DPGO_utils.h>
#include <DPGO/DPGO_robust.h>
#include <DPGO/PGOLogger.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

using namespace DPGO;

int main(int argc, char **argv) {
  if (argc >= 2 && !std::strcmp(argv[1], "chi2") && argc == 4) {  // chi2inv(quantile, dof), testUtils.cpp:56-70
    std::printf("%.15g\n", chi2inv(std::atof(argv[2]), static_cast<size_t>(std::atoi(argv[3]))));
    return 0;
  }
  const bool averaging_with_params = (argc == 6 && !std::strcmp(argv[1], "averaging"));
  if (argc != 3 && !(argc == 4 && !std::strcmp(argv[1], "logroundtrip")) && !averaging_with_params) {
    std::fprintf(stderr, "usage: %s parse|chordal <file.g2o> | logroundtrip <file.g2o> <dir/>\n", argv[0]);
    return 2;
  }
  if (!std::strcmp(argv[1], "averaging")) {
    std::FILE *f = std::fopen(argv[2], "r");
    if (!f) return 3;
    int cnt = 0, dim = 0;
    if (std::fscanf(f, "%d %d", &cnt, &dim) != 2) return 3;
    std::vector<Matrix> RVec;
    std::vector<Vector> tVec;
    for (int k = 0; k < cnt; ++k) {
      Matrix R(dim, dim);
      Vector t(dim, 1);
      for (int a = 0; a < dim; ++a)
        for (int b = 0; b < dim; ++b)
          if (std::fscanf(f, "%lf", &R(a, b)) != 1) return 3;
      for (int a = 0; a < dim; ++a)
        if (std::fscanf(f, "%lf", &t(a, 0)) != 1) return 3;
      RVec.push_back(R);
      tVec.push_back(t);
    }
    std::fclose(f);
    auto dump = [&](const char *tag, const Matrix &R, const Vector &t, const std::vector<size_t> &inl) {
      std::printf("%s", tag);
      for (std::ptrdiff_t a = 0; a < R.rows(); ++a)
        for (std::ptrdiff_t b = 0; b < R.cols(); ++b) std::printf(" %.17g", R(a, b));
      for (std::ptrdiff_t a = 0; a < t.rows(); ++a) std::printf(" %.17g", t(a, 0));
      std::printf(" |");
      for (size_t i : inl) std::printf(" %zu", i);
      std::printf("\n");
    };
    Matrix ROpt;
    Vector tOpt;
    std::vector<size_t> inl;
    // the two-stage alignment of PGOAgent::computeRobustNeighborTransformTwoStage
    // optional: <rotation threshold in rad> <kappa> <tau> (defaults = PGOAgent's alignment settings)
    const double rot_thr = averaging_with_params ? std::atof(argv[3]) : 0.5;
    const double kap = averaging_with_params ? std::atof(argv[4]) : 1.82;
    const double ta = averaging_with_params ? std::atof(argv[5]) : 0.01;
    robustSingleRotationAveraging(ROpt, inl, RVec, Vector(), angular2ChordalSO3(rot_thr));
    std::vector<Vector> tin;
    for (size_t i : inl) tin.push_back(tVec[i]);
    if (tin.empty()) tin = tVec;
    singleTranslationAveraging(tOpt, tin);
    dump("TWOSTAGE", ROpt, tOpt, inl);
    // the joint alignment of PGOAgent::computeRobustNeighborTransform
    Vector kappa(cnt, 1), tau(cnt, 1);
    for (int k = 0; k < cnt; ++k) { kappa(k) = kap; tau(k) = ta; }
    robustSinglePoseAveraging(ROpt, tOpt, inl, RVec, tVec, kappa, tau, RobustCost::computeErrorThresholdAtQuantile(0.9, 3));
    dump("JOINT", ROpt, tOpt, inl);
    return 0;
  }
  size_t n = 0;
  const std::vector<RelativeSEMeasurement> ms = read_g2o_file(argv[2], n);
  const size_t d = ms.empty() ? 0 : static_cast<size_t>(ms[0].t.size());
  if (!std::strcmp(argv[1], "parse")) {
    std::printf("%zu %zu %zu\n", n, ms.size(), d);
    for (const auto &m : ms) {
      std::printf("%zu %zu %d", m.p1, m.p2, m.fixedWeight ? 1 : 0);
      for (size_t a = 0; a < d; ++a)
        for (size_t b = 0; b < d; ++b) std::printf(" %.17g", m.R(a, b));
      for (size_t a = 0; a < d; ++a) std::printf(" %.17g", m.t(a, 0));
      std::printf(" %.17g %.17g\n", m.kappa, m.tau);
    }
    return 0;
  }
  if (!std::strcmp(argv[1], "chordal")) {   // needs a GPU (the relaxation is solved on the device)
    const Matrix T = chordalInitialization(ms).getData();
    std::printf("%td %td\n", T.rows(), T.cols());
    for (std::ptrdiff_t j = 0; j < T.cols(); ++j)
      for (std::ptrdiff_t i = 0; i < T.rows(); ++i) std::printf("%.17g\n", T(i, j));
    return 0;
  }
  if (!std::strcmp(argv[1], "logroundtrip")) {
    // measurements (with GNC weights / fixed flags) and the chordal trajectory through the CSV
    // files; prints the largest deviation after reloading
    std::vector<RelativeSEMeasurement> in = ms;
    for (size_t k = 0; k < in.size(); ++k) in[k].weight = 1.0 / (1.0 + static_cast<double>(k % 7));
    PGOLogger logger(argv[3]);
    logger.logMeasurements(in, "measurements.csv");
    // a trajectory to log without a GPU: the odometry chain composed from the identity
    std::vector<RelativeSEMeasurement> odom(n > 0 ? n - 1 : 0);
    size_t found = 0;
    for (const auto &e : ms)
      if (e.p2 == e.p1 + 1 && e.p1 + 1 < n) { odom[e.p1] = e; ++found; }
    if (found + 1 < n) return 4;
    const Matrix T = odometryInitialization(odom).getData();
    logger.logTrajectory(static_cast<unsigned>(d), static_cast<unsigned>(n), T, "trajectory.csv");
    const std::vector<RelativeSEMeasurement> out =
        PGOLogger::loadMeasurements(std::string(argv[3]) + "measurements.csv", true);
    const Matrix T2 = logger.loadTrajectory("trajectory.csv");
    if (out.size() != in.size() || T2.cols() != T.cols()) return 3;
    double devR = 0, devt = 0, devw = 0, devT = (T - T2).norm();
    bool flags = true;
    for (size_t k = 0; k < in.size(); ++k) {
      devR = std::max(devR, (in[k].R - out[k].R).norm());
      devt = std::max(devt, (in[k].t - out[k].t).norm());
      devw = std::max(devw, std::fabs(in[k].weight - out[k].weight));
      flags = flags && in[k].fixedWeight == out[k].fixedWeight && in[k].p1 == out[k].p1 && in[k].p2 == out[k].p2 &&
              in[k].r1 == out[k].r1 && in[k].r2 == out[k].r2;
    }
    std::printf("RESULT %zu %.3e %.3e %.3e %.3e %d\n", out.size(), devR, devt, devw, devT, flags ? 1 : 0);
    return 0;
  }
  return 2;
}
