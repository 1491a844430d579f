// Initialization and single-robot solve helpers (reference: include/DPGO/DPGO_solver.h).
// Host-side cold path: the two sparse least-squares problems of the chordal relaxation are
// solved by preconditioned CG on the normal equations (the reference uses SuiteSparseQR; the
// minimizers are unique once pose 0 is anchored).
#ifndef DPGO_B200_SOLVER_H
#define DPGO_B200_SOLVER_H

#include <DPGO/DPGO_robust.h>
#include <DPGO/DPGO_types.h>
#include <DPGO/DPGO_utils.h>
#include <DPGO/RelativeSEMeasurement.h>
#include <DPGO/manifold/Poses.h>

#include <vector>

namespace DPGO {

// ---- single translation / rotation / pose averaging (reference: src/DPGO_solver.cpp:23-218), used by
// the robust inter-robot frame alignment.  An empty weight vector means unit weights.
void singleTranslationAveraging(Vector &tOpt, const std::vector<Vector> &tVec, const Vector &tau = Vector());
void singleRotationAveraging(Matrix &ROpt, const std::vector<Matrix> &RVec, const Vector &kappa = Vector());
void singlePoseAveraging(Matrix &ROpt, Vector &tOpt, const std::vector<Matrix> &RVec, const std::vector<Vector> &tVec,
                         const Vector &kappa = Vector(), const Vector &tau = Vector());
/// GNC-TLS rotation averaging; inlierIndices = measurements whose final weight is 1
void robustSingleRotationAveraging(Matrix &ROpt, std::vector<size_t> &inlierIndices, const std::vector<Matrix> &RVec,
                                   const Vector &kappa = Vector(), double errorThreshold = 0.1);
/// GNC-TLS pose averaging (default precisions kappa = 10000, tau = 100 when none are given)
void robustSinglePoseAveraging(Matrix &ROpt, Vector &tOpt, std::vector<size_t> &inlierIndices,
                               const std::vector<Matrix> &RVec, const std::vector<Vector> &tVec,
                               const Vector &kappa = Vector(), const Vector &tau = Vector(),
                               double errorThreshold = 0.1);

/// chordal relaxation (reference: src/DPGO_solver.cpp:220-269), solved on the device (dpgo_chordal_initialization)
PoseArray chordalInitialization(const std::vector<RelativeSEMeasurement> &measurements);
/// compose odometry from the identity or a partial trajectory (reference :271-303)
PoseArray odometryInitialization(const std::vector<RelativeSEMeasurement> &odometry,
                                 const PoseArray *partial_trajectory = nullptr);
/// single-robot pose-graph optimization at rank r = d (reference :305-333)
PoseArray solvePGO(const std::vector<RelativeSEMeasurement> &measurements, const ROptParameters &params,
                   const PoseArray *T0 = nullptr);

struct solveRobustPGOParams {
  ROptParameters opt_params;
  RobustCostParameters robust_params;
  bool verbose;
  solveRobustPGOParams() : opt_params(), robust_params(RobustCostParameters::Type::GNC_TLS), verbose(true) {}
};
/// single-robot pose-graph optimization with graduated non-convexity: alternates solvePGO with GNC-TLS
/// re-weighting of the measurements whose weight is not fixed (reference :335-412); the weights are
/// left in `mutable_measurements`
PoseArray solveRobustPGO(std::vector<RelativeSEMeasurement> &mutable_measurements, const solveRobustPGOParams &params,
                         const PoseArray *T0 = nullptr);

}  // namespace DPGO
#endif
