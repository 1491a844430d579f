// Initialization and single-robot solve helpers (reference: include/DPGO/DPGO_solver.h).
// Host-side cold path: the two sparse least-squares problems of the chordal relaxation are
// solved by preconditioned CG on the normal equations (the reference uses SuiteSparseQR; the
// minimizers are unique once pose 0 is anchored).
#ifndef DPGO_B200_SOLVER_H
#define DPGO_B200_SOLVER_H

#include <DPGO/DPGO_types.h>
#include <DPGO/DPGO_utils.h>
#include <DPGO/RelativeSEMeasurement.h>
#include <DPGO/manifold/Poses.h>

#include <vector>

namespace DPGO {

/// chordal relaxation (reference: src/DPGO_solver.cpp:220-269)
PoseArray chordalInitialization(const std::vector<RelativeSEMeasurement> &measurements);
/// compose odometry from the identity or a partial trajectory (reference :271-303)
PoseArray odometryInitialization(const std::vector<RelativeSEMeasurement> &odometry,
                                 const PoseArray *partial_trajectory = nullptr);
/// single-robot pose-graph optimization at rank r = d (reference :305-333)
PoseArray solvePGO(const std::vector<RelativeSEMeasurement> &measurements, const ROptParameters &params,
                   const PoseArray *T0 = nullptr);

}  // namespace DPGO
#endif
