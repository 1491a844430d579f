// Public types of the DPGO drop-in API (B200 build).  Same names / fields / defaults as the
// reference's include/DPGO/DPGO_types.h so that code written against it compiles unchanged;
// Eigen / ROPTLIB / CHOLMOD types are replaced by self-contained ones.
#ifndef DPGO_B200_TYPES_H
#define DPGO_B200_TYPES_H

#include <DPGO/Matrix.h>

#include <cstddef>
#include <functional>
#include <iomanip>
#include <iostream>
#include <map>
#include <memory>
#include <string>
#include <tuple>
#include <utility>

namespace DPGO {

/// How a trajectory estimate is initialized (reference: DPGO_types.h:33-37)
enum class InitializationMethod { Odometry, Chordal, GNC_TLS };
std::string InitializationMethodToString(InitializationMethod method);

/// Termination status of the truncated-CG inner solver (stands in for
/// ROPTLIB::tCGstatusSet used at reference DPGO_types.h:106).
enum tCGstatusSet { TR_LCON = 0, TR_SCON, TR_NEGCURVTURE, TR_EXCREGION, TR_MAXITER, TCGSTATUSSETLENGTH };

/// Settings of the local Riemannian solve (defaults = reference DPGO_types.h:53-61)
class ROptParameters {
 public:
  enum class ROptMethod { RTR, RGD };

  ROptMethod method = ROptMethod::RTR;
  bool verbose = false;
  double gradnorm_tol = 1e-2;
  double RGD_stepsize = 1e-3;
  bool RGD_use_preconditioner = true;
  int RTR_iterations = 3;
  int RTR_tCG_iterations = 50;
  double RTR_initial_radius = 100;

  static std::string ROptMethodToString(ROptMethod method);

  friend std::ostream &operator<<(std::ostream &os, const ROptParameters &p) {
    os << "Riemannian optimization parameters: \n"
       << "Method: " << ROptMethodToString(p.method) << "\n"
       << "Gradient norm tol: " << p.gradnorm_tol << "\n"
       << "RGD stepsize: " << p.RGD_stepsize << "\n"
       << "RGD use preconditioner: " << p.RGD_use_preconditioner << "\n"
       << "RTR iterations: " << p.RTR_iterations << "\n"
       << "RTR tCG iterations: " << p.RTR_tCG_iterations << "\n"
       << "RTR initial radius: " << p.RTR_initial_radius << "\n";
    return os;
  }
};

/// Statistics of one local solve (reference DPGO_types.h:91-107)
struct ROPTResult {
  explicit ROPTResult(bool suc = false, double f0 = 0, double gn0 = 0, double fStar = 0, double gnStar = 0,
                      double ms = 0)
      : success(suc), fInit(f0), gradNormInit(gn0), fOpt(fStar), gradNormOpt(gnStar), elapsedMs(ms),
        tCGStatus(TR_MAXITER) {}
  bool success;
  double fInit, gradNormInit, fOpt, gradNormOpt, elapsedMs;
  tCGstatusSet tCGStatus;
};

/// A pose is identified by (robot, frame)
class PoseID {
 public:
  unsigned int robot_id, frame_id;
  explicit PoseID(unsigned int rid = 0, unsigned int fid = 0) : robot_id(rid), frame_id(fid) {}
  bool operator==(const PoseID &o) const { return robot_id == o.robot_id && frame_id == o.frame_id; }
};
struct ComparePoseID {
  bool operator()(const PoseID &a, const PoseID &b) const {
    return std::tie(a.robot_id, a.frame_id) < std::tie(b.robot_id, b.frame_id);
  }
};

/// A measurement is identified by an ordered pair of poses
class EdgeID {
 public:
  PoseID src_pose_id, dst_pose_id;
  EdgeID(const PoseID &src, const PoseID &dst) : src_pose_id(src), dst_pose_id(dst) {}
  bool operator==(const EdgeID &o) const { return src_pose_id == o.src_pose_id && dst_pose_id == o.dst_pose_id; }
  bool isOdometry() const {
    return src_pose_id.robot_id == dst_pose_id.robot_id && src_pose_id.frame_id + 1 == dst_pose_id.frame_id;
  }
  bool isPrivateLoopClosure() const {
    return src_pose_id.robot_id == dst_pose_id.robot_id && src_pose_id.frame_id + 1 != dst_pose_id.frame_id;
  }
  bool isSharedLoopClosure() const { return src_pose_id.robot_id != dst_pose_id.robot_id; }
};
struct CompareEdgeID {
  bool operator()(const EdgeID &a, const EdgeID &b) const {
    return std::make_tuple(a.src_pose_id.robot_id, a.dst_pose_id.robot_id, a.src_pose_id.frame_id,
                           a.dst_pose_id.frame_id) <
           std::make_tuple(b.src_pose_id.robot_id, b.dst_pose_id.robot_id, b.src_pose_id.frame_id,
                           b.dst_pose_id.frame_id);
  }
};
struct HashEdgeID {
  std::size_t operator()(const EdgeID &e) const {
    std::size_t seed = 0;
    auto mix = [&seed](unsigned v) { seed ^= std::hash<unsigned>()(v) + 0x9e3779b97f4a7c15ULL + (seed << 6) + (seed >> 2); };
    mix(e.src_pose_id.robot_id);
    mix(e.dst_pose_id.robot_id);
    mix(e.src_pose_id.frame_id);
    mix(e.dst_pose_id.frame_id);
    return seed;
  }
};

}  // namespace DPGO
#endif
