// One agent's pose graph: measurements on the host, data matrices on the GPU.
//
// Drop-in for the reference's PoseGraph (include/DPGO/PoseGraph.h).  What the reference stores
// as Eigen objects -- the sparse quadratic matrix Q, the dense linear matrix G and the CHOLMOD
// factorization of Q + 0.1 I (src/PoseGraph.cpp:381-613) -- lives in device memory behind a
// dpgo_handle (include/dpgo_b200.h): block-CSR Q, G built from the neighbour-pose buffer by a
// kernel, dense inverse of Q + 0.1 I.  The invalidation rules are the reference's:
// setNeighborPoses resets G only (:183-186); weight / active-neighbour / prior changes reset Q
// and the preconditioner (:199-207, :352-355).
#ifndef DPGO_B200_POSEGRAPH_H
#define DPGO_B200_POSEGRAPH_H

#include <DPGO/DPGO_types.h>
#include <DPGO/RelativeSEMeasurement.h>
#include <DPGO/manifold/Poses.h>

#include <cstdint>
#include <map>
#include <memory>
#include <set>
#include <unordered_map>
#include <vector>

struct dpgo_dev;  // C-ABI handle (include/dpgo_b200.h)

namespace DPGO {

class PoseGraph {
 public:
  /// statistics of loop-closure weights (meaningful for GNC_TLS)
  class Statistics {
   public:
    Statistics() : total_loop_closures(0), accept_loop_closures(0), reject_loop_closures(0), undecided_loop_closures(0) {}
    double total_loop_closures, accept_loop_closures, reject_loop_closures, undecided_loop_closures;
  };

  PoseGraph(unsigned int id, unsigned int r, unsigned int d);
  ~PoseGraph();
  PoseGraph(const PoseGraph &) = delete;
  PoseGraph &operator=(const PoseGraph &) = delete;

  unsigned int d() const { return d_; }
  unsigned int r() const { return r_; }
  unsigned int n() const { return n_; }
  unsigned int numOdometry() const { return static_cast<unsigned>(odometry_.size()); }
  unsigned int numPrivateLoopClosures() const { return static_cast<unsigned>(private_lcs_.size()); }
  unsigned int numSharedLoopClosures() const { return static_cast<unsigned>(shared_lcs_.size()); }
  unsigned int numMeasurements() const;

  void empty();
  void reset();
  void clearNeighborPoses();
  void setMeasurements(const std::vector<RelativeSEMeasurement> &measurements);
  void addMeasurement(const RelativeSEMeasurement &m);

  std::vector<RelativeSEMeasurement> odometry() const { return odometry_; }
  std::vector<RelativeSEMeasurement> privateLoopClosures() const { return private_lcs_; }
  std::vector<RelativeSEMeasurement> sharedLoopClosures() const { return shared_lcs_; }
  std::vector<RelativeSEMeasurement> sharedLoopClosuresWithRobot(unsigned int neighbor_id) const;
  std::vector<RelativeSEMeasurement> measurements() const;
  std::vector<RelativeSEMeasurement> localMeasurements() const;

  void clearPriors();
  void setPrior(unsigned index, const LiftedPose &Xi);
  void setNeighborPoses(const PoseDict &pose_dict);

  bool hasNeighbor(unsigned int robot_id) const;
  bool isNeighborActive(unsigned int neighbor_id) const;
  void setNeighborActive(unsigned int neighbor_id, bool active);
  bool requireNeighborPose(const PoseID &pose_id) const;
  bool hasMeasurement(const PoseID &srcID, const PoseID &dstID) const;
  RelativeSEMeasurement *findMeasurement(const PoseID &srcID, const PoseID &dstID);
  std::vector<RelativeSEMeasurement *> allLoopClosures();
  std::vector<RelativeSEMeasurement *> activeLoopClosures();
  std::vector<RelativeSEMeasurement *> inactiveLoopClosures();

  PoseSet myPublicPoseIDs() const { return local_shared_pose_ids_; }
  PoseSet neighborPublicPoseIDs() const { return nbr_shared_pose_ids_; }
  PoseSet activeNeighborPublicPoseIDs() const;
  std::set<unsigned> neighborIDs() const { return nbr_robot_ids_; }
  std::set<unsigned> activeNeighborIDs() const;
  size_t numNeighbors() const { return nbr_robot_ids_.size(); }
  size_t numActiveNeighbors() const;
  Statistics statistics() const;

  /// Build whatever is stale on the device (Q and/or G).  false when a required neighbour pose
  /// is missing (reference: src/PoseGraph.cpp:364-370, 417-430).
  bool constructDataMatrices();
  void clearDataMatrices();
  void clearQuadraticMatrix();
  void clearLinearMatrix();
  /// G as a host matrix (downloads it); Q as block-CSR triplets
  Matrix linearMatrix();
  void quadraticMatrixBSR(std::vector<int> &rowptr, std::vector<int> &colidx, std::vector<double> &blocks);
  /// build the preconditioner if needed; false if Q + 0.1 I is not positive definite
  bool hasPreconditioner();
  void useInactiveNeighbors(bool use = true);
  /// tell the graph that measurement weights were edited through findMeasurement()
  void weightsChanged() { clearDataMatrices(); }

  /// device side of this graph (valid after constructDataMatrices())
  dpgo_dev *device() { return dev_; }
  /// make sure the device handle exists for the current number of poses (no matrices built);
  /// the generation counter changes whenever the handle had to be re-created
  dpgo_dev *deviceHandle();
  unsigned deviceGeneration() const { return dev_generation_; }

 private:
  void addOdometry(const RelativeSEMeasurement &factor);
  void addPrivateLoopClosure(const RelativeSEMeasurement &factor);
  void addSharedLoopClosure(const RelativeSEMeasurement &factor);
  void updatePublicPoseIDs();
  bool ensureDevice();
  bool constructQ();
  bool constructG();
  bool selectSharedEdges(std::vector<size_t> &edges, std::vector<PoseID> &slots) const;

  unsigned int id_, r_, d_, n_;
  std::vector<RelativeSEMeasurement> odometry_, private_lcs_, shared_lcs_;
  std::unordered_map<EdgeID, size_t, HashEdgeID> edge_id_to_index_;
  PoseSet local_shared_pose_ids_, nbr_shared_pose_ids_;
  std::set<unsigned> nbr_robot_ids_;
  std::map<unsigned, bool> neighbor_active_;
  PoseDict neighbor_poses_;
  std::map<unsigned, LiftedPose> priors_;
  bool use_inactive_neighbors_;
  double prior_kappa_, prior_tau_;

  // device state
  dpgo_dev *dev_;
  unsigned dev_n_;
  unsigned dev_generation_ = 0;
  bool q_valid_, g_valid_, precon_valid_;
  std::vector<size_t> q_edges_;     // shared edges included in the current Q
  std::vector<PoseID> q_slots_;     // neighbour poses in slot order
  // what the device handle was last given, weights aside: when only the measurement weights changed (GNC) the
  // matrices are re-weighted on the device (dpgo_update_weights) instead of being assembled again
  struct UploadedGraph {
    std::vector<int32_t> p1, p2, my, slot, prior_idx;
    std::vector<uint8_t> outgoing;
    std::vector<double> Rp, tp, kp, taup, Rs, ts, ks, taus, prior_tiles;
    double prior_kappa = 0, prior_tau = 0;
    unsigned generation = 0;
    bool valid = false, had_precon = false;
    bool sameStructure(const UploadedGraph &o) const {
      return valid && o.valid && generation == o.generation && p1 == o.p1 && p2 == o.p2 && my == o.my && slot == o.slot &&
             prior_idx == o.prior_idx && outgoing == o.outgoing && Rp == o.Rp && tp == o.tp && kp == o.kp &&
             taup == o.taup && Rs == o.Rs && ts == o.ts && ks == o.ks && taus == o.taus &&
             prior_tiles == o.prior_tiles && prior_kappa == o.prior_kappa && prior_tau == o.prior_tau;
    }
  } uploaded_;
};

}  // namespace DPGO
#endif
