// PoseGraph shell: measurement bookkeeping on the host (same rules as the reference's
// src/PoseGraph.cpp), data matrices on the device through the C-ABI.
#include <DPGO/DPGO_utils.h>
#include <DPGO/PoseGraph.h>

#include <algorithm>
#include <cstdio>

#include "check.h"

namespace DPGO {

PoseGraph::PoseGraph(unsigned int id, unsigned int r, unsigned int d)
    : id_(id), r_(r), d_(d), n_(0), use_inactive_neighbors_(false), prior_kappa_(10000), prior_tau_(100),
      dev_(nullptr), dev_n_(0), q_valid_(false), g_valid_(false), precon_valid_(false) {
  DPGO_CHECK(r >= d);  // reference: src/PoseGraph.cpp:19
  DPGO_CHECK(d == 2 || d == 3);
  if (r > d + 3) {
    // the register-resident pose tiles of the CUDA kernels are instantiated for d <= r <= d + 3 (the reference
    // accepts any r >= d; its examples and dpgo_ros use r = d .. 5): refuse here, not at the first device call
    std::fprintf(stderr, "[DPGO] relaxation rank r = %u is not supported by the CUDA path (d = %u: %u <= r <= %u)\n",
                 r, d, d, d + 3);
    std::abort();
  }
  empty();
}

PoseGraph::~PoseGraph() {
  if (dev_) dpgo_destroy(dev_);
}

void PoseGraph::empty() {
  n_ = 0;
  edge_id_to_index_.clear();
  odometry_.clear();
  private_lcs_.clear();
  shared_lcs_.clear();
  local_shared_pose_ids_.clear();
  nbr_shared_pose_ids_.clear();
  nbr_robot_ids_.clear();
  neighbor_active_.clear();
  clearNeighborPoses();
  clearDataMatrices();
  clearPriors();
}

void PoseGraph::reset() {
  clearNeighborPoses();
  clearDataMatrices();
  clearPriors();
  for (unsigned nb : nbr_robot_ids_) neighbor_active_[nb] = true;
}

void PoseGraph::clearNeighborPoses() {
  neighbor_poses_.clear();
  g_valid_ = false;
}

unsigned int PoseGraph::numMeasurements() const {
  return numOdometry() + numPrivateLoopClosures() + numSharedLoopClosures();
}

void PoseGraph::setMeasurements(const std::vector<RelativeSEMeasurement> &measurements) {
  empty();
  for (const auto &m : measurements) addMeasurement(m);
}

void PoseGraph::addMeasurement(const RelativeSEMeasurement &m) {
  if (m.r1 != id_ && m.r2 != id_) {
    std::fprintf(stderr, "[PoseGraph] ignoring an edge that does not touch robot %u\n", id_);
    return;
  }
  if (m.r1 == id_ && m.r2 == id_) {
    if (m.p1 + 1 == m.p2) addOdometry(m);
    else addPrivateLoopClosure(m);
  } else {
    addSharedLoopClosure(m);
  }
  clearDataMatrices();
}

void PoseGraph::addOdometry(const RelativeSEMeasurement &f) {
  const PoseID src(static_cast<unsigned>(f.r1), static_cast<unsigned>(f.p1));
  const PoseID dst(static_cast<unsigned>(f.r2), static_cast<unsigned>(f.p2));
  if (hasMeasurement(src, dst)) return;
  DPGO_CHECK(f.r1 == id_ && f.r2 == id_ && f.p1 + 1 == f.p2);
  DPGO_CHECK(f.R.rows() == d_ && f.R.cols() == d_ && f.t.rows() == d_ && f.t.cols() == 1);
  n_ = std::max(n_, static_cast<unsigned>(f.p2) + 1);
  odometry_.push_back(f);
  edge_id_to_index_.emplace(EdgeID(src, dst), odometry_.size() - 1);
}

void PoseGraph::addPrivateLoopClosure(const RelativeSEMeasurement &f) {
  const PoseID src(static_cast<unsigned>(f.r1), static_cast<unsigned>(f.p1));
  const PoseID dst(static_cast<unsigned>(f.r2), static_cast<unsigned>(f.p2));
  if (hasMeasurement(src, dst)) return;
  DPGO_CHECK(f.r1 == id_ && f.r2 == id_);
  DPGO_CHECK(f.R.rows() == d_ && f.R.cols() == d_ && f.t.rows() == d_ && f.t.cols() == 1);
  n_ = std::max(n_, static_cast<unsigned>(std::max(f.p1, f.p2)) + 1);
  private_lcs_.push_back(f);
  edge_id_to_index_.emplace(EdgeID(src, dst), private_lcs_.size() - 1);
}

void PoseGraph::addSharedLoopClosure(const RelativeSEMeasurement &f) {
  const PoseID src(static_cast<unsigned>(f.r1), static_cast<unsigned>(f.p1));
  const PoseID dst(static_cast<unsigned>(f.r2), static_cast<unsigned>(f.p2));
  if (hasMeasurement(src, dst)) return;
  DPGO_CHECK(f.R.rows() == d_ && f.R.cols() == d_ && f.t.rows() == d_ && f.t.cols() == 1);
  const bool outgoing = (f.r1 == id_);
  if (outgoing) DPGO_CHECK(f.r2 != id_);
  else DPGO_CHECK(f.r2 == id_);
  const PoseID &mine = outgoing ? src : dst;
  const PoseID &theirs = outgoing ? dst : src;
  n_ = std::max(n_, mine.frame_id + 1);
  local_shared_pose_ids_.insert(mine);
  nbr_shared_pose_ids_.insert(theirs);
  nbr_robot_ids_.insert(theirs.robot_id);
  neighbor_active_[theirs.robot_id] = true;
  shared_lcs_.push_back(f);
  edge_id_to_index_.emplace(EdgeID(src, dst), shared_lcs_.size() - 1);
}

std::vector<RelativeSEMeasurement> PoseGraph::sharedLoopClosuresWithRobot(unsigned int nb) const {
  std::vector<RelativeSEMeasurement> out;
  for (const auto &m : shared_lcs_)
    if (m.r1 == nb || m.r2 == nb) out.push_back(m);
  return out;
}

std::vector<RelativeSEMeasurement> PoseGraph::measurements() const {
  std::vector<RelativeSEMeasurement> out = localMeasurements();
  out.insert(out.end(), shared_lcs_.begin(), shared_lcs_.end());
  return out;
}

std::vector<RelativeSEMeasurement> PoseGraph::localMeasurements() const {
  std::vector<RelativeSEMeasurement> out = odometry_;
  out.insert(out.end(), private_lcs_.begin(), private_lcs_.end());
  return out;
}

void PoseGraph::clearPriors() {
  if (!priors_.empty()) clearDataMatrices();
  priors_.clear();
}

void PoseGraph::setPrior(unsigned index, const LiftedPose &Xi) {
  DPGO_CHECK(index < n());
  DPGO_CHECK(d() == Xi.d() && r() == Xi.r());
  priors_[index] = Xi;
  clearDataMatrices();
}

void PoseGraph::setNeighborPoses(const PoseDict &pose_dict) {
  neighbor_poses_ = pose_dict;
  g_valid_ = false;  // reference: setting neighbour poses only invalidates the linear term
}

bool PoseGraph::hasNeighbor(unsigned int robot_id) const { return nbr_robot_ids_.count(robot_id) > 0; }

bool PoseGraph::isNeighborActive(unsigned int nb) const {
  if (!hasNeighbor(nb)) return false;
  return neighbor_active_.at(nb);
}

void PoseGraph::setNeighborActive(unsigned int nb, bool active) {
  if (!hasNeighbor(nb)) return;
  if (neighbor_active_.at(nb) != active) clearDataMatrices();
  neighbor_active_[nb] = active;
}

bool PoseGraph::requireNeighborPose(const PoseID &pose_id) const { return nbr_shared_pose_ids_.count(pose_id) > 0; }

bool PoseGraph::hasMeasurement(const PoseID &src, const PoseID &dst) const {
  return edge_id_to_index_.count(EdgeID(src, dst)) > 0;
}

RelativeSEMeasurement *PoseGraph::findMeasurement(const PoseID &src, const PoseID &dst) {
  const EdgeID eid(src, dst);
  auto it = edge_id_to_index_.find(eid);
  if (it == edge_id_to_index_.end()) return nullptr;
  RelativeSEMeasurement *edge = eid.isOdometry() ? &odometry_[it->second]
                                : eid.isPrivateLoopClosure() ? &private_lcs_[it->second]
                                                             : &shared_lcs_[it->second];
  DPGO_CHECK(edge->r1 == src.robot_id && edge->p1 == src.frame_id && edge->r2 == dst.robot_id &&
             edge->p2 == dst.frame_id);
  return edge;
}

std::vector<RelativeSEMeasurement *> PoseGraph::allLoopClosures() {
  std::vector<RelativeSEMeasurement *> out;
  for (auto &m : private_lcs_) out.push_back(&m);
  for (auto &m : shared_lcs_) out.push_back(&m);
  return out;
}

std::vector<RelativeSEMeasurement *> PoseGraph::activeLoopClosures() {
  std::vector<RelativeSEMeasurement *> out;
  for (auto &m : private_lcs_) out.push_back(&m);
  for (auto &m : shared_lcs_) {
    const unsigned nb = static_cast<unsigned>(m.r1 == id_ ? m.r2 : m.r1);
    if (isNeighborActive(nb)) out.push_back(&m);
  }
  return out;
}

std::vector<RelativeSEMeasurement *> PoseGraph::inactiveLoopClosures() {
  std::vector<RelativeSEMeasurement *> out;
  for (auto &m : shared_lcs_) {
    const unsigned nb = static_cast<unsigned>(m.r1 == id_ ? m.r2 : m.r1);
    if (!isNeighborActive(nb)) out.push_back(&m);
  }
  return out;
}

PoseSet PoseGraph::activeNeighborPublicPoseIDs() const {
  PoseSet out;
  for (const auto &pid : nbr_shared_pose_ids_)
    if (isNeighborActive(pid.robot_id)) out.insert(pid);
  return out;
}

std::set<unsigned> PoseGraph::activeNeighborIDs() const {
  std::set<unsigned> out;
  for (unsigned nb : nbr_robot_ids_)
    if (isNeighborActive(nb)) out.insert(nb);
  return out;
}

size_t PoseGraph::numActiveNeighbors() const { return activeNeighborIDs().size(); }

PoseGraph::Statistics PoseGraph::statistics() const {
  Statistics s;
  auto count = [&s](const RelativeSEMeasurement &m) {
    if (m.weight == 1) s.accept_loop_closures += 1;
    else if (m.weight == 0) s.reject_loop_closures += 1;
    s.total_loop_closures += 1;
  };
  for (const auto &m : private_lcs_) count(m);
  for (const auto &m : shared_lcs_) {
    const unsigned nb = static_cast<unsigned>(m.r1 == id_ ? m.r2 : m.r1);
    if (isNeighborActive(nb)) count(m);
  }
  s.undecided_loop_closures = s.total_loop_closures - s.accept_loop_closures - s.reject_loop_closures;
  return s;
}

void PoseGraph::updatePublicPoseIDs() {
  local_shared_pose_ids_.clear();
  nbr_shared_pose_ids_.clear();
  for (const auto &m : shared_lcs_) {
    const bool outgoing = (m.r1 == id_);
    local_shared_pose_ids_.emplace(id_, static_cast<unsigned>(outgoing ? m.p1 : m.p2));
    nbr_shared_pose_ids_.emplace(static_cast<unsigned>(outgoing ? m.r2 : m.r1),
                                 static_cast<unsigned>(outgoing ? m.p2 : m.p1));
  }
}

void PoseGraph::useInactiveNeighbors(bool use) {
  use_inactive_neighbors_ = use;
  clearDataMatrices();
}

void PoseGraph::clearQuadraticMatrix() {
  q_valid_ = false;
  precon_valid_ = false;  // the preconditioner depends on Q
}
void PoseGraph::clearLinearMatrix() { g_valid_ = false; }
void PoseGraph::clearDataMatrices() {
  clearQuadraticMatrix();
  clearLinearMatrix();
}

// ---- device side ---------------------------------------------------------------------------
bool PoseGraph::ensureDevice() {
  if (n_ == 0) return false;
  if (dev_ && dev_n_ == n_) return true;
  if (dev_) {
    dpgo_destroy(dev_);
    dev_ = nullptr;
  }
  DPGO_DEVICE_CALL(dpgo_create(defaultDevice(), static_cast<int>(n_), static_cast<int>(d_), static_cast<int>(r_), nullptr, &dev_));
  dev_n_ = n_;
  dev_generation_++;
  q_valid_ = g_valid_ = precon_valid_ = false;
  return true;
}

dpgo_dev *PoseGraph::deviceHandle() { return ensureDevice() ? dev_ : nullptr; }

// Which shared edges enter Q / G, and the neighbour poses they need (slot order = PoseID order).
// Mirrors the active / inactive / missing-pose rules of src/PoseGraph.cpp:403-458.
bool PoseGraph::selectSharedEdges(std::vector<size_t> &edges, std::vector<PoseID> &slots) const {
  edges.clear();
  PoseSet needed;
  for (size_t k = 0; k < shared_lcs_.size(); ++k) {
    const auto &m = shared_lcs_[k];
    const bool outgoing = (m.r1 == id_);
    const PoseID nID(static_cast<unsigned>(outgoing ? m.r2 : m.r1), static_cast<unsigned>(outgoing ? m.p2 : m.p1));
    const bool has_pose = neighbor_poses_.count(nID) > 0;
    if (isNeighborActive(nID.robot_id)) {
      if (!has_pose) {
        std::fprintf(stderr, "[PoseGraph] Missing active neighbor pose %u, %u\n", nID.robot_id, nID.frame_id);
        return false;
      }
    } else if (!use_inactive_neighbors_ || !has_pose) {
      continue;
    }
    edges.push_back(k);
    needed.insert(nID);
  }
  slots.assign(needed.begin(), needed.end());
  return true;
}

bool PoseGraph::constructQ() {
  if (!ensureDevice()) return false;
  std::vector<size_t> edges;
  std::vector<PoseID> slots;
  if (!selectSharedEdges(edges, slots)) return false;
  const int d = static_cast<int>(d_);
  auto pack = [d](const std::vector<const RelativeSEMeasurement *> &ms, std::vector<double> &R, std::vector<double> &t,
                  std::vector<double> &kappa, std::vector<double> &tau, std::vector<double> &w) {
    for (const auto *m : ms) {
      for (int a = 0; a < d; ++a)
        for (int b = 0; b < d; ++b) R.push_back(m->R(a, b));  // row-major per edge
      for (int a = 0; a < d; ++a) t.push_back(m->t(a, 0));
      kappa.push_back(m->kappa);
      tau.push_back(m->tau);
      w.push_back(m->weight);
    }
  };
  UploadedGraph g;
  std::vector<double> wp, ws;
  {  // private edges: odometry + private loop closures
    std::vector<const RelativeSEMeasurement *> ms;
    for (const auto &m : odometry_) ms.push_back(&m);
    for (const auto &m : private_lcs_) ms.push_back(&m);
    for (const auto *m : ms) {
      g.p1.push_back(static_cast<int32_t>(m->p1));
      g.p2.push_back(static_cast<int32_t>(m->p2));
    }
    pack(ms, g.Rp, g.tp, g.kp, g.taup, wp);
  }
  {  // shared edges
    std::map<PoseID, int, ComparePoseID> slot_of;
    for (size_t s = 0; s < slots.size(); ++s) slot_of[slots[s]] = static_cast<int>(s);
    std::vector<const RelativeSEMeasurement *> ms;
    for (size_t k : edges) {
      const auto &m = shared_lcs_[k];
      const bool out = (m.r1 == id_);
      ms.push_back(&m);
      g.my.push_back(static_cast<int32_t>(out ? m.p1 : m.p2));
      g.slot.push_back(slot_of.at(PoseID(static_cast<unsigned>(out ? m.r2 : m.r1), static_cast<unsigned>(out ? m.p2 : m.p1))));
      g.outgoing.push_back(out ? 1 : 0);
    }
    pack(ms, g.Rs, g.ts, g.ks, g.taus, ws);
  }
  for (const auto &kv : priors_) {
    g.prior_idx.push_back(static_cast<int32_t>(kv.first));
    const Matrix P = kv.second.getData();
    g.prior_tiles.insert(g.prior_tiles.end(), P.data(), P.data() + P.size());
  }
  g.prior_kappa = prior_kappa_;
  g.prior_tau = prior_tau_;
  g.generation = dev_generation_;
  g.valid = true;
  bool precon_built = false;
  if (uploaded_.sameStructure(g) && slots.size() == q_slots_.size()) {
    // only the weights changed (GNC, ref: src/PGOAgent.cpp:1104-1142): re-weight Q, the cross blocks and -- when the
    // handle had one -- the preconditioner on the device
    precon_built = uploaded_.had_precon;
    const int rc = dpgo_update_weights(dev_, wp.data(), ws.data(), precon_built ? 1 : 0);
    if (rc == DPGO_ENUMERIC && precon_built) {
      std::fprintf(stderr, "[PoseGraph] Failed to compute preconditioner: %s\n", dpgo_last_error());
      precon_built = false;
      DPGO_DEVICE_CALL(dpgo_update_weights(dev_, wp.data(), ws.data(), 0));
    } else {
      DPGO_DEVICE_CALL(rc);
    }
    g.had_precon = uploaded_.had_precon;
  } else {
    DPGO_DEVICE_CALL(dpgo_set_private_edges(dev_, static_cast<int>(g.p1.size()), g.p1.data(), g.p2.data(), g.Rp.data(),
                                            g.tp.data(), g.kp.data(), g.taup.data(), wp.data()));
    DPGO_DEVICE_CALL(dpgo_set_shared_edges(dev_, static_cast<int>(g.my.size()), static_cast<int>(slots.size()),
                                           g.my.data(), g.slot.data(), g.outgoing.data(), g.Rs.data(), g.ts.data(),
                                           g.ks.data(), g.taus.data(), ws.data()));
    DPGO_DEVICE_CALL(dpgo_set_priors(dev_, static_cast<int>(g.prior_idx.size()), g.prior_idx.data(), g.prior_tiles.data(),
                                     prior_kappa_, prior_tau_));
    DPGO_DEVICE_CALL(dpgo_finalize(dev_, 0));
  }
  uploaded_ = std::move(g);
  q_edges_ = edges;
  q_slots_ = slots;
  q_valid_ = true;
  precon_valid_ = precon_built;
  g_valid_ = false;
  return true;
}

bool PoseGraph::constructG() {
  if (!q_valid_) return false;
  std::vector<size_t> edges;
  std::vector<PoseID> slots;
  if (!selectSharedEdges(edges, slots)) return false;
  if (edges != q_edges_ || slots.size() != q_slots_.size() || !std::equal(slots.begin(), slots.end(), q_slots_.begin())) {
    // the set of usable shared edges changed (inactive neighbour poses appeared / vanished)
    if (!constructQ()) return false;
  }
  const size_t tile = static_cast<size_t>(r_) * (d_ + 1);
  std::vector<double> buf(std::max<size_t>(q_slots_.size(), 1) * tile, 0.0);
  for (size_t s = 0; s < q_slots_.size(); ++s) {
    const Matrix P = neighbor_poses_.at(q_slots_[s]).getData();
    std::copy(P.data(), P.data() + tile, buf.begin() + static_cast<std::ptrdiff_t>(s * tile));
  }
  DPGO_DEVICE_CALL(dpgo_set_neighbor_poses(dev_, buf.data()));
  g_valid_ = true;
  return true;
}

bool PoseGraph::constructDataMatrices() {
  if (!q_valid_ && !constructQ()) return false;
  if (!g_valid_ && !constructG()) return false;
  return true;
}

bool PoseGraph::hasPreconditioner() {
  if (!q_valid_ && !constructQ()) return false;
  if (precon_valid_) return true;
  const int rc = dpgo_finalize(dev_, 1);
  if (rc == DPGO_ENUMERIC) {
    std::fprintf(stderr, "[PoseGraph] Failed to compute preconditioner: %s\n", dpgo_last_error());
    return false;
  }
  DPGO_DEVICE_CALL(rc);
  precon_valid_ = true;
  uploaded_.had_precon = true;
  g_valid_ = false;  // finalize() re-initialises the linear term
  return true;
}

Matrix PoseGraph::linearMatrix() {
  DPGO_CHECK(constructDataMatrices());
  Matrix G(r_, static_cast<std::ptrdiff_t>(d_ + 1) * n_);
  DPGO_DEVICE_CALL(dpgo_get_G(dev_, G.data()));
  return G;
}

void PoseGraph::quadraticMatrixBSR(std::vector<int> &rowptr, std::vector<int> &colidx, std::vector<double> &blocks) {
  DPGO_CHECK(q_valid_ || constructQ());
  int nnzb = 0;
  DPGO_DEVICE_CALL(dpgo_get_Q_bsr(dev_, &nnzb, nullptr, nullptr, nullptr));
  rowptr.assign(n_ + 1, 0);
  colidx.assign(static_cast<size_t>(nnzb), 0);
  blocks.assign(static_cast<size_t>(nnzb) * (d_ + 1) * (d_ + 1), 0.0);
  DPGO_DEVICE_CALL(dpgo_get_Q_bsr(dev_, &nnzb, rowptr.data(), colidx.data(), blocks.data()));
}

}  // namespace DPGO
