// PGOAgent over the device hot path.  Control flow follows the reference's src/PGOAgent.cpp
// (cited per function); all per-pose arithmetic is delegated to CUDA kernels via the C-ABI.
#include <DPGO/DPGO_solver.h>
#include <DPGO/PGOAgent.h>
#include <DPGO/QuadraticOptimizer.h>
#include <DPGO/QuadraticProblem.h>

#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <random>

#include "check.h"

using std::lock_guard;
using std::mutex;

namespace DPGO {

PGOAgent::PGOAgent(unsigned ID, const PGOAgentParameters &params)
    : mID(ID), d(params.d), r(params.r), X(params.r, params.d, 1), mParams(params),
      mState(PGOAgentState::WAIT_FOR_DATA), mStatus(ID, PGOAgentState::WAIT_FOR_DATA, 0, 0, false, 0),
      mRobustCost(params.robustCostParams), mLogger(params.logDirectory), mPoseGraph(std::make_shared<PoseGraph>(ID, params.r, params.d)),
      mInstanceNumber(0), mIterationNumber(0), gamma(0), alpha(0), Y(params.r, params.d, 1) {
  if (mID == 0) setLiftingMatrix(fixedStiefelVariable(d, r));  // reference: src/PGOAgent.cpp:45
  mTeamRobotActive.assign(mParams.numRobots, true);
}

PGOAgent::~PGOAgent() { endOptimizationLoop(); }

// ---- device slot helpers ---------------------------------------------------------------------
void PGOAgent::uploadState() {
  dpgo_dev *h = mPoseGraph->deviceHandle();
  DPGO_CHECK(h != nullptr);
  const Matrix Xm = X.getData();
  DPGO_DEVICE_CALL(dpgo_slot_set(h, DPGO_SLOT_X, Xm.data()));
  mDeviceStateValid = true;
  mHostXStale = false;
}

void PGOAgent::downloadX() {
  Matrix M(r, static_cast<std::ptrdiff_t>(d + 1) * num_poses());
  DPGO_DEVICE_CALL(dpgo_slot_get(mPoseGraph->deviceHandle(), DPGO_SLOT_X, M.data()));
  X.setData(M);
  mHostXStale = false;
}

void PGOAgent::downloadY() {
  Matrix M(r, static_cast<std::ptrdiff_t>(d + 1) * num_poses());
  DPGO_DEVICE_CALL(dpgo_slot_get(mPoseGraph->deviceHandle(), DPGO_SLOT_Y, M.data()));
  Y.setData(M);
  mHostYStale = false;
}

// The iterates live on the device; iterate() only marks the host mirrors stale and the first reader fetches
// them (callers hold mPosesMutex).  A round in which nobody reads the poses on the host moves no pose data.
LiftedPoseArray &PGOAgent::hostX() {
  if (mHostXStale) downloadX();
  return X;
}

LiftedPoseArray &PGOAgent::hostY() {
  if (mHostYStale) downloadY();
  return Y;
}

// ---- state -------------------------------------------------------------------------------------
void PGOAgent::setX(const Matrix &Xin) {  // reference :52-63
  lock_guard<mutex> lock(mPosesMutex);
  DPGO_CHECK(mState != PGOAgentState::WAIT_FOR_DATA);
  DPGO_CHECK(static_cast<unsigned>(Xin.rows()) == relaxation_rank());
  DPGO_CHECK(static_cast<unsigned>(Xin.cols()) == (dimension() + 1) * num_poses());
  mState = PGOAgentState::INITIALIZED;
  X = LiftedPoseArray(relaxation_rank(), dimension(), num_poses());
  X.setData(Xin);
  uploadState();
  if (mParams.acceleration) initializeAcceleration();
}

void PGOAgent::setXToInitialGuess() {
  DPGO_CHECK(mState != PGOAgentState::WAIT_FOR_DATA);
  DPGO_CHECK(XInit.has_value());
  lock_guard<mutex> lock(mPosesMutex);
  X = XInit.value();
  uploadState();
}

bool PGOAgent::getX(Matrix &Mout) {
  lock_guard<mutex> lock(mPosesMutex);
  Mout = hostX().getData();
  return true;
}

bool PGOAgent::getSharedPose(unsigned index, Matrix &Mout) {
  if (mState != PGOAgentState::INITIALIZED) return false;
  lock_guard<mutex> lock(mPosesMutex);
  if (index >= num_poses()) return false;
  Mout = hostX().pose(index);
  return true;
}

bool PGOAgent::getAuxSharedPose(unsigned index, Matrix &Mout) {
  DPGO_CHECK(mParams.acceleration);
  if (mState != PGOAgentState::INITIALIZED) return false;
  lock_guard<mutex> lock(mPosesMutex);
  if (index >= num_poses()) return false;
  Mout = hostY().pose(index);
  return true;
}

bool PGOAgent::getSharedPoseDict(PoseDict &map) {  // reference :97-110
  if (mState != PGOAgentState::INITIALIZED) return false;
  map.clear();
  lock_guard<mutex> lock(mPosesMutex);
  for (const auto &pid : mPoseGraph->myPublicPoseIDs()) {
    DPGO_CHECK(pid.robot_id == getID());
    map.emplace(pid, LiftedPose(hostX().pose(pid.frame_id)));
  }
  return true;
}

bool PGOAgent::getSharedPoseDictWithNeighbor(PoseDict &map, unsigned neighborID) {  // :112-130
  if (mState != PGOAgentState::INITIALIZED) return false;
  map.clear();
  lock_guard<mutex> lock(mPosesMutex);
  for (const auto &m : mPoseGraph->sharedLoopClosuresWithRobot(neighborID)) {
    const PoseID pid(getID(), static_cast<unsigned>(m.r1 == getID() ? m.p1 : m.p2));
    map.emplace(pid, LiftedPose(hostX().pose(pid.frame_id)));
  }
  return true;
}

bool PGOAgent::getAuxSharedPoseDict(PoseDict &map) {  // :132-146
  DPGO_CHECK(mParams.acceleration);
  if (mState != PGOAgentState::INITIALIZED) return false;
  map.clear();
  lock_guard<mutex> lock(mPosesMutex);
  for (const auto &pid : mPoseGraph->myPublicPoseIDs()) {
    DPGO_CHECK(pid.robot_id == getID());
    map.emplace(pid, LiftedPose(hostY().pose(pid.frame_id)));
  }
  return true;
}

bool PGOAgent::getAuxSharedPoseDictWithNeighbor(PoseDict &map, unsigned neighborID) {
  DPGO_CHECK(mParams.acceleration);
  if (mState != PGOAgentState::INITIALIZED) return false;
  map.clear();
  lock_guard<mutex> lock(mPosesMutex);
  for (const auto &m : mPoseGraph->sharedLoopClosuresWithRobot(neighborID)) {
    const PoseID pid(getID(), static_cast<unsigned>(m.r1 == getID() ? m.p1 : m.p2));
    map.emplace(pid, LiftedPose(hostY().pose(pid.frame_id)));
  }
  return true;
}

void PGOAgent::setLiftingMatrix(const Matrix &M) {
  DPGO_CHECK(static_cast<unsigned>(M.rows()) == r && static_cast<unsigned>(M.cols()) == d);
  YLift.emplace(M);
}

bool PGOAgent::getLiftingMatrix(Matrix &M) const {
  if (!YLift.has_value()) return false;
  M = YLift.value();
  return true;
}

void PGOAgent::setGlobalAnchor(const Matrix &M) {
  DPGO_CHECK(static_cast<unsigned>(M.rows()) == relaxation_rank());
  DPGO_CHECK(static_cast<unsigned>(M.cols()) == dimension() + 1);
  LiftedPose Xa(r, d);
  Xa.pose() = M;
  globalAnchor.emplace(Xa);
}

// ---- measurements / initialisation ---------------------------------------------------------------
void PGOAgent::addMeasurement(const RelativeSEMeasurement &factor) {
  if (mState != PGOAgentState::WAIT_FOR_DATA) {
    std::fprintf(stderr, "[PGOAgent] Robot state is not WAIT_FOR_DATA. Ignore new measurements!\n");
    return;
  }
  lock_guard<mutex> lock(mMeasurementsMutex);
  mPoseGraph->addMeasurement(factor);
}

void PGOAgent::setMeasurements(const std::vector<RelativeSEMeasurement> &inputOdometry,
                               const std::vector<RelativeSEMeasurement> &inputPrivateLoopClosures,
                               const std::vector<RelativeSEMeasurement> &inputSharedLoopClosures) {  // :185-197
  DPGO_CHECK(!isOptimizationRunning());
  DPGO_CHECK(mState == PGOAgentState::WAIT_FOR_DATA);
  if (inputOdometry.empty()) return;
  mPoseGraph = std::make_shared<PoseGraph>(mID, r, d);
  mDeviceStateValid = false;
  std::vector<RelativeSEMeasurement> all = inputOdometry;
  all.insert(all.end(), inputPrivateLoopClosures.begin(), inputPrivateLoopClosures.end());
  all.insert(all.end(), inputSharedLoopClosures.begin(), inputSharedLoopClosures.end());
  mPoseGraph->setMeasurements(all);
}

void PGOAgent::initialize(const PoseArray *TInitPtr) {  // reference :199-306
  if (mState != PGOAgentState::WAIT_FOR_DATA) return;
  endOptimizationLoop();
  if (mPoseGraph->n() == 0) return;

  if (TInitPtr && TInitPtr->d() == dimension() && TInitPtr->n() == num_poses()) {
    TLocalInit.emplace(*TInitPtr);
  } else {
    PoseArray T(dimension(), num_poses());
    switch (mParams.localInitializationMethod) {
      case InitializationMethod::Odometry: T = odometryInitialization(mPoseGraph->odometry()); break;
      case InitializationMethod::Chordal: T = chordalInitialization(mPoseGraph->localMeasurements()); break;
      case InitializationMethod::GNC_TLS: {   // reference :233-262
        // robust single-robot solve from the odometry guess; local loop closures it rejects get weight 0
        solveRobustPGOParams rp;
        rp.verbose = mParams.verbose;
        rp.opt_params.verbose = false;
        rp.opt_params.gradnorm_tol = 1;
        rp.opt_params.RTR_iterations = 20;
        rp.robust_params.costType = RobustCostParameters::Type::GNC_TLS;
        rp.robust_params.GNCMaxNumIters = 10;
        rp.robust_params.GNCBarc = 5.0;
        rp.robust_params.GNCMuStep = 1.4;
        const PoseArray TOdom = odometryInitialization(mPoseGraph->odometry());
        std::vector<RelativeSEMeasurement> local = mPoseGraph->localMeasurements();
        T = solveRobustPGO(local, rp, &TOdom);
        int rejected = 0;
        for (const RelativeSEMeasurement &m : local)
          if (m.weight < 1e-8) {
            setMeasurementWeight(PoseID(m.r1, m.p1), PoseID(m.r2, m.p2), 0);
            rejected++;
          }
        if (mParams.verbose) std::printf("GNC_TLS initialization: rejected %d local loop closure(s)\n", rejected);
        break;
      }
    }
    DPGO_CHECK(T.d() == dimension());
    TLocalInit.emplace(T);
  }

  // express the local trajectory in the frame of its first pose
  PoseArray Tt(dimension(), num_poses());
  const Pose Tw0(TLocalInit.value().pose(0));
  const Pose T0w = Tw0.inverse();
  for (unsigned i = 0; i < num_poses(); ++i) Tt.pose(i) = (T0w * Pose(TLocalInit.value().pose(i))).pose();
  TLocalInit.emplace(Tt);

  X = LiftedPoseArray(relaxation_rank(), dimension(), num_poses());
  Y = X;
  mState = PGOAgentState::WAIT_FOR_INITIALIZATION;
  if (mID == 0 || !mParams.multirobotInitialization) initializeInGlobalFrame(Pose(d));
  if (mParams.asynchronous) startOptimizationLoop();
}

void PGOAgent::initializeInGlobalFrame(const Pose &T_world_robot) {  // reference :308-374
  DPGO_CHECK(YLift.has_value());
  DPGO_CHECK(T_world_robot.d() == dimension());
  checkRotationMatrix(T_world_robot.rotation());
  bool halted = false;
  if (isOptimizationRunning()) {
    halted = true;
    endOptimizationLoop();
  }
  {
    lock_guard<mutex> lock(mPosesMutex);
    clearNeighborPoses();
    PoseArray T = TLocalInit.value();
    for (unsigned i = 0; i < num_poses(); ++i) T.pose(i) = (T_world_robot * Pose(T.pose(i))).pose();
    TLocalInitInGlobal_ = T.getData();
    X.setData(YLift.value() * T.getData());
    XInit.emplace(X);
    mState = PGOAgentState::INITIALIZED;
    uploadState();
    if (mParams.acceleration) initializeAcceleration();
  }
  if (mParams.logData)  // reference :366-370
    mLogger.logTrajectory(dimension(), num_poses(), TLocalInitInGlobal_, "trajectory_initial.csv");
  // robust optimization starts from unit weights on every non-fixed loop closure (:348-352)
  if (mParams.robustCostParams.costType != RobustCostParameters::Type::L2) initializeRobustOptimization();
  if (halted) startOptimizationLoop();
}

// ---- the RBCD iteration --------------------------------------------------------------------------
bool PGOAgent::iterate(bool doOptimization) {  // reference :376-432
  mIterationNumber++;
  if (mParams.robustCostParams.costType != RobustCostParameters::Type::L2) mRobustOptInnerIter++;
  if (mState != PGOAgentState::INITIALIZED) return true;
  dpgo_dev *h = mPoseGraph->deviceHandle();
  if (!mDeviceStateValid) {
    lock_guard<mutex> lock(mPosesMutex);
    uploadState();
    if (mParams.acceleration) initializeAcceleration();
  }
  DPGO_DEVICE_CALL(dpgo_slot_copy(h, DPGO_SLOT_XPREV, DPGO_SLOT_X));  // XPrev = X
  bool success;
  if (mParams.acceleration) {
    updateGamma();
    updateAlpha();
    updateY();
    success = updateX(doOptimization, true);
    updateV();
    if (shouldRestart()) restartNesterovAcceleration(doOptimization);
  } else {
    success = updateX(doOptimization, false);
  }
  {
    lock_guard<mutex> lock(mPosesMutex);
    mHostXStale = true;
    if (mParams.acceleration) mHostYStale = true;
  }
  if (doOptimization) {
    mStatus.agentID = getID();
    mStatus.state = mState;
    mStatus.instanceNumber = instance_number();
    mStatus.iterationNumber = iteration_number();
    double change = 0;
    DPGO_DEVICE_CALL(dpgo_max_translation_distance(h, DPGO_SLOT_X, DPGO_SLOT_XPREV, &change));
    mStatus.relativeChange = change;
    // loose threshold during the first inner iterations of robust optimization (:410-415)
    double relative_change_tol = mParams.relChangeTol;
    if (mParams.robustCostParams.costType != RobustCostParameters::Type::L2 && mWeightUpdateCount == 0)
      relative_change_tol = 5;
    bool ready = success && !(change > relative_change_tol);
    const auto stat = mPoseGraph->statistics();
    if (stat.total_loop_closures > 0) {
      const double ratio = (stat.accept_loop_closures + stat.reject_loop_closures) / stat.total_loop_closures;
      if (ratio < mParams.robustOptMinConvergenceRatio) ready = false;
    }
    mStatus.readyToTerminate = ready;
  }
  if (doOptimization || mParams.acceleration) mPublishPublicPosesRequested = true;
  mPublishAsynchronousRequested = true;
  return success;
}

bool PGOAgent::updateX(bool doOptimization, bool acceleration) {  // reference :938-995
  std::unique_lock<mutex> tLock(mPosesMutex), mLock(mMeasurementsMutex), nLock(mNeighborPosesMutex);
  dpgo_dev *h = mPoseGraph->deviceHandle();
  if (!doOptimization) {
    if (acceleration) DPGO_DEVICE_CALL(dpgo_slot_copy(h, DPGO_SLOT_X, DPGO_SLOT_Y));  // X = Y
    return true;
  }
  if (acceleration) DPGO_CHECK(mParams.acceleration);
  DPGO_CHECK(mState == PGOAgentState::INITIALIZED);
  mPoseGraph->setNeighborPoses(acceleration ? neighborAuxPoseDict : neighborPoseDict);
  const ROptParameters &lp = mParams.localOptimizationParams;
  const bool rtr = lp.method == ROptParameters::ROptMethod::RTR;
  const bool want_precon = rtr || lp.RGD_use_preconditioner;
  if ((want_precon && !mPoseGraph->hasPreconditioner()) || !mPoseGraph->constructDataMatrices()) {
    std::fprintf(stderr, "[PGOAgent] Robot %u cannot construct data matrices... Skip optimization.\n", getID());
    mLocalOptResult = ROPTResult(false);
    return false;
  }
  dpgo_ropt_params prm;
  dpgo_default_params(&prm);
  prm.method = rtr ? 0 : 1;
  prm.verbose = mParams.verbose ? 1 : 0;
  prm.gradnorm_tol = lp.gradnorm_tol;
  prm.RGD_stepsize = lp.RGD_stepsize;
  prm.RGD_use_preconditioner = lp.RGD_use_preconditioner ? 1 : 0;
  prm.RTR_iterations = lp.RTR_iterations;
  prm.RTR_tCG_iterations = lp.RTR_tCG_iterations;
  prm.RTR_initial_radius = lp.RTR_initial_radius;
  dpgo_ropt_result res;
  DPGO_DEVICE_CALL(dpgo_optimize_slot(h, &prm, acceleration ? DPGO_SLOT_Y : DPGO_SLOT_X, &res));
  mLocalOptResult = ROPTResult(res.success != 0, res.f_init, res.gradnorm_init, res.f_opt, res.gradnorm_opt,
                               res.elapsed_ms);
  mLocalOptResult.tCGStatus = static_cast<tCGstatusSet>(res.tcg_status);
  if (mParams.verbose)
    std::printf("df: %f, init_gradnorm: %f, opt_gradnorm: %f. \n", res.f_init - res.f_opt, res.gradnorm_init,
                res.gradnorm_opt);
  return true;
}

// ---- Nesterov acceleration (reference :880-936) --------------------------------------------------
bool PGOAgent::shouldRestart() const {
  return mParams.acceleration && ((mIterationNumber + 1) % mParams.restartInterval == 0);
}

void PGOAgent::restartNesterovAcceleration(bool doOptimization) {
  if (!(mParams.acceleration && mState == PGOAgentState::INITIALIZED)) return;
  dpgo_dev *h = mPoseGraph->deviceHandle();
  DPGO_DEVICE_CALL(dpgo_slot_copy(h, DPGO_SLOT_X, DPGO_SLOT_XPREV));  // X = XPrev
  updateX(doOptimization, false);
  DPGO_DEVICE_CALL(dpgo_slot_copy(h, DPGO_SLOT_V, DPGO_SLOT_X));
  DPGO_DEVICE_CALL(dpgo_slot_copy(h, DPGO_SLOT_Y, DPGO_SLOT_X));
  gamma = 0;
  alpha = 0;
}

void PGOAgent::initializeAcceleration() {
  DPGO_CHECK(mParams.acceleration);
  if (mState != PGOAgentState::INITIALIZED) return;
  dpgo_dev *h = mPoseGraph->deviceHandle();
  DPGO_DEVICE_CALL(dpgo_slot_copy(h, DPGO_SLOT_XPREV, DPGO_SLOT_X));
  DPGO_DEVICE_CALL(dpgo_slot_copy(h, DPGO_SLOT_V, DPGO_SLOT_X));
  DPGO_DEVICE_CALL(dpgo_slot_copy(h, DPGO_SLOT_Y, DPGO_SLOT_X));
  if (mHostXStale) {
    mHostYStale = true;
  } else {
    Y = X;
    mHostYStale = false;
  }
  gamma = 0;
  alpha = 0;
}

void PGOAgent::updateGamma() {
  const double R = static_cast<double>(mParams.numRobots);
  gamma = (1 + std::sqrt(1 + 4 * R * R * gamma * gamma)) / (2 * R);
}
void PGOAgent::updateAlpha() { alpha = 1 / (gamma * static_cast<double>(mParams.numRobots)); }
void PGOAgent::updateY() { DPGO_DEVICE_CALL(dpgo_nesterov_update_Y(mPoseGraph->deviceHandle(), alpha)); }
void PGOAgent::updateV() { DPGO_DEVICE_CALL(dpgo_nesterov_update_V(mPoseGraph->deviceHandle(), gamma)); }

// ---- neighbours ----------------------------------------------------------------------------------
Pose PGOAgent::computeNeighborTransform(const RelativeSEMeasurement &m, const LiftedPose &nbr) {  // :515-560
  DPGO_CHECK(YLift.has_value());
  // T_world_robot = T_world_nbr * T_nbr_me (edge) * inverse(T_robot_me (local init))
  Pose dT(d);
  dT.rotation() = m.R;
  dT.translation() = m.t;
  Pose T_world_nbr(d);
  T_world_nbr.rotation() = projectToRotationGroup(YLift.value().transpose() * nbr.rotation());
  T_world_nbr.translation() = YLift.value().transpose() * nbr.translation();
  const bool outgoing = (m.r1 == getID());
  const Pose T_local(TLocalInit.value().pose(static_cast<unsigned>(outgoing ? m.p1 : m.p2)));
  const Pose T_world_me = outgoing ? T_world_nbr * dT.inverse() : T_world_nbr * dT;
  return T_world_me * T_local.inverse();
}

namespace {
// one candidate alignment per inter-robot loop closure whose neighbour pose is available
template <typename Transform>
void candidateAlignments(const std::vector<RelativeSEMeasurement> &edges, unsigned neighborID, const PoseDict &poseDict,
                         Transform transform, std::vector<Matrix> &RVec, std::vector<Vector> &tVec) {
  for (const RelativeSEMeasurement &m : edges) {
    const PoseID nbr(neighborID, static_cast<unsigned>(m.r1 == neighborID ? m.p1 : m.p2));
    const auto it = poseDict.find(nbr);
    if (it == poseDict.end()) continue;
    const Pose T = transform(m, it->second);
    RVec.emplace_back(T.rotation());
    tVec.emplace_back(T.translation());
  }
}
}  // namespace

bool PGOAgent::computeRobustNeighborTransformTwoStage(unsigned neighborID, const PoseDict &poseDict,
                                                      Pose *T_world_robot) {  // reference :551-604
  std::vector<Matrix> RVec;
  std::vector<Vector> tVec;
  candidateAlignments(mPoseGraph->sharedLoopClosuresWithRobot(neighborID), neighborID, poseDict,
                      [this](const RelativeSEMeasurement &m, const LiftedPose &p) { return computeNeighborTransform(m, p); },
                      RVec, tVec);
  if (RVec.empty()) return false;
  Matrix ROpt;
  Vector tOpt;
  std::vector<size_t> inliers;
  const double maxRotationError = angular2ChordalSO3(0.5);  // about 30 degrees
  robustSingleRotationAveraging(ROpt, inliers, RVec, Vector(), maxRotationError);
  std::printf("Robot %u attempts initialization from neighbor %u: finds %zu/%zu inliers.\n", getID(), neighborID,
              inliers.size(), RVec.size());
  if (inliers.size() < mParams.robustInitMinInliers) return false;
  std::vector<Vector> tIn;
  for (size_t i : inliers) tIn.push_back(tVec[i]);
  singleTranslationAveraging(tOpt, tIn);
  DPGO_CHECK(T_world_robot != nullptr && T_world_robot->d() == dimension());
  T_world_robot->rotation() = ROpt;
  T_world_robot->translation() = tOpt;
  return true;
}

bool PGOAgent::computeRobustNeighborTransform(unsigned neighborID, const PoseDict &poseDict,
                                              Pose *T_world_robot) {  // reference :606-648
  std::vector<Matrix> RVec;
  std::vector<Vector> tVec;
  candidateAlignments(mPoseGraph->sharedLoopClosuresWithRobot(neighborID), neighborID, poseDict,
                      [this](const RelativeSEMeasurement &m, const LiftedPose &p) { return computeNeighborTransform(m, p); },
                      RVec, tVec);
  if (RVec.empty()) return false;
  const std::ptrdiff_t m = static_cast<std::ptrdiff_t>(RVec.size());
  Vector kappa(m, 1), tau(m, 1);
  for (std::ptrdiff_t i = 0; i < m; ++i) {
    kappa(i) = 1.82;  // rotation stddev of about 30 degrees
    tau(i) = 0.01;    // translation stddev of 10 m
  }
  const double cbar = RobustCost::computeErrorThresholdAtQuantile(0.9, 3);
  Matrix ROpt;
  Vector tOpt;
  std::vector<size_t> inliers;
  robustSinglePoseAveraging(ROpt, tOpt, inliers, RVec, tVec, kappa, tau, cbar);
  std::printf("Robot %u attempts initialization from neighbor %u: finds %zu/%zu inliers.\n", getID(), neighborID,
              inliers.size(), RVec.size());
  if (inliers.size() < mParams.robustInitMinInliers) return false;
  DPGO_CHECK(T_world_robot != nullptr && T_world_robot->d() == dimension());
  T_world_robot->rotation() = ROpt;
  T_world_robot->translation() = tOpt;
  return true;
}

void PGOAgent::updateNeighborPoses(unsigned neighborID, const PoseDict &poseDict) {  // :650-678
  DPGO_CHECK(neighborID != mID);
  if (!YLift) return;
  if (!hasNeighborStatus(neighborID)) return;
  if (getNeighborStatus(neighborID).state != PGOAgentState::INITIALIZED) return;
  if (mState == PGOAgentState::WAIT_FOR_INITIALIZATION) {
    Pose T_world_robot(dimension());
    if (computeRobustNeighborTransformTwoStage(neighborID, poseDict, &T_world_robot))
      initializeInGlobalFrame(T_world_robot);
  }
  if (mState != PGOAgentState::INITIALIZED) return;
  lock_guard<mutex> lock(mNeighborPosesMutex);
  for (const auto &kv : poseDict) {
    DPGO_CHECK(kv.first.robot_id == neighborID);
    DPGO_CHECK(kv.second.r() == r && kv.second.d() == d);
    if (!mPoseGraph->requireNeighborPose(kv.first)) continue;
    neighborPoseDict[kv.first] = kv.second;
  }
}

void PGOAgent::updateAuxNeighborPoses(unsigned neighborID, const PoseDict &poseDict) {  // :680-702
  DPGO_CHECK(mParams.acceleration);
  DPGO_CHECK(neighborID != mID);
  if (!YLift) return;
  if (!hasNeighborStatus(neighborID)) return;
  if (getNeighborStatus(neighborID).state != PGOAgentState::INITIALIZED) return;
  if (mState != PGOAgentState::INITIALIZED) return;
  lock_guard<mutex> lock(mNeighborPosesMutex);
  for (const auto &kv : poseDict) {
    DPGO_CHECK(kv.first.robot_id == neighborID);
    DPGO_CHECK(kv.second.r() == r && kv.second.d() == d);
    if (!mPoseGraph->requireNeighborPose(kv.first)) continue;
    neighborAuxPoseDict[kv.first] = kv.second;
  }
}

void PGOAgent::clearNeighborPoses() {
  lock_guard<mutex> lock(mNeighborPosesMutex);
  neighborPoseDict.clear();
  neighborAuxPoseDict.clear();
}

void PGOAgent::clearActiveNeighborPoses() {
  lock_guard<mutex> lock(mNeighborPosesMutex);
  for (const auto &pid : mPoseGraph->activeNeighborPublicPoseIDs()) {
    neighborPoseDict.erase(pid);
    neighborAuxPoseDict.erase(pid);
  }
}

bool PGOAgent::hasNeighbor(unsigned neighborID) const { return mPoseGraph->hasNeighbor(neighborID); }

std::vector<unsigned> PGOAgent::getNeighbors() const {
  const auto ids = mPoseGraph->neighborIDs();
  return std::vector<unsigned>(ids.begin(), ids.end());
}

bool PGOAgent::isRobotActive(unsigned robot_id) const {
  return robot_id < mTeamRobotActive.size() && mTeamRobotActive[robot_id];
}

// ---- rounding (reference :718-767) ------------------------------------------------------------------
// Rounding runs on the device (dpgo_round_trajectory: one thread per pose, anchor rotation applied and the d x d
// block projected to SO(d) in registers); only the d x (d+1)n result crosses to the host.
bool PGOAgent::getTrajectoryInLocalFrame(Matrix &Trajectory) {
  if (mState != PGOAgentState::INITIALIZED) return false;
  lock_guard<mutex> lock(mPosesMutex);
  if (!mDeviceStateValid) uploadState();
  Matrix T(d, static_cast<std::ptrdiff_t>(d + 1) * num_poses());
  DPGO_DEVICE_CALL(dpgo_round_trajectory(mPoseGraph->deviceHandle(), DPGO_SLOT_X, nullptr, T.data()));
  Trajectory = T;
  return true;
}

bool PGOAgent::getTrajectoryInGlobalFrame(PoseArray &Trajectory) {
  if (!globalAnchor) return false;
  const LiftedPose Xa = globalAnchor.value();
  DPGO_CHECK(Xa.r() == relaxation_rank() && Xa.d() == dimension());
  if (mState != PGOAgentState::INITIALIZED) return false;
  lock_guard<mutex> lock(mPosesMutex);
  if (!mDeviceStateValid) uploadState();
  Matrix T(d, static_cast<std::ptrdiff_t>(d + 1) * num_poses());
  const Matrix anchor = Xa.getData();      // r x (d+1): rotation block and translation of the anchor
  DPGO_DEVICE_CALL(dpgo_round_trajectory(mPoseGraph->deviceHandle(), DPGO_SLOT_X, anchor.data(), T.data()));
  PoseArray out(d, num_poses());
  out.setData(T);
  Trajectory = out;
  return true;
}

bool PGOAgent::getTrajectoryInGlobalFrame(Matrix &Trajectory) {
  PoseArray T(d, num_poses());
  if (!getTrajectoryInGlobalFrame(T)) return false;
  Trajectory = T.getData();
  return true;
}

bool PGOAgent::getPoseInGlobalFrame(unsigned poseID, Matrix &T) {
  if (!globalAnchor) return false;
  const LiftedPose Xa = globalAnchor.value();
  if (mState != PGOAgentState::INITIALIZED) return false;
  lock_guard<mutex> lock(mPosesMutex);
  if (poseID >= num_poses()) return false;
  const Matrix Ya = Xa.rotation();
  const Matrix t0 = Ya.transpose() * Xa.translation();
  Matrix Ti = Ya.transpose() * hostX().pose(poseID);
  Ti.block(0, d, d, 1) -= t0;
  T = Ti;
  return true;
}

bool PGOAgent::getNeighborPoseInGlobalFrame(unsigned neighborID, unsigned poseID, Matrix &T) {
  if (!globalAnchor) return false;
  const LiftedPose Xa = globalAnchor.value();
  if (mState != PGOAgentState::INITIALIZED) return false;
  lock_guard<mutex> lock(mNeighborPosesMutex);
  const auto it = neighborPoseDict.find(PoseID(neighborID, poseID));
  if (it == neighborPoseDict.end()) return false;
  const Matrix Ya = Xa.rotation();
  const Matrix t0 = Ya.transpose() * Xa.translation();
  Matrix Ti = Ya.transpose() * it->second.pose();   // d x (d+1): rotation | translation
  Ti.block(0, d, d, 1) -= t0;
  T = Ti;
  return true;
}

Matrix PGOAgent::localPoseGraphOptimization() {  // reference :823-828
  ROptParameters pgo_params;
  pgo_params.verbose = true;
  return solvePGO(mPoseGraph->localMeasurements(), pgo_params).getData();
}

// ---- termination / reset -----------------------------------------------------------------------------
bool PGOAgent::shouldTerminate() {  // reference :844-878
  if (iteration_number() >= mParams.maxNumIters) return true;
  // not before the weights have been updated often enough (:853-857)
  if (mParams.robustCostParams.costType != RobustCostParameters::Type::L2 &&
      mWeightUpdateCount < static_cast<unsigned>(mParams.robustOptNumWeightUpdates))
    return false;
  for (unsigned robot_id = 0; robot_id < mParams.numRobots; ++robot_id) {
    if (!isRobotActive(robot_id)) continue;
    const auto it = mTeamStatus.find(robot_id);
    if (it == mTeamStatus.end()) return false;
    if (it->second.state != PGOAgentState::INITIALIZED) return false;
    if (!it->second.readyToTerminate) return false;
  }
  return true;
}

// End of one optimization instance (reference :434-473): stop the worker thread, leave the logs of the finished
// instance, then return to WAIT_FOR_DATA under the next instance number with nothing carried over.
void PGOAgent::reset() {
  endOptimizationLoop();
  if (mParams.logData) {
    const std::string &dir = mParams.logDirectory;
    std::vector<RelativeSEMeasurement> weighted = mPoseGraph->measurements();   // with their final GNC weights
    mLogger.logMeasurements(weighted, "measurements.csv");
    Matrix rounded;
    if (getTrajectoryInGlobalFrame(rounded)) {
      mLogger.logTrajectory(dimension(), num_poses(), rounded, "trajectory_optimized.csv");
      std::cout << "Saved optimized trajectory to " << dir << std::endl;
    }
    writeMatrixToFile(hostX().getData(), dir + "X.txt");
  }

  // what the agent knew about this instance: frames, initial guesses, neighbours, measurements, device state
  globalAnchor.reset();
  XInit.reset();
  TLocalInit.reset();
  clearNeighborPoses();
  mPoseGraph->reset();
  mDeviceStateValid = false;
  mTeamStatus.clear();
  mTeamRobotActive.assign(mParams.numRobots, false);
  mPublishPublicPosesRequested = mPublishAsynchronousRequested = false;

  // counters of the new instance
  ++mInstanceNumber;
  mIterationNumber = mLatestWeightUpdateIteration = 0;
  mRobustOptInnerIter = mWeightUpdateCount = mTrajectoryResetCount = 0;
  mState = PGOAgentState::WAIT_FOR_DATA;
  mStatus = PGOAgentStatus(getID(), mState, mInstanceNumber, mIterationNumber, false, 0);
}

// ---- asynchronous mode (reference :475-513) ------------------------------------------------------------
void PGOAgent::startOptimizationLoop() {
  DPGO_CHECK(!mParams.acceleration);  // "Asynchronous mode does not support acceleration!"
  if (isOptimizationRunning()) return;
  mOptimizationThread = std::make_unique<std::thread>(&PGOAgent::runOptimizationLoop, this);
}

void PGOAgent::runOptimizationLoop() {
  std::random_device rd;
  std::mt19937 rng(rd());
  std::exponential_distribution<double> wait(mParams.asynchronousOptimizationRate);
  while (true) {
    iterate(true);
    usleep(static_cast<useconds_t>(1e6 * wait(rng)));
    if (mEndLoopRequested) break;
  }
}

void PGOAgent::endOptimizationLoop() {
  if (!isOptimizationRunning()) return;
  mEndLoopRequested = true;
  mOptimizationThread->join();
  mOptimizationThread.reset(nullptr);
  mEndLoopRequested = false;
}

bool PGOAgent::isOptimizationRunning() { return mOptimizationThread != nullptr; }

// ---- robust optimization (reference :997-1215) --------------------------------------------------------
bool PGOAgent::shouldUpdateMeasurementWeights() const {  // :997-1046
  if (mParams.robustCostParams.costType == RobustCostParameters::Type::L2) return false;
  if (mWeightUpdateCount >= static_cast<unsigned>(mParams.robustOptNumWeightUpdates)) return false;
  if (mRobustOptInnerIter >= mParams.robustOptInnerIters) return true;
  for (unsigned robot_id = 0; robot_id < mParams.numRobots; ++robot_id) {
    if (!isRobotActive(robot_id)) continue;
    const auto it = mTeamStatus.find(robot_id);
    if (it == mTeamStatus.end()) return false;
    const PGOAgentStatus &st = it->second;
    DPGO_CHECK(st.agentID == robot_id);
    if (st.iterationNumber < mLatestWeightUpdateIteration) return false;  // outdated status
    if (st.state != PGOAgentState::INITIALIZED) return false;
    if (!st.readyToTerminate) return false;
  }
  return true;
}

void PGOAgent::initializeRobustOptimization() {  // :1048-1060
  mRobustCost.reset();
  lock_guard<mutex> lock(mMeasurementsMutex);
  for (RelativeSEMeasurement *m : mPoseGraph->activeLoopClosures())
    if (!m->fixedWeight) m->weight = 1.0;
  mPoseGraph->weightsChanged();
}

bool PGOAgent::computeMeasurementResidual(const RelativeSEMeasurement &m, double *residual) const {  // :1062-1102
  if (mState != PGOAgentState::INITIALIZED) return false;
  DPGO_CHECK(residual != nullptr);
  // host mirror of the device iterate, fetched if iterate() ran since the last read
  const LiftedPoseArray &Xh = const_cast<PGOAgent *>(this)->hostX();
  Matrix Y1, p1, Y2, p2;
  if (m.r1 == m.r2) {
    Y1 = Xh.rotation(m.p1); p1 = Xh.translation(m.p1);
    Y2 = Xh.rotation(m.p2); p2 = Xh.translation(m.p2);
  } else if (m.r1 == getID()) {
    Y1 = Xh.rotation(m.p1); p1 = Xh.translation(m.p1);
    const auto it = neighborPoseDict.find(PoseID(m.r2, m.p2));
    if (it == neighborPoseDict.end()) return false;
    Y2 = it->second.rotation(); p2 = it->second.translation();
  } else {
    Y2 = Xh.rotation(m.p2); p2 = Xh.translation(m.p2);
    const auto it = neighborPoseDict.find(PoseID(m.r1, m.p1));
    if (it == neighborPoseDict.end()) return false;
    Y1 = it->second.rotation(); p1 = it->second.translation();
  }
  *residual = std::sqrt(computeMeasurementError(m, Y1, p1, Y2, p2));
  return true;
}

void PGOAgent::updateMeasurementWeights() {  // :1104-1142
  if (mState != PGOAgentState::INITIALIZED) return;
  {
    lock_guard<mutex> lock(mMeasurementsMutex);
    double residual = 0;
    for (RelativeSEMeasurement *m : mPoseGraph->activeLoopClosures()) {
      if (m->fixedWeight) continue;
      if (computeMeasurementResidual(*m, &residual)) m->weight = mRobustCost.weight(residual);
    }
  }
  mWeightUpdateCount++;
  mLatestWeightUpdateIteration = iteration_number();
  mRobustOptInnerIter = 0;
  mPoseGraph->clearDataMatrices();  // Q and the preconditioner are rebuilt on the device at the next solve
  mRobustCost.update();
  mTeamStatus.clear();
  mStatus.readyToTerminate = false;
  mStatus.relativeChange = 0;
  if (mTrajectoryResetCount < static_cast<unsigned>(mParams.robustOptNumResets)) {
    mTrajectoryResetCount++;
    setXToInitialGuess();
    clearNeighborPoses();
  }
  if (mParams.acceleration) initializeAcceleration();
}

bool PGOAgent::setMeasurementWeight(const PoseID &src_ID, const PoseID &dst_ID, double weight, bool fixed_weight) {
  RelativeSEMeasurement *m = mPoseGraph->findMeasurement(src_ID, dst_ID);
  if (!m) return false;
  lock_guard<mutex> lock(mMeasurementsMutex);
  m->weight = weight;
  m->fixedWeight = fixed_weight;
  mPoseGraph->weightsChanged();
  return true;
}

bool PGOAgent::isRobotInitialized(unsigned robot_id) const {
  if (robot_id == getID()) return mState == PGOAgentState::INITIALIZED;
  if (!hasNeighborStatus(robot_id)) return false;
  return getNeighborStatus(robot_id).state == PGOAgentState::INITIALIZED;
}

void PGOAgent::setRobotActive(unsigned robot_id, bool active) {
  if (robot_id >= mParams.numRobots) return;
  mTeamRobotActive[robot_id] = active;
  if (mPoseGraph->hasNeighbor(robot_id)) mPoseGraph->setNeighborActive(robot_id, active);
}

size_t PGOAgent::numActiveRobots() const {
  size_t num_active = 0;
  for (unsigned robot_id = 0; robot_id < mParams.numRobots; ++robot_id)
    if (isRobotActive(robot_id)) num_active++;
  return num_active;
}

bool PGOAgent::anchorFirstPose() {
  if (num_poses() == 0) return false;
  LiftedPose prior(relaxation_rank(), dimension());
  prior.setData(hostX().pose(0));
  mPoseGraph->setPrior(0, prior);
  return true;
}

bool PGOAgent::anchorFirstPose(const LiftedPose &prior) {
  DPGO_CHECK(prior.d() == dimension());
  DPGO_CHECK(prior.r() == relaxation_rank());
  mPoseGraph->setPrior(0, prior);
  return true;
}

}  // namespace DPGO
