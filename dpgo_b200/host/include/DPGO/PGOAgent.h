// RBCD agent -- drop-in for the reference's PGOAgent (include/DPGO/PGOAgent.h) over the B200
// hot path.  The iterate X and the Nesterov sequences Y, V, XPrev live in device slots of the
// agent's PoseGraph handle; iterate() runs the polar projections and the local solve as CUDA
// kernels (dpgo_nesterov_update_Y/V, dpgo_optimize_slot) and mirrors X (and Y) back to the host
// so that getX / getSharedPoseDict serve host PoseDicts exactly like the reference.
//
// In scope: everything examples/MultiRobotExample.cpp and the reference's agent tests call
// (construction, measurements, initialize, setX/getX, iterate with and without acceleration,
// periodic restart, shared / auxiliary pose dictionaries, neighbour poses, status, anchors,
// rounding, reset, the asynchronous optimization thread), the robust-optimization methods (GNC weight
// updates), the robust inter-robot frame alignment used when an agent is initialised from a
// neighbour, and CSV logging (PGOLogger).
#ifndef DPGO_B200_PGOAGENT_H
#define DPGO_B200_PGOAGENT_H

#include <DPGO/DPGO_robust.h>
#include <DPGO/PGOLogger.h>
#include <DPGO/DPGO_types.h>
#include <DPGO/DPGO_utils.h>
#include <DPGO/PoseGraph.h>
#include <DPGO/RelativeSEMeasurement.h>
#include <DPGO/manifold/Poses.h>

#include <atomic>
#include <memory>
#include <mutex>
#include <optional>
#include <string>
#include <thread>
#include <vector>

namespace DPGO {

/// parameters of an agent; same fields and defaults as the reference (PGOAgent.h:47-179)
class PGOAgentParameters {
 public:
  unsigned d, r, numRobots;
  bool asynchronous;
  double asynchronousOptimizationRate;
  ROptParameters localOptimizationParams;
  InitializationMethod localInitializationMethod;
  bool multirobotInitialization;
  bool acceleration;
  unsigned restartInterval;
  RobustCostParameters robustCostParams;
  int robustOptNumWeightUpdates, robustOptNumResets, robustOptInnerIters;
  double robustOptMinConvergenceRatio;
  unsigned robustInitMinInliers;
  unsigned maxNumIters;
  double relChangeTol;
  bool verbose, logData;
  std::string logDirectory;

  PGOAgentParameters(unsigned dIn, unsigned rIn, unsigned numRobotsIn = 1,
                     ROptParameters local_opt_params = ROptParameters(), bool accel = false,
                     unsigned restartInt = 30, RobustCostParameters costParams = RobustCostParameters(),
                     int robust_opt_num_weight_updates = 10, int robust_opt_num_resets = 0,
                     int robust_opt_inner_iters = 30, double robust_opt_min_convergence_ratio = 0.8,
                     unsigned robust_init_min_inliers = 2, unsigned maxIters = 500, double changeTol = 5e-3,
                     bool v = false, bool log = false, std::string logDir = "")
      : d(dIn), r(rIn), numRobots(numRobotsIn), asynchronous(false), asynchronousOptimizationRate(1),
        localOptimizationParams(local_opt_params), localInitializationMethod(InitializationMethod::Odometry),
        multirobotInitialization(true), acceleration(accel), restartInterval(restartInt),
        robustCostParams(costParams), robustOptNumWeightUpdates(robust_opt_num_weight_updates),
        robustOptNumResets(robust_opt_num_resets), robustOptInnerIters(robust_opt_inner_iters),
        robustOptMinConvergenceRatio(robust_opt_min_convergence_ratio),
        robustInitMinInliers(robust_init_min_inliers), maxNumIters(maxIters), relChangeTol(changeTol),
        verbose(v), logData(log), logDirectory(std::move(logDir)) {}
};

enum PGOAgentState { WAIT_FOR_DATA, WAIT_FOR_INITIALIZATION, INITIALIZED };

struct PGOAgentStatus {
  unsigned agentID;
  PGOAgentState state;
  unsigned instanceNumber, iterationNumber;
  bool readyToTerminate;
  double relativeChange;
  explicit PGOAgentStatus(unsigned id = 0, PGOAgentState s = PGOAgentState::WAIT_FOR_DATA, unsigned instance = 0,
                          unsigned iteration = 0, bool ready_to_terminate = false, double relative_change = 0)
      : agentID(id), state(s), instanceNumber(instance), iterationNumber(iteration),
        readyToTerminate(ready_to_terminate), relativeChange(relative_change) {}
};

class PGOAgent {
 public:
  PGOAgent(unsigned ID, const PGOAgentParameters &params);
  virtual ~PGOAgent();

  void addMeasurement(const RelativeSEMeasurement &factor);
  void setMeasurements(const std::vector<RelativeSEMeasurement> &inputOdometry,
                       const std::vector<RelativeSEMeasurement> &inputPrivateLoopClosures,
                       const std::vector<RelativeSEMeasurement> &inputSharedLoopClosures);
  void initialize(const PoseArray *TInitPtr = nullptr);
  void initializeInGlobalFrame(const Pose &T_world_robot);
  bool iterate(bool doOptimization = true);
  virtual void reset();

  unsigned getID() const { return mID; }
  unsigned num_poses() const { return mPoseGraph->n(); }
  unsigned dimension() const { return mPoseGraph->d(); }
  unsigned relaxation_rank() const { return mPoseGraph->r(); }
  unsigned instance_number() const { return mInstanceNumber; }
  unsigned iteration_number() const { return mIterationNumber; }
  PGOAgentParameters getParams() const { return mParams; }
  PGOAgentState getState() const { return mState; }

  PGOAgentStatus getStatus() {
    mStatus.agentID = getID();
    mStatus.state = mState;
    mStatus.instanceNumber = instance_number();
    mStatus.iterationNumber = iteration_number();
    return mStatus;
  }
  bool hasNeighborStatus(unsigned neighborID) const { return mTeamStatus.count(neighborID) > 0; }
  PGOAgentStatus getNeighborStatus(unsigned neighborID) const { return mTeamStatus.at(neighborID); }
  void setNeighborStatus(const PGOAgentStatus &status) { mTeamStatus[status.agentID] = status; }

  bool hasNeighbor(unsigned neighborID) const;
  std::vector<unsigned> getNeighbors() const;

  bool getTrajectoryInLocalFrame(Matrix &Trajectory);
  bool getTrajectoryInGlobalFrame(Matrix &Trajectory);
  bool getTrajectoryInGlobalFrame(PoseArray &Trajectory);
  bool getPoseInGlobalFrame(unsigned poseID, Matrix &T);
  /// a cached neighbour pose expressed in the global frame fixed by the anchor (reference :789-810)
  bool getNeighborPoseInGlobalFrame(unsigned neighborID, unsigned poseID, Matrix &T);
  bool getSharedPose(unsigned index, Matrix &Mout);
  bool getAuxSharedPose(unsigned index, Matrix &Mout);
  bool getSharedPoseDict(PoseDict &map);
  bool getSharedPoseDictWithNeighbor(PoseDict &map, unsigned neighborID);
  bool getAuxSharedPoseDict(PoseDict &map);
  bool getAuxSharedPoseDictWithNeighbor(PoseDict &map, unsigned neighborID);

  void setX(const Matrix &Xin);
  void setXToInitialGuess();
  bool getX(Matrix &Mout);

  bool shouldTerminate();
  bool shouldRestart() const;
  void restartNesterovAcceleration(bool doOptimization);

  void startOptimizationLoop();
  void endOptimizationLoop();
  bool isOptimizationRunning();

  bool getLiftingMatrix(Matrix &M) const;
  void setLiftingMatrix(const Matrix &M);
  void setGlobalAnchor(const Matrix &M);

  void updateNeighborPoses(unsigned neighborID, const PoseDict &poseDict);
  void updateAuxNeighborPoses(unsigned neighborID, const PoseDict &poseDict);
  void clearNeighborPoses();
  void clearActiveNeighborPoses();

  Matrix localPoseGraphOptimization();
  ROPTResult getLocalOptResult() const { return mLocalOptResult; }

 protected:
  unsigned mID, d, r;
  LiftedPoseArray X;  // host mirror of the device iterate
  PGOAgentParameters mParams;
  PGOAgentState mState;
  PGOAgentStatus mStatus;
  RobustCost mRobustCost;
  PGOLogger mLogger;
  std::shared_ptr<PoseGraph> mPoseGraph;
  unsigned mInstanceNumber, mIterationNumber;
  std::optional<Matrix> YLift;
  std::optional<LiftedPose> globalAnchor;
  std::optional<PoseArray> TLocalInit;
  std::optional<LiftedPoseArray> XInit;
  PoseDict neighborPoseDict, neighborAuxPoseDict;
  std::map<unsigned, PGOAgentStatus> mTeamStatus;
  std::vector<bool> mTeamRobotActive;
  ROPTResult mLocalOptResult;
  bool mPublishPublicPosesRequested = false, mPublishAsynchronousRequested = false;

  std::mutex mPosesMutex, mMeasurementsMutex, mNeighborPosesMutex;
  std::unique_ptr<std::thread> mOptimizationThread;
  std::atomic<bool> mEndLoopRequested{false};

  // Nesterov acceleration (reference: PGOAgent.h:734-768)
  double gamma, alpha;
  LiftedPoseArray Y;  // host mirror of the auxiliary sequence
  void initializeAcceleration();
  void updateGamma();
  void updateAlpha();
  void updateY();
  void updateV();
  bool updateX(bool doOptimization = false, bool acceleration = false);

  void runOptimizationLoop();
  Pose computeNeighborTransform(const RelativeSEMeasurement &measurement, const LiftedPose &neighbor_pose);
  /// robust alignment of my local frame to the world frame from all loop closures with one neighbour:
  /// GNC rotation averaging, then translation averaging on the inliers (reference :551-604) ...
  bool computeRobustNeighborTransformTwoStage(unsigned neighborID, const PoseDict &poseDict, Pose *T_world_robot);
  /// ... or joint GNC pose averaging (reference :606-648)
  bool computeRobustNeighborTransform(unsigned neighborID, const PoseDict &poseDict, Pose *T_world_robot);

  // robust optimization (GNC / M-estimators), reference: include/DPGO/PGOAgent.h:676-708,
  // src/PGOAgent.cpp:997-1155.  Like the reference these are driven by a subclass / the ROS layer.
  unsigned mWeightUpdateCount = 0, mTrajectoryResetCount = 0, mLatestWeightUpdateIteration = 0;
  int mRobustOptInnerIter = 0;
  void initializeRobustOptimization();
  bool shouldUpdateMeasurementWeights() const;
  void updateMeasurementWeights();
  bool computeMeasurementResidual(const RelativeSEMeasurement &measurement, double *residual) const;
  bool setMeasurementWeight(const PoseID &src_ID, const PoseID &dst_ID, double weight, bool fixed_weight = false);
  bool isRobotInitialized(unsigned robot_id) const;
  bool isRobotActive(unsigned robot_id) const;
  void setRobotActive(unsigned robot_id, bool active = true);
  size_t numActiveRobots() const;
  bool anchorFirstPose();
  bool anchorFirstPose(const LiftedPose &prior);

 private:
  // device-slot bookkeeping
  bool mDeviceStateValid = false;  // slots X / Y / V / XPREV hold the agent's sequences
  Matrix TLocalInitInGlobal_;      // initial trajectory in the global frame (logged when logData)
  void uploadState();              // host X -> device slots (after setX / re-initialisation)
  void downloadX();
  void downloadY();
  // host mirrors X / Y of the device iterates, fetched on demand (iterate() marks them stale); a subclass that
  // reads the protected members X / Y directly must go through these instead
  LiftedPoseArray &hostX();
  LiftedPoseArray &hostY();
  bool mHostXStale = false, mHostYStale = false;
};

}  // namespace DPGO
#endif
