// CSV logging of trajectories and measurements, file-format compatible with the reference's
// PGOLogger (include/DPGO/PGOLogger.h, src/PGOLogger.cpp): same header rows, column order,
// quaternion convention (x, y, z, w) and default stream formatting, so logs written by either
// implementation are read by the other (and by dpgo_ros tooling).  3-D data only, like the
// reference.  Pure host code, no CUDA.
#ifndef DPGO_B200_PGOLOGGER_H
#define DPGO_B200_PGOLOGGER_H

#include <DPGO/DPGO_types.h>
#include <DPGO/RelativeSEMeasurement.h>

#include <string>
#include <vector>

namespace DPGO {

class PGOLogger {
 public:
  /// logDir is prepended verbatim to every file name (reference: "logDirectory + filename")
  explicit PGOLogger(std::string logDir);
  ~PGOLogger();

  /// pose_index,qx,qy,qz,qw,tx,ty,tz -- T is d x (d+1)n; nothing is written for d == 2
  void logTrajectory(unsigned d, unsigned n, const Matrix &T, const std::string &filename);
  /// robot_src,pose_src,robot_dst,pose_dst,qx,qy,qz,qw,tx,ty,tz,kappa,tau,is_known_inlier,weight
  void logMeasurements(std::vector<RelativeSEMeasurement> &measurements, const std::string &filename);
  /// 3 x 4n matrix (empty when the file cannot be opened); quaternions are normalised on load
  Matrix loadTrajectory(const std::string &filename);
  /// `filename` is used as given (not prefixed), as in the reference
  static std::vector<RelativeSEMeasurement> loadMeasurements(const std::string &filename, bool load_weight = false);

 private:
  std::string logDirectory;
};

}  // namespace DPGO
#endif
