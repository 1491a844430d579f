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
  /// `directory` is prepended verbatim to every file name passed to the non-static members
  explicit PGOLogger(std::string directory);
  ~PGOLogger();

  // ---- writers (3-D only: nothing is written for d == 2)
  /// columns: pose_index,qx,qy,qz,qw,tx,ty,tz ; `trajectory` is d x (d+1)n
  void logTrajectory(unsigned d, unsigned n, const Matrix &trajectory, const std::string &file);
  /// columns: robot_src,pose_src,robot_dst,pose_dst,qx,qy,qz,qw,tx,ty,tz,kappa,tau,is_known_inlier,weight
  void logMeasurements(std::vector<RelativeSEMeasurement> &edges, const std::string &file);

  // ---- readers (quaternions are normalised on load)
  /// 3 x 4n matrix, empty when the file cannot be opened
  Matrix loadTrajectory(const std::string &file);
  /// `path` is used as given (NOT prefixed with the log directory), as in the reference
  static std::vector<RelativeSEMeasurement> loadMeasurements(const std::string &path, bool with_weights = false);

 private:
  std::string logDirectory;
};

}  // namespace DPGO
#endif
