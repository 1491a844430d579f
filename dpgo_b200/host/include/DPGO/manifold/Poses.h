// Host-side pose containers of the public API (reference: include/DPGO/manifold/Poses.h).  The
// column-major r x (d+1)n array they wrap is exactly the layout the CUDA kernels consume.
#ifndef DPGO_B200_POSES_H
#define DPGO_B200_POSES_H

#include <DPGO/DPGO_types.h>

#include <map>
#include <set>

namespace DPGO {

/// n lifted poses X_i = [Y_i p_i], Y_i in St(d, r), stored as one r x (d+1)n matrix
class LiftedPoseArray {
 public:
  LiftedPoseArray(unsigned int r, unsigned int d, unsigned int n);
  unsigned int r() const { return r_; }
  unsigned int d() const { return d_; }
  unsigned int n() const { return n_; }
  Matrix getData() const { return X_; }
  void setData(const Matrix &X);
  void checkData() const;
  const double *raw() const { return X_.data(); }
  double *raw() { return X_.data(); }

  BlockRef pose(unsigned int index);
  Matrix pose(unsigned int index) const;
  BlockRef rotation(unsigned int index);
  Matrix rotation(unsigned int index) const;
  BlockRef translation(unsigned int index);
  Matrix translation(unsigned int index) const;

  static double averageTranslationDistance(const LiftedPoseArray &a, const LiftedPoseArray &b);
  static double maxTranslationDistance(const LiftedPoseArray &a, const LiftedPoseArray &b);

 protected:
  unsigned int r_, d_, n_;
  Matrix X_;
};

/// n poses in SE(d): the special case r = d
class PoseArray : public LiftedPoseArray {
 public:
  PoseArray(unsigned int d, unsigned int n) : LiftedPoseArray(d, d, n) {}
};

/// a single lifted pose
class LiftedPose : public LiftedPoseArray {
 public:
  LiftedPose() : LiftedPose(3, 3) {}
  LiftedPose(unsigned int r, unsigned int d) : LiftedPoseArray(r, d, 1) {}
  explicit LiftedPose(const Matrix &X) : LiftedPose(static_cast<unsigned>(X.rows()), static_cast<unsigned>(X.cols() - 1)) {
    setData(X);
  }
  BlockRef pose() { return LiftedPoseArray::pose(0); }
  Matrix pose() const { return LiftedPoseArray::pose(0); }
  BlockRef rotation() { return LiftedPoseArray::rotation(0); }
  Matrix rotation() const { return LiftedPoseArray::rotation(0); }
  BlockRef translation() { return LiftedPoseArray::translation(0); }
  Matrix translation() const { return LiftedPoseArray::translation(0); }
};

/// a single pose in SE(d)
class Pose : public LiftedPose {
 public:
  Pose() : Pose(3) {}
  explicit Pose(unsigned int d) : LiftedPose(d, d) {}
  explicit Pose(const Matrix &T);
  static Pose Identity(unsigned int d);
  Pose identity() const { return Identity(d()); }
  Pose inverse() const;
  Pose operator*(const Pose &other) const;
  Matrix matrix() const;  ///< homogeneous (d+1) x (d+1)
};

typedef std::map<PoseID, LiftedPose, ComparePoseID> PoseDict;
typedef std::set<PoseID, ComparePoseID> PoseSet;

}  // namespace DPGO
#endif
