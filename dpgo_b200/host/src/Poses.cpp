#include <DPGO/manifold/Poses.h>

#include <cstdio>
#include <cstdlib>

#include "check.h"

namespace DPGO {

LiftedPoseArray::LiftedPoseArray(unsigned int r, unsigned int d, unsigned int n) : r_(r), d_(d), n_(n) {
  X_ = Matrix::Zero(r, static_cast<std::ptrdiff_t>(d + 1) * n);
  for (unsigned i = 0; i < n; ++i)
    for (unsigned k = 0; k < d && k < r; ++k) X_(k, static_cast<std::ptrdiff_t>(i) * (d + 1) + k) = 1.0;
}

void LiftedPoseArray::setData(const Matrix &X) {
  DPGO_CHECK(X.rows() == static_cast<std::ptrdiff_t>(r_));
  DPGO_CHECK(X.cols() == static_cast<std::ptrdiff_t>(d_ + 1) * n_);
  X_ = X;
}

void LiftedPoseArray::checkData() const {
  for (unsigned i = 0; i < n_; ++i) {
    const Matrix Y = rotation(i);
    const double err = (Y.transpose() * Y - Matrix::Identity(d_, d_)).norm();
    if (err > 1e-5) std::fprintf(stderr, "[LiftedPoseArray] pose %u is off the Stiefel manifold by %g\n", i, err);
  }
}

BlockRef LiftedPoseArray::pose(unsigned int i) {
  DPGO_CHECK(i < n_);
  return X_.block(0, static_cast<std::ptrdiff_t>(i) * (d_ + 1), r_, d_ + 1);
}
Matrix LiftedPoseArray::pose(unsigned int i) const {
  DPGO_CHECK(i < n_);
  return X_.block(0, static_cast<std::ptrdiff_t>(i) * (d_ + 1), r_, d_ + 1);
}
BlockRef LiftedPoseArray::rotation(unsigned int i) {
  DPGO_CHECK(i < n_);
  return X_.block(0, static_cast<std::ptrdiff_t>(i) * (d_ + 1), r_, d_);
}
Matrix LiftedPoseArray::rotation(unsigned int i) const {
  DPGO_CHECK(i < n_);
  return X_.block(0, static_cast<std::ptrdiff_t>(i) * (d_ + 1), r_, d_);
}
BlockRef LiftedPoseArray::translation(unsigned int i) {
  DPGO_CHECK(i < n_);
  return X_.block(0, static_cast<std::ptrdiff_t>(i) * (d_ + 1) + d_, r_, 1);
}
Matrix LiftedPoseArray::translation(unsigned int i) const {
  DPGO_CHECK(i < n_);
  return X_.block(0, static_cast<std::ptrdiff_t>(i) * (d_ + 1) + d_, r_, 1);
}

double LiftedPoseArray::averageTranslationDistance(const LiftedPoseArray &a, const LiftedPoseArray &b) {
  DPGO_CHECK(a.r() == b.r() && a.d() == b.d() && a.n() == b.n());
  double s = 0;
  for (unsigned i = 0; i < a.n(); ++i) s += (a.translation(i) - b.translation(i)).norm();
  return s / static_cast<double>(a.n());
}

double LiftedPoseArray::maxTranslationDistance(const LiftedPoseArray &a, const LiftedPoseArray &b) {
  DPGO_CHECK(a.r() == b.r() && a.d() == b.d() && a.n() == b.n());
  double m = 0;
  for (unsigned i = 0; i < a.n(); ++i) {
    const double v = (a.translation(i) - b.translation(i)).norm();
    if (v > m) m = v;
  }
  return m;
}

Pose::Pose(const Matrix &T) : LiftedPose(static_cast<unsigned>(T.rows()), static_cast<unsigned>(T.rows())) {
  DPGO_CHECK(T.cols() == T.rows() + 1);
  setData(T);
}

Pose Pose::Identity(unsigned int d) { return Pose(d); }

Pose Pose::inverse() const {
  // [R t]^-1 = [R^T  -R^T t]
  const Matrix Rt = rotation().transpose();
  Pose out(d());
  out.rotation() = Rt;
  out.translation() = -(Rt * translation());
  return out;
}

Pose Pose::operator*(const Pose &other) const {
  DPGO_CHECK(d() == other.d());
  Pose out(d());
  out.rotation() = rotation() * other.rotation();
  out.translation() = rotation() * other.translation() + translation();
  return out;
}

Matrix Pose::matrix() const {
  Matrix T = Matrix::Identity(d() + 1, d() + 1);
  T.block(0, 0, d(), d()) = rotation();
  T.block(0, d(), d(), 1) = translation();
  return T;
}

}  // namespace DPGO
