#include <DPGO/QuadraticProblem.h>

#include "check.h"

namespace DPGO {

QuadraticProblem::QuadraticProblem(const std::shared_ptr<PoseGraph> &pose_graph) : pose_graph_(pose_graph) {
  DPGO_CHECK(pose_graph_ != nullptr);
}

void QuadraticProblem::checkShape(const Matrix &Y) const {
  DPGO_CHECK(static_cast<unsigned>(Y.rows()) == relaxation_rank());
  DPGO_CHECK(static_cast<unsigned>(Y.cols()) == (dimension() + 1) * num_poses());
}

dpgo_dev *QuadraticProblem::device() const {
  DPGO_CHECK(pose_graph_->constructDataMatrices());  // reference: lazy quadraticMatrix()/linearMatrix()
  return pose_graph_->device();
}

double QuadraticProblem::f(const Matrix &Y) const {
  checkShape(Y);
  double v = 0;
  DPGO_DEVICE_CALL(dpgo_f(device(), Y.data(), &v));
  return v;
}

Matrix QuadraticProblem::EucGrad(const Matrix &Y) const {
  checkShape(Y);
  Matrix out(Y.rows(), Y.cols());
  DPGO_DEVICE_CALL(dpgo_egrad(device(), Y.data(), out.data()));
  return out;
}

Matrix QuadraticProblem::HessianEta(const Matrix &Y, const Matrix &V) const {
  checkShape(Y);
  checkShape(V);
  Matrix out(Y.rows(), Y.cols());
  DPGO_DEVICE_CALL(dpgo_hessvec(device(), Y.data(), V.data(), out.data()));
  return out;
}

Matrix QuadraticProblem::PreConditioner(const Matrix &Y, const Matrix &V) const {
  checkShape(Y);
  checkShape(V);
  Matrix out(Y.rows(), Y.cols());
  if (!pose_graph_->hasPreconditioner()) {
    std::fprintf(stderr, "[QuadraticProblem] Failed to compute preconditioner.\n");
    DPGO_DEVICE_CALL(dpgo_tangent_project(device(), Y.data(), V.data(), out.data()));
    return out;
  }
  DPGO_DEVICE_CALL(dpgo_precon(device(), Y.data(), V.data(), out.data()));
  return out;
}

Matrix QuadraticProblem::RieGrad(const Matrix &Y) const {
  checkShape(Y);
  Matrix out(Y.rows(), Y.cols());
  DPGO_DEVICE_CALL(dpgo_rgrad(device(), Y.data(), out.data(), nullptr));
  return out;
}

double QuadraticProblem::RieGradNorm(const Matrix &Y) const {
  checkShape(Y);
  double nrm = 0;
  DPGO_DEVICE_CALL(dpgo_rgrad(device(), Y.data(), nullptr, &nrm));
  return nrm;
}

}  // namespace DPGO
