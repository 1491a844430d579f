// f(X) = 0.5 <Q, X^T X> + <X, G> on (St(d,r) x R^r)^n -- drop-in for the reference's
// QuadraticProblem (include/DPGO/QuadraticProblem.h:39-103).  The reference subclasses
// ROPTLIB::Problem and evaluates with Eigen; here every evaluation is a CUDA kernel behind the
// C-ABI (dpgo_f / dpgo_rgrad / dpgo_egrad / dpgo_hessvec / dpgo_precon), and the ROPTLIB
// Variable* overloads (used only by ROPTLIB itself) are replaced by Matrix overloads.
#ifndef DPGO_B200_QUADRATICPROBLEM_H
#define DPGO_B200_QUADRATICPROBLEM_H

#include <DPGO/DPGO_types.h>
#include <DPGO/PoseGraph.h>

#include <memory>

namespace DPGO {

class QuadraticProblem {
 public:
  explicit QuadraticProblem(const std::shared_ptr<PoseGraph> &pose_graph);
  virtual ~QuadraticProblem() = default;

  unsigned int num_poses() const { return pose_graph_->n(); }
  unsigned int dimension() const { return pose_graph_->d(); }
  unsigned int relaxation_rank() const { return pose_graph_->r(); }

  /// cost (reference: src/QuadraticProblem.cpp:29-41)
  double f(const Matrix &Y) const;
  /// Euclidean gradient X Q + G (reference :43-47)
  Matrix EucGrad(const Matrix &Y) const;
  /// Riemannian Hessian-vector product at Y (reference :49-54 + ROPTLIB's Stiefel correction)
  Matrix HessianEta(const Matrix &Y, const Matrix &V) const;
  /// preconditioner: (Q + 0.1 I)^-1 V projected on the tangent space at Y (reference :56-69)
  Matrix PreConditioner(const Matrix &Y, const Matrix &V) const;
  /// Riemannian gradient and its norm (reference :71-83)
  Matrix RieGrad(const Matrix &Y) const;
  double RieGradNorm(const Matrix &Y) const;

  const std::shared_ptr<PoseGraph> &poseGraph() const { return pose_graph_; }

 private:
  void checkShape(const Matrix &Y) const;
  dpgo_dev *device() const;
  std::shared_ptr<PoseGraph> pose_graph_;
};

}  // namespace DPGO
#endif
