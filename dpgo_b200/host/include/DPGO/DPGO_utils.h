// Utility functions of the public API (reference: include/DPGO/DPGO_utils.h): g2o parsing, small
// dense projections, timers.  Host-side, cold path.
#ifndef DPGO_B200_UTILS_H
#define DPGO_B200_UTILS_H

#include <DPGO/DPGO_types.h>
#include <DPGO/RelativeSEMeasurement.h>

#include <chrono>
#include <string>
#include <vector>

namespace DPGO {

class SimpleTimer {
 public:
  void tic() { t0_ = Tic(); }
  double toc() { return Toc(t0_); }  ///< milliseconds since tic()
  static std::chrono::time_point<std::chrono::high_resolution_clock> Tic() {
    return std::chrono::high_resolution_clock::now();
  }
  static double Toc(const std::chrono::time_point<std::chrono::high_resolution_clock> &start) {
    return std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - start).count();
  }

 private:
  std::chrono::time_point<std::chrono::high_resolution_clock> t0_;
};

void writeMatrixToFile(const Matrix &M, const std::string &filename);

/// parse EDGE_SE2 / EDGE_SE3:QUAT lines of a .g2o file (reference: src/DPGO_utils.cpp:113-257)
std::vector<RelativeSEMeasurement> read_g2o_file(const std::string &filename, size_t &num_poses);

void get_dimension_and_num_poses(const std::vector<RelativeSEMeasurement> &measurements, size_t &dimension,
                                 size_t &num_poses);

/// nearest rotation (SVD, det = +1) -- reference :464-478
Matrix projectToRotationGroup(const Matrix &M);
/// polar factor U V^T of an r x d matrix -- reference :480-486
Matrix projectToStiefelManifold(const Matrix &M);
/// the lifting matrix every agent shares: a fixed element of St(d, r).  The reference draws it
/// from ROPTLIB's RNG after srand(1) (:488-493); any fixed Stiefel element yields the same
/// objective values, so a deterministic closed form is used here.
Matrix fixedStiefelVariable(unsigned d, unsigned r);
Matrix randomStiefelVariable(unsigned d, unsigned r);

/// kappa |R1 R - R2|^2 + tau |t2 - t1 - R1 t|^2   (reference :501-507)
double computeMeasurementError(const RelativeSEMeasurement &m, const Matrix &R1, const Matrix &t1,
                               const Matrix &R2, const Matrix &t2);
double chi2inv(double quantile, size_t dof);
double angular2ChordalSO3(double rad);
void checkRotationMatrix(const Matrix &R);
void checkStiefelMatrix(const Matrix &Y);

/// thin SVD of a small r x d matrix (d <= r): M = U diag(S) V^T, singular values descending
void smallSVD(const Matrix &M, Matrix &U, Matrix &S, Matrix &V);

/// CUDA device used by PoseGraph objects created afterwards (default 0)
void setDefaultDevice(int device);
int defaultDevice();

}  // namespace DPGO
#endif
