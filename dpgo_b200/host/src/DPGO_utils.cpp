#include <DPGO/DPGO_robust.h>
#include <DPGO/DPGO_utils.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <fstream>
#include <random>
#include <sstream>

#include "check.h"

namespace DPGO {

static int g_default_device = 0;
void setDefaultDevice(int device) { g_default_device = device; }
int defaultDevice() { return g_default_device; }

std::string ROptParameters::ROptMethodToString(ROptParameters::ROptMethod method) {
  return method == ROptMethod::RTR ? "RTR" : "RGD";
}

std::string InitializationMethodToString(InitializationMethod method) {
  switch (method) {
    case InitializationMethod::Odometry: return "Odometry";
    case InitializationMethod::Chordal: return "Chordal";
    case InitializationMethod::GNC_TLS: return "GNC_TLS";
  }
  return "";
}

void writeMatrixToFile(const Matrix &M, const std::string &filename) {
  std::ofstream file(filename);
  if (!file.is_open()) {
    std::printf("Cannot write to specified file: %s\n", filename.c_str());
    return;
  }
  file << std::setprecision(17);
  for (std::ptrdiff_t i = 0; i < M.rows(); ++i) {
    for (std::ptrdiff_t j = 0; j < M.cols(); ++j) file << (j ? ", " : "") << M(i, j);
    file << "\n";
  }
}

// ---- small dense helpers ---------------------------------------------------------------------
static Matrix inverseSmall(const Matrix &A) {  // 2x2 / 3x3, used for the g2o information blocks
  const double det = A.determinant();
  Matrix inv(A.rows(), A.cols());
  if (A.rows() == 2) {
    inv(0, 0) = A(1, 1) / det; inv(0, 1) = -A(0, 1) / det;
    inv(1, 0) = -A(1, 0) / det; inv(1, 1) = A(0, 0) / det;
    return inv;
  }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      const int a = (j + 1) % 3, b = (j + 2) % 3, c = (i + 1) % 3, e = (i + 2) % 3;
      inv(i, j) = (A(a, c) * A(b, e) - A(a, e) * A(b, c)) / det;  // cofactor of (j,i)
    }
  return inv;
}

void smallSVD(const Matrix &M, Matrix &U, Matrix &S, Matrix &V) {
  // one-sided Jacobi on the columns of A = M: A V = U diag(S)
  const std::ptrdiff_t r = M.rows(), d = M.cols();
  DPGO_CHECK(r >= d);
  Matrix A = M;
  V = Matrix::Identity(d, d);
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0;
    for (std::ptrdiff_t p = 0; p + 1 < d; ++p)
      for (std::ptrdiff_t q = p + 1; q < d; ++q) {
        double al = 0, be = 0, ga = 0;
        for (std::ptrdiff_t k = 0; k < r; ++k) {
          al += A(k, p) * A(k, p);
          be += A(k, q) * A(k, q);
          ga += A(k, p) * A(k, q);
        }
        const double lim = std::sqrt(al * be);
        const double rel = lim > 0 ? std::fabs(ga) / lim : 0;
        off = std::max(off, rel);
        if (rel <= 1e-16) continue;
        const double zeta = (be - al) / (2 * ga);
        const double t = std::copysign(1.0, zeta) / (std::fabs(zeta) + std::sqrt(1 + zeta * zeta));
        const double c = 1 / std::sqrt(1 + t * t), s = c * t;
        for (std::ptrdiff_t k = 0; k < r; ++k) {
          const double x = A(k, p), y = A(k, q);
          A(k, p) = c * x - s * y;
          A(k, q) = s * x + c * y;
        }
        for (std::ptrdiff_t k = 0; k < d; ++k) {
          const double x = V(k, p), y = V(k, q);
          V(k, p) = c * x - s * y;
          V(k, q) = s * x + c * y;
        }
      }
    if (off <= 1e-15) break;
  }
  std::vector<double> sv(static_cast<size_t>(d));
  std::vector<std::ptrdiff_t> order(static_cast<size_t>(d));
  for (std::ptrdiff_t j = 0; j < d; ++j) {
    double nn = 0;
    for (std::ptrdiff_t k = 0; k < r; ++k) nn += A(k, j) * A(k, j);
    sv[static_cast<size_t>(j)] = std::sqrt(nn);
    order[static_cast<size_t>(j)] = j;
  }
  std::sort(order.begin(), order.end(), [&](std::ptrdiff_t a, std::ptrdiff_t b) { return sv[a] > sv[b]; });
  U = Matrix(r, d);
  S = Matrix(d, 1);
  Matrix Vs(d, d);
  for (std::ptrdiff_t j = 0; j < d; ++j) {
    const std::ptrdiff_t src = order[static_cast<size_t>(j)];
    const double s = sv[static_cast<size_t>(src)];
    S(j, 0) = s;
    for (std::ptrdiff_t k = 0; k < r; ++k) U(k, j) = s > 0 ? A(k, src) / s : 0.0;
    for (std::ptrdiff_t k = 0; k < d; ++k) Vs(k, j) = V(k, src);
  }
  V = Vs;
}

Matrix projectToRotationGroup(const Matrix &M) {
  DPGO_CHECK(M.rows() == M.cols());
  Matrix U, S, V;
  smallSVD(M, U, S, V);
  // a rank-deficient input leaves zero columns in U: complete it to an orthonormal basis
  const std::ptrdiff_t d = M.rows();
  if (S(d - 1, 0) <= 1e-300) {
    if (d == 3) {
      U(0, 2) = U(1, 0) * U(2, 1) - U(2, 0) * U(1, 1);
      U(1, 2) = U(2, 0) * U(0, 1) - U(0, 0) * U(2, 1);
      U(2, 2) = U(0, 0) * U(1, 1) - U(1, 0) * U(0, 1);
    } else {
      U(0, 1) = -U(1, 0);
      U(1, 1) = U(0, 0);
    }
  }
  if (U.determinant() * V.determinant() <= 0)
    for (std::ptrdiff_t k = 0; k < d; ++k) U(k, d - 1) = -U(k, d - 1);
  return U * V.transpose();
}

Matrix projectToStiefelManifold(const Matrix &M) {
  DPGO_CHECK(M.rows() >= M.cols());
  Matrix U, S, V;
  smallSVD(M, U, S, V);
  return U * V.transpose();
}

static Matrix qfOf(const Matrix &A) {  // thin QR with positive diagonal (Gram-Schmidt, twice)
  Matrix Q = A;
  for (std::ptrdiff_t k = 0; k < Q.cols(); ++k) {
    for (int pass = 0; pass < 2; ++pass)
      for (std::ptrdiff_t p = 0; p < k; ++p) {
        double s = 0;
        for (std::ptrdiff_t i = 0; i < Q.rows(); ++i) s += Q(i, p) * Q(i, k);
        for (std::ptrdiff_t i = 0; i < Q.rows(); ++i) Q(i, k) -= s * Q(i, p);
      }
    double nn = 0;
    for (std::ptrdiff_t i = 0; i < Q.rows(); ++i) nn += Q(i, k) * Q(i, k);
    nn = std::sqrt(nn);
    for (std::ptrdiff_t i = 0; i < Q.rows(); ++i) Q(i, k) /= nn;
  }
  return Q;
}

Matrix fixedStiefelVariable(unsigned d, unsigned r) {
  Matrix A(r, d);
  for (unsigned i = 0; i < r; ++i)
    for (unsigned j = 0; j < d; ++j) A(i, j) = std::cos(1.0 + 0.7 * i + 1.3 * j) + (i == j ? 1.0 : 0.0);
  return qfOf(A);
}

Matrix randomStiefelVariable(unsigned d, unsigned r) {
  static std::mt19937_64 rng(std::random_device{}());
  std::normal_distribution<double> N(0.0, 1.0);
  Matrix A(r, d);
  for (unsigned j = 0; j < d; ++j)
    for (unsigned i = 0; i < r; ++i) A(i, j) = N(rng);
  return qfOf(A);
}

double computeMeasurementError(const RelativeSEMeasurement &m, const Matrix &R1, const Matrix &t1,
                               const Matrix &R2, const Matrix &t2) {
  const double rot = (R1 * m.R - R2).squaredNorm();
  const double tra = (t2 - t1 - R1 * m.t).squaredNorm();
  return m.kappa * rot + m.tau * tra;
}

// regularized lower incomplete gamma P(a, x)
static double gammaP(double a, double x) {
  if (x <= 0) return 0;
  const double gln = std::lgamma(a);
  if (x < a + 1) {
    double ap = a, sum = 1 / a, del = sum;
    for (int n = 0; n < 1000; ++n) {
      ap += 1;
      del *= x / ap;
      sum += del;
      if (std::fabs(del) < std::fabs(sum) * 1e-16) break;
    }
    return sum * std::exp(-x + a * std::log(x) - gln);
  }
  double b = x + 1 - a, c = 1e300, dd = 1 / b, h = dd;
  for (int i = 1; i < 1000; ++i) {
    const double an = -i * (i - a);
    b += 2;
    dd = an * dd + b;
    if (std::fabs(dd) < 1e-300) dd = 1e-300;
    c = b + an / c;
    if (std::fabs(c) < 1e-300) c = 1e-300;
    dd = 1 / dd;
    const double del = dd * c;
    h *= del;
    if (std::fabs(del - 1) < 1e-16) break;
  }
  return 1 - std::exp(-x + a * std::log(x) - gln) * h;
}

double chi2inv(double quantile, size_t dof) {
  DPGO_CHECK(quantile >= 0 && quantile < 1);
  double lo = 0, hi = std::max<double>(10.0, 4.0 * dof);
  while (gammaP(0.5 * dof, 0.5 * hi) < quantile) hi *= 2;
  for (int it = 0; it < 200; ++it) {
    const double mid = 0.5 * (lo + hi);
    if (gammaP(0.5 * dof, 0.5 * mid) < quantile) lo = mid; else hi = mid;
  }
  return 0.5 * (lo + hi);
}

double angular2ChordalSO3(double rad) { return 2 * std::sqrt(2.0) * std::sin(rad / 2); }

void checkRotationMatrix(const Matrix &R) {
  const auto d = R.rows();
  DPGO_CHECK(R.cols() == d);
  const double err_det = std::fabs(R.determinant() - 1.0);
  const double err_norm = (R.transpose() * R - Matrix::Identity(d, d)).norm();
  if (err_det > 1e-5 || err_norm > 1e-5)
    std::fprintf(stderr, "[checkRotationMatrix] Invalid rotation: err_det=%g, err_norm=%g\n", err_det, err_norm);
}

void checkStiefelMatrix(const Matrix &Y) {
  const auto d = Y.cols();
  const double err_norm = (Y.transpose() * Y - Matrix::Identity(d, d)).norm();
  if (err_norm > 1e-5) std::fprintf(stderr, "[checkStiefelMatrix] Invalid Stiefel: err_norm=%g\n", err_norm);
}

// ---- g2o ---------------------------------------------------------------------------------------
std::vector<RelativeSEMeasurement> read_g2o_file(const std::string &filename, size_t &num_poses) {
  std::vector<RelativeSEMeasurement> out;
  std::ifstream in(filename);
  if (!in.is_open()) {
    std::fprintf(stderr, "[read_g2o_file] cannot open %s\n", filename.c_str());
    std::abort();
  }
  num_poses = 0;
  std::string line, tag;
  while (std::getline(in, line)) {
    std::stringstream ss(line);
    if (!(ss >> tag)) continue;
    RelativeSEMeasurement m;
    m.weight = 1.0;
    m.r1 = m.r2 = 0;
    if (tag == "EDGE_SE2") {
      double dx, dy, dth, I11, I12, I13, I22, I23, I33;
      ss >> m.p1 >> m.p2 >> dx >> dy >> dth >> I11 >> I12 >> I13 >> I22 >> I23 >> I33;
      m.t = Matrix(2, 1);
      m.t(0, 0) = dx; m.t(1, 0) = dy;
      m.R = Matrix(2, 2);
      m.R(0, 0) = std::cos(dth); m.R(0, 1) = -std::sin(dth);
      m.R(1, 0) = std::sin(dth); m.R(1, 1) = std::cos(dth);
      Matrix C(2, 2);
      C(0, 0) = I11; C(0, 1) = I12; C(1, 0) = I12; C(1, 1) = I22;
      m.tau = 2 / inverseSmall(C).trace();
      m.kappa = I33;
    } else if (tag == "EDGE_SE3:QUAT") {
      double dx, dy, dz, qx, qy, qz, qw, I[21];
      ss >> m.p1 >> m.p2 >> dx >> dy >> dz >> qx >> qy >> qz >> qw;
      for (double &v : I) ss >> v;  // upper triangle, row by row: I11..I16, I22..I26, I33.., I66
      m.t = Matrix(3, 1);
      m.t(0, 0) = dx; m.t(1, 0) = dy; m.t(2, 0) = dz;
      // quaternion (w,x,y,z) -> rotation, no normalisation (Eigen::Quaterniond::toRotationMatrix)
      const double tx = 2 * qx, ty = 2 * qy, tz = 2 * qz;
      const double twx = tx * qw, twy = ty * qw, twz = tz * qw;
      const double txx = tx * qx, txy = ty * qx, txz = tz * qx, tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
      m.R = Matrix(3, 3);
      m.R(0, 0) = 1 - (tyy + tzz); m.R(0, 1) = txy - twz; m.R(0, 2) = txz + twy;
      m.R(1, 0) = txy + twz; m.R(1, 1) = 1 - (txx + tzz); m.R(1, 2) = tyz - twx;
      m.R(2, 0) = txz - twy; m.R(2, 1) = tyz + twx; m.R(2, 2) = 1 - (txx + tyy);
      Matrix Ct(3, 3), Cr(3, 3);
      Ct(0, 0) = I[0]; Ct(0, 1) = Ct(1, 0) = I[1]; Ct(0, 2) = Ct(2, 0) = I[2];
      Ct(1, 1) = I[6]; Ct(1, 2) = Ct(2, 1) = I[7]; Ct(2, 2) = I[11];
      Cr(0, 0) = I[15]; Cr(0, 1) = Cr(1, 0) = I[16]; Cr(0, 2) = Cr(2, 0) = I[17];
      Cr(1, 1) = I[18]; Cr(1, 2) = Cr(2, 1) = I[19]; Cr(2, 2) = I[20];
      m.tau = 3 / inverseSmall(Ct).trace();
      m.kappa = 3 / (2 * inverseSmall(Cr).trace());
    } else if (tag == "VERTEX_SE2" || tag == "VERTEX_SE3:QUAT") {
      continue;
    } else {
      std::fprintf(stderr, "[read_g2o_file] Error: unrecognized type: %s!\n", tag.c_str());
      std::abort();
    }
    m.fixedWeight = (m.p1 + 1 == m.p2);
    num_poses = std::max(num_poses, std::max(m.p1, m.p2));
    out.push_back(m);
  }
  num_poses++;  // zero-based ids
  return out;
}

void get_dimension_and_num_poses(const std::vector<RelativeSEMeasurement> &measurements, size_t &dimension,
                                 size_t &num_poses) {
  DPGO_CHECK(!measurements.empty());
  dimension = static_cast<size_t>(measurements[0].t.size());
  DPGO_CHECK(dimension == 2 || dimension == 3);
  num_poses = 0;
  for (const auto &m : measurements) num_poses = std::max(num_poses, std::max(m.p1, m.p2) + 1);
}

}  // namespace DPGO
