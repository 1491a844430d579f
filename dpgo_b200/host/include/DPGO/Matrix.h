// Minimal column-major dense double matrix standing in for Eigen::MatrixXd in the public DPGO
// API (the reference typedefs `Matrix = Eigen::MatrixXd`, include/DPGO/DPGO_types.h:24; Eigen is
// not a dependency here).  It covers what the drop-in surface and the reference's drivers use:
// construction, (i,j) access, block() as l- and r-value, products, sums, norm(), transpose().
// Storage is column-major like Eigen's default, so data() can be handed to the C-ABI directly.
#ifndef DPGO_B200_MATRIX_H
#define DPGO_B200_MATRIX_H

#include <cassert>
#include <cmath>
#include <cstddef>
#include <iomanip>
#include <iostream>
#include <vector>

namespace DPGO {

class Matrix;

// View of a rectangular sub-block of a Matrix (what Eigen's .block() returns).
class BlockRef {
 public:
  BlockRef(double *base, std::ptrdiff_t ld, std::ptrdiff_t r, std::ptrdiff_t c) : p_(base), ld_(ld), r_(r), c_(c) {}
  std::ptrdiff_t rows() const { return r_; }
  std::ptrdiff_t cols() const { return c_; }
  double &operator()(std::ptrdiff_t i, std::ptrdiff_t j) { return p_[i + j * ld_]; }
  double operator()(std::ptrdiff_t i, std::ptrdiff_t j) const { return p_[i + j * ld_]; }
  BlockRef &operator=(const Matrix &m);
  BlockRef &operator=(const BlockRef &o);
  BlockRef &operator+=(const Matrix &m);
  BlockRef &operator-=(const Matrix &m);
  BlockRef &operator*=(double s) {
    for (std::ptrdiff_t j = 0; j < c_; ++j)
      for (std::ptrdiff_t i = 0; i < r_; ++i) (*this)(i, j) *= s;
    return *this;
  }
  void setZero() { *this *= 0.0; }
  Matrix eval() const;
  Matrix transpose() const;
  double norm() const;
  double determinant() const;

 private:
  double *p_;
  std::ptrdiff_t ld_, r_, c_;
};

class Matrix {
 public:
  Matrix() : r_(0), c_(0) {}
  Matrix(std::ptrdiff_t r, std::ptrdiff_t c) : r_(r), c_(c), v_(static_cast<size_t>(r * c), 0.0) {}
  Matrix(const BlockRef &b) : r_(b.rows()), c_(b.cols()), v_(static_cast<size_t>(b.rows() * b.cols())) {  // NOLINT
    for (std::ptrdiff_t j = 0; j < c_; ++j)
      for (std::ptrdiff_t i = 0; i < r_; ++i) (*this)(i, j) = b(i, j);
  }
  static Matrix Zero(std::ptrdiff_t r, std::ptrdiff_t c) { return Matrix(r, c); }
  static Matrix Identity(std::ptrdiff_t r, std::ptrdiff_t c) {
    Matrix m(r, c);
    for (std::ptrdiff_t i = 0; i < (r < c ? r : c); ++i) m(i, i) = 1.0;
    return m;
  }

  std::ptrdiff_t rows() const { return r_; }
  std::ptrdiff_t cols() const { return c_; }
  std::ptrdiff_t size() const { return r_ * c_; }
  double *data() { return v_.data(); }
  const double *data() const { return v_.data(); }
  void resize(std::ptrdiff_t r, std::ptrdiff_t c) {
    r_ = r;
    c_ = c;
    v_.assign(static_cast<size_t>(r * c), 0.0);
  }
  void setZero() { v_.assign(v_.size(), 0.0); }

  double &operator()(std::ptrdiff_t i, std::ptrdiff_t j) {
    assert(i >= 0 && i < r_ && j >= 0 && j < c_);
    return v_[static_cast<size_t>(i + j * r_)];
  }
  double operator()(std::ptrdiff_t i, std::ptrdiff_t j) const {
    assert(i >= 0 && i < r_ && j >= 0 && j < c_);
    return v_[static_cast<size_t>(i + j * r_)];
  }
  double &operator()(std::ptrdiff_t i) { return v_[static_cast<size_t>(i)]; }  // vector access
  double operator()(std::ptrdiff_t i) const { return v_[static_cast<size_t>(i)]; }

  BlockRef block(std::ptrdiff_t i, std::ptrdiff_t j, std::ptrdiff_t p, std::ptrdiff_t q) {
    assert(i >= 0 && j >= 0 && i + p <= r_ && j + q <= c_);
    return BlockRef(v_.data() + i + j * r_, r_, p, q);
  }
  Matrix block(std::ptrdiff_t i, std::ptrdiff_t j, std::ptrdiff_t p, std::ptrdiff_t q) const {
    assert(i >= 0 && j >= 0 && i + p <= r_ && j + q <= c_);
    Matrix m(p, q);
    for (std::ptrdiff_t b = 0; b < q; ++b)
      for (std::ptrdiff_t a = 0; a < p; ++a) m(a, b) = (*this)(i + a, j + b);
    return m;
  }
  BlockRef col(std::ptrdiff_t j) { return block(0, j, r_, 1); }
  Matrix col(std::ptrdiff_t j) const { return block(0, j, r_, 1); }

  Matrix transpose() const {
    Matrix t(c_, r_);
    for (std::ptrdiff_t j = 0; j < c_; ++j)
      for (std::ptrdiff_t i = 0; i < r_; ++i) t(j, i) = (*this)(i, j);
    return t;
  }
  double squaredNorm() const {
    double s = 0.0;
    for (double x : v_) s += x * x;
    return s;
  }
  double norm() const { return std::sqrt(squaredNorm()); }
  double sum() const {
    double s = 0.0;
    for (double x : v_) s += x;
    return s;
  }
  double trace() const {
    double s = 0.0;
    for (std::ptrdiff_t i = 0; i < (r_ < c_ ? r_ : c_); ++i) s += (*this)(i, i);
    return s;
  }
  double determinant() const;  // 2x2 / 3x3 only

  Matrix &operator+=(const Matrix &o) {
    assert(r_ == o.r_ && c_ == o.c_);
    for (size_t k = 0; k < v_.size(); ++k) v_[k] += o.v_[k];
    return *this;
  }
  Matrix &operator-=(const Matrix &o) {
    assert(r_ == o.r_ && c_ == o.c_);
    for (size_t k = 0; k < v_.size(); ++k) v_[k] -= o.v_[k];
    return *this;
  }
  Matrix &operator*=(double s) {
    for (double &x : v_) x *= s;
    return *this;
  }

 private:
  std::ptrdiff_t r_, c_;
  std::vector<double> v_;
};

typedef Matrix Vector;  // reference: typedef Eigen::VectorXd Vector (a 1-column Matrix here)

inline Matrix operator+(Matrix a, const Matrix &b) { return a += b; }
inline Matrix operator-(Matrix a, const Matrix &b) { return a -= b; }
inline Matrix operator*(Matrix a, double s) { return a *= s; }
inline Matrix operator*(double s, Matrix a) { return a *= s; }
inline Matrix operator-(Matrix a) { return a *= -1.0; }
inline Matrix operator*(const Matrix &a, const Matrix &b) {
  assert(a.cols() == b.rows());
  Matrix c(a.rows(), b.cols());
  for (std::ptrdiff_t j = 0; j < b.cols(); ++j)
    for (std::ptrdiff_t k = 0; k < a.cols(); ++k) {
      const double bkj = b(k, j);
      for (std::ptrdiff_t i = 0; i < a.rows(); ++i) c(i, j) += a(i, k) * bkj;
    }
  return c;
}

inline double Matrix::determinant() const {
  assert(r_ == c_ && (r_ == 2 || r_ == 3));
  const Matrix &m = *this;
  if (r_ == 2) return m(0, 0) * m(1, 1) - m(0, 1) * m(1, 0);
  return m(0, 0) * (m(1, 1) * m(2, 2) - m(1, 2) * m(2, 1)) - m(0, 1) * (m(1, 0) * m(2, 2) - m(1, 2) * m(2, 0)) +
         m(0, 2) * (m(1, 0) * m(2, 1) - m(1, 1) * m(2, 0));
}

inline BlockRef &BlockRef::operator=(const Matrix &m) {
  assert(m.rows() == r_ && m.cols() == c_);
  for (std::ptrdiff_t j = 0; j < c_; ++j)
    for (std::ptrdiff_t i = 0; i < r_; ++i) (*this)(i, j) = m(i, j);
  return *this;
}
inline BlockRef &BlockRef::operator=(const BlockRef &o) { return *this = Matrix(o); }
inline BlockRef &BlockRef::operator+=(const Matrix &m) {
  for (std::ptrdiff_t j = 0; j < c_; ++j)
    for (std::ptrdiff_t i = 0; i < r_; ++i) (*this)(i, j) += m(i, j);
  return *this;
}
inline BlockRef &BlockRef::operator-=(const Matrix &m) {
  for (std::ptrdiff_t j = 0; j < c_; ++j)
    for (std::ptrdiff_t i = 0; i < r_; ++i) (*this)(i, j) -= m(i, j);
  return *this;
}
inline Matrix BlockRef::eval() const { return Matrix(*this); }
inline Matrix BlockRef::transpose() const { return Matrix(*this).transpose(); }
inline double BlockRef::norm() const { return Matrix(*this).norm(); }
inline double BlockRef::determinant() const { return Matrix(*this).determinant(); }

inline Matrix operator*(const BlockRef &a, const Matrix &b) { return Matrix(a) * b; }
inline Matrix operator*(const Matrix &a, const BlockRef &b) { return a * Matrix(b); }
inline Matrix operator*(const BlockRef &a, const BlockRef &b) { return Matrix(a) * Matrix(b); }
inline Matrix operator+(const BlockRef &a, const Matrix &b) { return Matrix(a) + b; }
inline Matrix operator-(const BlockRef &a, const Matrix &b) { return Matrix(a) - b; }
inline Matrix operator-(const Matrix &a, const BlockRef &b) { return a - Matrix(b); }
inline Matrix operator-(const BlockRef &a, const BlockRef &b) { return Matrix(a) - Matrix(b); }

inline std::ostream &operator<<(std::ostream &os, const Matrix &m) {
  for (std::ptrdiff_t i = 0; i < m.rows(); ++i) {
    for (std::ptrdiff_t j = 0; j < m.cols(); ++j) os << (j ? " " : "") << m(i, j);
    if (i + 1 < m.rows()) os << "\n";
  }
  return os;
}

}  // namespace DPGO
#endif
