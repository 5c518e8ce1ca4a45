// eigen_shim.hpp — the reference's host API is written against Eigen (hpipm-cpp/include/hpipm-cpp/ocp_qp.hpp:15-177).
// Eigen is not available in this build image, so the facades compile against this minimal column-major
// stand-in; when <Eigen/Core> exists the real library is used instead and nothing here is compiled.
#pragma once
#if __has_include(<Eigen/Core>) && !defined(SRBD_FORCE_EIGEN_SHIM)   // (the reference-wrapper build of test_hpipm_compat.cpp puts an empty Eigen/Core stub on the path)
#include <Eigen/Core>
#else
#include <cstddef>
#include <vector>

namespace Eigen {

class MatrixXd {
 public:
  MatrixXd() = default;
  MatrixXd(std::ptrdiff_t r, std::ptrdiff_t c) : r_(r), c_(c), v_(static_cast<size_t>(r * c), 0.0) {}
  void resize(std::ptrdiff_t r, std::ptrdiff_t c) { r_ = r; c_ = c; v_.assign(static_cast<size_t>(r * c), 0.0); }
  std::ptrdiff_t rows() const { return r_; }
  std::ptrdiff_t cols() const { return c_; }
  std::ptrdiff_t size() const { return r_ * c_; }
  double* data() { return v_.data(); }
  const double* data() const { return v_.data(); }
  double& operator()(std::ptrdiff_t i, std::ptrdiff_t j) { return v_[static_cast<size_t>(i + r_ * j)]; }
  double operator()(std::ptrdiff_t i, std::ptrdiff_t j) const { return v_[static_cast<size_t>(i + r_ * j)]; }
  void setZero() { v_.assign(v_.size(), 0.0); }
  void setIdentity() { setZero(); for (std::ptrdiff_t i = 0; i < (r_ < c_ ? r_ : c_); ++i) (*this)(i, i) = 1.0; }
  static MatrixXd Zero(std::ptrdiff_t r, std::ptrdiff_t c) { return MatrixXd(r, c); }
  static MatrixXd Identity(std::ptrdiff_t r, std::ptrdiff_t c) { MatrixXd m(r, c); m.setIdentity(); return m; }

 private:
  std::ptrdiff_t r_ = 0, c_ = 0;
  std::vector<double> v_;
};

class VectorXd {
 public:
  VectorXd() = default;
  explicit VectorXd(std::ptrdiff_t n) : v_(static_cast<size_t>(n), 0.0) {}
  void resize(std::ptrdiff_t n) { v_.assign(static_cast<size_t>(n), 0.0); }
  std::ptrdiff_t size() const { return static_cast<std::ptrdiff_t>(v_.size()); }
  std::ptrdiff_t rows() const { return size(); }
  std::ptrdiff_t cols() const { return 1; }
  double* data() { return v_.data(); }
  const double* data() const { return v_.data(); }
  double& operator()(std::ptrdiff_t i) { return v_[static_cast<size_t>(i)]; }
  double operator()(std::ptrdiff_t i) const { return v_[static_cast<size_t>(i)]; }
  double& operator[](std::ptrdiff_t i) { return v_[static_cast<size_t>(i)]; }
  double operator[](std::ptrdiff_t i) const { return v_[static_cast<size_t>(i)]; }
  void setZero() { v_.assign(v_.size(), 0.0); }
  void fill(double x) { v_.assign(v_.size(), x); }
  static VectorXd Zero(std::ptrdiff_t n) { return VectorXd(n); }

 private:
  std::vector<double> v_;
};

}  // namespace Eigen
#endif
