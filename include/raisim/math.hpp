// raisim::Vec / Mat / VecDyn / MatDyn -- the small math types on the reference's public API
// (upstream include/raisim/math.hpp, SURVEY.md 8a row a11 [RECALL]; not in the reference snapshot).
// Re-authored from scratch: plain storage + operator[] / operator(); `.e()` Eigen maps are offered
// only when <Eigen/Core> is available (Eigen is the reference's one declared dependency,
// /root/reference/.travis.yml:7, and is absent from this image).
#pragma once
#include <array>
#include <cmath>
#include <cstddef>
#include <vector>

#if defined(__has_include)
#if __has_include(<Eigen/Core>)
#include <Eigen/Core>
#define RAISIM_B200_HAS_EIGEN 1
#endif
#endif

namespace raisim {

template <size_t n>
class Vec {
 public:
  double v[n] = {};
  Vec() = default;
  Vec(std::initializer_list<double> l) { size_t i = 0; for (double x : l) if (i < n) v[i++] = x; }
  double& operator[](size_t i) { return v[i]; }
  double operator[](size_t i) const { return v[i]; }
  double* ptr() { return v; }
  const double* ptr() const { return v; }
  static constexpr size_t size() { return n; }
  void setZero() { for (double& x : v) x = 0; }
  double norm() const { double s = 0; for (double x : v) s += x * x; return std::sqrt(s); }
#ifdef RAISIM_B200_HAS_EIGEN
  Eigen::Map<Eigen::Matrix<double, int(n), 1>> e() { return Eigen::Map<Eigen::Matrix<double, int(n), 1>>(v); }
#endif
};

template <size_t n, size_t m>
class Mat {   // column-major like the reference (so that .e() maps without a copy)
 public:
  double v[n * m] = {};
  double& operator()(size_t i, size_t j) { return v[i + n * j]; }
  double operator()(size_t i, size_t j) const { return v[i + n * j]; }
  double* ptr() { return v; }
  void setZero() { for (double& x : v) x = 0; }
  void setIdentity() { setZero(); for (size_t i = 0; i < (n < m ? n : m); i++) (*this)(i, i) = 1; }
#ifdef RAISIM_B200_HAS_EIGEN
  Eigen::Map<Eigen::Matrix<double, int(n), int(m)>> e() { return Eigen::Map<Eigen::Matrix<double, int(n), int(m)>>(v); }
#endif
};

class VecDyn {
 public:
  std::vector<double> v;
  size_t n = 0;
  VecDyn() = default;
  explicit VecDyn(size_t size) { resize(size); }
  void resize(size_t size) { n = size; v.assign(size, 0.0); }
  void setZero(size_t size) { resize(size); }
  void setZero() { for (double& x : v) x = 0; }
  double& operator[](size_t i) { return v[i]; }
  double operator[](size_t i) const { return v[i]; }
  double* ptr() { return v.data(); }
  const double* ptr() const { return v.data(); }
  size_t size() const { return n; }
#ifdef RAISIM_B200_HAS_EIGEN
  Eigen::Map<Eigen::VectorXd> e() { return Eigen::Map<Eigen::VectorXd>(v.data(), n); }
  VecDyn& operator=(const Eigen::VectorXd& x) { resize(x.size()); for (size_t i = 0; i < n; i++) v[i] = x[i]; return *this; }
#endif
};

class MatDyn {   // column-major
 public:
  std::vector<double> v;
  size_t n = 0, m = 0;
  void resize(size_t rows, size_t cols) { n = rows; m = cols; v.assign(rows * cols, 0.0); }
  double& operator()(size_t i, size_t j) { return v[i + n * j]; }
  double operator()(size_t i, size_t j) const { return v[i + n * j]; }
  double* ptr() { return v.data(); }
  size_t rows() const { return n; }
  size_t cols() const { return m; }
#ifdef RAISIM_B200_HAS_EIGEN
  Eigen::Map<Eigen::MatrixXd> e() { return Eigen::Map<Eigen::MatrixXd>(v.data(), n, m); }
#endif
};

inline void quatToRotMat(const Vec<4>& q, Mat<3, 3>& R) {
  double w = q[0], x = q[1], y = q[2], z = q[3];
  R(0, 0) = 1 - 2 * (y * y + z * z); R(0, 1) = 2 * (x * y - w * z); R(0, 2) = 2 * (x * z + w * y);
  R(1, 0) = 2 * (x * y + w * z); R(1, 1) = 1 - 2 * (x * x + z * z); R(1, 2) = 2 * (y * z - w * x);
  R(2, 0) = 2 * (x * z - w * y); R(2, 1) = 2 * (y * z + w * x); R(2, 2) = 1 - 2 * (x * x + y * y);
}

inline void rotMatToQuat(const Mat<3, 3>& R, Vec<4>& q) {
  double tr = R(0, 0) + R(1, 1) + R(2, 2);
  if (tr > 0) { double s = std::sqrt(tr + 1.0) * 2; q[0] = 0.25 * s; q[1] = (R(2, 1) - R(1, 2)) / s; q[2] = (R(0, 2) - R(2, 0)) / s; q[3] = (R(1, 0) - R(0, 1)) / s; }
  else if (R(0, 0) > R(1, 1) && R(0, 0) > R(2, 2)) { double s = std::sqrt(1.0 + R(0, 0) - R(1, 1) - R(2, 2)) * 2; q[0] = (R(2, 1) - R(1, 2)) / s; q[1] = 0.25 * s; q[2] = (R(0, 1) + R(1, 0)) / s; q[3] = (R(0, 2) + R(2, 0)) / s; }
  else if (R(1, 1) > R(2, 2)) { double s = std::sqrt(1.0 + R(1, 1) - R(0, 0) - R(2, 2)) * 2; q[0] = (R(0, 2) - R(2, 0)) / s; q[1] = (R(0, 1) + R(1, 0)) / s; q[2] = 0.25 * s; q[3] = (R(1, 2) + R(2, 1)) / s; }
  else { double s = std::sqrt(1.0 + R(2, 2) - R(0, 0) - R(1, 1)) * 2; q[0] = (R(1, 0) - R(0, 1)) / s; q[1] = (R(0, 2) + R(2, 0)) / s; q[2] = (R(1, 2) + R(2, 1)) / s; q[3] = 0.25 * s; }
}

}  // namespace raisim
