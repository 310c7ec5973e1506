// pcl_eigen_min.hpp — the few PCL / Eigen types the NanoGICP class surface mentions, for builds where the
// real libraries are absent (this repo's image has neither).  With real PCL/Eigen on the include path
// nano_gicp.hpp uses those instead and this file is not included.  Layouts match the originals:
// pcl::PointXYZI is 32 bytes {x,y,z,1,intensity,pad[3]}, Eigen matrices are column-major PODs.
#pragma once
#include <cstddef>
#include <cstdint>
#include <memory>
#include <vector>

namespace Eigen {

template <class T, int R, int C>
struct Matrix {
  T v[R * C];
  Matrix() { for (int i = 0; i < R * C; i++) v[i] = T(0); }
  T& operator()(int r, int c) { return v[c * R + r]; }
  const T& operator()(int r, int c) const { return v[c * R + r]; }
  T& operator[](int i) { return v[i]; }                 // vectors
  const T& operator[](int i) const { return v[i]; }
  T* data() { return v; }
  const T* data() const { return v; }
  // m.block(r0, c0, nr, nc) as an rvalue (what OdomNode does: `Eigen::Vector3f p = T.block(0,3,3,1)`)
  struct ConstBlock {
    const Matrix& m; int r0, c0;
    template <int RR, int CC> operator Matrix<T, RR, CC>() const {
      Matrix<T, RR, CC> o;
      for (int c = 0; c < CC; c++) for (int r = 0; r < RR; r++) o(r, c) = m(r0 + r, c0 + c);
      return o;
    }
  };
  ConstBlock block(int r0, int c0, int, int) const { return ConstBlock{*this, r0, c0}; }
  static Matrix Identity() { Matrix m; for (int i = 0; i < (R < C ? R : C); i++) m(i, i) = T(1); return m; }
  static Matrix Zero() { return Matrix(); }
  void setIdentity() { *this = Identity(); }
  template <class U> Matrix<U, R, C> cast() const { Matrix<U, R, C> o; for (int i = 0; i < R * C; i++) o.v[i] = (U)v[i]; return o; }
  Matrix operator*(const Matrix& b) const {
    static_assert(R == C, "square only in the shim");
    Matrix o;
    for (int c = 0; c < C; c++) for (int r = 0; r < R; r++) { T s = T(0); for (int k = 0; k < C; k++) s += (*this)(r, k) * b(k, c); o(r, c) = s; }
    return o;
  }
};
typedef Matrix<float, 4, 4> Matrix4f;
typedef Matrix<double, 4, 4> Matrix4d;
typedef Matrix<float, 3, 1> Vector3f;

template <class T>
struct aligned_allocator : public std::allocator<T> {
  template <class U> struct rebind { typedef aligned_allocator<U> other; };
  aligned_allocator() {}
  template <class U> aligned_allocator(const aligned_allocator<U>&) {}
};

}  // namespace Eigen

#ifndef EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#endif

namespace pcl {

struct alignas(16) PointXYZI {
  union { float data[4]; struct { float x, y, z; }; };
  union { float data_c[4]; struct { float intensity; }; };
  PointXYZI() : data{0.f, 0.f, 0.f, 1.f}, data_c{0.f, 0.f, 0.f, 0.f} {}
};
static_assert(sizeof(PointXYZI) == 32, "pcl::PointXYZI is 32 bytes");

template <class PointT>
struct PointCloud {
  typedef std::shared_ptr<PointCloud<PointT>> Ptr;
  typedef std::shared_ptr<const PointCloud<PointT>> ConstPtr;
  std::vector<PointT> points;
  uint32_t width = 0, height = 1;
  bool is_dense = true;
  size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  PointT& at(size_t i) { return points.at(i); }
  const PointT& at(size_t i) const { return points.at(i); }
  PointT& operator[](size_t i) { return points[i]; }
  const PointT& operator[](size_t i) const { return points[i]; }
  void resize(size_t n) { points.resize(n); width = (uint32_t)n; height = 1; }
  void push_back(const PointT& p) { points.push_back(p); width = (uint32_t)points.size(); }
  PointCloud& operator+=(const PointCloud& o) { points.insert(points.end(), o.points.begin(), o.points.end()); width = (uint32_t)points.size(); return *this; }
};

}  // namespace pcl
