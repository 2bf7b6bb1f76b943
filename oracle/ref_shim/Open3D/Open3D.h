// Stand-in for <Open3D/Open3D.h> + the few Eigen types poseRefine::process (LL.cpp:27-155) names, so
// that the reference's translation unit compiles.  TEST INFRASTRUCTURE ONLY.  poseRefine is NOT part
// of the compiled match path: every Open3D entry point aborts when reached.  (Open3D itself is an
// external, unpinned dependency of the reference; the ICP oracle lives in oracle/icp_oracle.py.)
#pragma once
#include <initializer_list>
#include <memory>
#include <vector>

#include "../cv_shim.h"

namespace Eigen {
enum { RowMajorBit = 0x1 };
template <typename T, int R, int C, int O = 0, int MR = R, int MC = C>
struct Matrix {
  enum { Flags = 0 };  // column-major
  T d[R * C];
  Matrix() { for (int i = 0; i < R * C; ++i) d[i] = T(0); }
  Matrix(T a, T b, T c) { static_assert(R * C == 3, "vec3"); d[0] = a; d[1] = b; d[2] = c; }
  static Matrix Zero() { return Matrix(); }
  static Matrix Identity(int = R, int = C) { Matrix m; for (int i = 0; i < (R < C ? R : C); ++i) m.d[i * R + i] = T(1); return m; }
  int rows() const { return R; }
  int cols() const { return C; }
  int stride() const { return R; }
  const T* data() const { return d; }
  T* data() { return d; }
  Matrix& operator+=(const Matrix& o) { for (int i = 0; i < R * C; ++i) d[i] += o.d[i]; return *this; }
  Matrix& operator/=(double s) { for (int i = 0; i < R * C; ++i) d[i] = T(d[i] / s); return *this; }
  Matrix operator-(const Matrix& o) const { Matrix m; for (int i = 0; i < R * C; ++i) m.d[i] = d[i] - o.d[i]; return m; }
  template <int C2> Matrix<T, R, C2> operator*(const Matrix<T, C, C2>& o) const {
    Matrix<T, R, C2> m;
    for (int r = 0; r < R; ++r) for (int c = 0; c < C2; ++c) { T s = 0; for (int k = 0; k < C; ++k) s += d[k * R + r] * o.d[c * C + k]; m.d[c * R + r] = s; }
    return m;
  }
  template <typename U> Matrix<U, R, C> cast() const { Matrix<U, R, C> m; for (int i = 0; i < R * C; ++i) m.d[i] = U(d[i]); return m; }
  struct BlockRef {
    Matrix* m; int r0, c0;
    template <int R2, int C2> BlockRef& operator=(const Matrix<T, R2, C2>& v) {
      for (int r = 0; r < R2; ++r) for (int c = 0; c < C2; ++c) m->d[(c0 + c) * R + r0 + r] = v.d[c * R2 + r];
      return *this;
    }
  };
  BlockRef block(int r0, int c0, int, int) { return BlockRef{this, r0, c0}; }
};
typedef Matrix<float, 4, 4> Matrix4f;
typedef Matrix<double, 4, 4> Matrix4d;
typedef Matrix<double, 3, 1> Vector3d;
template <typename M> struct Map : M {
  template <typename T> explicit Map(T* p) { for (int i = 0; i < this->rows() * this->cols(); ++i) this->d[i] = p[i]; }
};
}  // namespace Eigen

namespace open3d {
namespace geometry {
class PointCloud {
 public:
  std::vector<Eigen::Vector3d> points_, normals_;
  std::shared_ptr<PointCloud> VoxelDownSample(double) const { cv::shim_unreachable("open3d::PointCloud::VoxelDownSample"); }
  bool EstimateNormals() { cv::shim_unreachable("open3d::PointCloud::EstimateNormals"); }
  void PaintUniformColor(const Eigen::Vector3d&) {}
  void Transform(const Eigen::Matrix4d&) {}
};
}  // namespace geometry
namespace registration {
struct RegistrationResult {
  Eigen::Matrix4d transformation_;
  double fitness_ = 0, inlier_rmse_ = 0;
};
struct TransformationEstimationPointToPlane {};
struct TransformationEstimationPointToPoint {};
inline RegistrationResult EvaluateRegistration(const geometry::PointCloud&, const geometry::PointCloud&, double, const Eigen::Matrix4d&) {
  cv::shim_unreachable("open3d::registration::EvaluateRegistration");
}
template <typename E>
inline RegistrationResult RegistrationICP(const geometry::PointCloud&, const geometry::PointCloud&, double, const Eigen::Matrix4d&, const E&) {
  cv::shim_unreachable("open3d::registration::RegistrationICP");
}
}  // namespace registration
namespace visualization {
inline void DrawGeometries(std::initializer_list<std::shared_ptr<geometry::PointCloud>>) {}
}  // namespace visualization
}  // namespace open3d
