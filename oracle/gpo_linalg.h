// ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into, imported by, or called from the
// product path (gpslam_b200/, include/).  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may use it, and only as the checker.
//
// Tiny fixed-size, column-major matrix type (the reference works on Eigen fixed-size
// matrices, which are column-major: SURVEY.md §8 notation).  No Eigen/Boost/GTSAM exist
// in the build container, so the oracle carries its own.
#pragma once
#include <cmath>
#include <cstring>
#include <limits>

namespace gpo {

template <int R, int C>
struct Mat {
  double a[R * C];
  static constexpr int rows = R, cols = C;
  double& operator()(int r, int c) { return a[r + c * R]; }
  double operator()(int r, int c) const { return a[r + c * R]; }
  double& operator[](int i) { return a[i]; }
  double operator[](int i) const { return a[i]; }
  static Mat Zero() { Mat m; for (int i = 0; i < R * C; i++) m.a[i] = 0.0; return m; }
  static Mat Identity() { Mat m = Zero(); for (int i = 0; i < (R < C ? R : C); i++) m(i, i) = 1.0; return m; }
  Mat<C, R> t() const { Mat<C, R> m; for (int r = 0; r < R; r++) for (int c = 0; c < C; c++) m(c, r) = (*this)(r, c); return m; }
  template <int BR, int BC> Mat<BR, BC> block(int r0, int c0) const {
    Mat<BR, BC> m; for (int r = 0; r < BR; r++) for (int c = 0; c < BC; c++) m(r, c) = (*this)(r0 + r, c0 + c); return m;
  }
  template <int BR, int BC> void set(int r0, int c0, const Mat<BR, BC>& b) {
    for (int r = 0; r < BR; r++) for (int c = 0; c < BC; c++) (*this)(r0 + r, c0 + c) = b(r, c);
  }
  double norm() const { double s = 0; for (int i = 0; i < R * C; i++) s += a[i] * a[i]; return std::sqrt(s); }
  double dot(const Mat& o) const { double s = 0; for (int i = 0; i < R * C; i++) s += a[i] * o.a[i]; return s; }
  double maxabs() const { double s = 0; for (int i = 0; i < R * C; i++) s = std::fmax(s, std::fabs(a[i])); return s; }
};

template <int R, int C> Mat<R, C> operator+(const Mat<R, C>& x, const Mat<R, C>& y) { Mat<R, C> m; for (int i = 0; i < R * C; i++) m.a[i] = x.a[i] + y.a[i]; return m; }
template <int R, int C> Mat<R, C> operator-(const Mat<R, C>& x, const Mat<R, C>& y) { Mat<R, C> m; for (int i = 0; i < R * C; i++) m.a[i] = x.a[i] - y.a[i]; return m; }
template <int R, int C> Mat<R, C> operator-(const Mat<R, C>& x) { Mat<R, C> m; for (int i = 0; i < R * C; i++) m.a[i] = -x.a[i]; return m; }
template <int R, int C> Mat<R, C> operator*(double s, const Mat<R, C>& x) { Mat<R, C> m; for (int i = 0; i < R * C; i++) m.a[i] = s * x.a[i]; return m; }
template <int R, int C> Mat<R, C> operator*(const Mat<R, C>& x, double s) { return s * x; }
template <int R, int C> Mat<R, C> operator/(const Mat<R, C>& x, double s) { Mat<R, C> m; for (int i = 0; i < R * C; i++) m.a[i] = x.a[i] / s; return m; }
template <int R, int K, int C> Mat<R, C> operator*(const Mat<R, K>& x, const Mat<K, C>& y) {
  Mat<R, C> m;
  for (int r = 0; r < R; r++) for (int c = 0; c < C; c++) { double s = 0; for (int k = 0; k < K; k++) s += x(r, k) * y(k, c); m(r, c) = s; }
  return m;
}

using Vec2 = Mat<2, 1>; using Vec3 = Mat<3, 1>; using Vec6 = Mat<6, 1>; using Vec12 = Mat<12, 1>;
using Mat2 = Mat<2, 2>; using Mat3 = Mat<3, 3>; using Mat6 = Mat<6, 6>; using Mat12 = Mat<12, 12>;

template <int N> Mat<N, 1> vec(const double* p) { Mat<N, 1> v; for (int i = 0; i < N; i++) v.a[i] = p[i]; return v; }
inline Vec3 V3(double x, double y, double z) { Vec3 v; v[0] = x; v[1] = y; v[2] = z; return v; }
inline Vec2 V2(double x, double y) { Vec2 v; v[0] = x; v[1] = y; return v; }
inline Vec3 cross(const Vec3& a, const Vec3& b) { return V3(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]); }

// Generic SPD inverse / Cholesky for the small noise matrices (N <= 12).
// chol_upper: returns upper-triangular U with U^T U = A (GTSAM noiseModel::Gaussian keeps
// R = llt(information).matrixU(): SURVEY.md Appendix A.4).
template <int N> bool chol_upper(const Mat<N, N>& A, Mat<N, N>& U) {
  U = Mat<N, N>::Zero();
  for (int j = 0; j < N; j++) {
    double d = A(j, j);
    for (int k = 0; k < j; k++) d -= U(k, j) * U(k, j);
    if (!(d > 0)) return false;
    U(j, j) = std::sqrt(d);
    for (int c = j + 1; c < N; c++) {
      double s = A(j, c);
      for (int k = 0; k < j; k++) s -= U(k, j) * U(k, c);
      U(j, c) = s / U(j, j);
    }
  }
  return true;
}
template <int N> Mat<N, N> inverse_gj(Mat<N, N> A) {  // Gauss-Jordan with partial pivoting
  Mat<N, N> I = Mat<N, N>::Identity();
  for (int c = 0; c < N; c++) {
    int p = c; for (int r = c + 1; r < N; r++) if (std::fabs(A(r, c)) > std::fabs(A(p, c))) p = r;
    if (p != c) for (int k = 0; k < N; k++) { double t = A(c, k); A(c, k) = A(p, k); A(p, k) = t; t = I(c, k); I(c, k) = I(p, k); I(p, k) = t; }
    double d = 1.0 / A(c, c);
    for (int k = 0; k < N; k++) { A(c, k) *= d; I(c, k) *= d; }
    for (int r = 0; r < N; r++) if (r != c) { double f = A(r, c); if (f != 0) for (int k = 0; k < N; k++) { A(r, k) -= f * A(c, k); I(r, k) -= f * I(c, k); } }
  }
  return I;
}

}  // namespace gpo
