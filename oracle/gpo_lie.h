// ORACLE — TEST INFRASTRUCTURE ONLY (see gpo_linalg.h).
//
// CPU restatement of the GTSAM geometry the reference calls into (SURVEY.md §8c and
// Appendix A).  GTSAM is a third-party dependency of gtrll/gpslam (README.md:12,
// "GTSAM >= 4.0 alpha", version unpinned, not vendored: .SUBMODULES.json:8) and is NOT
// present in the build container, so its published algorithms are restated here and pinned
// by the reference's own unit tests (tests/test_oracle_golden.py ports them).
//
// Conventions (each pinned by a reference test):
//   * right perturbations  T (+) d = T * Exp(d)
//   * Pose3 tangent  xi = (omega, v)  rotation first   (gp/Pose3utils.cpp:116)
//   * Pose2 tangent  (vx, vy, omega)  translation first (gp/tests/testGaussianProcessPriorPose2.cpp:69-72)
//   * Rot3::Ypr(y,p,r) = Rz(y) Ry(p) Rx(r)              (gp/tests/testPose3Utils.cpp:126-133)
// Wire layouts: Rot3 = 9 doubles column-major; Pose3 = [R(9) col-major, t(3)];
// Pose2 = (x, y, theta).
#pragma once
#include "gpo_linalg.h"

namespace gpo {

// ---------------------------------------------------------------- SO(3)
inline Mat3 skew(const Vec3& w) {  // gtsam::skewSymmetric
  Mat3 m = Mat3::Zero();
  m(0, 1) = -w[2]; m(0, 2) = w[1]; m(1, 0) = w[2]; m(1, 2) = -w[0]; m(2, 0) = -w[1]; m(2, 1) = w[0];
  return m;
}

// gtsam::SO3::Expmap (Rodrigues; near-zero branch theta^2 <= eps -> I + W)
inline Mat3 so3_expmap(const Vec3& w) {
  const double theta2 = w.dot(w);
  const Mat3 W = skew(w);
  if (theta2 <= std::numeric_limits<double>::epsilon()) return Mat3::Identity() + W;
  const double theta = std::sqrt(theta2);
  const double s = std::sin(theta), s2 = std::sin(0.5 * theta);
  const double one_minus_cos = 2.0 * s2 * s2;
  const Mat3 K = W / theta;
  return Mat3::Identity() + s * K + one_minus_cos * (K * K);
}

// gtsam::SO3::Logmap (matrix form).  Near-identity Taylor uses the corrected series
// 1/2 - (tr-3)/12 (GTSAM >= 4.1); GTSAM 4.0 had (tr-3)^2/12 there, a difference of
// < 1e-11 rad on the branch's domain (theta < 3.2e-4) — stated, not pinned by any test.
inline Vec3 so3_logmap(const Mat3& R) {
  const double R11 = R(0, 0), R12 = R(0, 1), R13 = R(0, 2);
  const double R21 = R(1, 0), R22 = R(1, 1), R23 = R(1, 2);
  const double R31 = R(2, 0), R32 = R(2, 1), R33 = R(2, 2);
  const double tr = R11 + R22 + R33;
  if (std::fabs(tr + 1.0) < 1e-10) {
    if (std::fabs(R33 + 1.0) > 1e-10) return (M_PI / std::sqrt(2.0 + 2.0 * R33)) * V3(R13, R23, 1.0 + R33);
    if (std::fabs(R22 + 1.0) > 1e-10) return (M_PI / std::sqrt(2.0 + 2.0 * R22)) * V3(R12, 1.0 + R22, R32);
    return (M_PI / std::sqrt(2.0 + 2.0 * R11)) * V3(1.0 + R11, R21, R31);
  }
  double magnitude;
  const double tr_3 = tr - 3.0;
  if (tr_3 < -1e-7) {
    const double theta = std::acos((tr - 1.0) / 2.0);
    magnitude = theta / (2.0 * std::sin(theta));
  } else {
    magnitude = 0.5 - tr_3 / 12.0;
  }
  return magnitude * V3(R32 - R23, R13 - R31, R21 - R12);
}

// gp/Pose3utils.cpp:203-212  (== gtsam::SO3::ExpmapDerivative)
inline Mat3 rightJacobianRot3(const Vec3& omega) {
  const double theta2 = omega.dot(omega);
  if (theta2 <= std::numeric_limits<double>::epsilon()) return Mat3::Identity();
  const double theta = std::sqrt(theta2);
  const Mat3 Y = skew(omega) / theta;
  return Mat3::Identity() - ((1 - std::cos(theta)) / theta) * Y + (1 - std::sin(theta) / theta) * (Y * Y);
}
// gp/Pose3utils.cpp:215-224  (== gtsam::SO3::LogmapDerivative)
inline Mat3 rightJacobianRot3inv(const Vec3& omega) {
  const double theta2 = omega.dot(omega);
  if (theta2 <= std::numeric_limits<double>::epsilon()) return Mat3::Identity();
  const double theta = std::sqrt(theta2);
  const Mat3 X = skew(omega);
  return Mat3::Identity() + 0.5 * X + (1 / (theta * theta) - (1 + std::cos(theta)) / (2 * theta * std::sin(theta))) * (X * X);
}
// gp/Pose3utils.cpp:136-148
inline Mat3 leftJacobianRot3(const Vec3& omega) {
  const double theta2 = omega.dot(omega);
  if (theta2 <= std::numeric_limits<double>::epsilon()) return Mat3::Identity();
  const double theta = std::sqrt(theta2);
  const Vec3 dir = omega / theta;
  const double sin_theta = std::sin(theta);
  const Mat3 A = skew(omega) / theta;
  return (sin_theta / theta) * Mat3::Identity() + (1 - sin_theta / theta) * (dir * dir.t()) + ((1 - std::cos(theta)) / theta) * A;
}
// gp/Pose3utils.cpp:151-164
inline Mat3 leftJacobianRot3inv(const Vec3& omega) {
  const double theta2 = omega.dot(omega);
  if (theta2 <= std::numeric_limits<double>::epsilon()) return Mat3::Identity();
  const double theta = std::sqrt(theta2);
  const Vec3 dir = omega / theta;
  const double theta_2 = theta / 2.0;
  const double cot_theta_2 = 1.0 / std::tan(theta_2);
  const Mat3 A = skew(omega) / theta;
  return (theta_2 * cot_theta_2) * Mat3::Identity() + (1 - theta_2 * cot_theta_2) * (dir * dir.t()) - theta_2 * A;
}

inline Mat3 rot_ypr(double y, double p, double r) {  // gtsam::Rot3::Ypr = Rz(y) Ry(p) Rx(r)
  const double cy = std::cos(y), sy = std::sin(y), cp = std::cos(p), sp = std::sin(p), cr = std::cos(r), sr = std::sin(r);
  Mat3 Rz = Mat3::Identity(), Ry = Mat3::Identity(), Rx = Mat3::Identity();
  Rz(0, 0) = cy; Rz(0, 1) = -sy; Rz(1, 0) = sy; Rz(1, 1) = cy;
  Ry(0, 0) = cp; Ry(0, 2) = sp; Ry(2, 0) = -sp; Ry(2, 2) = cp;
  Rx(1, 1) = cr; Rx(1, 2) = -sr; Rx(2, 1) = sr; Rx(2, 2) = cr;
  return Rz * Ry * Rx;
}

// ---------------------------------------------------------------- SE(3)
struct Pose3 {
  Mat3 R; Vec3 t;
  Pose3() : R(Mat3::Identity()), t(Vec3::Zero()) {}
  Pose3(const Mat3& R_, const Vec3& t_) : R(R_), t(t_) {}
  static Pose3 from(const double* p) { Pose3 T; for (int i = 0; i < 9; i++) T.R.a[i] = p[i]; for (int i = 0; i < 3; i++) T.t.a[i] = p[9 + i]; return T; }
  void to(double* p) const { for (int i = 0; i < 9; i++) p[i] = R.a[i]; for (int i = 0; i < 3; i++) p[9 + i] = t.a[i]; }
  Pose3 inverse() const { Mat3 Rt = R.t(); return Pose3(Rt, -(Rt * t)); }
  Pose3 compose(const Pose3& o) const { return Pose3(R * o.R, R * o.t + t); }
  // gtsam::Pose3::AdjointMap = [[R,0],[[t]x R, R]]
  Mat6 Adjoint() const {
    Mat6 A = Mat6::Zero();
    A.set(0, 0, R); A.set(3, 0, skew(t) * R); A.set(3, 3, R);
    return A;
  }
};
// compose(A,B): d/dA = Ad(B^-1), d/dB = I ; inverse(A): -Ad(A)   (gtsam LieGroup)
inline Mat6 pose3_Hcompose1(const Pose3& B) { return B.inverse().Adjoint(); }
inline Mat6 pose3_Hinverse(const Pose3& A) { return -A.Adjoint(); }

// gtsam::Pose3::Expmap
inline Pose3 pose3_expmap(const Vec6& xi) {
  const Vec3 omega = V3(xi[0], xi[1], xi[2]), v = V3(xi[3], xi[4], xi[5]);
  const Mat3 R = so3_expmap(omega);
  const double theta2 = omega.dot(omega);
  if (theta2 > std::numeric_limits<double>::epsilon()) {
    const Vec3 t_parallel = omega * omega.dot(v);
    const Vec3 omega_cross_v = cross(omega, v);
    const Vec3 t = (omega_cross_v - R * omega_cross_v + t_parallel) / theta2;
    return Pose3(R, t);
  }
  return Pose3(R, v);
}
// gtsam::Pose3::Logmap (Agrawal06iros eq. 14)
inline Vec6 pose3_logmap(const Pose3& p) {
  const Vec3 w = so3_logmap(p.R);
  const Vec3 T = p.t;
  const double t = w.norm();
  Vec6 out;
  if (t < 1e-10) {
    for (int i = 0; i < 3; i++) { out[i] = w[i]; out[3 + i] = T[i]; }
    return out;
  }
  const Mat3 W = skew(w / t);
  const double Tan = std::tan(0.5 * t);
  const Vec3 WT = W * T;
  const Vec3 u = T - (0.5 * t) * WT + (1 - t / (2. * Tan)) * (W * WT);
  for (int i = 0; i < 3; i++) { out[i] = w[i]; out[3 + i] = u[i]; }
  return out;
}

// gp/Pose3utils.cpp:92-113  (== gtsam computeQforExpmapDerivative)
inline Mat3 rightJacobianPose3Q(const Vec6& xi) {
  const Vec3 omega = V3(xi[0], xi[1], xi[2]), rho = V3(xi[3], xi[4], xi[5]);
  const double theta = omega.norm();
  const Mat3 X = skew(omega), Y = skew(rho);
  const Mat3 XY = X * Y, YX = Y * X, XYX = X * YX;
  if (std::fabs(theta) > 1e-5) {
    const double sin_theta = std::sin(theta), cos_theta = std::cos(theta);
    const double theta2 = theta * theta, theta3 = theta2 * theta, theta4 = theta3 * theta, theta5 = theta4 * theta;
    return -0.5 * Y + ((theta - sin_theta) / theta3) * (XY + YX - XYX)
        + ((1.0 - 0.5 * theta2 - cos_theta) / theta4) * (X * XY + YX * X - 3.0 * XYX)
        - (0.5 * ((1.0 - 0.5 * theta2 - cos_theta) / theta4 - 3.0 * (theta - sin_theta - theta3 / 6.0) / theta5)) * (XYX * X + X * XYX);
  }
  return -0.5 * Y + (1.0 / 6.0) * (XY + YX - XYX) + (1.0 / 24.0) * (X * XY + YX * X - 3.0 * XYX)
      - (0.5 * (1.0 / 24.0 + 3.0 / 120.0)) * (XYX * X + X * XYX);
}
// gp/Pose3utils.cpp:68-89
inline Mat3 leftJacobianPose3Q(const Vec6& xi) {
  const Vec3 omega = V3(xi[0], xi[1], xi[2]), rho = V3(xi[3], xi[4], xi[5]);
  const double theta = omega.norm();
  const Mat3 X = skew(omega), Y = skew(rho);
  const Mat3 XY = X * Y, YX = Y * X, XYX = X * YX;
  if (std::fabs(theta) > 1e-5) {
    const double sin_theta = std::sin(theta), cos_theta = std::cos(theta);
    const double theta2 = theta * theta, theta3 = theta2 * theta, theta4 = theta3 * theta, theta5 = theta4 * theta;
    return 0.5 * Y + ((theta - sin_theta) / theta3) * (XY + YX + XYX)
        - ((1.0 - 0.5 * theta2 - cos_theta) / theta4) * (X * XY + YX * X - 3.0 * XYX)
        - (0.5 * ((1.0 - 0.5 * theta2 - cos_theta) / theta4 - 3.0 * (theta - sin_theta - theta3 / 6.0) / theta5)) * (XYX * X + X * XYX);
  }
  return 0.5 * Y + (1.0 / 6.0) * (XY + YX + XYX) - (1.0 / 24.0) * (X * XY + YX * X - 3.0 * XYX)
      - (0.5 * (1.0 / 24.0 + 3.0 / 120.0)) * (XYX * X + X * XYX);
}
inline Mat6 blk6(const Mat3& a, const Mat3& c, const Mat3& d) {  // [[a,0],[c,d]]
  Mat6 J = Mat6::Zero(); J.set(0, 0, a); J.set(3, 0, c); J.set(3, 3, d); return J;
}
// gp/Pose3utils.cpp:182-189  (== gtsam::Pose3::ExpmapDerivative)
inline Mat6 rightJacobianPose3(const Vec6& xi) {
  const Vec3 w = V3(xi[0], xi[1], xi[2]);
  const Mat3 Jw = rightJacobianRot3(w);
  return blk6(Jw, rightJacobianPose3Q(xi), Jw);
}
// gp/Pose3utils.cpp:192-200  (== gtsam::Pose3::LogmapDerivative)
inline Mat6 rightJacobianPose3inv(const Vec6& xi) {
  const Vec3 w = V3(xi[0], xi[1], xi[2]);
  const Mat3 Jw = rightJacobianRot3inv(w);
  const Mat3 Q = rightJacobianPose3Q(xi);
  return blk6(Jw, -(Jw * Q * Jw), Jw);
}
// gp/Pose3utils.cpp:116-123, 126-133
inline Mat6 leftJacobianPose3(const Vec6& xi) {
  const Vec3 w = V3(xi[0], xi[1], xi[2]);
  const Mat3 J = leftJacobianRot3(w);
  return blk6(J, leftJacobianPose3Q(xi), J);
}
inline Mat6 leftJacobianPose3inv(const Vec6& xi) {
  const Vec3 w = V3(xi[0], xi[1], xi[2]);
  const Mat3 Jinv = leftJacobianRot3inv(w);
  return blk6(Jinv, -(Jinv * leftJacobianPose3Q(xi) * Jinv), Jinv);
}
// gp/Pose3utils.cpp:167-179 — central difference of f(xi) * x, dxi = 1e-6 by default
// (gp/Pose3utils.h declares the default).  12 evaluations of func.
template <class F> Mat6 jacobianMethodNumercialDiff(F func, const Vec6& xi, const Vec6& x, double dxi = 1e-6) {
  Mat6 Diff = Mat6::Zero();
  for (int i = 0; i < 6; i++) {
    Vec6 xi_dxip = xi, xi_dxin = xi;
    xi_dxip[i] += dxi;
    const Mat6 Jdiffp = func(xi_dxip);
    xi_dxin[i] -= dxi;
    const Mat6 Jdiffn = func(xi_dxin);
    const Vec6 col = ((Jdiffp - Jdiffn) / (2.0 * dxi)) * x;
    for (int r = 0; r < 6; r++) Diff(r, i) = col[r];
  }
  return Diff;
}
// gp/Pose3utils.cpp:17-24
inline Vec6 getBodyCentricVb(const Pose3& p1, const Pose3& p2, double dt) { return pose3_logmap(p1.inverse().compose(p2)) / dt; }
inline Vec6 getBodyCentricVs(const Pose3& p1, const Pose3& p2, double dt) { return pose3_logmap(p2.compose(p1.inverse())) / dt; }

// gtsam::Pose3::range(point, H1, H2): q = R^T (p - t), r = |q|,
// d/dT = q^T/|q| [ [q]x , -I ],  d/dp = q^T/|q| R^T
inline double pose3_range(const Pose3& T, const Vec3& p, Mat<1, 6>* H1, Mat<1, 3>* H2) {
  const Mat3 Rt = T.R.t();
  const Vec3 q = Rt * (p - T.t);
  const double r = q.norm();
  if (H1 || H2) {
    const Mat<1, 3> D = q.t() / r;
    if (H1) { Mat<3, 6> Dp = Mat<3, 6>::Zero(); Dp.set(0, 0, skew(q)); Dp.set(0, 3, -Mat3::Identity()); *H1 = D * Dp; }
    if (H2) *H2 = D * Rt;
  }
  return r;
}

// gtsam::Pose3::translation(H): t, d/dT = [0, R]
inline Vec3 pose3_translation(const Pose3& T, Mat<3, 6>* H) {
  if (H) { *H = Mat<3, 6>::Zero(); H->set(0, 3, T.R); }
  return T.t;
}

// gtsam::PinholeCamera<Cal3_S2>(pose, K).project(point, Dpose, Dpoint)  (GTSAM 4.0 PinholePose / CalibratedCamera):
// q = R^T (p - t); cheirality: q.z <= 0 -> behind the camera (returns false; GTSAM throws CheiralityException);
// pn = (q.x, q.y) / q.z; pi = (fx pn.x + s pn.y + u0, fy pn.y + v0);
// Dpn/Dpose = [[uv, -1-uu, v, -d, 0, du], [1+vv, -uv, -u, 0, -d, dv]], d = 1/q.z;  Dpn/Dpoint = d [[1,0,-u],[0,1,-v]] R^T.
// K = (fx, fy, s, u0, v0).
inline bool pinhole_project(const Pose3& T, const double* K, const Vec3& p, Vec2& pi, Mat<2, 6>* Dpose, Mat<2, 3>* Dpoint) {
  const Mat3 Rt = T.R.t();
  const Vec3 q = Rt * (p - T.t);
  if (q[2] <= 0) return false;
  const double d = 1.0 / q[2], u = q[0] * d, v = q[1] * d;
  pi[0] = K[0] * u + K[2] * v + K[3];
  pi[1] = K[1] * v + K[4];
  Mat2 Dpi = Mat2::Zero();
  Dpi(0, 0) = K[0]; Dpi(0, 1) = K[2]; Dpi(1, 1) = K[1];
  if (Dpose) {
    Mat<2, 6> Dn;
    Dn(0, 0) = u * v; Dn(0, 1) = -1 - u * u; Dn(0, 2) = v; Dn(0, 3) = -d; Dn(0, 4) = 0; Dn(0, 5) = d * u;
    Dn(1, 0) = 1 + v * v; Dn(1, 1) = -u * v; Dn(1, 2) = -u; Dn(1, 3) = 0; Dn(1, 4) = -d; Dn(1, 5) = d * v;
    *Dpose = Dpi * Dn;
  }
  if (Dpoint) {
    Mat<2, 3> Dq;
    Dq(0, 0) = d; Dq(0, 1) = 0; Dq(0, 2) = -d * u; Dq(1, 0) = 0; Dq(1, 1) = d; Dq(1, 2) = -d * v;
    *Dpoint = Dpi * (Dq * Rt);
  }
  return true;
}

// ---------------------------------------------------------------- SE(2)
struct Pose2 {
  double x, y, th;
  Pose2() : x(0), y(0), th(0) {}
  Pose2(double x_, double y_, double th_) : x(x_), y(y_), th(th_) {}
  double c() const { return std::cos(th); }
  double s() const { return std::sin(th); }
  // gtsam stores Rot2 as (c,s); theta() = atan2(s,c) in (-pi, pi]
  double theta() const { return std::atan2(std::sin(th), std::cos(th)); }
  Pose2 inverse() const { const double C = c(), S = s(); return Pose2(-(C * x + S * y), -(-S * x + C * y), -th); }
  Pose2 compose(const Pose2& o) const { const double C = c(), S = s(); return Pose2(x + C * o.x - S * o.y, y + S * o.x + C * o.y, th + o.th); }
  // gtsam::Pose2::AdjointMap = [[c,-s, y],[s, c,-x],[0,0,1]]
  Mat3 Adjoint() const {
    const double C = c(), S = s();
    Mat3 A = Mat3::Identity();
    A(0, 0) = C; A(0, 1) = -S; A(0, 2) = y; A(1, 0) = S; A(1, 1) = C; A(1, 2) = -x;
    return A;
  }
};
inline Pose2 pose2_expmap(const Vec3& xi) {  // gtsam::Pose2::Expmap
  const double w = xi[2];
  if (std::fabs(w) < 1e-10) return Pose2(xi[0], xi[1], xi[2]);
  const double C = std::cos(w), S = std::sin(w);
  const double ox = -xi[1], oy = xi[0];              // v_ortho = R_PI_2 * v
  const double rx = C * ox - S * oy, ry = S * ox + C * oy;
  return Pose2((ox - rx) / w, (oy - ry) / w, w);
}
inline Vec3 pose2_logmap(const Pose2& p) {  // gtsam::Pose2::Logmap
  const double w = p.theta();
  if (std::fabs(w) < 1e-10) return V3(p.x, p.y, w);
  const double C = std::cos(w), S = std::sin(w);
  const double c_1 = C - 1.0, det = c_1 * c_1 + S * S;
  const double ux = C * p.x + S * p.y - p.x, uy = -S * p.x + C * p.y - p.y;  // R.unrotate(t) - t
  const double px = -uy, py = ux;                                             // R_PI_2 * (.)
  return V3((w / det) * px, (w / det) * py, w);
}
inline Mat3 pose2_ExpmapDerivative(const Vec3& v) {  // gtsam::Pose2::ExpmapDerivative
  const double alpha = v[2];
  Mat3 J = Mat3::Identity();
  if (std::fabs(alpha) > 1e-5) {
    const double sZalpha = std::sin(alpha) / alpha, c_1Zalpha = (std::cos(alpha) - 1) / alpha;
    const double v1Zalpha = v[0] / alpha, v2Zalpha = v[1] / alpha;
    J(0, 0) = sZalpha; J(0, 1) = -c_1Zalpha; J(0, 2) = v1Zalpha + v2Zalpha * c_1Zalpha - v1Zalpha * sZalpha;
    J(1, 0) = c_1Zalpha; J(1, 1) = sZalpha; J(1, 2) = -v1Zalpha * c_1Zalpha + v2Zalpha - v2Zalpha * sZalpha;
  } else {
    J(0, 2) = -0.5 * v[1]; J(1, 2) = 0.5 * v[0];
  }
  return J;
}
inline Mat3 pose2_LogmapDerivative(const Pose2& p) {  // gtsam::Pose2::LogmapDerivative
  const Vec3 v = pose2_logmap(p);
  const double alpha = v[2];
  Mat3 J = Mat3::Identity();
  if (std::fabs(alpha) > 1e-5) {
    const double alphaInv = 1 / alpha;
    const double halfCotHalfAlpha = 0.5 * std::sin(alpha) / (1 - std::cos(alpha));
    const double v1 = v[0], v2 = v[1];
    J(0, 0) = alpha * halfCotHalfAlpha; J(0, 1) = -0.5 * alpha; J(0, 2) = v1 * alphaInv - v1 * halfCotHalfAlpha + 0.5 * v2;
    J(1, 0) = 0.5 * alpha; J(1, 1) = alpha * halfCotHalfAlpha; J(1, 2) = v2 * alphaInv - 0.5 * v1 - v2 * halfCotHalfAlpha;
  } else {
    J(0, 2) = 0.5 * v[1]; J(1, 2) = -0.5 * v[0];
  }
  return J;
}
// gtsam::Pose2::range(point,H1,H2): d = p - t; H1 = d^T/|d| [[-c, s, 0],[-s,-c,0]]; H2 = d^T/|d|
inline double pose2_range(const Pose2& T, const Vec2& p, Mat<1, 3>* H1, Mat<1, 2>* H2) {
  const double dx = p[0] - T.x, dy = p[1] - T.y;
  const double r = std::sqrt(dx * dx + dy * dy);
  if (H1 || H2) {
    Mat<1, 2> D; D[0] = dx / r; D[1] = dy / r;
    if (H1) {
      const double C = T.c(), S = T.s();
      Mat<2, 3> Dp = Mat<2, 3>::Zero();
      Dp(0, 0) = -C; Dp(0, 1) = S; Dp(1, 0) = -S; Dp(1, 1) = -C;
      *H1 = D * Dp;
    }
    if (H2) *H2 = D;
  }
  return r;
}
// gtsam::Point2::norm(H): H = d^T/|d|, or (1,1) when |d| <= 1e-10
inline double point2_norm(const Vec2& d, Mat<1, 2>* H) {
  const double r = std::sqrt(d[0] * d[0] + d[1] * d[1]);
  if (H) { if (std::fabs(r) > 1e-10) { (*H)[0] = d[0] / r; (*H)[1] = d[1] / r; } else { (*H)[0] = 1; (*H)[1] = 1; } }
  return r;
}

}  // namespace gpo
