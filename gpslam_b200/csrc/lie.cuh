// Lie-group arithmetic for the batched factor kernels (SO(3), SE(3), SE(2)), FP64.
//
// Conventions are gpslam's / GTSAM's: right perturbations, Pose3 tangent
// (omega, v) rotation first, Pose2 tangent (vx, vy, omega).  Branch thresholds follow the
// reference so results agree with its CPU path:
//   gp/Pose3utils.cpp:203-224 (SO(3) right Jacobian and inverse, theta^2 <= eps -> I)
//   gp/Pose3utils.cpp:92-113  (SE(3) Q block, Taylor branch when theta <= 1e-5)
//
// The one deliberate departure: the reference differentiates rightJacobianPose3inv(xi)*v
// numerically (24 evaluations per factor, gp/Pose3utils.cpp:167-179).  Here the derivative
// is closed-form, from  Jr^-1(xi) = I + ad/2 + alpha(theta) ad^2 + beta(theta) ad^4
// (ad = ad_xi satisfies ad (ad^2 + theta^2)^2 = 0 on se(3)); it agrees with the reference's
// central differences to ~1e-8, well inside the reference's own 1e-6 Jacobian tolerance.
//
// Everything is __host__ __device__ so tests can run the same arithmetic on the CPU
// (tests/hostmath); the product only calls it from CUDA kernels.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define GPB_HD __host__ __device__ __forceinline__
#else
#define GPB_HD inline
#endif

namespace gpb {

struct V3 { double x, y, z; };
struct M3 { double m[9]; };  // row-major: m[3*r+c]

GPB_HD V3 v3(double x, double y, double z) { V3 v; v.x = x; v.y = y; v.z = z; return v; }
GPB_HD V3 operator+(const V3& a, const V3& b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
GPB_HD V3 operator-(const V3& a, const V3& b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
GPB_HD V3 operator-(const V3& a) { return v3(-a.x, -a.y, -a.z); }
GPB_HD V3 operator*(double s, const V3& a) { return v3(s * a.x, s * a.y, s * a.z); }
GPB_HD double dot(const V3& a, const V3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
GPB_HD V3 cross(const V3& a, const V3& b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
GPB_HD double elem(const V3& a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

GPB_HD M3 m3_zero() { M3 r; for (int i = 0; i < 9; i++) r.m[i] = 0.0; return r; }
GPB_HD M3 m3_identity() { M3 r = m3_zero(); r.m[0] = r.m[4] = r.m[8] = 1.0; return r; }
GPB_HD M3 skew(const V3& w) { M3 r; r.m[0] = 0; r.m[1] = -w.z; r.m[2] = w.y; r.m[3] = w.z; r.m[4] = 0; r.m[5] = -w.x; r.m[6] = -w.y; r.m[7] = w.x; r.m[8] = 0; return r; }
GPB_HD M3 operator+(const M3& a, const M3& b) { M3 r; for (int i = 0; i < 9; i++) r.m[i] = a.m[i] + b.m[i]; return r; }
GPB_HD M3 operator-(const M3& a, const M3& b) { M3 r; for (int i = 0; i < 9; i++) r.m[i] = a.m[i] - b.m[i]; return r; }
GPB_HD M3 operator-(const M3& a) { M3 r; for (int i = 0; i < 9; i++) r.m[i] = -a.m[i]; return r; }
GPB_HD M3 operator*(double s, const M3& a) { M3 r; for (int i = 0; i < 9; i++) r.m[i] = s * a.m[i]; return r; }
GPB_HD M3 operator*(const M3& a, const M3& b) {
  M3 r;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) r.m[3 * i + j] = a.m[3 * i] * b.m[j] + a.m[3 * i + 1] * b.m[3 + j] + a.m[3 * i + 2] * b.m[6 + j];
  return r;
}
GPB_HD V3 operator*(const M3& a, const V3& v) {
  return v3(a.m[0] * v.x + a.m[1] * v.y + a.m[2] * v.z, a.m[3] * v.x + a.m[4] * v.y + a.m[5] * v.z, a.m[6] * v.x + a.m[7] * v.y + a.m[8] * v.z);
}
GPB_HD M3 transpose(const M3& a) { M3 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[3 * i + j] = a.m[3 * j + i]; return r; }
GPB_HD V3 tmul(const M3& a, const V3& v) {  // a^T v
  return v3(a.m[0] * v.x + a.m[3] * v.y + a.m[6] * v.z, a.m[1] * v.x + a.m[4] * v.y + a.m[7] * v.z, a.m[2] * v.x + a.m[5] * v.y + a.m[8] * v.z);
}
GPB_HD M3 outer(const V3& a, const V3& b) { M3 r; r.m[0] = a.x * b.x; r.m[1] = a.x * b.y; r.m[2] = a.x * b.z; r.m[3] = a.y * b.x; r.m[4] = a.y * b.y; r.m[5] = a.y * b.z; r.m[6] = a.z * b.x; r.m[7] = a.z * b.y; r.m[8] = a.z * b.z; return r; }
// wire layout of a rotation: 9 doubles column-major
GPB_HD M3 m3_from_wire(const double* p) { M3 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[3 * i + j] = p[i + 3 * j]; return r; }
GPB_HD void m3_to_wire(const M3& a, double* p) { for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) p[i + 3 * j] = a.m[3 * i + j]; }

#define GPB_EPS 2.220446049250313e-16

// ---------------------------------------------------------------- SO(3)
GPB_HD M3 so3_expmap(const V3& w) {  // gtsam::SO3::Expmap
  const double theta2 = dot(w, w);
  const M3 W = skew(w);
  if (theta2 <= GPB_EPS) return m3_identity() + W;
  const double theta = sqrt(theta2);
  const double s = sin(theta), s2 = sin(0.5 * theta);
  const M3 K = (1.0 / theta) * W;
  return m3_identity() + s * K + (2.0 * s2 * s2) * (K * K);
}
GPB_HD V3 so3_logmap(const M3& R) {  // gtsam::SO3::Logmap
  const double tr = R.m[0] + R.m[4] + R.m[8];
  if (fabs(tr + 1.0) < 1e-10) {
    if (fabs(R.m[8] + 1.0) > 1e-10) return (M_PI / sqrt(2.0 + 2.0 * R.m[8])) * v3(R.m[2], R.m[5], 1.0 + R.m[8]);
    if (fabs(R.m[4] + 1.0) > 1e-10) return (M_PI / sqrt(2.0 + 2.0 * R.m[4])) * v3(R.m[1], 1.0 + R.m[4], R.m[7]);
    return (M_PI / sqrt(2.0 + 2.0 * R.m[0])) * v3(1.0 + R.m[0], R.m[3], R.m[6]);
  }
  double magnitude;
  const double tr_3 = tr - 3.0;
  if (tr_3 < -1e-7) {
    const double theta = acos((tr - 1.0) / 2.0);
    magnitude = theta / (2.0 * sin(theta));
  } else {
    magnitude = 0.5 - tr_3 / 12.0;
  }
  return magnitude * v3(R.m[7] - R.m[5], R.m[2] - R.m[6], R.m[3] - R.m[1]);
}
GPB_HD M3 so3_jr(const V3& w) {  // gp/Pose3utils.cpp:203-212
  const double theta2 = dot(w, w);
  if (theta2 <= GPB_EPS) return m3_identity();
  const double theta = sqrt(theta2);
  const M3 Y = (1.0 / theta) * skew(w);
  return m3_identity() - ((1 - cos(theta)) / theta) * Y + (1 - sin(theta) / theta) * (Y * Y);
}
GPB_HD M3 so3_jrinv(const V3& w) {  // gp/Pose3utils.cpp:215-224
  const double theta2 = dot(w, w);
  if (theta2 <= GPB_EPS) return m3_identity();
  const double theta = sqrt(theta2);
  const M3 X = skew(w);
  return m3_identity() + 0.5 * X + (1 / (theta * theta) - (1 + cos(theta)) / (2 * theta * sin(theta))) * (X * X);
}

// ---------------------------------------------------------------- SE(3)
struct P3 { M3 R; V3 t; };
GPB_HD P3 p3_from_wire(const double* p) { P3 T; T.R = m3_from_wire(p); T.t = v3(p[9], p[10], p[11]); return T; }
GPB_HD void p3_to_wire(const P3& T, double* p) { m3_to_wire(T.R, p); p[9] = T.t.x; p[10] = T.t.y; p[11] = T.t.z; }
GPB_HD P3 p3_compose(const P3& a, const P3& b) { P3 r; r.R = a.R * b.R; r.t = a.R * b.t + a.t; return r; }
GPB_HD P3 p3_between(const P3& a, const P3& b) { P3 r; r.R = transpose(a.R) * b.R; r.t = tmul(a.R, b.t - a.t); return r; }  // a^-1 b
GPB_HD P3 p3_inverse(const P3& a) { P3 r; r.R = transpose(a.R); r.t = -tmul(a.R, a.t); return r; }

struct X6 { V3 w, v; };  // se(3) vector (omega, v)
GPB_HD X6 x6(const V3& w, const V3& v) { X6 r; r.w = w; r.v = v; return r; }
GPB_HD X6 x6_from(const double* p) { return x6(v3(p[0], p[1], p[2]), v3(p[3], p[4], p[5])); }
GPB_HD X6 operator+(const X6& a, const X6& b) { return x6(a.w + b.w, a.v + b.v); }
GPB_HD X6 operator-(const X6& a, const X6& b) { return x6(a.w - b.w, a.v - b.v); }
GPB_HD X6 operator*(double s, const X6& a) { return x6(s * a.w, s * a.v); }
GPB_HD double elem(const X6& a, int i) { return i < 3 ? elem(a.w, i) : elem(a.v, i - 3); }
// ad_xi y = (w x a, v x a + w x b)
GPB_HD X6 ad_apply(const X6& xi, const X6& y) { return x6(cross(xi.w, y.w), cross(xi.v, y.w) + cross(xi.w, y.v)); }

GPB_HD P3 se3_expmap(const X6& xi) {  // gtsam::Pose3::Expmap
  P3 T;
  T.R = so3_expmap(xi.w);
  const double theta2 = dot(xi.w, xi.w);
  if (theta2 > GPB_EPS) {
    const V3 t_parallel = dot(xi.w, xi.v) * xi.w;
    const V3 wxv = cross(xi.w, xi.v);
    T.t = (1.0 / theta2) * (wxv - T.R * wxv + t_parallel);
  } else {
    T.t = xi.v;
  }
  return T;
}
GPB_HD X6 se3_logmap(const P3& p) {  // gtsam::Pose3::Logmap
  const V3 w = so3_logmap(p.R);
  const double t = sqrt(dot(w, w));
  if (t < 1e-10) return x6(w, p.t);
  const V3 n = (1.0 / t) * w;
  const double Tan = tan(0.5 * t);
  const V3 WT = cross(n, p.t);
  const V3 u = p.t - (0.5 * t) * WT + (1 - t / (2. * Tan)) * cross(n, WT);
  return x6(w, u);
}

// 6x6 matrices of the class [[A,0],[B,C]] (3x3 blocks).  ad matrices and SE(3) Jacobians have C == A.
struct L6 { M3 A, B, C; };
GPB_HD L6 operator*(const L6& x, const L6& y) { L6 r; r.A = x.A * y.A; r.B = x.B * y.A + x.C * y.B; r.C = x.C * y.C; return r; }
GPB_HD L6 operator+(const L6& x, const L6& y) { L6 r; r.A = x.A + y.A; r.B = x.B + y.B; r.C = x.C + y.C; return r; }
GPB_HD L6 operator-(const L6& x, const L6& y) { L6 r; r.A = x.A - y.A; r.B = x.B - y.B; r.C = x.C - y.C; return r; }
GPB_HD L6 operator*(double s, const L6& x) { L6 r; r.A = s * x.A; r.B = s * x.B; r.C = s * x.C; return r; }
GPB_HD L6 l6_ad(const X6& y) { L6 r; r.A = skew(y.w); r.B = skew(y.v); r.C = r.A; return r; }
GPB_HD X6 operator*(const L6& x, const X6& y) { return x6(x.A * y.w, x.B * y.w + x.C * y.v); }
GPB_HD double elem(const L6& x, int r, int c) {  // dense accessor
  if (r < 3) return c < 3 ? x.A.m[3 * r + c] : 0.0;
  return c < 3 ? x.B.m[3 * (r - 3) + c] : x.C.m[3 * (r - 3) + (c - 3)];
}
GPB_HD L6 l6_adjoint(const P3& T) { L6 r; r.A = T.R; r.B = skew(T.t) * T.R; r.C = T.R; return r; }  // gtsam::Pose3::AdjointMap

// gp/Pose3utils.cpp:92-113
GPB_HD M3 se3_Q(const X6& xi) {
  const double theta = sqrt(dot(xi.w, xi.w));
  const M3 X = skew(xi.w), Y = skew(xi.v);
  const M3 XY = X * Y, YX = Y * X, XYX = X * YX;
  if (fabs(theta) > 1e-5) {
    const double sin_theta = sin(theta), cos_theta = cos(theta);
    const double theta2 = theta * theta, theta3 = theta2 * theta, theta4 = theta3 * theta, theta5 = theta4 * theta;
    return -0.5 * Y + ((theta - sin_theta) / theta3) * (XY + YX - XYX)
        + ((1.0 - 0.5 * theta2 - cos_theta) / theta4) * (X * XY + YX * X - 3.0 * XYX)
        - (0.5 * ((1.0 - 0.5 * theta2 - cos_theta) / theta4 - 3.0 * (theta - sin_theta - theta3 / 6.0) / theta5)) * (XYX * X + X * XYX);
  }
  return -0.5 * Y + (1.0 / 6.0) * (XY + YX - XYX) + (1.0 / 24.0) * (X * XY + YX * X - 3.0 * XYX)
      - (0.5 * (1.0 / 24.0 + 3.0 / 120.0)) * (XYX * X + X * XYX);
}
GPB_HD L6 se3_jr(const X6& xi) {  // gp/Pose3utils.cpp:182-189
  L6 r; r.A = so3_jr(xi.w); r.B = se3_Q(xi); r.C = r.A; return r;
}

// Coefficients of Jr^-1 = I + ad/2 + alpha ad^2 + beta ad^4 and alpha'/theta, beta'/theta.
// Series below theta = 0.3 (truncation < 1e-15 relative), closed forms above (cancellation < 2e-13).
struct JrinvCoef { double alpha, beta, dalpha, dbeta; };
GPB_HD JrinvCoef se3_jrinv_coef(double theta2) {
  JrinvCoef c;
  if (theta2 < 0.09) {
    const double t2 = theta2, t4 = t2 * t2, t6 = t4 * t2, t8 = t4 * t4;
    c.alpha = 1.0 / 12 - t4 / 30240 - t6 / 604800 - t8 / 15966720;
    c.beta = -1.0 / 720 - t2 / 15120 - t4 / 403200 - t6 / 11975040 - 691.0 * t8 / 261534873600.0;
    c.dalpha = -t2 / 7560 - t4 / 100800 - t6 / 1995840 - 691.0 * t8 / 32691859200.0;
    c.dbeta = -1.0 / 7560 - t2 / 100800 - t4 / 1995840 - 691.0 * t6 / 32691859200.0 - t8 / 1245404160.0;
  } else {
    const double th = sqrt(theta2), h = 0.5 * th;
    const double sh = sin(h), ch = cos(h), cot = ch / sh, s2 = sh * sh;
    const double gam = h * cot, kap = -0.5 * cot + th / (4 * s2);
    c.beta = (2 * (1 - gam) / th - kap) / (2 * th * theta2);
    c.alpha = (1 - gam) / theta2 + c.beta * theta2;
    const double N = th * theta2 * ch / (sh * s2) + 3 * theta2 / s2 + 6 * th * cot - 32;
    c.dalpha = N / (8 * theta2 * theta2);
    c.dbeta = N / (8 * theta2 * theta2 * theta2);
  }
  return c;
}
// Jr^-1(xi) (== gp/Pose3utils.cpp:192-200 up to rounding) as an L6 with C == A.
GPB_HD L6 se3_jrinv(const X6& xi, const JrinvCoef& c) {
  const double theta2 = dot(xi.w, xi.w);
  L6 r;
  if (theta2 <= GPB_EPS) {
    // the reference's degenerate branch: rightJacobianRot3inv returns I (gp/Pose3utils.cpp:219) and the Q block
    // is the Taylor form (gp/Pose3utils.cpp:108-112), so Jr^-1 = [[I,0],[-Q,I]] there; kept for residual parity.
    r.A = m3_identity(); r.B = -se3_Q(xi); r.C = r.A;
    return r;
  }
  const M3 X = skew(xi.w), Y = skew(xi.v);
  const M3 X2 = X * X, S = X * Y + Y * X;
  r.A = m3_identity() + 0.5 * X + (c.alpha - c.beta * theta2) * X2;
  r.B = 0.5 * Y + c.alpha * S + c.beta * (X2 * S + S * X2);
  r.C = r.A;
  return r;
}
// D = d( Jr^-1(xi) v ) / d xi  — replaces jacobianMethodNumercialDiff(rightJacobianPose3inv, xi, v)
// (gp/GaussianProcessPriorPose3.h:81-82).  Blocks: [[D.A, 0],[D.B, D.C]].
GPB_HD L6 se3_djrinv(const X6& xi, const X6& v, const JrinvCoef& c) {
  const L6 A = l6_ad(xi);
  const X6 w1 = ad_apply(xi, v), w2 = ad_apply(xi, w1), w3 = ad_apply(xi, w2), w4 = ad_apply(xi, w3);
  const L6 adv = l6_ad(v);
  const L6 M = c.alpha * adv + c.beta * l6_ad(w2) + A * (c.beta * l6_ad(w1) + A * (c.beta * adv));
  L6 D = (-0.5) * adv - c.alpha * l6_ad(w1) - c.beta * l6_ad(w3) - A * M;
  const X6 g = c.dalpha * w2 + c.dbeta * w4;  // rank-one term  g [w^T, 0]
  D.A = D.A + outer(g.w, xi.w);
  D.B = D.B + outer(g.v, xi.w);
  return D;
}

// ---------------------------------------------------------------- SE(2)
struct P2 { double x, y, th; };
GPB_HD P2 p2(double x, double y, double th) { P2 p; p.x = x; p.y = y; p.th = th; return p; }
GPB_HD P2 p2_compose(const P2& a, const P2& b) { const double c = cos(a.th), s = sin(a.th); return p2(a.x + c * b.x - s * b.y, a.y + s * b.x + c * b.y, a.th + b.th); }
GPB_HD P2 p2_between(const P2& a, const P2& b) { const double c = cos(a.th), s = sin(a.th); const double dx = b.x - a.x, dy = b.y - a.y; return p2(c * dx + s * dy, -s * dx + c * dy, b.th - a.th); }
GPB_HD P2 p2_inverse(const P2& a) { const double c = cos(a.th), s = sin(a.th); return p2(-(c * a.x + s * a.y), -(-s * a.x + c * a.y), -a.th); }
GPB_HD double p2_theta(const P2& a) { return atan2(sin(a.th), cos(a.th)); }
GPB_HD M3 p2_adjoint(const P2& a) {  // gtsam::Pose2::AdjointMap
  const double c = cos(a.th), s = sin(a.th);
  M3 r = m3_identity(); r.m[0] = c; r.m[1] = -s; r.m[2] = a.y; r.m[3] = s; r.m[4] = c; r.m[5] = -a.x; return r;
}
GPB_HD P2 se2_expmap(const V3& xi) {  // gtsam::Pose2::Expmap
  const double w = xi.z;
  if (fabs(w) < 1e-10) return p2(xi.x, xi.y, xi.z);
  const double c = cos(w), s = sin(w);
  const double ox = -xi.y, oy = xi.x;
  return p2((ox - (c * ox - s * oy)) / w, (oy - (s * ox + c * oy)) / w, w);
}
GPB_HD V3 se2_logmap(const P2& p) {  // gtsam::Pose2::Logmap
  const double w = p2_theta(p);
  if (fabs(w) < 1e-10) return v3(p.x, p.y, w);
  const double c = cos(w), s = sin(w);
  const double c_1 = c - 1.0, det = c_1 * c_1 + s * s;
  const double ux = c * p.x + s * p.y - p.x, uy = -s * p.x + c * p.y - p.y;
  return v3((w / det) * (-uy), (w / det) * ux, w);
}
GPB_HD M3 se2_dexp(const V3& v) {  // gtsam::Pose2::ExpmapDerivative
  const double alpha = v.z;
  M3 J = m3_identity();
  if (fabs(alpha) > 1e-5) {
    const double sZ = sin(alpha) / alpha, cZ = (cos(alpha) - 1) / alpha;
    const double v1Z = v.x / alpha, v2Z = v.y / alpha;
    J.m[0] = sZ; J.m[1] = -cZ; J.m[2] = v1Z + v2Z * cZ - v1Z * sZ;
    J.m[3] = cZ; J.m[4] = sZ; J.m[5] = -v1Z * cZ + v2Z - v2Z * sZ;
  } else {
    J.m[2] = -0.5 * v.y; J.m[5] = 0.5 * v.x;
  }
  return J;
}
GPB_HD M3 se2_dlog(const V3& v) {  // gtsam::Pose2::LogmapDerivative, v = Logmap(p)
  const double alpha = v.z;
  M3 J = m3_identity();
  if (fabs(alpha) > 1e-5) {
    const double alphaInv = 1 / alpha;
    const double hc = 0.5 * sin(alpha) / (1 - cos(alpha));
    J.m[0] = alpha * hc; J.m[1] = -0.5 * alpha; J.m[2] = v.x * alphaInv - v.x * hc + 0.5 * v.y;
    J.m[3] = 0.5 * alpha; J.m[4] = alpha * hc; J.m[5] = v.y * alphaInv - 0.5 * v.x - v.y * hc;
  } else {
    J.m[2] = 0.5 * v.y; J.m[5] = -0.5 * v.x;
  }
  return J;
}

// gtsam::Unit3::basis(): columns b1, b2 of the tangent basis of unit vector n
GPB_HD void unit3_basis(const V3& n, V3& b1, V3& b2) {
  const double mx = fabs(n.x), my = fabs(n.y), mz = fabs(n.z);
  V3 axis = v3(0, 0, 1);
  if (mx <= my && mx <= mz) axis = v3(1, 0, 0);
  else if (my <= mx && my <= mz) axis = v3(0, 1, 0);
  b1 = cross(n, axis);
  b1 = (1.0 / sqrt(dot(b1, b1))) * b1;
  b2 = cross(n, b1);
}

}  // namespace gpb
