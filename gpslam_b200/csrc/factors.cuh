// Per-factor residual + Jacobian arithmetic of the gpslam factor types, written for one
// thread per factor (all loops compile-time unrolled so every small matrix stays in registers).
//
// Reference semantics reproduced (file:line in /root/reference/gpslam):
//   gp/GaussianProcessPriorPose3.h:60-98, ...Pose2.h:58-82, ...Rot3.h:58-79, ...Linear.h:63-83
//   gp/GaussianProcessInterpolatorPose3.h:57-105 (+Pose2/Rot3/Linear), GPutils.h:54-71
//   slam/GPInterpolatedRangeFactor{Pose3,Pose2,2DLinear}.h, slam/GPInterpolatedAttitudeFactorRot3.h
//   slam/{RangeFactor2DLinear,RangeBearingFactor2DLinear,OdometryFactor2DLinear,RangeFactorPose2}.h
//   GTSAM PriorFactor / BetweenFactor
//
// Output convention: the whitened GTSAM JacobianFactor payload  A = R*H,  b = -R*e  with R the
// upper-triangular square-root information.  For GP priors R = chol(calcQ_inv(Qc,dt)) =
// U (x) Rq with U = chol([[12/dt^3,-6/dt^2],[-6/dt^2,4/dt]]) and Rq = chol(Qc^-1)
// (gp/GPutils.h:33-41, gp/GaussianProcessPriorPose3.h:46).
#pragma once
#include "lie.cuh"

namespace gpb {

// G_POSE3VW: SE(3) states whose velocity is [v_world(3) | w_world(3)] (the reference's "VW" family, gp/GaussianProcessPriorPose3VW.h);
// a kernel class of G_POSE3, not a graph group of its own - record layout, block sizes, assembly, solver and retraction are G_POSE3's.
enum Group { G_POSE3 = 0, G_POSE2 = 1, G_ROT3 = 2, G_LINEAR = 3, G_POSE3VW = 4 };

// Layout of the whitened GP-prior JacobianFactors [A|b] in HBM: tiles of AB_TF consecutive factors; inside a tile
// [row pair p = column * D + rp][factor in tile][2 doubles].  A linearise CTA (AB_TF threads) writes ONE contiguous tile
// (NP x 2 KB: its 128-bit stores are base + compile-time offsets, and the tile sits in one or two 2 MB pages instead of NP of
// them); an assembly CTA reads 16-byte entries of 8-9 consecutive factors per row pair, all inside one tile.
constexpr int AB_TF = 128;
GPB_HD size_t ab_off(int p, int f, int NP) { return ((size_t)(f / AB_TF) * NP + p) * (size_t)(AB_TF * 2) + (size_t)(f % AB_TF) * 2; }

template <int I, int N, class F> GPB_HD void static_for(F&& f) {
  if constexpr (I < N) { f(std::integral_constant<int, I>{}); static_for<I + 1, N>(f); }
}

template <int G> struct GroupTraits;
template <> struct GroupTraits<G_POSE3> { static constexpr int D = 6, PS = 12, DL = 3; };
template <> struct GroupTraits<G_POSE3VW> { static constexpr int D = 6, PS = 12, DL = 3; };
template <> struct GroupTraits<G_POSE2> { static constexpr int D = 3, PS = 3, DL = 2; };
template <> struct GroupTraits<G_ROT3> { static constexpr int D = 3, PS = 9, DL = 0; };
template <> struct GroupTraits<G_LINEAR> { static constexpr int D = 3, PS = 3, DL = 2; };  // Linear<3> ("2DLinear" states)

// chol of the 2x2 GP kernel inverse: u11,u12,u22
struct GpWhiten { double u11, u12, u22; };
GPB_HD GpWhiten gp_whiten(double dt) {
  GpWhiten w;
  w.u11 = sqrt(12.0 / (dt * dt * dt));
  w.u12 = (-6.0 / (dt * dt)) / w.u11;
  w.u22 = sqrt(4.0 / dt - w.u12 * w.u12);
  return w;
}
// Hermite interpolation scalars: Lambda_12, Psi_11, Psi_12 of gp/GPutils.h:54-71
struct InterpCoef { double lam12, psi11, psi12; };
GPB_HD InterpCoef interp_coef(double dt, double tau) {
  const double s = tau / dt, s2 = s * s, s3 = s2 * s;
  InterpCoef c;
  c.psi11 = 3 * s2 - 2 * s3;
  c.psi12 = dt * (s3 - s2);
  c.lam12 = tau - c.psi11 * dt - c.psi12;
  return c;
}

// y = Rq x for upper-triangular Rq (DxD column-major), vectors as arrays
template <int D> GPB_HD void triu_mul(const double* Rq, const double* x, double* y) {
#pragma unroll
  for (int r = 0; r < D; r++) {
    double s = 0;
#pragma unroll
    for (int k = r; k < D; k++) s += Rq[r + k * D] * x[k];
    y[r] = s;
  }
}
// whiten one (top,bot) column pair of a GP prior: out[0:D] = Rq(u11 top + u12 bot), out[D:2D] = u22 Rq bot
template <int D> GPB_HD void gp_whiten_col(const GpWhiten& w, const double* Rq, const double* top, const double* bot, double sign, double* out) {
  double t[D], y[D];
#pragma unroll
  for (int k = 0; k < D; k++) t[k] = sign * (w.u11 * top[k] + w.u12 * bot[k]);
  triu_mul<D>(Rq, t, y);
#pragma unroll
  for (int k = 0; k < D; k++) out[k] = y[k];
#pragma unroll
  for (int k = 0; k < D; k++) t[k] = sign * w.u22 * bot[k];
  triu_mul<D>(Rq, t, y);
#pragma unroll
  for (int k = 0; k < D; k++) out[D + k] = y[k];
}

// ================================================================= GP prior, SE(3)
struct GpPose3 {
  X6 e_top, e_bot;
  L6 a, b, Da, Db;  // a = Jl^-1(r) (H1_top = -a), b = Jr^-1(r) (H3_top), Da = D a, Db = D b
};
// state record: [pose(12) | vel(6)]
GPB_HD void gp_prior_pose3_eval(const double* s1, const double* s2, double dt, bool wantJ, GpPose3& o) {
  const P3 T1 = p3_from_wire(s1), T2 = p3_from_wire(s2);
  const X6 v1 = x6_from(s1 + 12), v2 = x6_from(s2 + 12);
  const X6 r = se3_logmap(p3_between(T1, T2));
  const JrinvCoef c = se3_jrinv_coef(dot(r.w, r.w));
  o.b = se3_jrinv(r, c);
  o.e_top = r - dt * v1;
  o.e_bot = o.b * v2 - v1;
  if (wantJ) {
    o.a = o.b - l6_ad(r);  // Jl^-1 = Jr^-1 - ad
    const L6 D = se3_djrinv(r, v2, c);
    o.Da = D * o.a;
    o.Db = D * o.b;
  }
}
// whitened column c (0..24; 24 = rhs) of the 12x25 [A|b]; CV must be a compile-time constant after unrolling
template <int VAR, int C> GPB_HD void gp_prior_pose3_col(const GpPose3& o, const GpWhiten& w, const double* Rq, double dt, double* out) {
  double top[6], bot[6];
  if (VAR == 0) {  // T1: [-a ; -Da]
#pragma unroll
    for (int k = 0; k < 6; k++) { top[k] = elem(o.a, k, C); bot[k] = elem(o.Da, k, C); }
    gp_whiten_col<6>(w, Rq, top, bot, -1.0, out);
  } else if (VAR == 1) {  // v1: [-dt I ; -I]
#pragma unroll
    for (int k = 0; k < 6; k++) { top[k] = (k == C) ? dt : 0.0; bot[k] = (k == C) ? 1.0 : 0.0; }
    gp_whiten_col<6>(w, Rq, top, bot, -1.0, out);
  } else if (VAR == 2) {  // T2: [b ; Db]
#pragma unroll
    for (int k = 0; k < 6; k++) { top[k] = elem(o.b, k, C); bot[k] = elem(o.Db, k, C); }
    gp_whiten_col<6>(w, Rq, top, bot, 1.0, out);
  } else if (VAR == 3) {  // v2: [0 ; b]
#pragma unroll
    for (int k = 0; k < 6; k++) { top[k] = 0.0; bot[k] = elem(o.b, k, C); }
    gp_whiten_col<6>(w, Rq, top, bot, 1.0, out);
  } else {  // rhs = -R e
#pragma unroll
    for (int k = 0; k < 6; k++) { top[k] = elem(o.e_top, k); bot[k] = elem(o.e_bot, k); }
    gp_whiten_col<6>(w, Rq, top, bot, -1.0, out);
  }
}

// ---- production emitter of the SE(3) prior's 25 whitened columns (k_lin_gp<G_POSE3>).  Same arithmetic as
// gp_prior_pose3_eval + gp_prior_pose3_col, organised around what is structurally zero and what is live:
//  * every Jacobian block is [[A,0],[B,C]]: the translation columns (C >= 3) have zero rotation rows, so their whitening
//    sums start at k = 3; the v1 columns are plain columns of Rq; the v2 columns need ONE product Rq b[:,C];
//  * DIAG: every Qc model of the graph is diagonal (the reference's tests and scripts use Qc = sigma^2 I throughout), so Rq is
//    diagonal and whitening a column is an element-wise scale - identical results, ~half the FP64 work of the kernel;
//  * order rhs, v1, v2, T2, T1: b and D are live throughout, D b only for T2, a = b - ad(r) and D a only for T1 (the
//    all-at-once struct keeps four 6x6 blocks live and pins the kernel at 255 registers).
// sink(c, col): column c (0..24, variable order T1 v1 T2 v2 rhs) as 12 doubles.
template <int D, int K0, bool DIAG> GPB_HD void rq_mul(const double* __restrict__ Rq, const double* t, double* y) {
#pragma unroll
  for (int r = 0; r < D; r++) {
    if (DIAG) {
      y[r] = r >= K0 ? Rq[r + r * D] * t[r] : 0.0;
    } else {
      const int k0 = r > K0 ? r : K0;
      double s = Rq[r + k0 * D] * t[k0];
#pragma unroll
      for (int k = k0 + 1; k < D; k++) s += Rq[r + k * D] * t[k];
      y[r] = s;
    }
  }
}
// column C of sign * [[u11 I, u12 I],[0, u22 I]] (x) Rq applied to [Top; Bot][:, C]
template <int C, bool DIAG> GPB_HD void gp_col_pair(const L6& Top, const L6& Bot, double sign, const GpWhiten& w, const double* __restrict__ Rq, double* out) {
  constexpr int K0 = C < 3 ? 0 : 3;
  double t[6], tb[6];
#pragma unroll
  for (int k = 0; k < 6; k++) {
    if (k >= K0) { const double tp = elem(Top, k, C), bt = elem(Bot, k, C); t[k] = sign * (w.u11 * tp + w.u12 * bt); tb[k] = (sign * w.u22) * bt; }
    else { t[k] = 0.0; tb[k] = 0.0; }
  }
  rq_mul<6, K0, DIAG>(Rq, t, out);
  rq_mul<6, K0, DIAG>(Rq, tb, out + 6);
}
template <bool DIAG, class Sink>
GPB_HD double gp_prior_pose3_emit(const double* s1, const double* s2, double dt, bool wantJ, const GpWhiten& w, const double* __restrict__ Rq, Sink&& sink) {
  const P3 T1 = p3_from_wire(s1), T2 = p3_from_wire(s2);
  const X6 v1 = x6_from(s1 + 12), v2 = x6_from(s2 + 12);
  const X6 r = se3_logmap(p3_between(T1, T2));
  const JrinvCoef c = se3_jrinv_coef(dot(r.w, r.w));
  const L6 b = se3_jrinv(r, c);
  double col[12];
  double err = 0.0;
  {  // rhs = -R e,  e = [r - dt v1; Jr^-1(r) v2 - v1]
    const X6 et = r - dt * v1, eb = b * v2 - v1;
    double t[6], tb[6];
#pragma unroll
    for (int k = 0; k < 6; k++) { t[k] = -(w.u11 * elem(et, k) + w.u12 * elem(eb, k)); tb[k] = -w.u22 * elem(eb, k); }
    rq_mul<6, 0, DIAG>(Rq, t, col); rq_mul<6, 0, DIAG>(Rq, tb, col + 6);
#pragma unroll
    for (int k = 0; k < 12; k++) err += col[k] * col[k];
    if (!wantJ) return err;
    sink(24, col);
  }
  {  // v1: [-dt I; -I] -> column C of Rq scaled
    const double st = -(w.u11 * dt + w.u12), sb = -w.u22;
    static_for<0, 6>([&](auto cc) {
      constexpr int C = decltype(cc)::value;
#pragma unroll
      for (int rr = 0; rr < 6; rr++) {
        if (DIAG ? rr == C : rr <= C) { const double q = Rq[rr + C * 6]; col[rr] = st * q; col[6 + rr] = sb * q; }
        else { col[rr] = 0.0; col[6 + rr] = 0.0; }
      }
      sink(6 + C, col);
    });
  }
  // v2: [0; b] -> one product y = Rq b[:, C], top u12 y, bottom u22 y
  static_for<0, 6>([&](auto cc) {
    constexpr int C = decltype(cc)::value; constexpr int K0 = C < 3 ? 0 : 3;
    double bc[6], y[6];
#pragma unroll
    for (int k = 0; k < 6; k++) bc[k] = k >= K0 ? elem(b, k, C) : 0.0;
    rq_mul<6, K0, DIAG>(Rq, bc, y);
#pragma unroll
    for (int k = 0; k < 6; k++) {
      if (DIAG && k < K0) { col[k] = 0.0; col[6 + k] = 0.0; }
      else { col[k] = w.u12 * y[k]; col[6 + k] = w.u22 * y[k]; }
    }
    sink(18 + C, col);
  });
  const L6 Dm = se3_djrinv(r, v2, c);
  {  // T2: [b; D b]
    const L6 Db = Dm * b;
    static_for<0, 6>([&](auto cc) { constexpr int C = decltype(cc)::value; gp_col_pair<C, DIAG>(b, Db, 1.0, w, Rq, col); sink(12 + C, col); });
  }
  {  // T1: -[a; D a],  a = Jl^-1(r) = Jr^-1(r) - ad(r)
    const L6 a = b - l6_ad(r);
    const L6 Da = Dm * a;
    static_for<0, 6>([&](auto cc) { constexpr int C = decltype(cc)::value; gp_col_pair<C, DIAG>(a, Da, -1.0, w, Rq, col); sink(C, col); });
  }
  return err;
}

// ================================================================= GP prior, SE(3) "VW" (gp/GaussianProcessPriorPose3VW.h:62-117)
// state record: [pose(12) | v_world(3) | w_world(3)], tangent order of a state [pose(6) | v(3) | w(3)].  convertVWtoVb
// (gp/Pose3utils.cpp:47-64): vb = [R^T w; R^T v], d vb / d pose = [[vb.w]x 0; [vb.v]x 0], d vb / d v = [0; R^T], d vb / d w = [R^T; 0].
// With those, the 12x25 [A|b] keeps the body-velocity factor's shape: p.a / p.Da / p.Db carry the extra pose terms
// (H1 = -[a + dt H1p; D a + H1p], H4 = [b; D b + b H2p]) and only the velocity columns need the rotations.
struct GpPose3VW { GpPose3 p; M3 R1, R2; };
GPB_HD X6 vw_to_vb(const M3& R, const double* vw) { return x6(tmul(R, v3(vw[3], vw[4], vw[5])), tmul(R, v3(vw[0], vw[1], vw[2]))); }
GPB_HD L6 vw_dpose(const X6& vb) { L6 r; r.A = skew(vb.w); r.B = skew(vb.v); r.C = m3_zero(); return r; }
GPB_HD void gp_prior_pose3vw_eval(const double* s1, const double* s2, double dt, bool wantJ, GpPose3VW& o) {
  const P3 T1 = p3_from_wire(s1), T2 = p3_from_wire(s2);
  const X6 v1 = vw_to_vb(T1.R, s1 + 12), v2 = vw_to_vb(T2.R, s2 + 12);
  const X6 r = se3_logmap(p3_between(T1, T2));
  const JrinvCoef c = se3_jrinv_coef(dot(r.w, r.w));
  o.p.b = se3_jrinv(r, c);
  o.p.e_top = r - dt * v1;
  o.p.e_bot = o.p.b * v2 - v1;
  if (wantJ) {
    o.R1 = T1.R; o.R2 = T2.R;
    const L6 a = o.p.b - l6_ad(r);
    const L6 D = se3_djrinv(r, v2, c);
    const L6 H1p = vw_dpose(v1), H2p = vw_dpose(v2);
    o.p.a = a + dt * H1p;
    o.p.Da = D * a + H1p;
    o.p.Db = D * o.p.b + o.p.b * H2p;
  }
}
template <int VAR, int C> GPB_HD void gp_prior_pose3vw_col(const GpPose3VW& o, const GpWhiten& w, const double* Rq, double dt, double* out) {
  if (VAR == 1 || VAR == 3) {  // [v | w] columns: u = d vb / d(v_C) = [0; R^T e_C]  or  d vb / d(w_{C-3}) = [R^T e_{C-3}; 0]
    const M3& R = VAR == 1 ? o.R1 : o.R2;
    const V3 rc = v3(R.m[3 * (C % 3)], R.m[3 * (C % 3) + 1], R.m[3 * (C % 3) + 2]);  // row C of R = column C of R^T
    const X6 u = C < 3 ? x6(v3(0, 0, 0), rc) : x6(rc, v3(0, 0, 0));
    double top[6], bot[6];
    if (VAR == 1) {  // [-dt u; -u]
#pragma unroll
      for (int k = 0; k < 6; k++) { top[k] = dt * elem(u, k); bot[k] = elem(u, k); }
      gp_whiten_col<6>(w, Rq, top, bot, -1.0, out);
    } else {  // [0; Jr^-1 u]
      const X6 bu = o.p.b * u;
#pragma unroll
      for (int k = 0; k < 6; k++) { top[k] = 0.0; bot[k] = elem(bu, k); }
      gp_whiten_col<6>(w, Rq, top, bot, 1.0, out);
    }
  } else {
    gp_prior_pose3_col<VAR, C>(o.p, w, Rq, dt, out);
  }
}

// ================================================================= GP prior, D = 3 groups
// Unwhitened: J_top = [Ja, -dt I, Jb, 0], J_bot = [0, -I, 0, +I] (Pose2/Rot3);  Linear: J_top = [I, dt I, -I, 0], J_bot = [0, I, 0, -I]
struct GpD3 { V3 e_top, e_bot; M3 Ja, Jb; double s_v1, s_v2; };  // s_v1: sign of the v1 blocks (-1 Lie, +1 linear); s_v2 likewise for v2 bottom
template <int G> GPB_HD void gp_prior_d3_eval(const double* s1, const double* s2, double dt, bool wantJ, GpD3& o) {
  constexpr int PS = GroupTraits<G>::PS;
  const V3 v1 = v3(s1[PS], s1[PS + 1], s1[PS + 2]), v2 = v3(s2[PS], s2[PS + 1], s2[PS + 2]);
  if (G == G_POSE2) {
    const P2 T1 = p2(s1[0], s1[1], s1[2]), T2 = p2(s2[0], s2[1], s2[2]);
    const P2 T12 = p2_between(T1, T2);
    const V3 r = se2_logmap(T12);
    o.e_top = r - dt * v1; o.e_bot = v2 - v1;
    if (wantJ) {
      const M3 Hlog = se2_dlog(r);
      // Hlog * Hcomp1 * Hinv = Hlog * Ad(T2^-1) * (-Ad(T1)) = -Hlog * Ad(T12^-1)
      o.Ja = -(Hlog * p2_adjoint(p2_inverse(T12)));
      o.Jb = Hlog;
    }
    o.s_v1 = -1.0; o.s_v2 = 1.0;
  } else if (G == G_ROT3) {
    const M3 R1 = m3_from_wire(s1), R2 = m3_from_wire(s2);
    const M3 R12 = transpose(R1) * R2;
    const V3 r = so3_logmap(R12);
    o.e_top = r - dt * v1; o.e_bot = v2 - v1;
    if (wantJ) {
      const M3 Hlog = so3_jrinv(r);
      o.Ja = -(Hlog * transpose(R12));  // Hlog * R2^T * (-R1)
      o.Jb = Hlog;
    }
    o.s_v1 = -1.0; o.s_v2 = 1.0;
  } else {  // linear: e = Phi(dt) x1 - x2
    const V3 p1 = v3(s1[0], s1[1], s1[2]), p2_ = v3(s2[0], s2[1], s2[2]);
    o.e_top = p1 + dt * v1 - p2_; o.e_bot = v1 - v2;
    o.Ja = m3_identity(); o.Jb = -m3_identity();
    o.s_v1 = 1.0; o.s_v2 = -1.0;
  }
}
template <int VAR, int C> GPB_HD void gp_prior_d3_col(const GpD3& o, const GpWhiten& w, const double* Rq, double dt, double* out) {
  double top[3], bot[3];
  if (VAR == 0) {
#pragma unroll
    for (int k = 0; k < 3; k++) { top[k] = o.Ja.m[3 * k + C]; bot[k] = 0.0; }
  } else if (VAR == 1) {
#pragma unroll
    for (int k = 0; k < 3; k++) { top[k] = (k == C) ? o.s_v1 * dt : 0.0; bot[k] = (k == C) ? o.s_v1 : 0.0; }
  } else if (VAR == 2) {
#pragma unroll
    for (int k = 0; k < 3; k++) { top[k] = o.Jb.m[3 * k + C]; bot[k] = 0.0; }
  } else if (VAR == 3) {
#pragma unroll
    for (int k = 0; k < 3; k++) { top[k] = 0.0; bot[k] = (k == C) ? o.s_v2 : 0.0; }
  } else {
#pragma unroll
    for (int k = 0; k < 3; k++) { top[k] = -elem(o.e_top, k); bot[k] = -elem(o.e_bot, k); }
  }
  gp_whiten_col<3>(w, Rq, top, bot, 1.0, out);
}

// ================================================================= "extra" factors (measurement rows attached to an interval)
// Every extra factor produces m whitened rows over [state_a (2D) | state_b (2D) | landmark (DL) | rhs].
// state_a / state_b are the two support states (i, i+1) of its interval (or (i, j) for a loop closure).
enum ExtraKind {
  X_INTERP_RANGE = 1, X_INTERP_ATTITUDE = 2, X_PRIOR_POSE = 3, X_PRIOR_VEL = 4, X_PRIOR_LANDMARK = 5,
  X_BETWEEN = 6, X_RANGE_2D = 7, X_RANGE_BEARING_2D = 8, X_ODOMETRY_2D = 9, X_INTERP_GPS = 10, X_INTERP_PROJECTION = 11,
  X_INTERP_GPS_VW = 12  // GPInterpolatedGPSFactorPose3VW (velocities [v_world | w_world])
};
constexpr int XP_STRIDE = 56;  // doubles of parameters per extra factor (see ExtraParams below)
// parameter record (doubles): [0] delta_t [1] tau [2] z [3] z2 [4..15] aux (sensor pose / nZ,bRef / value) [16] has_sensor
// [17] between: arguments swapped  [20..55] sqrt information R (m x m column-major, upper triangular); for scalar factors
// R[0] = 1/sigma.  GPS / projection factors (R needs 9 / 4 entries): [40..42] measured point, [43..47] Cal3_S2 (fx, fy, s, u0, v0).

GPB_HD X6 rowmul(const X6& h, const L6& M) { return x6(tmul(M.A, h.w) + tmul(M.B, h.v), tmul(M.C, h.v)); }  // h^T M as a row

// slam/GPInterpolatedRangeFactorPose3.h:64-98 — unwhitened residual and 1x6 Jacobian rows, 1x3 landmark row
struct Range3Out { double e; X6 H1, H2, H3, H4; V3 H5; };
GPB_HD void interp_range_pose3(const double* s1, const double* s2, const double* land, const double* prm, bool wantJ, Range3Out& o) {
  const double dt = prm[0], tau = prm[1], z = prm[2];
  const P3 T1 = p3_from_wire(s1), T2 = p3_from_wire(s2);
  const X6 v1 = x6_from(s1 + 12), v2 = x6_from(s2 + 12);
  const X6 r = se3_logmap(p3_between(T1, T2));
  const JrinvCoef c = se3_jrinv_coef(dot(r.w, r.w));
  const L6 b = se3_jrinv(r, c);
  const X6 f = b * v2;
  const InterpCoef ic = interp_coef(dt, tau);
  const X6 xi = ic.lam12 * v1 + ic.psi11 * r + ic.psi12 * f;
  const P3 dT = se3_expmap(xi);
  P3 T = p3_compose(T1, dT);
  const bool has_sensor = prm[16] != 0.0;
  P3 S;
  if (has_sensor) { S = p3_from_wire(prm + 4); T = p3_compose(T, S); }
  const V3 q = tmul(T.R, v3(land[0], land[1], land[2]) - T.t);
  const double rng = sqrt(dot(q, q));
  o.e = rng - z;
  if (!wantJ) return;
  const V3 qh = (1.0 / rng) * q;
  X6 hpose = x6(v3(0, 0, 0), -qh);  // qh^T [ [q]x , -I ] = [0, -qh^T]
  if (has_sensor) hpose = rowmul(hpose, l6_adjoint(p3_inverse(S)));
  const X6 g = rowmul(hpose, se3_jr(xi));  // Hpose * Hexp
  const L6 a = b - l6_ad(r);
  const L6 D = se3_djrinv(r, v2, c);
  const X6 gD = rowmul(g, D);
  const X6 k = ic.psi11 * g + ic.psi12 * gD;
  o.H1 = rowmul(hpose, l6_adjoint(p3_inverse(dT))) - rowmul(k, a);
  o.H2 = ic.lam12 * g;
  o.H3 = rowmul(k, b);
  o.H4 = ic.psi12 * rowmul(g, b);
  o.H5 = T.R * qh;  // (qh^T R^T)^T
}

// Shared pipeline of the SE(3) interpolated measurement factors with several residual rows (GPS, projection): the interpolated
// pose T(tau) [x body_P_sensor] once, then the four 1x6 Jacobian rows of ANY row `hpose` of d h / d T(tau) through
// updatePoseJacobians (gp/GaussianProcessInterpolatorPose3.h:57-116), exactly as interp_range_pose3 does for its single row.
struct Interp3 {
  P3 T, dT, S; bool has_sensor;
  X6 r, xi; L6 b, a, Dd, jr, adS, adT; InterpCoef ic;
};
GPB_HD void interp3_setup(const double* s1, const double* s2, const double* prm, bool wantJ, Interp3& c) {
  const double dt = prm[0], tau = prm[1];
  const P3 T1 = p3_from_wire(s1), T2 = p3_from_wire(s2);
  const X6 v1 = x6_from(s1 + 12), v2 = x6_from(s2 + 12);
  c.r = se3_logmap(p3_between(T1, T2));
  const JrinvCoef cf = se3_jrinv_coef(dot(c.r.w, c.r.w));
  c.b = se3_jrinv(c.r, cf);
  const X6 f = c.b * v2;
  c.ic = interp_coef(dt, tau);
  c.xi = c.ic.lam12 * v1 + c.ic.psi11 * c.r + c.ic.psi12 * f;
  c.dT = se3_expmap(c.xi);
  c.T = p3_compose(T1, c.dT);
  c.has_sensor = prm[16] != 0.0;
  if (c.has_sensor) { c.S = p3_from_wire(prm + 4); c.T = p3_compose(c.T, c.S); }
  if (!wantJ) return;
  if (c.has_sensor) c.adS = l6_adjoint(p3_inverse(c.S));
  c.jr = se3_jr(c.xi);
  c.a = c.b - l6_ad(c.r);
  c.Dd = se3_djrinv(c.r, v2, cf);
  c.adT = l6_adjoint(p3_inverse(c.dT));
}
// hpose: one row of d h / d(sensor pose); out: the row's Jacobians wrt (x1, v1, x2, v2)
GPB_HD void interp3_row(const Interp3& c, X6 hpose, X6& H1, X6& H2, X6& H3, X6& H4) {
  if (c.has_sensor) hpose = rowmul(hpose, c.adS);
  const X6 g = rowmul(hpose, c.jr);
  const X6 gD = rowmul(g, c.Dd);
  const X6 k = c.ic.psi11 * g + c.ic.psi12 * gD;
  H1 = rowmul(hpose, c.adT) - rowmul(k, c.a);
  H2 = c.ic.lam12 * g;
  H3 = rowmul(k, c.b);
  H4 = c.ic.psi12 * rowmul(g, c.b);
}
// slam/GPInterpolatedGPSFactorPose3.h:67-95: e = translation(T(tau) [body_P_sensor]) - measured; d translation / dT = [0, R]
struct Gps3Out { V3 e; X6 H[3][4]; };
GPB_HD void interp_gps_pose3(const double* s1, const double* s2, const double* prm, bool wantJ, Gps3Out& o) {
  Interp3 c;
  interp3_setup(s1, s2, prm, wantJ, c);
  o.e = c.T.t - v3(prm[40], prm[41], prm[42]);
  if (!wantJ) return;
#pragma unroll
  for (int k = 0; k < 3; k++) interp3_row(c, x6(v3(0, 0, 0), v3(c.T.R.m[3 * k], c.T.R.m[3 * k + 1], c.T.R.m[3 * k + 2])), o.H[k][0], o.H[k][1], o.H[k][2], o.H[k][3]);
}
// The same pipeline for the "VW" states (gp/GaussianProcessInterpolatorPose3VW.h:58-124): body velocities through convertVWtoVb,
// two extra pose terms (Hvel1 H1p, Hvel2 H2p) and the velocity rows rotated into the world frame.  Output rows are in the
// tangent order of a VW state: H2 = [d/dv1 | d/dw1], H4 = [d/dv2 | d/dw2] (stored in the (w, v) slots of an X6 in that order).
struct Interp3VW { Interp3 c; M3 R1, R2; L6 H1p, bH2p; };
GPB_HD void interp3vw_setup(const double* s1, const double* s2, const double* prm, bool wantJ, Interp3VW& o) {
  Interp3& c = o.c;
  const double dt = prm[0], tau = prm[1];
  const P3 T1 = p3_from_wire(s1), T2 = p3_from_wire(s2);
  const X6 v1 = vw_to_vb(T1.R, s1 + 12), v2 = vw_to_vb(T2.R, s2 + 12);
  c.r = se3_logmap(p3_between(T1, T2));
  const JrinvCoef cf = se3_jrinv_coef(dot(c.r.w, c.r.w));
  c.b = se3_jrinv(c.r, cf);
  const X6 f = c.b * v2;
  c.ic = interp_coef(dt, tau);
  c.xi = c.ic.lam12 * v1 + c.ic.psi11 * c.r + c.ic.psi12 * f;
  c.dT = se3_expmap(c.xi);
  c.T = p3_compose(T1, c.dT);
  c.has_sensor = prm[16] != 0.0;
  if (c.has_sensor) { c.S = p3_from_wire(prm + 4); c.T = p3_compose(c.T, c.S); }
  if (!wantJ) return;
  if (c.has_sensor) c.adS = l6_adjoint(p3_inverse(c.S));
  c.jr = se3_jr(c.xi);
  c.a = c.b - l6_ad(c.r);
  c.Dd = se3_djrinv(c.r, v2, cf);
  c.adT = l6_adjoint(p3_inverse(c.dT));
  o.R1 = T1.R; o.R2 = T2.R;
  o.H1p = vw_dpose(v1);
  o.bH2p = c.b * vw_dpose(v2);
}
GPB_HD void interp3vw_row(const Interp3VW& o, X6 hpose, X6& H1, X6& H2, X6& H3, X6& H4) {
  const Interp3& c = o.c;
  if (c.has_sensor) hpose = rowmul(hpose, c.adS);
  const X6 g = rowmul(hpose, c.jr);
  const X6 gD = rowmul(g, c.Dd);
  const X6 k = c.ic.psi11 * g + c.ic.psi12 * gD;
  H1 = rowmul(hpose, c.adT) - rowmul(k, c.a) + c.ic.lam12 * rowmul(g, o.H1p);
  H2 = c.ic.lam12 * x6(o.R1 * g.v, o.R1 * g.w);          // row (g.v)^T R1^T over v1, (g.w)^T R1^T over w1
  H3 = rowmul(k, c.b) + c.ic.psi12 * rowmul(g, o.bH2p);
  const X6 gb = rowmul(g, c.b);
  H4 = c.ic.psi12 * x6(o.R2 * gb.v, o.R2 * gb.w);
}
// slam/GPInterpolatedGPSFactorPose3VW.h:71-106
GPB_HD void interp_gps_pose3vw(const double* s1, const double* s2, const double* prm, bool wantJ, Gps3Out& o) {
  Interp3VW c;
  interp3vw_setup(s1, s2, prm, wantJ, c);
  o.e = c.c.T.t - v3(prm[40], prm[41], prm[42]);
  if (!wantJ) return;
#pragma unroll
  for (int k = 0; k < 3; k++) interp3vw_row(c, x6(v3(0, 0, 0), v3(c.c.T.R.m[3 * k], c.c.T.R.m[3 * k + 1], c.c.T.R.m[3 * k + 2])), o.H[k][0], o.H[k][1], o.H[k][2], o.H[k][3]);
}
// slam/GPInterpolatedProjectionFactorPose3.h:82-139 with Cal3_S2: e = K(pi(T^-1 l)) - measured; a landmark behind the camera
// (gtsam::CheiralityException, :123-138) gives zero Jacobians and the residual (2 fx, 2 fx).
struct Proj3Out { double e[2]; X6 H[2][4]; V3 H5[2]; };
GPB_HD void interp_projection_pose3(const double* s1, const double* s2, const double* land, const double* prm, bool wantJ, Proj3Out& o) {
  Interp3 c;
  interp3_setup(s1, s2, prm, wantJ, c);
  const double fx = prm[43], fy = prm[44], sk = prm[45], u0 = prm[46], v0 = prm[47];
  const V3 q = tmul(c.T.R, v3(land[0], land[1], land[2]) - c.T.t);
  if (!(q.z > 0.0)) {
    o.e[0] = o.e[1] = 2.0 * fx;
#pragma unroll
    for (int k = 0; k < 2; k++) { for (int v = 0; v < 4; v++) o.H[k][v] = x6(v3(0, 0, 0), v3(0, 0, 0)); o.H5[k] = v3(0, 0, 0); }
    return;
  }
  const double d = 1.0 / q.z, u = q.x * d, v = q.y * d;
  o.e[0] = fx * u + sk * v + u0 - prm[40];
  o.e[1] = fy * v + v0 - prm[41];
  if (!wantJ) return;
  // d pn / d pose = [[uv, -1-uu, v, -d, 0, du], [1+vv, -uv, -u, 0, -d, dv]];  d pn / d q = d [[1, 0, -u], [0, 1, -v]]
  const X6 n0 = x6(v3(u * v, -1.0 - u * u, v), v3(-d, 0.0, d * u)), n1 = x6(v3(1.0 + v * v, -u * v, -u), v3(0.0, -d, d * v));
  const V3 q0 = v3(d, 0.0, -d * u), q1 = v3(0.0, d, -d * v);
  const X6 h0 = fx * n0 + sk * n1, h1 = fy * n1;
  interp3_row(c, h0, o.H[0][0], o.H[0][1], o.H[0][2], o.H[0][3]);
  interp3_row(c, h1, o.H[1][0], o.H[1][1], o.H[1][2], o.H[1][3]);
  o.H5[0] = c.T.R * (fx * q0 + sk * q1);  // (row * R^T)^T
  o.H5[1] = c.T.R * (fy * q1);
}

// slam/GPInterpolatedRangeFactorPose2.h:64-98 and slam/GPInterpolatedRangeFactor2DLinear.h:60-88
struct Range2Out { double e; V3 H1, H2, H3, H4; double H5[2]; };
template <int G> GPB_HD void interp_range_2d(const double* s1, const double* s2, const double* land, const double* prm, bool wantJ, Range2Out& o) {
  const double dt = prm[0], tau = prm[1], z = prm[2];
  const InterpCoef ic = interp_coef(dt, tau);
  const V3 v1 = v3(s1[3], s1[4], s1[5]), v2 = v3(s2[3], s2[4], s2[5]);
  if (G == G_POSE2) {
    const P2 T1 = p2(s1[0], s1[1], s1[2]), T2 = p2(s2[0], s2[1], s2[2]);
    const P2 T12 = p2_between(T1, T2);
    const V3 r = se2_logmap(T12);
    const V3 xi = ic.lam12 * v1 + ic.psi11 * r + ic.psi12 * v2;
    const P2 dT = se2_expmap(xi);
    P2 T = p2_compose(T1, dT);
    const bool has_sensor = prm[16] != 0.0;
    P2 S = p2(prm[4], prm[5], prm[6]);
    if (has_sensor) T = p2_compose(T, S);
    const double dx = land[0] - T.x, dy = land[1] - T.y;
    const double rng = sqrt(dx * dx + dy * dy);
    o.e = rng - z;
    if (!wantJ) return;
    const double hx = dx / rng, hy = dy / rng, c = cos(T.th), s = sin(T.th);
    V3 hpose = v3(-(hx * c + hy * s), hx * s - hy * c, 0.0);  // D_r_d * [[-c, s, 0],[-s,-c,0]]
    if (has_sensor) hpose = tmul(p2_adjoint(p2_inverse(S)), hpose);
    const V3 g = tmul(se2_dexp(xi), hpose);  // row * Hexp
    const M3 Hlog = se2_dlog(r);
    const V3 gl = tmul(Hlog, g);
    o.H1 = tmul(p2_adjoint(p2_inverse(dT)), hpose) - ic.psi11 * tmul(p2_adjoint(p2_inverse(T12)), gl);
    o.H2 = ic.lam12 * g;
    o.H3 = ic.psi11 * gl;
    o.H4 = ic.psi12 * g;
    o.H5[0] = hx; o.H5[1] = hy;
  } else {  // 2DLinear: pose = Lambda_1 x1 + Psi_1 x2 ; theta ignored
    const double lam11 = 1.0 - ic.psi11;
    const double px = lam11 * s1[0] + ic.lam12 * v1.x + ic.psi11 * s2[0] + ic.psi12 * v2.x;
    const double py = lam11 * s1[1] + ic.lam12 * v1.y + ic.psi11 * s2[1] + ic.psi12 * v2.y;
    const double dx = land[0] - px, dy = land[1] - py;
    const double rng = sqrt(dx * dx + dy * dy);
    o.e = rng - z;
    if (!wantJ) return;
    double hx, hy;
    if (fabs(rng) > 1e-10) { hx = dx / rng; hy = dy / rng; } else { hx = 1; hy = 1; }  // gtsam::Point2::norm
    const V3 hpose = v3(-hx, -hy, 0.0);
    o.H1 = lam11 * hpose; o.H2 = ic.lam12 * hpose; o.H3 = ic.psi11 * hpose; o.H4 = ic.psi12 * hpose;
    o.H5[0] = hx; o.H5[1] = hy;
  }
}

// slam/GPInterpolatedAttitudeFactorRot3.h:61-83 (+ gtsam::AttitudeFactor::attitudeError): 2 rows
struct AttOut { double e[2]; V3 H1[2], H2[2], H3[2], H4[2]; };
GPB_HD void interp_attitude_rot3(const double* s1, const double* s2, const double* prm, bool wantJ, AttOut& o) {
  const double dt = prm[0], tau = prm[1];
  const V3 nZ = v3(prm[4], prm[5], prm[6]), bRef = v3(prm[7], prm[8], prm[9]);
  const InterpCoef ic = interp_coef(dt, tau);
  const M3 R1 = m3_from_wire(s1), R2 = m3_from_wire(s2);
  const V3 v1 = v3(s1[9], s1[10], s1[11]), v2 = v3(s2[9], s2[10], s2[11]);
  const M3 R12 = transpose(R1) * R2;
  const V3 r = so3_logmap(R12);
  const V3 xi = ic.lam12 * v1 + ic.psi11 * r + ic.psi12 * v2;
  const M3 dR = so3_expmap(xi);
  const M3 R = R1 * dR;
  const V3 nRef = R * bRef;
  V3 z1, z2; unit3_basis(nZ, z1, z2);
  o.e[0] = dot(z1, nRef); o.e[1] = dot(z2, nRef);
  if (!wantJ) return;
  // Hrot = Bz^T Bq Bq^T (-R [bRef]x);  Bq Bq^T = I - nRef nRef^T
  const M3 dn = -(R * skew(bRef));
  const M3 P = m3_identity() - outer(nRef, nRef);
  const M3 PM = P * dn;
  const M3 Hexp = so3_jr(xi), Hlog = so3_jrinv(r);
  const V3 zz[2] = {z1, z2};
#pragma unroll
  for (int k = 0; k < 2; k++) {
    const V3 hrot = tmul(PM, zz[k]);
    const V3 g = tmul(Hexp, hrot);
    const V3 gl = tmul(Hlog, g);
    o.H1[k] = tmul(transpose(dR), hrot) - ic.psi11 * tmul(transpose(R12), gl);  // Hcomp21 = dR^T ; Hlog*R2^T*(-R1) = -Hlog R12^T
    o.H2[k] = ic.lam12 * g;
    o.H3[k] = ic.psi11 * gl;
    o.H4[k] = ic.psi12 * g;
  }
}

// ================================================================= interpolatePose as a query (all groups)
// GaussianProcessInterpolator{Pose3,Pose3VW,Pose2,Rot3,Linear}::interpolatePose (gp/GaussianProcessInterpolatorPose3.h:57-105,
// ...Pose3VW.h:58-108, ...Pose2.h:56-89, ...Rot3.h:56-86, ...Linear.h:70-90): the pose at tau from the two support states, and
// on request the four D x D Jacobians Hint1..Hint4 (column-major, one after the other) - the same pipelines the interpolated
// measurement factors use, evaluated on unit rows.  pose_out: wire layout (12 / 9 / 3 doubles).
template <int G> GPB_HD void interp_pose(const double* s1, const double* s2, double dt, double tau, bool wantJ, double* pose_out, double* H) {
  constexpr int D = GroupTraits<G>::D;
  if constexpr (G == G_POSE3 || G == G_POSE3VW) {
    double prm[20];
#pragma unroll
    for (int k = 0; k < 20; k++) prm[k] = 0.0;
    prm[0] = dt; prm[1] = tau;
    auto rows = [&](auto& c, auto rowfn) {
#pragma unroll
      for (int k = 0; k < 6; k++) {
        const X6 ek = x6(v3(k == 0 ? 1.0 : 0.0, k == 1 ? 1.0 : 0.0, k == 2 ? 1.0 : 0.0), v3(k == 3 ? 1.0 : 0.0, k == 4 ? 1.0 : 0.0, k == 5 ? 1.0 : 0.0));
        X6 h[4];
        rowfn(c, ek, h[0], h[1], h[2], h[3]);
#pragma unroll
        for (int v = 0; v < 4; v++)
#pragma unroll
          for (int col = 0; col < 6; col++) H[v * 36 + k + 6 * col] = elem(h[v], col);
      }
    };
    if constexpr (G == G_POSE3) {
      Interp3 c; interp3_setup(s1, s2, prm, wantJ, c);
      p3_to_wire(c.T, pose_out);
      if (wantJ) rows(c, [](const Interp3& cc, const X6& e, X6& a, X6& b, X6& cx, X6& d) { interp3_row(cc, e, a, b, cx, d); });
    } else {
      Interp3VW c; interp3vw_setup(s1, s2, prm, wantJ, c);
      p3_to_wire(c.c.T, pose_out);
      if (wantJ) rows(c, [](const Interp3VW& cc, const X6& e, X6& a, X6& b, X6& cx, X6& d) { interp3vw_row(cc, e, a, b, cx, d); });
    }
  } else {
    const InterpCoef ic = interp_coef(dt, tau);
    constexpr int PS = GroupTraits<G>::PS;
    const V3 v1 = v3(s1[PS], s1[PS + 1], s1[PS + 2]), v2 = v3(s2[PS], s2[PS + 1], s2[PS + 2]);
    M3 H1, H2, H3, H4;
    if constexpr (G == G_POSE2) {
      const P2 T1 = p2(s1[0], s1[1], s1[2]), T2 = p2(s2[0], s2[1], s2[2]);
      const P2 T12 = p2_between(T1, T2);
      const V3 r = se2_logmap(T12);
      const V3 xi = ic.lam12 * v1 + ic.psi11 * r + ic.psi12 * v2;
      const P2 dT = se2_expmap(xi);
      const P2 T = p2_compose(T1, dT);
      pose_out[0] = T.x; pose_out[1] = T.y; pose_out[2] = T.th;
      if (wantJ) {
        const M3 Hexp = se2_dexp(xi), HexpHlog = Hexp * se2_dlog(r);
        H1 = p2_adjoint(p2_inverse(dT)) - ic.psi11 * (HexpHlog * p2_adjoint(p2_inverse(T12)));
        H2 = ic.lam12 * Hexp; H3 = ic.psi11 * HexpHlog; H4 = ic.psi12 * Hexp;
      }
    } else if constexpr (G == G_ROT3) {
      const M3 R1 = m3_from_wire(s1), R2 = m3_from_wire(s2);
      const M3 R12 = transpose(R1) * R2;
      const V3 r = so3_logmap(R12);
      const V3 xi = ic.lam12 * v1 + ic.psi11 * r + ic.psi12 * v2;
      const M3 dR = so3_expmap(xi);
      m3_to_wire(R1 * dR, pose_out);
      if (wantJ) {
        const M3 Hexp = so3_jr(xi), HexpHlog = Hexp * so3_jrinv(r);
        H1 = transpose(dR) - ic.psi11 * (HexpHlog * transpose(R12));
        H2 = ic.lam12 * Hexp; H3 = ic.psi11 * HexpHlog; H4 = ic.psi12 * Hexp;
      }
    } else {
      const double lam11 = 1.0 - ic.psi11;
#pragma unroll
      for (int k = 0; k < 3; k++) pose_out[k] = lam11 * s1[k] + ic.lam12 * s1[3 + k] + ic.psi11 * s2[k] + ic.psi12 * s2[3 + k];
      if (wantJ) { const M3 I = m3_identity(); H1 = lam11 * I; H2 = ic.lam12 * I; H3 = ic.psi11 * I; H4 = ic.psi12 * I; }
    }
    if (wantJ) {
      const M3* Hs[4] = {&H1, &H2, &H3, &H4};
#pragma unroll
      for (int v = 0; v < 4; v++)
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
          for (int col = 0; col < 3; col++) H[v * 9 + r + 3 * col] = Hs[v]->m[3 * r + col];
    }
  }
  (void)D;
}

}  // namespace gpb
