// ORACLE — TEST INFRASTRUCTURE ONLY (see gpo_linalg.h).
//
// CPU restatement of the reference's optimiser loop, which lives in GTSAM (SURVEY.md §3.2,
// §8a row 15): NonlinearFactorGraph::linearize -> whitened JacobianFactors -> (LM damping)
// -> Cholesky elimination -> back-substitution -> Values::retract -> graph.error.
// GTSAM's multifrontal elimination of a GP chain with landmarks ordered last is restated as
// a sequential bordered block-tridiagonal Cholesky (chain cliques {x_i,v_i | x_{i+1},v_{i+1},
// landmarks}); states touched by loop closures join the border.
//
// PARITY UNPINNED items (no reference test pins them): LM lambda schedule
// (GTSAM 4.0 defaults restated from its published algorithm), retraction flavour
// (Pose3/Rot3: Expmap, GTSAM >= 4.1 default; Pose2: GTSAM's default first-order chart),
// PriorFactor Jacobian = I and BetweenFactor without the Local Jacobian (GTSAM defaults),
// whitening matrix R = upper Cholesky of the information matrix, elimination order.
// PINNED: every factor's residual/Jacobian and the GN-converged 2-state solutions, by the
// reference's own unit tests ported in tests/test_oracle_golden.py.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include <thread>
#include <time.h>
#include "gpo_factors.h"

using namespace gpo;

namespace {

enum Group { G_POSE3 = 0, G_POSE2 = 1, G_ROT3 = 2, G_LINEAR = 3 };
enum FactorKind {
  F_GP_PRIOR = 0, F_INTERP_RANGE = 1, F_INTERP_ATTITUDE = 2, F_PRIOR_POSE = 3, F_PRIOR_VEL = 4,
  F_PRIOR_LANDMARK = 5, F_BETWEEN = 6, F_RANGE_2D = 7, F_RANGE_BEARING_2D = 8, F_ODOMETRY_2D = 9,
  F_INTERP_GPS = 10, F_INTERP_PROJECTION = 11,
  F_GP_PRIOR_VW = 12, F_INTERP_GPS_VW = 13  // Pose3 "VW" family: the velocity variable of a state is [v_world; w_world]
};

struct Factor {
  int kind;
  int i = 0, j = 0, l = 0;  // state index / second state / landmark
  int qc = 0;               // Qc model id
  double delta_t = 0, tau = 0, z = 0, z2 = 0;
  bool has_sensor = false;
  double aux[12] = {0};     // body_P_sensor (Pose3 12 / Pose2 3) | nZ,bRef | measurement value
  double meas[3] = {0};     // GPS point / image point
  double K[5] = {0};        // Cal3_S2 (fx, fy, s, u0, v0)
  double R[36] = {0};       // sqrt information (m x m, column-major, upper triangular)
  int m = 0;
};

// one whitened JacobianFactor: b and A blocks per variable
struct VarRef { int type; int idx; };  // type 0 pose x_i, 1 vel v_i, 2 landmark l
struct Lin {
  int m = 0, nv = 0;
  VarRef v[5];
  int d[5];
  double A[5][12 * 6];  // column-major m x d
  double b[12];
};

struct Graph {
  int group, D, PS, DL, N, L;
  std::vector<std::vector<double>> Qc;  // D*D each
  std::vector<Factor> factors;
  std::vector<double> poses, vels, lands;
  std::string err;
  int threads = 1;
  // solver scratch
  std::vector<int> chain_of_state, border_of_state;
};

int pose_storage(int group, int D) { return group == G_POSE3 ? 12 : group == G_ROT3 ? 9 : group == G_POSE2 ? 3 : D; }
int land_dim(int group, int D) { return group == G_POSE3 ? 3 : group == G_ROT3 ? 0 : 2; }

template <int R, int C> void put(double* dst, const Mat<R, C>& m) { for (int i = 0; i < R * C; i++) dst[i] = m.a[i]; }
template <int R, int C> Mat<R, C> get(const double* src) { Mat<R, C> m; for (int i = 0; i < R * C; i++) m.a[i] = src[i]; return m; }

Mat3 rot3_from(const double* p) { return get<3, 3>(p); }
Pose2 pose2_from(const double* p) { return Pose2(p[0], p[1], p[2]); }

// Unwhitened residual + Jacobians of one factor (the reference's evaluateError).  e: m doubles.
// H[k]: m x d[k] column-major, or skipped when wantH is false.
template <int D>
void eval_linear_group(const Graph& g, const Factor& f, const double* P, const double* V, const double* Lm, bool wantH, Lin& out, double* e);

void eval_factor(const Graph& g, const Factor& f, const double* P, const double* V, const double* Lm, bool wantH, Lin& out, double* e) {
  const int D = g.D, PS = g.PS, DL = g.DL;
  out.nv = 0;
  auto addv = [&](int type, int idx, int d) { out.v[out.nv] = {type, idx}; out.d[out.nv] = d; return out.nv++; };
  if (g.group == G_LINEAR) {
    switch (D) {
      case 1: eval_linear_group<1>(g, f, P, V, Lm, wantH, out, e); return;
      case 2: eval_linear_group<2>(g, f, P, V, Lm, wantH, out, e); return;
      case 3: eval_linear_group<3>(g, f, P, V, Lm, wantH, out, e); return;
      case 6: eval_linear_group<6>(g, f, P, V, Lm, wantH, out, e); return;
      default: std::abort();
    }
  }
  switch (f.kind) {
    case F_GP_PRIOR: {
      const int i = f.i, j = f.i + 1;
      addv(0, i, D); addv(1, i, D); addv(0, j, D); addv(1, j, D);
      if (g.group == G_POSE3) {
        out.m = 12;
        Mat<12, 6> H1, H2, H3, H4;
        const Vec12 r = gpPriorPose3(Pose3::from(P + i * PS), vec<6>(V + i * D), Pose3::from(P + j * PS), vec<6>(V + j * D), f.delta_t,
                                     wantH ? &H1 : nullptr, wantH ? &H2 : nullptr, wantH ? &H3 : nullptr, wantH ? &H4 : nullptr);
        put(e, r);
        if (wantH) { put(out.A[0], H1); put(out.A[1], H2); put(out.A[2], H3); put(out.A[3], H4); }
      } else {
        out.m = 6;
        Mat<6, 3> H1, H2, H3, H4;
        Vec6 r;
        if (g.group == G_POSE2)
          r = gpPriorPose2(pose2_from(P + i * PS), vec<3>(V + i * D), pose2_from(P + j * PS), vec<3>(V + j * D), f.delta_t,
                           wantH ? &H1 : nullptr, wantH ? &H2 : nullptr, wantH ? &H3 : nullptr, wantH ? &H4 : nullptr);
        else
          r = gpPriorRot3(rot3_from(P + i * PS), vec<3>(V + i * D), rot3_from(P + j * PS), vec<3>(V + j * D), f.delta_t,
                          wantH ? &H1 : nullptr, wantH ? &H2 : nullptr, wantH ? &H3 : nullptr, wantH ? &H4 : nullptr);
        put(e, r);
        if (wantH) { put(out.A[0], H1); put(out.A[1], H2); put(out.A[2], H3); put(out.A[3], H4); }
      }
      return;
    }
    case F_INTERP_RANGE: {
      const int i = f.i, j = f.i + 1;
      addv(0, i, D); addv(1, i, D); addv(0, j, D); addv(1, j, D); addv(2, f.l, DL);
      out.m = 1;
      if (g.group == G_POSE3) {
        const InterpolatorPose3 gp(get<6, 6>(g.Qc[f.qc].data()), f.delta_t, f.tau);
        Pose3 sensor; if (f.has_sensor) sensor = Pose3::from(f.aux);
        Mat<1, 6> H1, H2, H3, H4; Mat<1, 3> H5;
        e[0] = gpRangePose3(gp, f.z, f.has_sensor ? &sensor : nullptr, Pose3::from(P + i * PS), vec<6>(V + i * D), Pose3::from(P + j * PS),
                            vec<6>(V + j * D), vec<3>(Lm + f.l * DL), wantH ? &H1 : nullptr, wantH ? &H2 : nullptr, wantH ? &H3 : nullptr,
                            wantH ? &H4 : nullptr, wantH ? &H5 : nullptr);
        if (wantH) { put(out.A[0], H1); put(out.A[1], H2); put(out.A[2], H3); put(out.A[3], H4); put(out.A[4], H5); }
      } else if (g.group == G_POSE2) {
        const InterpolatorPose2 gp(get<3, 3>(g.Qc[f.qc].data()), f.delta_t, f.tau);
        Pose2 sensor; if (f.has_sensor) sensor = pose2_from(f.aux);
        Mat<1, 3> H1, H2, H3, H4; Mat<1, 2> H5;
        e[0] = gpRangePose2(gp, f.z, f.has_sensor ? &sensor : nullptr, pose2_from(P + i * PS), vec<3>(V + i * D), pose2_from(P + j * PS),
                            vec<3>(V + j * D), vec<2>(Lm + f.l * DL), wantH ? &H1 : nullptr, wantH ? &H2 : nullptr, wantH ? &H3 : nullptr,
                            wantH ? &H4 : nullptr, wantH ? &H5 : nullptr);
        if (wantH) { put(out.A[0], H1); put(out.A[1], H2); put(out.A[2], H3); put(out.A[3], H4); put(out.A[4], H5); }
      } else std::abort();
      return;
    }
    case F_INTERP_GPS: {  // slam/GPInterpolatedGPSFactorPose3.h
      const int i = f.i, j = f.i + 1;
      addv(0, i, D); addv(1, i, D); addv(0, j, D); addv(1, j, D);
      out.m = 3;
      if (g.group != G_POSE3) std::abort();
      const InterpolatorPose3 gp(get<6, 6>(g.Qc[f.qc].data()), f.delta_t, f.tau);
      Pose3 sensor; if (f.has_sensor) sensor = Pose3::from(f.aux);
      Mat<3, 6> H1, H2, H3, H4;
      put(e, gpGPSPose3(gp, vec<3>(f.meas), f.has_sensor ? &sensor : nullptr, Pose3::from(P + i * PS), vec<6>(V + i * D), Pose3::from(P + j * PS), vec<6>(V + j * D),
                        wantH ? &H1 : nullptr, wantH ? &H2 : nullptr, wantH ? &H3 : nullptr, wantH ? &H4 : nullptr));
      if (wantH) { put(out.A[0], H1); put(out.A[1], H2); put(out.A[2], H3); put(out.A[3], H4); }
      return;
    }
    case F_GP_PRIOR_VW: {  // gp/GaussianProcessPriorPose3VW.h
      const int i = f.i, j = f.i + 1;
      addv(0, i, D); addv(1, i, D); addv(0, j, D); addv(1, j, D);
      out.m = 12;
      if (g.group != G_POSE3) std::abort();
      Mat<12, 6> H1, H2, H3, H4;
      put(e, gpPriorPose3VW(Pose3::from(P + i * PS), vec<3>(V + i * D), vec<3>(V + i * D + 3), Pose3::from(P + j * PS), vec<3>(V + j * D), vec<3>(V + j * D + 3), f.delta_t,
                            wantH ? &H1 : nullptr, wantH ? &H2 : nullptr, wantH ? &H3 : nullptr, wantH ? &H4 : nullptr));
      if (wantH) { put(out.A[0], H1); put(out.A[1], H2); put(out.A[2], H3); put(out.A[3], H4); }
      return;
    }
    case F_INTERP_GPS_VW: {  // slam/GPInterpolatedGPSFactorPose3VW.h
      const int i = f.i, j = f.i + 1;
      addv(0, i, D); addv(1, i, D); addv(0, j, D); addv(1, j, D);
      out.m = 3;
      if (g.group != G_POSE3) std::abort();
      const InterpolatorPose3VW gp(get<6, 6>(g.Qc[f.qc].data()), f.delta_t, f.tau);
      Pose3 sensor; if (f.has_sensor) sensor = Pose3::from(f.aux);
      Mat<3, 6> H1, H2, H3, H4;
      put(e, gpGPSPose3VW(gp, vec<3>(f.meas), f.has_sensor ? &sensor : nullptr, Pose3::from(P + i * PS), vec<3>(V + i * D), vec<3>(V + i * D + 3), Pose3::from(P + j * PS),
                          vec<3>(V + j * D), vec<3>(V + j * D + 3), wantH ? &H1 : nullptr, wantH ? &H2 : nullptr, wantH ? &H3 : nullptr, wantH ? &H4 : nullptr));
      if (wantH) { put(out.A[0], H1); put(out.A[1], H2); put(out.A[2], H3); put(out.A[3], H4); }
      return;
    }
    case F_INTERP_PROJECTION: {  // slam/GPInterpolatedProjectionFactorPose3.h
      const int i = f.i, j = f.i + 1;
      addv(0, i, D); addv(1, i, D); addv(0, j, D); addv(1, j, D); addv(2, f.l, DL);
      out.m = 2;
      if (g.group != G_POSE3) std::abort();
      const InterpolatorPose3 gp(get<6, 6>(g.Qc[f.qc].data()), f.delta_t, f.tau);
      Pose3 sensor; if (f.has_sensor) sensor = Pose3::from(f.aux);
      Mat<2, 6> H1, H2, H3, H4; Mat<2, 3> H5;
      put(e, gpProjectionPose3(gp, vec<2>(f.meas), f.K, f.has_sensor ? &sensor : nullptr, Pose3::from(P + i * PS), vec<6>(V + i * D), Pose3::from(P + j * PS),
                               vec<6>(V + j * D), vec<3>(Lm + f.l * DL), wantH ? &H1 : nullptr, wantH ? &H2 : nullptr, wantH ? &H3 : nullptr,
                               wantH ? &H4 : nullptr, wantH ? &H5 : nullptr));
      if (wantH) { put(out.A[0], H1); put(out.A[1], H2); put(out.A[2], H3); put(out.A[3], H4); put(out.A[4], H5); }
      return;
    }
    case F_INTERP_ATTITUDE: {
      const int i = f.i, j = f.i + 1;
      addv(0, i, D); addv(1, i, D); addv(0, j, D); addv(1, j, D);
      out.m = 2;
      const InterpolatorRot3 gp(get<3, 3>(g.Qc[f.qc].data()), f.delta_t, f.tau);
      Mat<2, 3> H1, H2, H3, H4;
      const Vec2 r = gpAttitudeRot3(gp, vec<3>(f.aux), vec<3>(f.aux + 3), rot3_from(P + i * PS), vec<3>(V + i * D), rot3_from(P + j * PS),
                                    vec<3>(V + j * D), wantH ? &H1 : nullptr, wantH ? &H2 : nullptr, wantH ? &H3 : nullptr, wantH ? &H4 : nullptr);
      put(e, r);
      if (wantH) { put(out.A[0], H1); put(out.A[1], H2); put(out.A[2], H3); put(out.A[3], H4); }
      return;
    }
    case F_PRIOR_POSE: {
      // gtsam::PriorFactor<T>: e = -Local(x, prior) = Logmap(prior^-1 x) (Pose2: first-order chart), H = I
      addv(0, f.i, D); out.m = D;
      if (g.group == G_POSE3) put(e, pose3_logmap(Pose3::from(f.aux).inverse().compose(Pose3::from(P + f.i * PS))));
      else if (g.group == G_ROT3) put(e, so3_logmap(rot3_from(f.aux).t() * rot3_from(P + f.i * PS)));
      else { const Pose2 d = pose2_from(P + f.i * PS).inverse().compose(pose2_from(f.aux)); e[0] = -d.x; e[1] = -d.y; e[2] = -d.theta(); }
      if (wantH) { for (int k = 0; k < D * D; k++) out.A[0][k] = 0; for (int k = 0; k < D; k++) out.A[0][k + k * D] = 1; }
      return;
    }
    case F_PRIOR_VEL: case F_PRIOR_LANDMARK: {
      const bool isv = f.kind == F_PRIOR_VEL;
      const int d = isv ? D : DL;
      addv(isv ? 1 : 2, isv ? f.i : f.l, d); out.m = d;
      const double* x = isv ? V + f.i * D : Lm + f.l * DL;
      for (int k = 0; k < d; k++) e[k] = x[k] - f.aux[k];
      if (wantH) { for (int k = 0; k < d * d; k++) out.A[0][k] = 0; for (int k = 0; k < d; k++) out.A[0][k + k * d] = 1; }
      return;
    }
    case F_BETWEEN: {
      // gtsam::BetweenFactor<T>: hx = x1^-1 x2, e = Local(measured, hx); H1 = -Ad(hx^-1), H2 = I
      addv(0, f.i, D); addv(0, f.j, D); out.m = D;
      if (g.group == G_POSE3) {
        const Pose3 hx = Pose3::from(P + f.i * PS).inverse().compose(Pose3::from(P + f.j * PS));
        put(e, pose3_logmap(Pose3::from(f.aux).inverse().compose(hx)));
        if (wantH) { put(out.A[0], -hx.inverse().Adjoint()); put(out.A[1], Mat6::Identity()); }
      } else if (g.group == G_ROT3) {
        const Mat3 hx = rot3_from(P + f.i * PS).t() * rot3_from(P + f.j * PS);
        put(e, so3_logmap(rot3_from(f.aux).t() * hx));
        if (wantH) { put(out.A[0], -hx.t()); put(out.A[1], Mat3::Identity()); }
      } else {
        const Pose2 hx = pose2_from(P + f.i * PS).inverse().compose(pose2_from(P + f.j * PS));
        const Pose2 d = pose2_from(f.aux).inverse().compose(hx);
        e[0] = d.x; e[1] = d.y; e[2] = d.theta();
        if (wantH) { put(out.A[0], -hx.inverse().Adjoint()); put(out.A[1], Mat3::Identity()); }
      }
      return;
    }
    case F_RANGE_2D: {  // slam/RangeFactorPose2.h:15 = gtsam::RangeFactor<Pose2,Point2>
      if (g.group != G_POSE2) std::abort();
      addv(0, f.i, D); addv(2, f.l, DL); out.m = 1;
      Mat<1, 3> H1; Mat<1, 2> H2;
      e[0] = pose2_range(pose2_from(P + f.i * PS), vec<2>(Lm + f.l * DL), wantH ? &H1 : nullptr, wantH ? &H2 : nullptr) - f.z;
      if (wantH) { put(out.A[0], H1); put(out.A[1], H2); }
      return;
    }
    default: std::abort();
  }
}

template <int D>
void eval_linear_group(const Graph& g, const Factor& f, const double* P, const double* V, const double* Lm, bool wantH, Lin& out, double* e) {
  const int DL = g.DL;
  auto addv = [&](int type, int idx, int d) { out.v[out.nv] = {type, idx}; out.d[out.nv] = d; return out.nv++; };
  switch (f.kind) {
    case F_GP_PRIOR: {
      const int i = f.i, j = f.i + 1;
      addv(0, i, D); addv(1, i, D); addv(0, j, D); addv(1, j, D); out.m = 2 * D;
      Mat<2 * D, D> H1, H2, H3, H4;
      put(e, gpPriorLinear<D>(vec<D>(P + i * D), vec<D>(V + i * D), vec<D>(P + j * D), vec<D>(V + j * D), f.delta_t,
                              wantH ? &H1 : nullptr, wantH ? &H2 : nullptr, wantH ? &H3 : nullptr, wantH ? &H4 : nullptr));
      if (wantH) { put(out.A[0], H1); put(out.A[1], H2); put(out.A[2], H3); put(out.A[3], H4); }
      return;
    }
    case F_PRIOR_POSE: case F_PRIOR_VEL: case F_PRIOR_LANDMARK: {
      const int type = f.kind == F_PRIOR_POSE ? 0 : f.kind == F_PRIOR_VEL ? 1 : 2;
      const int d = type == 2 ? DL : D;
      addv(type, type == 2 ? f.l : f.i, d); out.m = d;
      const double* x = type == 0 ? P + f.i * D : type == 1 ? V + f.i * D : Lm + f.l * DL;
      for (int k = 0; k < d; k++) e[k] = x[k] - f.aux[k];
      if (wantH) { for (int k = 0; k < d * d; k++) out.A[0][k] = 0; for (int k = 0; k < d; k++) out.A[0][k + k * d] = 1; }
      return;
    }
    case F_BETWEEN: {  // vector-space BetweenFactor: e = (x2 - x1) - measured
      addv(0, f.i, D); addv(0, f.j, D); out.m = D;
      for (int k = 0; k < D; k++) e[k] = (P[f.j * D + k] - P[f.i * D + k]) - f.aux[k];
      if (wantH) { for (int k = 0; k < D * D; k++) { out.A[0][k] = 0; out.A[1][k] = 0; } for (int k = 0; k < D; k++) { out.A[0][k + k * D] = -1; out.A[1][k + k * D] = 1; } }
      return;
    }
    default: break;
  }
  if constexpr (D == 3) {
    switch (f.kind) {
      case F_INTERP_RANGE: {
        const int i = f.i, j = f.i + 1;
        addv(0, i, 3); addv(1, i, 3); addv(0, j, 3); addv(1, j, 3); addv(2, f.l, 2); out.m = 1;
        const InterpolatorLinear<3> gp(get<3, 3>(g.Qc[f.qc].data()), f.delta_t, f.tau);
        Mat<1, 3> H1, H2, H3, H4; Mat<1, 2> H5;
        e[0] = gpRange2DLinear(gp, f.z, vec<3>(P + i * 3), vec<3>(V + i * 3), vec<3>(P + j * 3), vec<3>(V + j * 3), vec<2>(Lm + f.l * 2),
                               wantH ? &H1 : nullptr, wantH ? &H2 : nullptr, wantH ? &H3 : nullptr, wantH ? &H4 : nullptr, wantH ? &H5 : nullptr);
        if (wantH) { put(out.A[0], H1); put(out.A[1], H2); put(out.A[2], H3); put(out.A[3], H4); put(out.A[4], H5); }
        return;
      }
      case F_RANGE_2D: {
        addv(0, f.i, 3); addv(2, f.l, 2); out.m = 1;
        Mat<1, 3> H1; Mat<1, 2> H2;
        e[0] = range2DLinear(f.z, vec<3>(P + f.i * 3), vec<2>(Lm + f.l * 2), wantH ? &H1 : nullptr, wantH ? &H2 : nullptr);
        if (wantH) { put(out.A[0], H1); put(out.A[1], H2); }
        return;
      }
      case F_RANGE_BEARING_2D: {
        addv(0, f.i, 3); addv(2, f.l, 2); out.m = 2;
        Mat<2, 3> H1; Mat2 H2;
        put(e, rangeBearing2DLinear(f.z, f.z2, vec<3>(P + f.i * 3), vec<2>(Lm + f.l * 2), wantH ? &H1 : nullptr, wantH ? &H2 : nullptr));
        if (wantH) { put(out.A[0], H1); put(out.A[1], H2); }
        return;
      }
      case F_ODOMETRY_2D: {
        addv(0, f.i, 3); addv(0, f.j, 3); out.m = 3;
        Mat3 H1, H2;
        put(e, odometry2DLinear(vec<3>(f.aux), vec<3>(P + f.i * 3), vec<3>(P + f.j * 3), wantH ? &H1 : nullptr, wantH ? &H2 : nullptr));
        if (wantH) { put(out.A[0], H1); put(out.A[1], H2); }
        return;
      }
      default: break;
    }
  }
  std::abort();
}

// noise model of a factor: sqrt information R (m x m upper, column-major).  GP priors:
// Gaussian::Covariance(calcQ(Qc, dt)) (gp/GaussianProcessPriorPose3.h:46) -> R = chol_upper(Q^-1);
// Q^-1 = calcQ_inv (gp/GPutils.h:33-41).
template <int D> void gp_prior_R(const double* Qc, double dt, double* R) {
  const Mat<2 * D, 2 * D> Qi = calcQ_inv<D>(get<D, D>(Qc), dt);
  Mat<2 * D, 2 * D> U;
  if (!chol_upper(Qi, U)) std::abort();
  put(R, U);
}
void factor_R(const Graph& g, const Factor& f, double* R /*12x12 max*/) {
  if (f.kind == F_GP_PRIOR || f.kind == F_GP_PRIOR_VW) {
    switch (g.D) {
      case 1: gp_prior_R<1>(g.Qc[f.qc].data(), f.delta_t, R); break;
      case 2: gp_prior_R<2>(g.Qc[f.qc].data(), f.delta_t, R); break;
      case 3: gp_prior_R<3>(g.Qc[f.qc].data(), f.delta_t, R); break;
      case 6: gp_prior_R<6>(g.Qc[f.qc].data(), f.delta_t, R); break;
      default: std::abort();
    }
  } else {
    for (int k = 0; k < f.m * f.m; k++) R[k] = f.R[k];
  }
}

// Whitened linearisation of one factor: A_k = R H_k, b = -R e  (NoiseModelFactor::linearize +
// noiseModel::Gaussian::WhitenSystem).  Returns 0.5 |R e|^2.
double linearize_factor(const Graph& g, const Factor& f, const double* P, const double* V, const double* Lm, bool wantH, Lin& out) {
  double e[12], R[144];
  eval_factor(g, f, P, V, Lm, wantH, out, e);
  const int m = out.m;
  factor_R(g, f, R);
  double err = 0;
  for (int r = 0; r < m; r++) { double s = 0; for (int k = r; k < m; k++) s += R[r + k * m] * e[k]; out.b[r] = -s; err += s * s; }
  if (wantH) {
    double tmp[12 * 6];
    for (int v = 0; v < out.nv; v++) {
      const int d = out.d[v];
      for (int c = 0; c < d; c++) for (int r = 0; r < m; r++) { double s = 0; for (int k = r; k < m; k++) s += R[r + k * m] * out.A[v][k + c * m]; tmp[r + c * m] = s; }
      for (int k = 0; k < m * d; k++) out.A[v][k] = tmp[k];
    }
  }
  return 0.5 * err;
}


// parallel sum over [0,n) with `threads` std::threads (static contiguous partition; the per-thread
// partials are added in thread order, so the result is deterministic for a given thread count)
template <class F> double parallel_sum(int n, int threads, F fn) {
  if (threads <= 1 || n < 2 * threads) { double t = 0; for (int k = 0; k < n; k++) t += fn(k); return t; }
  std::vector<double> part(threads, 0.0);
  std::vector<std::thread> pool;
  for (int w = 0; w < threads; w++)
    pool.emplace_back([&, w]() { const int lo = (int)((long long)n * w / threads), hi = (int)((long long)n * (w + 1) / threads); double t = 0; for (int k = lo; k < hi; k++) t += fn(k); part[w] = t; });
  for (auto& th : pool) th.join();
  double t = 0; for (double p : part) t += p; return t;
}

double graph_error(const Graph& g, const double* P, const double* V, const double* Lm) {
  const int nf = (int)g.factors.size();
  return parallel_sum(nf, g.threads, [&](int k) { Lin lin; return linearize_factor(g, g.factors[k], P, V, Lm, false, lin); });
}

// retract: Pose3/Rot3 Expmap (right), Pose2 GTSAM default chart x * Pose2(v0,v1,v2), vectors add.
void retract(const Graph& g, const double* P, const double* V, const double* Lm, const double* dchain, const double* dborder,
             const std::vector<int>& chain_of_state, const std::vector<int>& border_of_state, int land_off,
             double* Pn, double* Vn, double* Ln) {
  const int D = g.D, PS = g.PS, DL = g.DL, bs = 2 * D;
  for (int i = 0; i < g.N; i++) {
    const double* d = chain_of_state[i] >= 0 ? dchain + (size_t)chain_of_state[i] * bs : dborder + border_of_state[i];
    if (g.group == G_POSE3) Pose3::from(P + i * PS).compose(pose3_expmap(vec<6>(d))).to(Pn + i * PS);
    else if (g.group == G_ROT3) put(Pn + i * PS, rot3_from(P + i * PS) * so3_expmap(vec<3>(d)));
    else if (g.group == G_POSE2) { const Pose2 q = pose2_from(P + i * PS).compose(Pose2(d[0], d[1], d[2])); Pn[i * PS] = q.x; Pn[i * PS + 1] = q.y; Pn[i * PS + 2] = q.th; }
    else for (int k = 0; k < D; k++) Pn[i * PS + k] = P[i * PS + k] + d[k];
    for (int k = 0; k < D; k++) Vn[i * D + k] = V[i * D + k] + d[D + k];
  }
  for (int l = 0; l < g.L; l++) for (int k = 0; k < DL; k++) Ln[l * DL + k] = Lm[l * DL + k] + dborder[land_off + l * DL + k];
}

// ---------------------------------------------------------------- normal equations + bordered block-tridiagonal Cholesky
struct System {
  int nc = 0, bs = 0, nb = 0;          // chain blocks, block size, border dim
  std::vector<double> Dg, E, B, gc;    // Dg[nc][bs*bs], E[nc-1][bs*bs] (rows k+1, cols k), B[nc][bs*nb], gc[nc][bs]  (column-major blocks)
  std::vector<double> C, gb;           // C[nb*nb], gb[nb]
  double c0 = 0;                       // 0.5 b^T b
};

struct Solver {
  Graph* g;
  std::vector<int> chain_of_state, border_of_state, state_of_chain;
  int land_off = 0;
  System sys;
  std::vector<Lin> lins;
  std::vector<double> Lf, Le, Y, yb, dchain, dborder;  // factor storage
  double lin_time = 0, solve_time = 0;

  void setup() {
    Graph& G = *g;
    const int N = G.N, bs = 2 * G.D;
    std::vector<char> bordered(N, 0);
    for (const Factor& f : G.factors) {
      const bool two = f.kind == F_BETWEEN || f.kind == F_ODOMETRY_2D;
      if (two && std::abs(f.i - f.j) != 1) { bordered[f.i] = 1; bordered[f.j] = 1; }
    }
    chain_of_state.assign(N, -1); border_of_state.assign(N, -1); state_of_chain.clear();
    int nb = 0;
    for (int i = 0; i < N; i++) {
      if (bordered[i]) { border_of_state[i] = nb; nb += bs; }
      else { chain_of_state[i] = (int)state_of_chain.size(); state_of_chain.push_back(i); }
    }
    land_off = nb; nb += G.L * G.DL;
    sys.nc = (int)state_of_chain.size(); sys.bs = bs; sys.nb = nb;
    sys.Dg.resize((size_t)sys.nc * bs * bs); sys.E.resize((size_t)std::max(0, sys.nc - 1) * bs * bs);
    sys.B.resize((size_t)sys.nc * bs * nb); sys.gc.resize((size_t)sys.nc * bs); sys.C.resize((size_t)nb * nb); sys.gb.resize(nb);
    lins.resize(G.factors.size());
  }

  // locate variable: returns chain block (>=0) + offset, or -1 and border offset
  inline void locate(const VarRef& v, int& chain, int& off) const {
    const int D = g->D;
    if (v.type == 2) { chain = -1; off = land_off + v.idx * g->DL; return; }
    const int c = chain_of_state[v.idx];
    if (c >= 0) { chain = c; off = v.type == 1 ? D : 0; }
    else { chain = -1; off = border_of_state[v.idx] + (v.type == 1 ? D : 0); }
  }

  double linearize(const double* P, const double* V, const double* Lm) {
    Graph& G = *g;
    const int nf = (int)G.factors.size();
    return parallel_sum(nf, G.threads, [&](int k) { return linearize_factor(G, G.factors[k], P, V, Lm, true, lins[k]); });
  }

  void assemble() {
    const int bs = sys.bs, nb = sys.nb;
    std::fill(sys.Dg.begin(), sys.Dg.end(), 0.0); std::fill(sys.E.begin(), sys.E.end(), 0.0);
    std::fill(sys.B.begin(), sys.B.end(), 0.0); std::fill(sys.gc.begin(), sys.gc.end(), 0.0);
    std::fill(sys.C.begin(), sys.C.end(), 0.0); std::fill(sys.gb.begin(), sys.gb.end(), 0.0);
    sys.c0 = 0;
    for (const Lin& f : lins) {
      const int m = f.m;
      int ch[5], off[5];
      for (int v = 0; v < f.nv; v++) locate(f.v[v], ch[v], off[v]);
      for (int r = 0; r < m; r++) sys.c0 += 0.5 * f.b[r] * f.b[r];
      for (int v = 0; v < f.nv; v++) {
        // rhs
        for (int c = 0; c < f.d[v]; c++) {
          double s = 0; for (int r = 0; r < m; r++) s += f.A[v][r + c * m] * f.b[r];
          if (ch[v] >= 0) sys.gc[(size_t)ch[v] * bs + off[v] + c] += s; else sys.gb[off[v] + c] += s;
        }
        for (int w = 0; w < f.nv; w++) {
          for (int c1 = 0; c1 < f.d[v]; c1++) for (int c2 = 0; c2 < f.d[w]; c2++) {
            double s = 0; for (int r = 0; r < m; r++) s += f.A[v][r + c1 * m] * f.A[w][r + c2 * m];
            const int r1 = off[v] + c1, r2 = off[w] + c2;
            if (ch[v] >= 0 && ch[w] >= 0) {
              if (ch[v] == ch[w]) sys.Dg[(size_t)ch[v] * bs * bs + r1 + r2 * bs] += s;
              else if (ch[v] == ch[w] + 1) sys.E[(size_t)ch[w] * bs * bs + r1 + r2 * bs] += s;  // rows k+1, cols k
              else if (ch[w] == ch[v] + 1) { /* upper mirror: skipped */ }
              else { std::fprintf(stderr, "gpo: non-adjacent chain coupling\n"); std::abort(); }
            } else if (ch[v] >= 0 && ch[w] < 0) {
              sys.B[(size_t)ch[v] * bs * nb + r1 + r2 * bs] += s;
            } else if (ch[v] < 0 && ch[w] < 0) {
              sys.C[r1 + (size_t)r2 * nb] += s;
            }
          }
        }
      }
    }
  }

  // Solve (H + lambda I) delta = g. Returns false if not positive definite.
  bool solve(double lambda) {
    const int nc = sys.nc, bs = sys.bs, nb = sys.nb, w = nb + 1;
    Lf.resize((size_t)nc * bs * bs); Le.resize((size_t)std::max(0, nc - 1) * bs * bs); Y.resize((size_t)nc * bs * w);
    dchain.assign((size_t)nc * bs, 0.0); dborder.assign(nb, 0.0);
    std::vector<double> Dk(bs * bs), S((size_t)nb * nb), sb(nb), Pk((size_t)bs * w), Pn((size_t)bs * w);
    for (int r = 0; r < nb; r++) { for (int c = 0; c < nb; c++) S[r + (size_t)c * nb] = sys.C[r + (size_t)c * nb]; S[r + (size_t)r * nb] += lambda; sb[r] = sys.gb[r]; }
    std::vector<double> Dnext(bs * bs);
    bool have_next = false;
    for (int k = 0; k < nc; k++) {
      const double* Din = sys.Dg.data() + (size_t)k * bs * bs;
      if (have_next) for (int t = 0; t < bs * bs; t++) Dk[t] = Dnext[t];
      else { for (int t = 0; t < bs * bs; t++) Dk[t] = Din[t]; for (int r = 0; r < bs; r++) Dk[r + r * bs] += lambda; }
      if (!have_next) { for (int c = 0; c < nb; c++) for (int r = 0; r < bs; r++) Pk[r + (size_t)c * bs] = sys.B[(size_t)k * bs * nb + r + (size_t)c * bs]; for (int r = 0; r < bs; r++) Pk[r + (size_t)nb * bs] = sys.gc[(size_t)k * bs + r]; }
      else Pk.swap(Pn);
      // Cholesky (lower) of Dk
      double* Lk = Lf.data() + (size_t)k * bs * bs;
      for (int t = 0; t < bs * bs; t++) Lk[t] = 0;
      for (int j = 0; j < bs; j++) {
        double d = Dk[j + j * bs];
        for (int t = 0; t < j; t++) d -= Lk[j + t * bs] * Lk[j + t * bs];
        if (!(d > 0)) return false;
        const double ljj = std::sqrt(d);
        Lk[j + j * bs] = ljj;
        for (int r = j + 1; r < bs; r++) { double s = Dk[r + j * bs]; for (int t = 0; t < j; t++) s -= Lk[r + t * bs] * Lk[j + t * bs]; Lk[r + j * bs] = s / ljj; }
      }
      // Y_k = L^-1 [B_k | g_k]
      double* Yk = Y.data() + (size_t)k * bs * w;
      for (int c = 0; c < w; c++) for (int r = 0; r < bs; r++) { double s = Pk[r + (size_t)c * bs]; for (int t = 0; t < r; t++) s -= Lk[r + t * bs] * Yk[t + (size_t)c * bs]; Yk[r + (size_t)c * bs] = s / Lk[r + r * bs]; }
      // border Schur: S -= Yb^T Yb ; sb -= Yb^T y
      for (int c2 = 0; c2 < nb; c2++) for (int c1 = c2; c1 < nb; c1++) { double s = 0; for (int r = 0; r < bs; r++) s += Yk[r + (size_t)c1 * bs] * Yk[r + (size_t)c2 * bs]; S[c1 + (size_t)c2 * nb] -= s; }
      for (int c1 = 0; c1 < nb; c1++) { double s = 0; for (int r = 0; r < bs; r++) s += Yk[r + (size_t)c1 * bs] * Yk[r + (size_t)nb * bs]; sb[c1] -= s; }
      have_next = false;
      if (k + 1 < nc) {
        // Le = E_k L^-T  (E_k: rows k+1, cols k)
        const double* Ek = sys.E.data() + (size_t)k * bs * bs;
        double* Lek = Le.data() + (size_t)k * bs * bs;
        for (int r = 0; r < bs; r++) for (int c = 0; c < bs; c++) { double s = Ek[r + c * bs]; for (int t = 0; t < c; t++) s -= Lek[r + t * bs] * Lk[c + t * bs]; Lek[r + c * bs] = s / Lk[c + c * bs]; }
        const double* Dn = sys.Dg.data() + (size_t)(k + 1) * bs * bs;
        for (int c = 0; c < bs; c++) for (int r = 0; r < bs; r++) { double s = 0; for (int t = 0; t < bs; t++) s += Lek[r + t * bs] * Lek[c + t * bs]; Dnext[r + c * bs] = Dn[r + c * bs] - s + (r == c ? lambda : 0.0); }
        for (int c = 0; c < w; c++) for (int r = 0; r < bs; r++) {
          double s = (c < nb) ? sys.B[(size_t)(k + 1) * bs * nb + r + (size_t)c * bs] : sys.gc[(size_t)(k + 1) * bs + r];
          for (int t = 0; t < bs; t++) s -= Lek[r + t * bs] * Yk[t + (size_t)c * bs];
          Pn[r + (size_t)c * bs] = s;
        }
        have_next = true;
      }
    }
    // border solve: S (lower stored) = Lb Lb^T
    for (int j = 0; j < nb; j++) {
      double d = S[j + (size_t)j * nb];
      for (int t = 0; t < j; t++) d -= S[j + (size_t)t * nb] * S[j + (size_t)t * nb];
      if (!(d > 0)) return false;
      const double ljj = std::sqrt(d);
      S[j + (size_t)j * nb] = ljj;
      for (int r = j + 1; r < nb; r++) { double s = S[r + (size_t)j * nb]; for (int t = 0; t < j; t++) s -= S[r + (size_t)t * nb] * S[j + (size_t)t * nb]; S[r + (size_t)j * nb] = s / ljj; }
    }
    for (int r = 0; r < nb; r++) { double s = sb[r]; for (int t = 0; t < r; t++) s -= S[r + (size_t)t * nb] * dborder[t]; dborder[r] = s / S[r + (size_t)r * nb]; }
    for (int r = nb - 1; r >= 0; r--) { double s = dborder[r]; for (int t = r + 1; t < nb; t++) s -= S[t + (size_t)r * nb] * dborder[t]; dborder[r] = s / S[r + (size_t)r * nb]; }
    // back-substitute chain
    std::vector<double> rhs(bs);
    for (int k = nc - 1; k >= 0; k--) {
      const double* Lk = Lf.data() + (size_t)k * bs * bs;
      const double* Yk = Y.data() + (size_t)k * bs * w;
      for (int r = 0; r < bs; r++) { double s = Yk[r + (size_t)nb * bs]; for (int c = 0; c < nb; c++) s -= Yk[r + (size_t)c * bs] * dborder[c]; rhs[r] = s; }
      if (k + 1 < nc) { const double* Lek = Le.data() + (size_t)k * bs * bs; for (int c = 0; c < bs; c++) { double s = 0; for (int r = 0; r < bs; r++) s += Lek[r + c * bs] * dchain[(size_t)(k + 1) * bs + r]; rhs[c] -= s; } }
      for (int r = bs - 1; r >= 0; r--) { double s = rhs[r]; for (int t = r + 1; t < bs; t++) s -= Lk[t + r * bs] * dchain[(size_t)k * bs + t]; dchain[(size_t)k * bs + r] = s / Lk[r + r * bs]; }
    }
    return true;
  }

  // linearised error at delta: 0.5 |A delta - b|^2 = c0 - g^T d + 0.5 d^T H d   (undamped H)
  double linear_error() const {
    const int nc = sys.nc, bs = sys.bs, nb = sys.nb;
    double gd = 0, dHd = 0;
    for (int k = 0; k < nc; k++) {
      const double* d = dchain.data() + (size_t)k * bs;
      for (int r = 0; r < bs; r++) gd += sys.gc[(size_t)k * bs + r] * d[r];
      const double* Dk = sys.Dg.data() + (size_t)k * bs * bs;
      for (int c = 0; c < bs; c++) for (int r = 0; r < bs; r++) dHd += d[r] * Dk[r + c * bs] * d[c];
      if (k + 1 < nc) { const double* Ek = sys.E.data() + (size_t)k * bs * bs; const double* dn = dchain.data() + (size_t)(k + 1) * bs; for (int c = 0; c < bs; c++) for (int r = 0; r < bs; r++) dHd += 2.0 * dn[r] * Ek[r + c * bs] * d[c]; }
      const double* Bk = sys.B.data() + (size_t)k * bs * nb;
      for (int c = 0; c < nb; c++) for (int r = 0; r < bs; r++) dHd += 2.0 * d[r] * Bk[r + (size_t)c * bs] * dborder[c];
    }
    for (int r = 0; r < nb; r++) gd += sys.gb[r] * dborder[r];
    // C holds full symmetric entries
    for (int c = 0; c < nb; c++) for (int r = 0; r < nb; r++) dHd += dborder[r] * sys.C[r + (size_t)c * nb] * dborder[c];
    return sys.c0 - gd + 0.5 * dHd;
  }
};

struct Params {
  int max_iterations = 100; double rel_tol = 1e-5, abs_tol = 1e-5, err_tol = 0.0;
  double lambda_initial = 1e-5, lambda_factor = 10.0, lambda_upper = 1e5, lambda_lower = 0.0, min_model_fidelity = 1e-3;
  int use_lm = 1;
};
struct Stats { int iterations; double error_initial, error_final, lambda; double lin_seconds, solve_seconds, total_seconds; int status; };

double now() {
  timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

struct Optimizer {
  Graph* g; Solver s; Params p; double lambda; double error; int iterations = 0;
  std::vector<double> Pn, Vn, Ln;
  double t_lin = 0, t_solve = 0;
  void init(Graph* g_, const Params& p_) {
    g = g_; p = p_; s.g = g; s.setup(); lambda = p.lambda_initial;
    Pn.resize(g->poses.size()); Vn.resize(g->vels.size()); Ln.resize(g->lands.size());
    error = graph_error(*g, g->poses.data(), g->vels.data(), g->lands.data());
  }
  bool try_step(double lam, double& newError, double& linErr) {
    double t0 = now();
    const bool ok = s.solve(lam);
    t_solve += now() - t0;
    if (!ok) return false;
    linErr = s.linear_error();
    retract(*g, g->poses.data(), g->vels.data(), g->lands.data(), s.dchain.data(), s.dborder.data(), s.chain_of_state, s.border_of_state, s.land_off,
            Pn.data(), Vn.data(), Ln.data());
    t0 = now();
    newError = graph_error(*g, Pn.data(), Vn.data(), Ln.data());
    t_lin += now() - t0;
    return true;
  }
  void accept(double newError) { g->poses.swap(Pn); g->vels.swap(Vn); g->lands.swap(Ln); error = newError; }
  // one GN or LM iteration; returns 0 ok, 1 failed (indeterminate / lambda bound)
  int iterate() {
    double t0 = now();
    s.linearize(g->poses.data(), g->vels.data(), g->lands.data());
    s.assemble();
    t_lin += now() - t0;
    int status = 0;
    if (!p.use_lm) {
      double ne, le;
      if (!try_step(0.0, ne, le)) status = 1; else accept(ne);
    } else {
      while (true) {  // gtsam::LevenbergMarquardtOptimizer::iterate (4.0)
        double newError = 0, linErr = 0;
        bool success = false, stop = false;
        const bool solved = try_step(lambda, newError, linErr);
        if (solved) {
          const double linearizedCostChange = error - linErr;
          if (linearizedCostChange >= 0) {
            const double costChange = error - newError;
            double modelFidelity = 0;
            if (linearizedCostChange > 1e-20) modelFidelity = costChange / linearizedCostChange;
            success = modelFidelity > p.min_model_fidelity;
            const double minAbsTol = p.rel_tol * error;
            if (std::fabs(costChange) < minAbsTol) stop = true;
          }
        }
        if (success) { accept(newError); lambda = std::max(p.lambda_lower, lambda / p.lambda_factor); break; }
        else if (!stop) { lambda *= p.lambda_factor; if (lambda >= p.lambda_upper) { status = 1; break; } }
        else break;
      }
    }
    iterations++;
    return status;
  }
  int optimize() {  // gtsam::NonlinearOptimizer::defaultOptimize
    if (error <= p.err_tol) return 0;
    int status = 0;
    double currentError;
    do {
      currentError = error;
      status = iterate();
      if (status) break;
      const double newError = error;
      if (newError <= p.err_tol) break;
      const double absDec = currentError - newError, relDec = absDec / currentError;
      if ((p.rel_tol && relDec <= p.rel_tol) || absDec <= p.abs_tol) break;
    } while (iterations < p.max_iterations);
    return status;
  }
};

}  // namespace

// ==================================================================== C API (ctypes)
extern "C" {

struct gpo_params { int max_iterations; double rel_tol, abs_tol, err_tol, lambda_initial, lambda_factor, lambda_upper, lambda_lower, min_model_fidelity; int use_lm; };
struct gpo_stats { int iterations; double error_initial, error_final, lambda, lin_seconds, solve_seconds, total_seconds; int status; };

void* gpo_graph_create(int group, int dim, int n_states, int n_landmarks) {
  Graph* g = new Graph();
  g->group = group;
  g->D = group == G_POSE3 ? 6 : group == G_LINEAR ? dim : 3;
  g->PS = pose_storage(group, g->D); g->DL = land_dim(group, g->D);
  g->N = n_states; g->L = g->DL ? n_landmarks : 0;
  g->poses.assign((size_t)n_states * g->PS, 0.0); g->vels.assign((size_t)n_states * g->D, 0.0); g->lands.assign((size_t)g->L * g->DL, 0.0);
  return g;
}
void gpo_graph_destroy(void* h) { delete (Graph*)h; }
void gpo_set_threads(void* h, int t) { ((Graph*)h)->threads = t < 1 ? 1 : t; }
int gpo_add_qc_model(void* h, const double* Qc) { Graph* g = (Graph*)h; g->Qc.emplace_back(Qc, Qc + g->D * g->D); return (int)g->Qc.size() - 1; }

static void set_R(Factor& f, int m, const double* R) { f.m = m; for (int k = 0; k < m * m; k++) f.R[k] = R[k]; }

int gpo_add_gp_prior(void* h, int n, const int* i, const double* delta_t, int qc) {
  Graph* g = (Graph*)h;
  for (int k = 0; k < n; k++) { Factor f; f.kind = F_GP_PRIOR; f.i = i[k]; f.delta_t = delta_t[k]; f.qc = qc; f.m = 2 * g->D; g->factors.push_back(f); }
  return 0;
}
int gpo_add_interp_range(void* h, int n, const int* i, const int* l, const double* z, const double* sigma, const double* delta_t, const double* tau, int qc, const double* body_P_sensor) {
  Graph* g = (Graph*)h;
  const int ps = g->group == G_POSE3 ? 12 : 3;
  for (int k = 0; k < n; k++) {
    Factor f; f.kind = F_INTERP_RANGE; f.i = i[k]; f.l = l[k]; f.z = z[k]; f.delta_t = delta_t[k]; f.tau = tau[k]; f.qc = qc; f.m = 1; f.R[0] = 1.0 / sigma[k];
    if (body_P_sensor) { f.has_sensor = true; for (int t = 0; t < ps; t++) f.aux[t] = body_P_sensor[t]; }
    g->factors.push_back(f);
  }
  return 0;
}
int gpo_add_interp_gps(void* h, int n, const int* i, const double* meas, const double* sqrt_info, const double* delta_t, const double* tau, int qc, const double* body_P_sensor) {
  Graph* g = (Graph*)h;
  if (g->group != G_POSE3) return -1;
  for (int k = 0; k < n; k++) {
    Factor f; f.kind = F_INTERP_GPS; f.i = i[k]; f.delta_t = delta_t[k]; f.tau = tau[k]; f.qc = qc;
    for (int t = 0; t < 3; t++) f.meas[t] = meas[3 * k + t];
    set_R(f, 3, sqrt_info);
    if (body_P_sensor) { f.has_sensor = true; for (int t = 0; t < 12; t++) f.aux[t] = body_P_sensor[t]; }
    g->factors.push_back(f);
  }
  return 0;
}
// Pose3 "VW" family: vels of the graph are read as [v_world(3); w_world(3)] by these two factor kinds
int gpo_add_gp_prior_vw(void* h, int n, const int* i, const double* delta_t, int qc) {
  Graph* g = (Graph*)h;
  if (g->group != G_POSE3) return -1;
  for (int k = 0; k < n; k++) { Factor f; f.kind = F_GP_PRIOR_VW; f.i = i[k]; f.delta_t = delta_t[k]; f.qc = qc; f.m = 12; g->factors.push_back(f); }
  return 0;
}
int gpo_add_interp_gps_vw(void* h, int n, const int* i, const double* meas, const double* sqrt_info, const double* delta_t, const double* tau, int qc, const double* body_P_sensor) {
  Graph* g = (Graph*)h;
  if (g->group != G_POSE3) return -1;
  for (int k = 0; k < n; k++) {
    Factor f; f.kind = F_INTERP_GPS_VW; f.i = i[k]; f.delta_t = delta_t[k]; f.tau = tau[k]; f.qc = qc;
    for (int t = 0; t < 3; t++) f.meas[t] = meas[3 * k + t];
    set_R(f, 3, sqrt_info);
    if (body_P_sensor) { f.has_sensor = true; for (int t = 0; t < 12; t++) f.aux[t] = body_P_sensor[t]; }
    g->factors.push_back(f);
  }
  return 0;
}
int gpo_add_interp_projection(void* h, int n, const int* i, const int* l, const double* meas, const double* sqrt_info, const double* delta_t, const double* tau, int qc,
                              const double* K, const double* body_P_sensor) {
  Graph* g = (Graph*)h;
  if (g->group != G_POSE3) return -1;
  for (int k = 0; k < n; k++) {
    Factor f; f.kind = F_INTERP_PROJECTION; f.i = i[k]; f.l = l[k]; f.delta_t = delta_t[k]; f.tau = tau[k]; f.qc = qc;
    for (int t = 0; t < 2; t++) f.meas[t] = meas[2 * k + t];
    for (int t = 0; t < 5; t++) f.K[t] = K[t];
    set_R(f, 2, sqrt_info);
    if (body_P_sensor) { f.has_sensor = true; for (int t = 0; t < 12; t++) f.aux[t] = body_P_sensor[t]; }
    g->factors.push_back(f);
  }
  return 0;
}
int gpo_add_interp_attitude(void* h, int n, const int* i, const double* delta_t, const double* tau, int qc, const double* nZ, const double* bRef, const double* sigma) {
  Graph* g = (Graph*)h;
  for (int k = 0; k < n; k++) {
    Factor f; f.kind = F_INTERP_ATTITUDE; f.i = i[k]; f.delta_t = delta_t[k]; f.tau = tau[k]; f.qc = qc; f.m = 2;
    for (int t = 0; t < 3; t++) { f.aux[t] = nZ[3 * k + t]; f.aux[3 + t] = bRef[3 * k + t]; }
    f.R[0] = 1.0 / sigma[k]; f.R[3] = 1.0 / sigma[k];
    g->factors.push_back(f);
  }
  return 0;
}
int gpo_add_prior_pose(void* h, int i, const double* value, const double* sqrt_info) {
  Graph* g = (Graph*)h; Factor f; f.kind = F_PRIOR_POSE; f.i = i; for (int t = 0; t < g->PS; t++) f.aux[t] = value[t]; set_R(f, g->D, sqrt_info); g->factors.push_back(f); return 0;
}
int gpo_add_prior_vel(void* h, int i, const double* value, const double* sqrt_info) {
  Graph* g = (Graph*)h; Factor f; f.kind = F_PRIOR_VEL; f.i = i; for (int t = 0; t < g->D; t++) f.aux[t] = value[t]; set_R(f, g->D, sqrt_info); g->factors.push_back(f); return 0;
}
int gpo_add_prior_landmark(void* h, int l, const double* value, const double* sqrt_info) {
  Graph* g = (Graph*)h; Factor f; f.kind = F_PRIOR_LANDMARK; f.l = l; for (int t = 0; t < g->DL; t++) f.aux[t] = value[t]; set_R(f, g->DL, sqrt_info); g->factors.push_back(f); return 0;
}
int gpo_add_between(void* h, int i, int j, const double* meas, const double* sqrt_info) {
  Graph* g = (Graph*)h; Factor f; f.kind = F_BETWEEN; f.i = i; f.j = j; for (int t = 0; t < g->PS; t++) f.aux[t] = meas[t]; set_R(f, g->D, sqrt_info); g->factors.push_back(f); return 0;
}
int gpo_add_range_2d(void* h, int i, int l, double z, double sigma) {
  Graph* g = (Graph*)h; Factor f; f.kind = F_RANGE_2D; f.i = i; f.l = l; f.z = z; f.m = 1; f.R[0] = 1.0 / sigma; g->factors.push_back(f); return 0;
}
int gpo_add_range_bearing_2d(void* h, int i, int l, double range, double bearing, const double* sqrt_info) {
  Graph* g = (Graph*)h; Factor f; f.kind = F_RANGE_BEARING_2D; f.i = i; f.l = l; f.z = range; f.z2 = bearing; set_R(f, 2, sqrt_info); g->factors.push_back(f); return 0;
}
int gpo_add_odometry_2d(void* h, int i, int j, const double* meas, const double* sqrt_info) {
  Graph* g = (Graph*)h; Factor f; f.kind = F_ODOMETRY_2D; f.i = i; f.j = j; for (int t = 0; t < 3; t++) f.aux[t] = meas[t]; set_R(f, 3, sqrt_info); g->factors.push_back(f); return 0;
}
int gpo_set_values(void* h, const double* poses, const double* vels, const double* lands) {
  Graph* g = (Graph*)h;
  if (poses) g->poses.assign(poses, poses + g->poses.size());
  if (vels) g->vels.assign(vels, vels + g->vels.size());
  if (lands && g->L) g->lands.assign(lands, lands + g->lands.size());
  return 0;
}
int gpo_get_values(void* h, double* poses, double* vels, double* lands) {
  Graph* g = (Graph*)h;
  if (poses) std::copy(g->poses.begin(), g->poses.end(), poses);
  if (vels) std::copy(g->vels.begin(), g->vels.end(), vels);
  if (lands && g->L) std::copy(g->lands.begin(), g->lands.end(), lands);
  return 0;
}
int gpo_num_factors(void* h) { return (int)((Graph*)h)->factors.size(); }
double gpo_error(void* h) { Graph* g = (Graph*)h; return graph_error(*g, g->poses.data(), g->vels.data(), g->lands.data()); }

// Whitened [A|b] of factor k at the current values.  A_out: concatenated column-major blocks
// (m x d_v each, in the factor's variable order), b_out: m.  Returns m; dims_out[5] gets d_v (0 padded).
int gpo_linearize_factor(void* h, int k, double* A_out, double* b_out, int* dims_out) {
  Graph* g = (Graph*)h; Lin lin;
  linearize_factor(*g, g->factors[k], g->poses.data(), g->vels.data(), g->lands.data(), true, lin);
  int o = 0;
  for (int v = 0; v < 5; v++) dims_out[v] = 0;
  for (int v = 0; v < lin.nv; v++) { dims_out[v] = lin.d[v]; for (int t = 0; t < lin.m * lin.d[v]; t++) A_out[o++] = lin.A[v][t]; }
  for (int r = 0; r < lin.m; r++) b_out[r] = lin.b[r];
  return lin.m;
}
// Unwhitened evaluateError of factor k: e_out (m), H_out concatenated column-major blocks (or null).
int gpo_eval_factor(void* h, int k, double* e_out, double* H_out, int* dims_out) {
  Graph* g = (Graph*)h; Lin lin;
  eval_factor(*g, g->factors[k], g->poses.data(), g->vels.data(), g->lands.data(), H_out != nullptr, lin, e_out);
  for (int v = 0; v < 5; v++) dims_out[v] = v < lin.nv ? lin.d[v] : 0;
  if (H_out) { int o = 0; for (int v = 0; v < lin.nv; v++) for (int t = 0; t < lin.m * lin.d[v]; t++) H_out[o++] = lin.A[v][t]; }
  return lin.m;
}

// Step check at any size (test infrastructure; O(non-zeros), threaded over factors): given a step delta in the oracle's variable
// order (states [x_i, v_i] then landmarks), evaluates r = (J^T J + lambda I) delta - J^T b with the oracle's own whitened
// Jacobians at the current values.  out[0] = max |r|, out[1] = max |J^T b|, out[2] = max |delta|, out[3] = 0.5 |J delta - b|^2
// (the linearised error at delta), out[4] = max (|J|^T |J| |delta|) - the scale a backward-stable solver's residual is measured
// against (|r| <= c eps (|H| |delta| + |g|)).  This is how the
// full-size configs (C5: 10^6 states, 128 loop closures) are held to the oracle without the oracle's dense-border solve.
int gpo_check_step(void* h, const double* delta_states, const double* delta_lands, double lambda, double* out) {
  Graph* g = (Graph*)h;
  const int bs = 2 * g->D, nf = (int)g->factors.size();
  const size_t n = (size_t)g->N * bs + (size_t)g->L * g->DL;
  const int T = std::min(16, std::max(1, g->threads));  // per-thread accumulators of the full system size: bounded
  std::vector<std::vector<double>> acc(T, std::vector<double>(n, 0.0)), gacc(T, std::vector<double>(n, 0.0)), aacc(T, std::vector<double>(n, 0.0));
  std::vector<double> lin_err(T, 0.0);
  auto dptr = [&](const VarRef& v) -> const double* { return v.type == 2 ? delta_lands + (size_t)v.idx * g->DL : delta_states + (size_t)v.idx * bs + (v.type == 1 ? g->D : 0); };
  auto off = [&](const VarRef& v) -> size_t { return v.type == 2 ? (size_t)g->N * bs + (size_t)v.idx * g->DL : (size_t)v.idx * bs + (v.type == 1 ? g->D : 0); };
  auto work = [&](int w) {
    const int lo = (int)((long long)nf * w / T), hi = (int)((long long)nf * (w + 1) / T);
    for (int k = lo; k < hi; k++) {
      Lin lin; linearize_factor(*g, g->factors[k], g->poses.data(), g->vels.data(), g->lands.data(), true, lin);
      double jd[12], ja[12];
      for (int r = 0; r < lin.m; r++) { jd[r] = 0.0; ja[r] = 0.0; }
      for (int v = 0; v < lin.nv; v++) { const double* d = dptr(lin.v[v]); for (int c = 0; c < lin.d[v]; c++) for (int r = 0; r < lin.m; r++) { jd[r] += lin.A[v][r + c * lin.m] * d[c]; ja[r] += std::fabs(lin.A[v][r + c * lin.m] * d[c]); } }
      for (int r = 0; r < lin.m; r++) lin_err[w] += 0.5 * (jd[r] - lin.b[r]) * (jd[r] - lin.b[r]);
      for (int v = 0; v < lin.nv; v++) {
        const size_t o = off(lin.v[v]);
        for (int c = 0; c < lin.d[v]; c++) {
          double s1 = 0, s2 = 0, s3 = 0;
          for (int r = 0; r < lin.m; r++) { s1 += lin.A[v][r + c * lin.m] * jd[r]; s2 += lin.A[v][r + c * lin.m] * lin.b[r]; s3 += std::fabs(lin.A[v][r + c * lin.m]) * ja[r]; }
          acc[w][o + c] += s1; gacc[w][o + c] += s2; aacc[w][o + c] += s3;
        }
      }
    }
  };
  if (T == 1) work(0);
  else { std::vector<std::thread> pool; for (int w = 0; w < T; w++) pool.emplace_back(work, w); for (auto& t : pool) t.join(); }
  double rmax = 0, gmax = 0, dmax = 0, le = 0, hmax = 0;
  for (int w = 0; w < T; w++) le += lin_err[w];
  for (size_t t = 0; t < n; t++) {
    double a = 0, gg = 0, ha = 0;
    for (int w = 0; w < T; w++) { a += acc[w][t]; gg += gacc[w][t]; ha += aacc[w][t]; }
    hmax = std::max(hmax, ha);
    const double d = t < (size_t)g->N * bs ? delta_states[t] : delta_lands[t - (size_t)g->N * bs];
    rmax = std::max(rmax, std::fabs(a + lambda * d - gg)); gmax = std::max(gmax, std::fabs(gg)); dmax = std::max(dmax, std::fabs(d));
  }
  out[0] = rmax; out[1] = gmax; out[2] = dmax; out[3] = le; out[4] = hmax;
  return 0;
}

// Dense normal equations in the oracle's variable order (states [x_i, v_i] for i = 0..N-1, then
// landmarks): H (n x n col-major, full symmetric), g (n).  For small parity cases only.
int gpo_normal_equations_dense(void* h, double* H, double* gvec, int n_expected) {
  Graph* g = (Graph*)h;
  const int bs = 2 * g->D, n = g->N * bs + g->L * g->DL;
  if (n != n_expected) return -1;
  for (size_t t = 0; t < (size_t)n * n; t++) H[t] = 0;
  for (int t = 0; t < n; t++) gvec[t] = 0;
  for (const Factor& f : g->factors) {
    Lin lin; linearize_factor(*g, f, g->poses.data(), g->vels.data(), g->lands.data(), true, lin);
    int off[5];
    for (int v = 0; v < lin.nv; v++) off[v] = lin.v[v].type == 2 ? g->N * bs + lin.v[v].idx * g->DL : lin.v[v].idx * bs + (lin.v[v].type == 1 ? g->D : 0);
    for (int v = 0; v < lin.nv; v++) {
      for (int c = 0; c < lin.d[v]; c++) { double s = 0; for (int r = 0; r < lin.m; r++) s += lin.A[v][r + c * lin.m] * lin.b[r]; gvec[off[v] + c] += s; }
      for (int w = 0; w < lin.nv; w++) for (int c1 = 0; c1 < lin.d[v]; c1++) for (int c2 = 0; c2 < lin.d[w]; c2++) {
        double s = 0; for (int r = 0; r < lin.m; r++) s += lin.A[v][r + c1 * lin.m] * lin.A[w][r + c2 * lin.m];
        H[(off[v] + c1) + (size_t)(off[w] + c2) * n] += s;
      }
    }
  }
  return 0;
}

static Params to_params(const gpo_params* p) {
  Params q;
  if (p) { q.max_iterations = p->max_iterations; q.rel_tol = p->rel_tol; q.abs_tol = p->abs_tol; q.err_tol = p->err_tol; q.lambda_initial = p->lambda_initial;
    q.lambda_factor = p->lambda_factor; q.lambda_upper = p->lambda_upper; q.lambda_lower = p->lambda_lower; q.min_model_fidelity = p->min_model_fidelity; q.use_lm = p->use_lm; }
  return q;
}
void gpo_default_params(gpo_params* p, int use_lm) {
  Params q; p->max_iterations = q.max_iterations; p->rel_tol = q.rel_tol; p->abs_tol = q.abs_tol; p->err_tol = q.err_tol; p->lambda_initial = q.lambda_initial;
  p->lambda_factor = q.lambda_factor; p->lambda_upper = q.lambda_upper; p->lambda_lower = q.lambda_lower; p->min_model_fidelity = q.min_model_fidelity; p->use_lm = use_lm;
}
// run `n_iter` iterations exactly (n_iter > 0) or optimise to convergence (n_iter <= 0)
int gpo_optimize(void* h, const gpo_params* params, int n_iter, gpo_stats* st) {
  Graph* g = (Graph*)h;
  Optimizer opt; const double t0 = now();
  opt.init(g, to_params(params));
  const double e0 = opt.error;
  int status = 0;
  if (n_iter > 0) { for (int k = 0; k < n_iter && !status; k++) status = opt.iterate(); }
  else status = opt.optimize();
  if (st) { st->iterations = opt.iterations; st->error_initial = e0; st->error_final = opt.error; st->lambda = opt.lambda; st->lin_seconds = opt.t_lin; st->solve_seconds = opt.t_solve; st->total_seconds = now() - t0; st->status = status; }
  return status;
}

// ---- free-function entry points used by the golden tests (reference unit-test parity)
void gpo_pose3_expmap(const double* xi, double* T) { pose3_expmap(vec<6>(xi)).to(T); }
void gpo_pose3_logmap(const double* T, double* xi) { put(xi, pose3_logmap(Pose3::from(T))); }
void gpo_pose3_compose(const double* A, const double* B, double* C) { Pose3::from(A).compose(Pose3::from(B)).to(C); }
void gpo_pose3_inverse(const double* A, double* C) { Pose3::from(A).inverse().to(C); }
void gpo_rot3_expmap(const double* w, double* R) { put(R, so3_expmap(vec<3>(w))); }
void gpo_rot3_logmap(const double* R, double* w) { put(w, so3_logmap(get<3, 3>(R))); }
void gpo_rot3_ypr(double y, double p, double r, double* R) { put(R, rot_ypr(y, p, r)); }
void gpo_pose2_expmap(const double* xi, double* T) { const Pose2 p = pose2_expmap(vec<3>(xi)); T[0] = p.x; T[1] = p.y; T[2] = p.th; }
void gpo_pose2_logmap(const double* T, double* xi) { put(xi, pose2_logmap(Pose2(T[0], T[1], T[2]))); }
void gpo_pose2_compose(const double* A, const double* B, double* C) { const Pose2 p = Pose2(A[0], A[1], A[2]).compose(Pose2(B[0], B[1], B[2])); C[0] = p.x; C[1] = p.y; C[2] = p.th; }
void gpo_pose2_inverse(const double* A, double* C) { const Pose2 p = Pose2(A[0], A[1], A[2]).inverse(); C[0] = p.x; C[1] = p.y; C[2] = p.th; }
// which: 0 rightJacobianRot3, 1 rightJacobianRot3inv, 2 leftJacobianRot3, 3 leftJacobianRot3inv
void gpo_so3_jacobian(int which, const double* w, double* J) {
  const Vec3 o = vec<3>(w);
  put(J, which == 0 ? rightJacobianRot3(o) : which == 1 ? rightJacobianRot3inv(o) : which == 2 ? leftJacobianRot3(o) : leftJacobianRot3inv(o));
}
// which: 0 rightJacobianPose3, 1 rightJacobianPose3inv, 2 leftJacobianPose3, 3 leftJacobianPose3inv
void gpo_se3_jacobian(int which, const double* xi, double* J) {
  const Vec6 x = vec<6>(xi);
  put(J, which == 0 ? rightJacobianPose3(x) : which == 1 ? rightJacobianPose3inv(x) : which == 2 ? leftJacobianPose3(x) : leftJacobianPose3inv(x));
}
void gpo_body_centric(int spatial, const double* T1, const double* T2, double dt, double* v) {
  put(v, spatial ? getBodyCentricVs(Pose3::from(T1), Pose3::from(T2), dt) : getBodyCentricVb(Pose3::from(T1), Pose3::from(T2), dt));
}
void gpo_pose2_derivs(const double* xi, double* Jexp, double* Jlog) { put(Jexp, pose2_ExpmapDerivative(vec<3>(xi))); put(Jlog, pose2_LogmapDerivative(pose2_expmap(vec<3>(xi)))); }
// Lambda/Psi (2D x 2D col-major) for D in {1,2,3,6}
void gpo_lambda_psi(int D, const double* Qc, double delta_t, double tau, double* Lambda, double* Psi) {
  switch (D) {
    case 1: put(Lambda, calcLambda<1>(get<1, 1>(Qc), delta_t, tau)); put(Psi, calcPsi<1>(get<1, 1>(Qc), delta_t, tau)); break;
    case 2: put(Lambda, calcLambda<2>(get<2, 2>(Qc), delta_t, tau)); put(Psi, calcPsi<2>(get<2, 2>(Qc), delta_t, tau)); break;
    case 3: put(Lambda, calcLambda<3>(get<3, 3>(Qc), delta_t, tau)); put(Psi, calcPsi<3>(get<3, 3>(Qc), delta_t, tau)); break;
    case 6: put(Lambda, calcLambda<6>(get<6, 6>(Qc), delta_t, tau)); put(Psi, calcPsi<6>(get<6, 6>(Qc), delta_t, tau)); break;
    default: std::abort();
  }
}
void gpo_calcQ(int D, const double* Qc, double tau, double* Q, double* Qinv) {
  switch (D) {
    case 3: put(Q, calcQ<3>(get<3, 3>(Qc), tau)); put(Qinv, calcQ_inv<3>(get<3, 3>(Qc), tau)); break;
    case 6: put(Q, calcQ<6>(get<6, 6>(Qc), tau)); put(Qinv, calcQ_inv<6>(get<6, 6>(Qc), tau)); break;
    default: std::abort();
  }
}
// interpolatePose for any group at the current wire formats; H1..H4 (D x D col-major) or null
void gpo_interpolate(int group, int D, const double* Qc, double delta_t, double tau, const double* p1, const double* v1, const double* p2, const double* v2,
                     double* pose_out, double* H /* 4*D*D or null */) {
  if (group == 4) {  // Pose3 "VW": v1, v2 are [v_world; w_world]; H = [H1 | H2,H3 | H4 | H5,H6]
    const InterpolatorPose3VW gp(get<6, 6>(Qc), delta_t, tau); Mat6 h[4];
    gp.interpolatePose(Pose3::from(p1), vec<3>(v1), vec<3>(v1 + 3), Pose3::from(p2), vec<3>(v2), vec<3>(v2 + 3), H ? &h[0] : nullptr, H ? &h[1] : nullptr, H ? &h[2] : nullptr,
                       H ? &h[3] : nullptr).to(pose_out);
    if (H) for (int k = 0; k < 4; k++) put(H + 36 * k, h[k]);
  } else if (group == G_POSE3) {
    const InterpolatorPose3 gp(get<6, 6>(Qc), delta_t, tau); Mat6 h[4];
    gp.interpolatePose(Pose3::from(p1), vec<6>(v1), Pose3::from(p2), vec<6>(v2), H ? &h[0] : nullptr, H ? &h[1] : nullptr, H ? &h[2] : nullptr, H ? &h[3] : nullptr).to(pose_out);
    if (H) for (int k = 0; k < 4; k++) put(H + 36 * k, h[k]);
  } else if (group == G_POSE2) {
    const InterpolatorPose2 gp(get<3, 3>(Qc), delta_t, tau); Mat3 h[4];
    const Pose2 p = gp.interpolatePose(pose2_from(p1), vec<3>(v1), pose2_from(p2), vec<3>(v2), H ? &h[0] : nullptr, H ? &h[1] : nullptr, H ? &h[2] : nullptr, H ? &h[3] : nullptr);
    pose_out[0] = p.x; pose_out[1] = p.y; pose_out[2] = p.th;
    if (H) for (int k = 0; k < 4; k++) put(H + 9 * k, h[k]);
  } else if (group == G_ROT3) {
    const InterpolatorRot3 gp(get<3, 3>(Qc), delta_t, tau); Mat3 h[4];
    put(pose_out, gp.interpolatePose(rot3_from(p1), vec<3>(v1), rot3_from(p2), vec<3>(v2), H ? &h[0] : nullptr, H ? &h[1] : nullptr, H ? &h[2] : nullptr, H ? &h[3] : nullptr));
    if (H) for (int k = 0; k < 4; k++) put(H + 9 * k, h[k]);
  } else {
    if (D != 3) std::abort();
    const InterpolatorLinear<3> gp(get<3, 3>(Qc), delta_t, tau); Mat3 h[4];
    put(pose_out, gp.interpolatePose(vec<3>(p1), vec<3>(v1), vec<3>(p2), vec<3>(v2), H ? &h[0] : nullptr, H ? &h[1] : nullptr, H ? &h[2] : nullptr, H ? &h[3] : nullptr));
    if (H) for (int k = 0; k < 4; k++) put(H + 9 * k, h[k]);
  }
}
// gp/Pose3utils.cpp:27-64
void gpo_convert_vw_to_vb(const double* v, const double* w, const double* pose, double* v6) { put(v6, convertVWtoVb(vec<3>(v), vec<3>(w), Pose3::from(pose), nullptr, nullptr, nullptr)); }
void gpo_convert_vb_to_vw(const double* v6, const double* pose, double* v, double* w) { Vec3 a, b; convertVbtoVW(vec<6>(v6), Pose3::from(pose), a, b); put(v, a); put(w, b); }
int gpo_hardware_threads() { const unsigned n = std::thread::hardware_concurrency(); return n ? (int)n : 1; }

}  // extern "C"
