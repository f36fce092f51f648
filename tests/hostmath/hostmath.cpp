// TEST HARNESS ONLY: compiles the product's __host__ __device__ factor arithmetic
// (gpslam_b200/csrc/{lie,factors}.cuh) with g++ so CPU tests can compare it with the oracle
// without a GPU.  Not part of the product path.
#include "../../gpslam_b200/csrc/factors.cuh"
using namespace gpb;

template <int VAR, int C, int N> struct EmitP3 {
  static void run(const GpPose3& o, const GpWhiten& w, const double* Rq, double dt, double* out) {
    gp_prior_pose3_col<VAR, C>(o, w, Rq, dt, out + (VAR * 6 + C) * 12);
    EmitP3<VAR, C + 1, N>::run(o, w, Rq, dt, out);
  }
};
template <int VAR, int N> struct EmitP3<VAR, N, N> { static void run(const GpPose3&, const GpWhiten&, const double*, double, double*) {} };
template <int VAR, int C, int N> struct EmitD3 {
  static void run(const GpD3& o, const GpWhiten& w, const double* Rq, double dt, double* out) {
    gp_prior_d3_col<VAR, C>(o, w, Rq, dt, out + (VAR * 3 + C) * 6);
    EmitD3<VAR, C + 1, N>::run(o, w, Rq, dt, out);
  }
};
template <int VAR, int N> struct EmitD3<VAR, N, N> { static void run(const GpD3&, const GpWhiten&, const double*, double, double*) {} };

extern "C" {
// out: m x (4D+1) column-major whitened [A|b]
void hm_gp_prior(int group, const double* s1, const double* s2, double dt, const double* Rq, double* out) {
  const GpWhiten w = gp_whiten(dt);
  if (group == G_POSE3) {  // the production emitter (k_lin_gp<G_POSE3>), dense-Rq instantiation
    gp_prior_pose3_emit<false>(s1, s2, dt, true, w, Rq, [&](int c, const double* col) { for (int k = 0; k < 12; k++) out[c * 12 + k] = col[k]; });
  } else if (group == 100) {  // same, diagonal-Rq instantiation (caller guarantees a diagonal Rq)
    gp_prior_pose3_emit<true>(s1, s2, dt, true, w, Rq, [&](int c, const double* col) { for (int k = 0; k < 12; k++) out[c * 12 + k] = col[k]; });
  } else if (group == 101) {  // the struct-based reference form of the same arithmetic (used by the VW kernel class)
    GpPose3 o; gp_prior_pose3_eval(s1, s2, dt, true, o);
    EmitP3<0, 0, 6>::run(o, w, Rq, dt, out); EmitP3<1, 0, 6>::run(o, w, Rq, dt, out);
    EmitP3<2, 0, 6>::run(o, w, Rq, dt, out); EmitP3<3, 0, 6>::run(o, w, Rq, dt, out);
    gp_prior_pose3_col<4, 0>(o, w, Rq, dt, out + 24 * 12);
  } else {
    GpD3 o;
    if (group == G_POSE2) gp_prior_d3_eval<G_POSE2>(s1, s2, dt, true, o);
    else if (group == G_ROT3) gp_prior_d3_eval<G_ROT3>(s1, s2, dt, true, o);
    else gp_prior_d3_eval<G_LINEAR>(s1, s2, dt, true, o);
    EmitD3<0, 0, 3>::run(o, w, Rq, dt, out); EmitD3<1, 0, 3>::run(o, w, Rq, dt, out);
    EmitD3<2, 0, 3>::run(o, w, Rq, dt, out); EmitD3<3, 0, 3>::run(o, w, Rq, dt, out);
    gp_prior_d3_col<4, 0>(o, w, Rq, dt, out + 12 * 6);
  }
}
// unwhitened rows: out = [H1(D) H2(D) H3(D) H4(D) H5(DL) e] per row
void hm_interp_range(int group, const double* s1, const double* s2, const double* land, const double* prm, double* out) {
  if (group == G_POSE3) {
    Range3Out o; interp_range_pose3(s1, s2, land, prm, true, o);
    for (int k = 0; k < 6; k++) { out[k] = elem(o.H1, k); out[6 + k] = elem(o.H2, k); out[12 + k] = elem(o.H3, k); out[18 + k] = elem(o.H4, k); }
    out[24] = o.H5.x; out[25] = o.H5.y; out[26] = o.H5.z; out[27] = o.e;
  } else {
    Range2Out o;
    if (group == G_POSE2) interp_range_2d<G_POSE2>(s1, s2, land, prm, true, o); else interp_range_2d<G_LINEAR>(s1, s2, land, prm, true, o);
    for (int k = 0; k < 3; k++) { out[k] = elem(o.H1, k); out[3 + k] = elem(o.H2, k); out[6 + k] = elem(o.H3, k); out[9 + k] = elem(o.H4, k); }
    out[12] = o.H5[0]; out[13] = o.H5[1]; out[14] = o.e;
  }
}
void hm_interp_attitude(const double* s1, const double* s2, const double* prm, double* out) {
  AttOut o; interp_attitude_rot3(s1, s2, prm, true, o);
  for (int r = 0; r < 2; r++) {
    for (int k = 0; k < 3; k++) { out[13 * r + k] = elem(o.H1[r], k); out[13 * r + 3 + k] = elem(o.H2[r], k); out[13 * r + 6 + k] = elem(o.H3[r], k); out[13 * r + 9 + k] = elem(o.H4[r], k); }
    out[13 * r + 12] = o.e[r];
  }
}
}

// ---- SE(3) "VW" family and the multi-row SE(3) interpolated factors
template <int VAR, int C, int N> struct EmitVW {
  static void run(const GpPose3VW& o, const GpWhiten& w, const double* Rq, double dt, double* out) {
    gp_prior_pose3vw_col<VAR, C>(o, w, Rq, dt, out + (VAR * 6 + C) * 12);
    EmitVW<VAR, C + 1, N>::run(o, w, Rq, dt, out);
  }
};
template <int VAR, int N> struct EmitVW<VAR, N, N> { static void run(const GpPose3VW&, const GpWhiten&, const double*, double, double*) {} };

extern "C" {
// out: 12 x 25 column-major whitened [A|b] over [x1(6) | v1,w1 | x2(6) | v2,w2 | rhs]
void hm_gp_prior_vw(const double* s1, const double* s2, double dt, const double* Rq, double* out) {
  const GpWhiten w = gp_whiten(dt);
  GpPose3VW o; gp_prior_pose3vw_eval(s1, s2, dt, true, o);
  EmitVW<0, 0, 6>::run(o, w, Rq, dt, out); EmitVW<1, 0, 6>::run(o, w, Rq, dt, out);
  EmitVW<2, 0, 6>::run(o, w, Rq, dt, out); EmitVW<3, 0, 6>::run(o, w, Rq, dt, out);
  gp_prior_pose3vw_col<4, 0>(o, w, Rq, dt, out + 24 * 12);
}
// unwhitened rows of the GPS factors: out[r] = [H1(6) H2(6) H3(6) H4(6) e] for r = 0..2; vw selects GPInterpolatedGPSFactorPose3VW
void hm_interp_gps(int vw, const double* s1, const double* s2, const double* prm, double* out) {
  Gps3Out o;
  if (vw) interp_gps_pose3vw(s1, s2, prm, true, o); else interp_gps_pose3(s1, s2, prm, true, o);
  const double e[3] = {o.e.x, o.e.y, o.e.z};
  for (int r = 0; r < 3; r++) {
    for (int v = 0; v < 4; v++) for (int k = 0; k < 6; k++) out[25 * r + 6 * v + k] = elem(o.H[r][v], k);
    out[25 * r + 24] = e[r];
  }
}
// out[r] = [H1 H2 H3 H4 (6 each) H5(3) e] for r = 0..1
void hm_interp_projection(const double* s1, const double* s2, const double* land, const double* prm, double* out) {
  Proj3Out o; interp_projection_pose3(s1, s2, land, prm, true, o);
  for (int r = 0; r < 2; r++) {
    for (int v = 0; v < 4; v++) for (int k = 0; k < 6; k++) out[28 * r + 6 * v + k] = elem(o.H[r][v], k);
    out[28 * r + 24] = o.H5[r].x; out[28 * r + 25] = o.H5[r].y; out[28 * r + 26] = o.H5[r].z;
    out[28 * r + 27] = o.e[r];
  }
}
}

// ---- interpolatePose as a query: pose (wire) and the four D x D Jacobians, group codes of include/gpb.h (4 = Pose3 VW)
extern "C" void hm_interp_pose(int group, const double* s1, const double* s2, double dt, double tau, double* pose_out, double* H) {
  switch (group) {
    case 0: interp_pose<G_POSE3>(s1, s2, dt, tau, H != nullptr, pose_out, H); break;
    case 1: interp_pose<G_POSE2>(s1, s2, dt, tau, H != nullptr, pose_out, H); break;
    case 2: interp_pose<G_ROT3>(s1, s2, dt, tau, H != nullptr, pose_out, H); break;
    case 3: interp_pose<G_LINEAR>(s1, s2, dt, tau, H != nullptr, pose_out, H); break;
    default: interp_pose<G_POSE3VW>(s1, s2, dt, tau, H != nullptr, pose_out, H); break;
  }
}
