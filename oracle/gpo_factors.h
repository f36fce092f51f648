// ORACLE — TEST INFRASTRUCTURE ONLY (see gpo_linalg.h).
//
// Line-by-line CPU restatement of the reference's hot-path factor arithmetic
// (SURVEY.md §8a rows 1-14), keeping the reference's cost structure on purpose — in
// particular the 2 x 12-evaluation central differences of rightJacobianPose3inv in the
// SE(3) prior and interpolator — because this file is also the timed CPU baseline.
#pragma once
#include "gpo_lie.h"

namespace gpo {

// ---------------------------------------------------------------- gp/GPutils.h
// gp/GPutils.h:24-30
template <int D> Mat<2 * D, 2 * D> calcQ(const Mat<D, D>& Qc, double tau) {
  Mat<2 * D, 2 * D> Q;
  Q.set(0, 0, (1.0 / 3 * std::pow(tau, 3.0)) * Qc); Q.set(0, D, (1.0 / 2 * std::pow(tau, 2.0)) * Qc);
  Q.set(D, 0, (1.0 / 2 * std::pow(tau, 2.0)) * Qc); Q.set(D, D, tau * Qc);
  return Q;
}
// gp/GPutils.h:33-41
template <int D> Mat<2 * D, 2 * D> calcQ_inv(const Mat<D, D>& Qc, double tau) {
  const Mat<D, D> Qc_inv = inverse_gj(Qc);
  Mat<2 * D, 2 * D> Q;
  Q.set(0, 0, (12.0 * std::pow(tau, -3.0)) * Qc_inv); Q.set(0, D, ((-6.0) * std::pow(tau, -2.0)) * Qc_inv);
  Q.set(D, 0, ((-6.0) * std::pow(tau, -2.0)) * Qc_inv); Q.set(D, D, (4.0 * std::pow(tau, -1.0)) * Qc_inv);
  return Q;
}
// gp/GPutils.h:44-51
template <int D> Mat<2 * D, 2 * D> calcPhi(double tau) {
  Mat<2 * D, 2 * D> Phi = Mat<2 * D, 2 * D>::Identity();
  Phi.set(0, D, tau * Mat<D, D>::Identity());
  return Phi;
}
// gp/GPutils.h:54-61
template <int D> Mat<2 * D, 2 * D> calcLambda(const Mat<D, D>& Qc, double delta_t, double tau) {
  return calcPhi<D>(tau) - calcQ<D>(Qc, tau) * (calcPhi<D>(delta_t - tau).t()) * calcQ_inv<D>(Qc, delta_t) * calcPhi<D>(delta_t);
}
// gp/GPutils.h:64-71
template <int D> Mat<2 * D, 2 * D> calcPsi(const Mat<D, D>& Qc, double delta_t, double tau) {
  return calcQ<D>(Qc, tau) * (calcPhi<D>(delta_t - tau).t()) * calcQ_inv<D>(Qc, delta_t);
}

// ---------------------------------------------------------------- GP priors
// gp/GaussianProcessPriorPose3.h:60-98.  H* may be null (cheap path, :73-74).
inline Vec12 gpPriorPose3(const Pose3& pose1, const Vec6& vel1, const Pose3& pose2, const Vec6& vel2, double delta_t,
                          Mat<12, 6>* H1, Mat<12, 6>* H2, Mat<12, 6>* H3, Mat<12, 6>* H4) {
  Mat6 Hinv, Hcomp1, Hcomp2, Hlogmap;
  const Pose3 T12 = pose1.inverse().compose(pose2);
  const Vec6 r = pose3_logmap(T12);
  if (H1 || H2 || H3 || H4) {
    Hinv = pose3_Hinverse(pose1); Hcomp1 = pose3_Hcompose1(pose2); Hcomp2 = Mat6::Identity();
    Hlogmap = rightJacobianPose3inv(r);  // gtsam::Pose3::LogmapDerivative
  }
  const Mat6 Jinv = rightJacobianPose3inv(r);
  if (H1) {
    const Mat6 J_Ti = Hlogmap * Hcomp1 * Hinv;
    const Mat6 Jdiff_Ti = jacobianMethodNumercialDiff(rightJacobianPose3inv, r, vel2) * J_Ti;
    H1->set(0, 0, J_Ti); H1->set(6, 0, Jdiff_Ti);
  }
  if (H2) { H2->set(0, 0, -delta_t * Mat6::Identity()); H2->set(6, 0, -Mat6::Identity()); }
  if (H3) {
    const Mat6 J_Ti1 = Hlogmap * Hcomp2;
    const Mat6 Jdiff_Ti1 = jacobianMethodNumercialDiff(rightJacobianPose3inv, r, vel2) * J_Ti1;
    H3->set(0, 0, J_Ti1); H3->set(6, 0, Jdiff_Ti1);
  }
  if (H4) { H4->set(0, 0, Mat6::Zero()); H4->set(6, 0, Jinv); }
  Vec12 e;
  e.set(0, 0, r - vel1 * delta_t);
  e.set(6, 0, Jinv * vel2 - vel1);
  return e;
}

// ---------------------------------------------------------------- Pose3 "VW" family
// State = (Pose3, linear velocity v, angular velocity w), v and w both expressed in the WORLD frame: convertVWtoVb rotates them
// into the body frame (gp/Pose3utils.cpp:47-64; gtsam::Rot3::unrotate: q = R^T p, dq/dR = [q]x, dq/dp = R^T).
// The oracle carries (v, w) as ONE 6-vector velocity variable [v; w] per state, so the reference's H2|H3 (and H5|H6) sit side
// by side in one 12x6 (or 6x6) block.
inline Vec6 convertVWtoVb(const Vec3& v, const Vec3& w, const Pose3& pose, Mat<6, 3>* Hv, Mat<6, 3>* Hw, Mat6* Hpose) {
  const Mat3 Rt = pose.R.t();
  const Vec3 qw = Rt * w, qv = Rt * v;
  Vec6 v6; v6.set(0, 0, qw); v6.set(3, 0, qv);
  if (Hv) { *Hv = Mat<6, 3>::Zero(); Hv->set(3, 0, Rt); }
  if (Hw) { *Hw = Mat<6, 3>::Zero(); Hw->set(0, 0, Rt); }
  if (Hpose) { *Hpose = Mat6::Zero(); Hpose->set(0, 0, skew(qw)); Hpose->set(3, 0, skew(qv)); }
  return v6;
}
// gp/Pose3utils.cpp:27-45
inline void convertVbtoVW(const Vec6& v6, const Pose3& pose, Vec3& v, Vec3& w) {
  v = pose.R * v6.block<3, 1>(3, 0); w = pose.R * v6.block<3, 1>(0, 0);
}

// gp/GaussianProcessPriorPose3VW.h:62-117.  Hvw1 = [H2 | H3] (12x6 over [v1; w1]), Hvw2 = [H5 | H6].
inline Vec12 gpPriorPose3VW(const Pose3& pose1, const Vec3& vel1, const Vec3& omega1, const Pose3& pose2, const Vec3& vel2, const Vec3& omega2,
                            double delta_t, Mat<12, 6>* H1, Mat<12, 6>* Hvw1, Mat<12, 6>* H4, Mat<12, 6>* Hvw2) {
  const bool wantH = H1 || Hvw1 || H4 || Hvw2;
  Mat6 Hinv, Hcomp1, Hcomp2, Hlogmap;
  const Vec6 r = pose3_logmap(pose1.inverse().compose(pose2));
  if (wantH) { Hinv = pose3_Hinverse(pose1); Hcomp1 = pose3_Hcompose1(pose2); Hcomp2 = Mat6::Identity(); Hlogmap = rightJacobianPose3inv(r); }
  const Mat6 Jinv = rightJacobianPose3inv(r);
  Mat<6, 3> H1v, H1w, H2v, H2w; Mat6 H1p, H2p;
  const Vec6 v1 = convertVWtoVb(vel1, omega1, pose1, wantH ? &H1v : nullptr, wantH ? &H1w : nullptr, wantH ? &H1p : nullptr);
  const Vec6 v2 = convertVWtoVb(vel2, omega2, pose2, wantH ? &H2v : nullptr, wantH ? &H2w : nullptr, wantH ? &H2p : nullptr);
  if (wantH) {
    Mat<12, 6> Hv1, Hv2;
    Hv1.set(0, 0, -delta_t * Mat6::Identity()); Hv1.set(6, 0, -Mat6::Identity());
    Hv2.set(0, 0, Mat6::Zero()); Hv2.set(6, 0, Jinv);
    if (H1) {
      const Mat6 J_Ti = Hlogmap * Hcomp1 * Hinv;
      const Mat6 Jdiff_Ti = jacobianMethodNumercialDiff(rightJacobianPose3inv, r, v2) * J_Ti;
      H1->set(0, 0, J_Ti - delta_t * H1p); H1->set(6, 0, Jdiff_Ti - H1p);
    }
    if (Hvw1) { Hvw1->set(0, 0, Hv1 * H1v); Hvw1->set(0, 3, Hv1 * H1w); }
    if (H4) {
      const Mat6 J_Ti1 = Hlogmap * Hcomp2;
      const Mat6 Jdiff_Ti1 = jacobianMethodNumercialDiff(rightJacobianPose3inv, r, v2) * J_Ti1;
      H4->set(0, 0, J_Ti1); H4->set(6, 0, Jdiff_Ti1 + Jinv * H2p);
    }
    if (Hvw2) { Hvw2->set(0, 0, Hv2 * H2v); Hvw2->set(0, 3, Hv2 * H2w); }
  }
  Vec12 e;
  e.set(0, 0, r - v1 * delta_t);
  e.set(6, 0, Jinv * v2 - v1);
  return e;
}

// gp/GaussianProcessPriorPose2.h:58-82
inline Vec6 gpPriorPose2(const Pose2& pose1, const Vec3& vel1, const Pose2& pose2, const Vec3& vel2, double delta_t,
                         Mat<6, 3>* H1, Mat<6, 3>* H2, Mat<6, 3>* H3, Mat<6, 3>* H4) {
  const Pose2 T12 = pose1.inverse().compose(pose2);
  const Vec3 r = pose2_logmap(T12);
  if (H1 || H2 || H3 || H4) {
    const Mat3 Hinv = -pose1.Adjoint(), Hcomp1 = pose2.inverse().Adjoint(), Hcomp2 = Mat3::Identity();
    const Mat3 Hlogmap = pose2_LogmapDerivative(T12);
    if (H1) { *H1 = Mat<6, 3>::Zero(); H1->set(0, 0, Hlogmap * Hcomp1 * Hinv); }
    if (H3) { *H3 = Mat<6, 3>::Zero(); H3->set(0, 0, Hlogmap * Hcomp2); }
  }
  if (H2) { H2->set(0, 0, -delta_t * Mat3::Identity()); H2->set(3, 0, -Mat3::Identity()); }
  if (H4) { H4->set(0, 0, Mat3::Zero()); H4->set(3, 0, Mat3::Identity()); }
  Vec6 e; e.set(0, 0, r - vel1 * delta_t); e.set(3, 0, vel2 - vel1);
  return e;
}

// gp/GaussianProcessPriorRot3.h:58-79
inline Vec6 gpPriorRot3(const Mat3& pose1, const Vec3& vel1, const Mat3& pose2, const Vec3& vel2, double delta_t,
                        Mat<6, 3>* H1, Mat<6, 3>* H2, Mat<6, 3>* H3, Mat<6, 3>* H4) {
  const Mat3 R12 = pose1.t() * pose2;
  const Vec3 r = so3_logmap(R12);
  if (H1 || H2 || H3 || H4) {
    // Rot3: inverse -> -R1 ; compose(A,B) d/dA = B^T, d/dB = I ; Logmap -> LogmapDerivative(r)
    const Mat3 Hinv = -pose1, Hcomp1 = pose2.t(), Hcomp2 = Mat3::Identity();
    const Mat3 Hlogmap = rightJacobianRot3inv(r);
    if (H1) { *H1 = Mat<6, 3>::Zero(); H1->set(0, 0, Hlogmap * Hcomp1 * Hinv); }
    if (H3) { *H3 = Mat<6, 3>::Zero(); H3->set(0, 0, Hlogmap * Hcomp2); }
  }
  if (H2) { H2->set(0, 0, -delta_t * Mat3::Identity()); H2->set(3, 0, -Mat3::Identity()); }
  if (H4) { H4->set(0, 0, Mat3::Zero()); H4->set(3, 0, Mat3::Identity()); }
  Vec6 e; e.set(0, 0, r - vel1 * delta_t); e.set(3, 0, vel2 - vel1);
  return e;
}

// gp/GaussianProcessPriorLinear.h:63-83
template <int D>
Mat<2 * D, 1> gpPriorLinear(const Mat<D, 1>& pose1, const Mat<D, 1>& vel1, const Mat<D, 1>& pose2, const Mat<D, 1>& vel2, double delta_t,
                            Mat<2 * D, D>* H1, Mat<2 * D, D>* H2, Mat<2 * D, D>* H3, Mat<2 * D, D>* H4) {
  Mat<2 * D, 1> x1, x2;
  x1.set(0, 0, pose1); x1.set(D, 0, vel1); x2.set(0, 0, pose2); x2.set(D, 0, vel2);
  const Mat<D, D> I = Mat<D, D>::Identity(), Z = Mat<D, D>::Zero();
  if (H1) { H1->set(0, 0, I); H1->set(D, 0, Z); }
  if (H2) { H2->set(0, 0, delta_t * I); H2->set(D, 0, I); }
  if (H3) { H3->set(0, 0, -1.0 * I); H3->set(D, 0, Z); }
  if (H4) { H4->set(0, 0, Z); H4->set(D, 0, -1.0 * I); }
  return calcPhi<D>(delta_t) * x1 - x2;
}

// ---------------------------------------------------------------- interpolators
// gp/GaussianProcessInterpolatorPose3.h:43-105
struct InterpolatorPose3 {
  double delta_t, tau; Mat6 Qc; Mat12 Lambda, Psi;
  InterpolatorPose3(const Mat6& Qc_, double dt, double tau_) : delta_t(dt), tau(tau_), Qc(Qc_) {
    Lambda = calcLambda<6>(Qc, dt, tau); Psi = calcPsi<6>(Qc, dt, tau);
  }
  Pose3 interpolatePose(const Pose3& pose1, const Vec6& vel1, const Pose3& pose2, const Vec6& vel2,
                        Mat6* H1, Mat6* H2, Mat6* H3, Mat6* H4) const {
    Vec12 r1 = Vec12::Zero(); r1.set(6, 0, vel1);
    const Pose3 T12 = pose1.inverse().compose(pose2);
    const Vec6 r = pose3_logmap(T12);
    const Mat6 Jinv = rightJacobianPose3inv(r);
    Vec12 r2; r2.set(0, 0, r); r2.set(6, 0, Jinv * vel2);
    const Mat<6, 12> Lam1 = Lambda.block<6, 12>(0, 0), Psi1 = Psi.block<6, 12>(0, 0);
    const Vec6 xi = Lam1 * r1 + Psi1 * r2;
    const Pose3 dT = pose3_expmap(xi);
    const Pose3 pose = pose1.compose(dT);
    if (H1 || H2 || H3 || H4) {
      const Mat6 Hinv = pose3_Hinverse(pose1), Hcomp11 = pose3_Hcompose1(pose2), Hcomp12 = Mat6::Identity();
      const Mat6 Hlogmap = rightJacobianPose3inv(r);
      const Mat6 Hexp = rightJacobianPose3(xi);  // gtsam::Pose3::ExpmapDerivative
      const Mat6 Hcomp21 = pose3_Hcompose1(dT), Hcomp22 = Mat6::Identity();
      const Mat6 Hexpr1 = Hcomp22 * Hexp;
      if (H1) {
        const Mat6 tmp = Hlogmap * Hcomp11 * Hinv;
        Mat<12, 6> dr2_dT1; dr2_dT1.set(0, 0, tmp);
        dr2_dT1.set(6, 0, jacobianMethodNumercialDiff(rightJacobianPose3inv, r, vel2) * tmp);
        *H1 = Hcomp21 + Hexpr1 * Psi1 * dr2_dT1;
      }
      if (H2) *H2 = Hexpr1 * Lambda.block<6, 6>(0, 6);
      if (H3) {
        const Mat6 tmp = Hlogmap * Hcomp12;
        Mat<12, 6> dr2_dT2; dr2_dT2.set(0, 0, tmp);
        dr2_dT2.set(6, 0, jacobianMethodNumercialDiff(rightJacobianPose3inv, r, vel2) * tmp);
        *H3 = Hexpr1 * Psi1 * dr2_dT2;
      }
      if (H4) *H4 = Hexpr1 * Psi.block<6, 6>(0, 6) * Jinv;
    }
    return pose;
  }
};

// gp/GaussianProcessInterpolatorPose3VW.h:43-124 (all six Jacobians requested or none - the only two ways the reference's factor
// calls it, slam/GPInterpolatedGPSFactorPose3VW.h:80-84).  Hvw1 = [H2 | H3], Hvw2 = [H5 | H6].
struct InterpolatorPose3VW {
  double delta_t, tau; Mat6 Qc; Mat12 Lambda, Psi;
  InterpolatorPose3VW(const Mat6& Qc_, double dt, double tau_) : delta_t(dt), tau(tau_), Qc(Qc_) {
    Lambda = calcLambda<6>(Qc, dt, tau); Psi = calcPsi<6>(Qc, dt, tau);
  }
  Pose3 interpolatePose(const Pose3& pose1, const Vec3& v1, const Vec3& omega1, const Pose3& pose2, const Vec3& v2, const Vec3& omega2,
                        Mat6* H1, Mat6* Hvw1, Mat6* H4, Mat6* Hvw2) const {
    const bool wantH = H1 || Hvw1 || H4 || Hvw2;
    const Vec6 r = pose3_logmap(pose1.inverse().compose(pose2));
    const Mat6 Jinv = rightJacobianPose3inv(r);
    Mat<6, 3> H1v, H1w, H2v, H2w; Mat6 H1p, H2p;
    const Vec6 vel1 = convertVWtoVb(v1, omega1, pose1, wantH ? &H1v : nullptr, wantH ? &H1w : nullptr, wantH ? &H1p : nullptr);
    const Vec6 vel2 = convertVWtoVb(v2, omega2, pose2, wantH ? &H2v : nullptr, wantH ? &H2w : nullptr, wantH ? &H2p : nullptr);
    Vec12 r1 = Vec12::Zero(); r1.set(6, 0, vel1);
    Vec12 r2; r2.set(0, 0, r); r2.set(6, 0, Jinv * vel2);
    const Mat<6, 12> Lam1 = Lambda.block<6, 12>(0, 0), Psi1 = Psi.block<6, 12>(0, 0);
    const Vec6 xi = Lam1 * r1 + Psi1 * r2;
    const Pose3 dT = pose3_expmap(xi);
    const Pose3 pose = pose1.compose(dT);
    if (wantH) {
      const Mat6 Hinv = pose3_Hinverse(pose1), Hcomp11 = pose3_Hcompose1(pose2), Hcomp12 = Mat6::Identity();
      const Mat6 Hlogmap = rightJacobianPose3inv(r);
      const Mat6 Hexp = rightJacobianPose3(xi);
      const Mat6 Hcomp21 = pose3_Hcompose1(dT), Hcomp22 = Mat6::Identity();
      const Mat6 Hexpr1 = Hcomp22 * Hexp;
      const Mat6 Hvel1 = Hexpr1 * Lambda.block<6, 6>(0, 6);
      const Mat6 Hvel2 = Hexpr1 * Psi.block<6, 6>(0, 6) * Jinv;
      if (H1) {
        const Mat6 tmp = Hlogmap * Hcomp11 * Hinv;
        Mat<12, 6> dr2_dT1; dr2_dT1.set(0, 0, tmp);
        dr2_dT1.set(6, 0, jacobianMethodNumercialDiff(rightJacobianPose3inv, r, vel2) * tmp);
        *H1 = Hcomp21 + Hexpr1 * Psi1 * dr2_dT1 + Hvel1 * H1p;
      }
      if (Hvw1) { Hvw1->set(0, 0, Hvel1 * H1v); Hvw1->set(0, 3, Hvel1 * H1w); }
      if (H4) {
        const Mat6 tmp = Hlogmap * Hcomp12;
        Mat<12, 6> dr2_dT2; dr2_dT2.set(0, 0, tmp);
        dr2_dT2.set(6, 0, jacobianMethodNumercialDiff(rightJacobianPose3inv, r, vel2) * tmp);
        *H4 = Hexpr1 * Psi1 * dr2_dT2 + Hvel2 * H2p;
      }
      if (Hvw2) { Hvw2->set(0, 0, Hvel2 * H2v); Hvw2->set(0, 3, Hvel2 * H2w); }
    }
    return pose;
  }
};

// gp/GaussianProcessInterpolatorPose2.h:43-89
struct InterpolatorPose2 {
  double delta_t, tau; Mat3 Qc; Mat6 Lambda, Psi;
  InterpolatorPose2(const Mat3& Qc_, double dt, double tau_) : delta_t(dt), tau(tau_), Qc(Qc_) {
    Lambda = calcLambda<3>(Qc, dt, tau); Psi = calcPsi<3>(Qc, dt, tau);
  }
  Pose2 interpolatePose(const Pose2& pose1, const Vec3& vel1, const Pose2& pose2, const Vec3& vel2,
                        Mat3* H1, Mat3* H2, Mat3* H3, Mat3* H4) const {
    Vec6 r1 = Vec6::Zero(); r1.set(3, 0, vel1);
    const Pose2 T12 = pose1.inverse().compose(pose2);
    const Vec3 r = pose2_logmap(T12);
    Vec6 r2; r2.set(0, 0, r); r2.set(3, 0, vel2);
    const Mat<3, 6> Lam1 = Lambda.block<3, 6>(0, 0), Psi1 = Psi.block<3, 6>(0, 0);
    const Vec3 xi = Lam1 * r1 + Psi1 * r2;
    const Pose2 dT = pose2_expmap(xi);
    const Pose2 pose = pose1.compose(dT);
    if (H1 || H2 || H3 || H4) {
      const Mat3 Hinv = -pose1.Adjoint(), Hcomp11 = pose2.inverse().Adjoint(), Hcomp12 = Mat3::Identity();
      const Mat3 Hlogmap = pose2_LogmapDerivative(T12);
      const Mat3 Hexp = pose2_ExpmapDerivative(xi);
      const Mat3 Hcomp21 = dT.inverse().Adjoint(), Hcomp22 = Mat3::Identity();
      const Mat3 Hexpr1 = Hcomp22 * Hexp;
      if (H1) *H1 = Hcomp21 + Hexpr1 * Psi.block<3, 3>(0, 0) * Hlogmap * Hcomp11 * Hinv;
      if (H2) *H2 = Hexpr1 * Lambda.block<3, 3>(0, 3);
      if (H3) *H3 = Hexpr1 * Psi.block<3, 3>(0, 0) * Hlogmap * Hcomp12;
      if (H4) *H4 = Hexpr1 * Psi.block<3, 3>(0, 3);
    }
    return pose;
  }
};

// gp/GaussianProcessInterpolatorRot3.h:43-86
struct InterpolatorRot3 {
  double delta_t, tau; Mat3 Qc; Mat6 Lambda, Psi;
  InterpolatorRot3(const Mat3& Qc_, double dt, double tau_) : delta_t(dt), tau(tau_), Qc(Qc_) {
    Lambda = calcLambda<3>(Qc, dt, tau); Psi = calcPsi<3>(Qc, dt, tau);
  }
  Mat3 interpolatePose(const Mat3& pose1, const Vec3& vel1, const Mat3& pose2, const Vec3& vel2,
                       Mat3* H1, Mat3* H2, Mat3* H3, Mat3* H4) const {
    Vec6 r1 = Vec6::Zero(); r1.set(3, 0, vel1);
    const Vec3 r = so3_logmap(pose1.t() * pose2);
    Vec6 r2; r2.set(0, 0, r); r2.set(3, 0, vel2);
    const Mat<3, 6> Lam1 = Lambda.block<3, 6>(0, 0), Psi1 = Psi.block<3, 6>(0, 0);
    const Vec3 xi = Lam1 * r1 + Psi1 * r2;
    const Mat3 dR = so3_expmap(xi);
    const Mat3 pose = pose1 * dR;
    if (H1 || H2 || H3 || H4) {
      const Mat3 Hinv = -pose1, Hcomp11 = pose2.t(), Hcomp12 = Mat3::Identity();
      const Mat3 Hlogmap = rightJacobianRot3inv(r);
      const Mat3 Hexp = rightJacobianRot3(xi);
      const Mat3 Hcomp21 = dR.t(), Hcomp22 = Mat3::Identity();
      const Mat3 Hexpr1 = Hcomp22 * Hexp;
      if (H1) *H1 = Hcomp21 + Hexpr1 * Psi.block<3, 3>(0, 0) * Hlogmap * Hcomp11 * Hinv;
      if (H2) *H2 = Hexpr1 * Lambda.block<3, 3>(0, 3);
      if (H3) *H3 = Hexpr1 * Psi.block<3, 3>(0, 0) * Hlogmap * Hcomp12;
      if (H4) *H4 = Hexpr1 * Psi.block<3, 3>(0, 3);
    }
    return pose;
  }
};

// gp/GaussianProcessInterpolatorLinear.h:51-126
template <int D> struct InterpolatorLinear {
  double delta_t, tau; Mat<D, D> Qc; Mat<2 * D, 2 * D> Lambda, Psi;
  InterpolatorLinear(const Mat<D, D>& Qc_, double dt, double tau_) : delta_t(dt), tau(tau_), Qc(Qc_) {
    Lambda = calcLambda<D>(Qc, dt, tau); Psi = calcPsi<D>(Qc, dt, tau);
  }
  Mat<D, 1> interp(int row0, const Mat<D, 1>& pose1, const Mat<D, 1>& vel1, const Mat<D, 1>& pose2, const Mat<D, 1>& vel2,
                   Mat<D, D>* H1, Mat<D, D>* H2, Mat<D, D>* H3, Mat<D, D>* H4) const {
    Mat<2 * D, 1> x1, x2;
    x1.set(0, 0, pose1); x1.set(D, 0, vel1); x2.set(0, 0, pose2); x2.set(D, 0, vel2);
    if (H1) *H1 = Lambda.template block<D, D>(row0, 0);
    if (H2) *H2 = Lambda.template block<D, D>(row0, D);
    if (H3) *H3 = Psi.template block<D, D>(row0, 0);
    if (H4) *H4 = Psi.template block<D, D>(row0, D);
    return Lambda.template block<D, 2 * D>(row0, 0) * x1 + Psi.template block<D, 2 * D>(row0, 0) * x2;
  }
  Mat<D, 1> interpolatePose(const Mat<D, 1>& p1, const Mat<D, 1>& v1, const Mat<D, 1>& p2, const Mat<D, 1>& v2,
                            Mat<D, D>* H1, Mat<D, D>* H2, Mat<D, D>* H3, Mat<D, D>* H4) const { return interp(0, p1, v1, p2, v2, H1, H2, H3, H4); }
  Mat<D, 1> interpolateVelocity(const Mat<D, 1>& p1, const Mat<D, 1>& v1, const Mat<D, 1>& p2, const Mat<D, 1>& v2,
                                Mat<D, D>* H1, Mat<D, D>* H2, Mat<D, D>* H3, Mat<D, D>* H4) const { return interp(D, p1, v1, p2, v2, H1, H2, H3, H4); }
};

// ---------------------------------------------------------------- interpolated measurement factors
// slam/GPInterpolatedRangeFactorPose3.h:64-98.  body_P_sensor may be null.
inline double gpRangePose3(const InterpolatorPose3& gp, double measured, const Pose3* body_P_sensor,
                           const Pose3& pose1, const Vec6& vel1, const Pose3& pose2, const Vec6& vel2, const Vec3& point,
                           Mat<1, 6>* H1, Mat<1, 6>* H2, Mat<1, 6>* H3, Mat<1, 6>* H4, Mat<1, 3>* H5) {
  const bool wantH = H1 || H2 || H3 || H4;
  Mat6 Hint1, Hint2, Hint3, Hint4;
  const Pose3 pose = wantH ? gp.interpolatePose(pose1, vel1, pose2, vel2, &Hint1, &Hint2, &Hint3, &Hint4)
                           : gp.interpolatePose(pose1, vel1, pose2, vel2, nullptr, nullptr, nullptr, nullptr);
  Mat<1, 6> Hpose;
  double hx;
  if (body_P_sensor) {
    const Pose3 sensor = pose.compose(*body_P_sensor);
    hx = pose3_range(sensor, point, &Hpose, H5);
    if (wantH) { const Mat6 H0 = pose3_Hcompose1(*body_P_sensor); Hpose = Hpose * H0; }
  } else {
    hx = pose3_range(pose, point, &Hpose, H5);
  }
  if (wantH) {  // updatePoseJacobians, gp/GaussianProcessInterpolatorPose3.h:108-116
    if (H1) *H1 = Hpose * Hint1; if (H2) *H2 = Hpose * Hint2; if (H3) *H3 = Hpose * Hint3; if (H4) *H4 = Hpose * Hint4;
  }
  return hx - measured;
}

// slam/GPInterpolatedGPSFactorPose3.h:67-95.  body_P_sensor may be null.
inline Vec3 gpGPSPose3(const InterpolatorPose3& gp, const Vec3& measured, const Pose3* body_P_sensor,
                       const Pose3& pose1, const Vec6& vel1, const Pose3& pose2, const Vec6& vel2,
                       Mat<3, 6>* H1, Mat<3, 6>* H2, Mat<3, 6>* H3, Mat<3, 6>* H4) {
  const bool wantH = H1 || H2 || H3 || H4;
  Mat6 Hint1, Hint2, Hint3, Hint4;
  const Pose3 pose = wantH ? gp.interpolatePose(pose1, vel1, pose2, vel2, &Hint1, &Hint2, &Hint3, &Hint4)
                           : gp.interpolatePose(pose1, vel1, pose2, vel2, nullptr, nullptr, nullptr, nullptr);
  Mat<3, 6> Hpose;
  Vec3 point_err;
  if (body_P_sensor) {
    point_err = pose3_translation(pose.compose(*body_P_sensor), &Hpose) - measured;
    if (wantH) Hpose = Hpose * pose3_Hcompose1(*body_P_sensor);
  } else {
    point_err = pose3_translation(pose, &Hpose) - measured;
  }
  if (wantH) {  // updatePoseJacobians, gp/GaussianProcessInterpolatorPose3.h:108-116
    if (H1) *H1 = Hpose * Hint1; if (H2) *H2 = Hpose * Hint2; if (H3) *H3 = Hpose * Hint3; if (H4) *H4 = Hpose * Hint4;
  }
  return point_err;
}

// slam/GPInterpolatedGPSFactorPose3VW.h:71-106.  Hvw1 = [H2 | H3], Hvw2 = [H5 | H6].
inline Vec3 gpGPSPose3VW(const InterpolatorPose3VW& gp, const Vec3& measured, const Pose3* body_P_sensor,
                         const Pose3& pose1, const Vec3& vel1, const Vec3& omega1, const Pose3& pose2, const Vec3& vel2, const Vec3& omega2,
                         Mat<3, 6>* H1, Mat<3, 6>* Hvw1, Mat<3, 6>* H4, Mat<3, 6>* Hvw2) {
  const bool wantH = H1 || Hvw1 || H4 || Hvw2;
  Mat6 Hint1, Hintvw1, Hint4, Hintvw2;
  const Pose3 pose = wantH ? gp.interpolatePose(pose1, vel1, omega1, pose2, vel2, omega2, &Hint1, &Hintvw1, &Hint4, &Hintvw2)
                           : gp.interpolatePose(pose1, vel1, omega1, pose2, vel2, omega2, nullptr, nullptr, nullptr, nullptr);
  Mat<3, 6> Hpose;
  Vec3 point_err;
  if (body_P_sensor) {
    point_err = pose3_translation(pose.compose(*body_P_sensor), &Hpose) - measured;
    if (wantH) Hpose = Hpose * pose3_Hcompose1(*body_P_sensor);
  } else {
    point_err = pose3_translation(pose, &Hpose) - measured;
  }
  if (wantH) {  // updatePoseJacobians, gp/GaussianProcessInterpolatorPose3VW.h:111-124
    if (H1) *H1 = Hpose * Hint1; if (Hvw1) *Hvw1 = Hpose * Hintvw1; if (H4) *H4 = Hpose * Hint4; if (Hvw2) *Hvw2 = Hpose * Hintvw2;
  }
  return point_err;
}

// slam/GPInterpolatedProjectionFactorPose3.h:82-139 with CALIBRATION = Cal3_S2, K = (fx, fy, s, u0, v0).  A landmark behind
// the camera (CheiralityException, :123-138): zero Jacobians and the residual (2 fx, 2 fx); throwCheirality is not restated.
inline Vec2 gpProjectionPose3(const InterpolatorPose3& gp, const Vec2& measured, const double* K, const Pose3* body_P_sensor,
                              const Pose3& pose1, const Vec6& vel1, const Pose3& pose2, const Vec6& vel2, const Vec3& point,
                              Mat<2, 6>* H1, Mat<2, 6>* H2, Mat<2, 6>* H3, Mat<2, 6>* H4, Mat<2, 3>* H5) {
  const bool wantH = H1 || H2 || H3 || H4;
  Mat6 Hint1, Hint2, Hint3, Hint4;
  const Pose3 pose = wantH ? gp.interpolatePose(pose1, vel1, pose2, vel2, &Hint1, &Hint2, &Hint3, &Hint4)
                           : gp.interpolatePose(pose1, vel1, pose2, vel2, nullptr, nullptr, nullptr, nullptr);
  Mat<2, 6> Hpose;
  Vec2 pi;
  const Pose3 cam = body_P_sensor ? pose.compose(*body_P_sensor) : pose;
  if (!pinhole_project(cam, K, point, pi, &Hpose, H5)) {
    if (H1) *H1 = Mat<2, 6>::Zero(); if (H2) *H2 = Mat<2, 6>::Zero(); if (H3) *H3 = Mat<2, 6>::Zero(); if (H4) *H4 = Mat<2, 6>::Zero();
    if (H5) *H5 = Mat<2, 3>::Zero();
    Vec2 e; e[0] = 2.0 * K[0]; e[1] = 2.0 * K[0];
    return e;
  }
  if (wantH) {
    if (body_P_sensor) Hpose = Hpose * pose3_Hcompose1(*body_P_sensor);
    if (H1) *H1 = Hpose * Hint1; if (H2) *H2 = Hpose * Hint2; if (H3) *H3 = Hpose * Hint3; if (H4) *H4 = Hpose * Hint4;
  }
  return pi - measured;
}

// slam/GPInterpolatedRangeFactorPose2.h:64-98
inline double gpRangePose2(const InterpolatorPose2& gp, double measured, const Pose2* body_P_sensor,
                           const Pose2& pose1, const Vec3& vel1, const Pose2& pose2, const Vec3& vel2, const Vec2& point,
                           Mat<1, 3>* H1, Mat<1, 3>* H2, Mat<1, 3>* H3, Mat<1, 3>* H4, Mat<1, 2>* H5) {
  const bool wantH = H1 || H2 || H3 || H4;
  Mat3 Hint1, Hint2, Hint3, Hint4;
  const Pose2 pose = wantH ? gp.interpolatePose(pose1, vel1, pose2, vel2, &Hint1, &Hint2, &Hint3, &Hint4)
                           : gp.interpolatePose(pose1, vel1, pose2, vel2, nullptr, nullptr, nullptr, nullptr);
  Mat<1, 3> Hpose;
  double hx;
  if (body_P_sensor) {
    const Pose2 sensor = pose.compose(*body_P_sensor);
    hx = pose2_range(sensor, point, &Hpose, H5);
    if (wantH) { const Mat3 H0 = body_P_sensor->inverse().Adjoint(); Hpose = Hpose * H0; }
  } else {
    hx = pose2_range(pose, point, &Hpose, H5);
  }
  if (wantH) { if (H1) *H1 = Hpose * Hint1; if (H2) *H2 = Hpose * Hint2; if (H3) *H3 = Hpose * Hint3; if (H4) *H4 = Hpose * Hint4; }
  return hx - measured;
}

// slam/GPInterpolatedRangeFactor2DLinear.h:60-88  (theta component ignored)
inline double gpRange2DLinear(const InterpolatorLinear<3>& gp, double measured,
                              const Vec3& pose1, const Vec3& vel1, const Vec3& pose2, const Vec3& vel2, const Vec2& point,
                              Mat<1, 3>* H1, Mat<1, 3>* H2, Mat<1, 3>* H3, Mat<1, 3>* H4, Mat<1, 2>* H5) {
  const bool wantH = H1 || H2 || H3 || H4;
  Mat3 Hint1, Hint2, Hint3, Hint4;
  const Vec3 pose = wantH ? gp.interpolatePose(pose1, vel1, pose2, vel2, &Hint1, &Hint2, &Hint3, &Hint4)
                          : gp.interpolatePose(pose1, vel1, pose2, vel2, nullptr, nullptr, nullptr, nullptr);
  const Vec2 d = V2(point[0] - pose[0], point[1] - pose[1]);
  Mat<1, 2> H;
  const double r = point2_norm(d, &H);
  if (wantH) {
    Mat<1, 3> Hpose; Hpose[0] = -H[0]; Hpose[1] = -H[1]; Hpose[2] = 0.0;
    if (H1) *H1 = Hpose * Hint1; if (H2) *H2 = Hpose * Hint2; if (H3) *H3 = Hpose * Hint3; if (H4) *H4 = Hpose * Hint4;
  }
  if (H5) *H5 = H;
  return r - measured;
}

// gtsam::Unit3::basis() (Appendix A.5): axis with the smallest |component|, b1 = n x axis / |.|, b2 = n x b1.
inline Mat<3, 2> unit3_basis(const Vec3& n) {
  const double mx = std::fabs(n[0]), my = std::fabs(n[1]), mz = std::fabs(n[2]);
  Vec3 axis = V3(0, 0, 1);
  if (mx <= my && mx <= mz) axis = V3(1, 0, 0);
  else if (my <= mx && my <= mz) axis = V3(0, 1, 0);
  Vec3 b1 = cross(n, axis); b1 = b1 / b1.norm();
  const Vec3 b2 = cross(n, b1);
  Mat<3, 2> B; B.set(0, 0, b1); B.set(0, 1, b2);
  return B;
}
// gtsam::AttitudeFactor::attitudeError(nRb, H): nRef = nRb * bRef ; e = nZ.error(nRef) = B(nZ)^T nRef ;
// H = B(nZ)^T B(nRef) * ( -B(nRef)^T R [bRef]x ).   PARITY UNPINNED: no reference test
// exercises this factor (SURVEY.md §4 "Untested in the reference").
inline Vec2 attitudeError(const Vec3& nZ, const Vec3& bRef, const Mat3& nRb, Mat<2, 3>* H) {
  const Vec3 nRef = nRb * bRef;
  const Mat<3, 2> Bz = unit3_basis(nZ);
  const Vec2 e = Bz.t() * nRef;
  if (H) {
    const Mat<3, 2> Bq = unit3_basis(nRef);
    const Mat<2, 3> D_nRef_R = -(Bq.t() * nRb * skew(bRef));
    const Mat2 D_e_nRef = Bz.t() * Bq;
    *H = D_e_nRef * D_nRef_R;
  }
  return e;
}
// slam/GPInterpolatedAttitudeFactorRot3.h:61-83
inline Vec2 gpAttitudeRot3(const InterpolatorRot3& gp, const Vec3& nZ, const Vec3& bRef,
                           const Mat3& pose1, const Vec3& vel1, const Mat3& pose2, const Vec3& vel2,
                           Mat<2, 3>* H1, Mat<2, 3>* H2, Mat<2, 3>* H3, Mat<2, 3>* H4) {
  if (H1 || H2 || H3 || H4) {
    Mat3 Hint1, Hint2, Hint3, Hint4;
    const Mat3 pose = gp.interpolatePose(pose1, vel1, pose2, vel2, &Hint1, &Hint2, &Hint3, &Hint4);
    Mat<2, 3> Hrot;
    const Vec2 err = attitudeError(nZ, bRef, pose, &Hrot);
    if (H1) *H1 = Hrot * Hint1; if (H2) *H2 = Hrot * Hint2; if (H3) *H3 = Hrot * Hint3; if (H4) *H4 = Hrot * Hint4;
    return err;
  }
  const Mat3 pose = gp.interpolatePose(pose1, vel1, pose2, vel2, nullptr, nullptr, nullptr, nullptr);
  return attitudeError(nZ, bRef, pose, nullptr);
}

// ---------------------------------------------------------------- plain 2-way factors on Vector3 "linear Pose2" states
// slam/RangeFactor2DLinear.h:43-56
inline double range2DLinear(double measured, const Vec3& pose, const Vec2& point, Mat<1, 3>* H1, Mat<1, 2>* H2) {
  const Vec2 d = V2(point[0] - pose[0], point[1] - pose[1]);
  Mat<1, 2> H;
  const double r = point2_norm(d, &H);
  if (H1) { (*H1)[0] = -H[0]; (*H1)[1] = -H[1]; (*H1)[2] = 0.0; }
  if (H2) *H2 = H;
  return r - measured;
}
// slam/RangeBearingFactor2DLinear.h:47-84.  bearing given as angle (Rot2::fromAngle).
inline Vec2 rangeBearing2DLinear(double range, double bearing, const Vec3& pose, const Vec2& point, Mat<2, 3>* H1, Mat2* H2) {
  const double c = std::cos(pose[2]), s = std::sin(pose[2]);
  const double dx = point[0] - pose[0], dy = point[1] - pose[1];
  const double rx = c * dx + s * dy, ry = -s * dx + c * dy;  // Pose2::transform_to
  // Rot2::atan2(y,x) normalises (x,y); bearing_.between(expect) = bearing^-1 * expect ; Rot2::Logmap = theta()
  const double n = std::sqrt(rx * rx + ry * ry);
  const double ec = rx / n, es = ry / n;
  const double bc = std::cos(bearing), bs = std::sin(bearing);
  const double rc = bc * ec + bs * es, rs = bc * es - bs * ec;
  Mat<1, 2> Hnorm;
  const double expect_d = point2_norm(V2(dx, dy), &Hnorm);
  if (H1 || H2) {
    Mat<1, 2> tmp = Mat<1, 2>::Zero();
    if (expect_d > 1e-5) { const double d2 = expect_d * expect_d; tmp[0] = -ry / d2; tmp[1] = rx / d2; }
    Mat2 Rt; Rt(0, 0) = c; Rt(0, 1) = s; Rt(1, 0) = -s; Rt(1, 1) = c;  // pose2.r().transpose()
    if (H1) {
      Mat<2, 3> M; M.set(0, 0, -Rt); M(0, 2) = ry; M(1, 2) = -rx;
      const Mat<1, 3> H11 = tmp * M;
      for (int k = 0; k < 3; k++) (*H1)(0, k) = H11[k];
      (*H1)(1, 0) = -Hnorm[0]; (*H1)(1, 1) = -Hnorm[1]; (*H1)(1, 2) = 0.0;
    }
    if (H2) {
      const Mat<1, 2> H12 = tmp * Rt;
      (*H2)(0, 0) = H12[0]; (*H2)(0, 1) = H12[1]; (*H2)(1, 0) = Hnorm[0]; (*H2)(1, 1) = Hnorm[1];
    }
  }
  return V2(std::atan2(rs, rc), expect_d - range);
}
// slam/OdometryFactor2DLinear.h:50-75
inline Vec3 odometry2DLinear(const Vec3& measured, const Vec3& pose1, const Vec3& pose2, Mat3* H1, Mat3* H2) {
  const Vec3 vd = pose2 - pose1;
  const double c = std::cos(pose1[2]), s = std::sin(pose1[2]);
  const double qx = c * vd[0] + s * vd[1], qy = -s * vd[0] + c * vd[1];  // Rot2::unrotate
  if (H1 || H2) {
    // Rot2::unrotate: Hrot = (q.y, -q.x)^T, Hp = R^T
    Mat2 Hp; Hp(0, 0) = c; Hp(0, 1) = s; Hp(1, 0) = -s; Hp(1, 1) = c;
    if (H1) { *H1 = Mat3::Zero(); H1->set(0, 0, -Hp); (*H1)(0, 2) = qy; (*H1)(1, 2) = -qx; (*H1)(2, 2) = -1; }
    if (H2) { *H2 = Mat3::Zero(); H2->set(0, 0, Hp); (*H2)(2, 2) = 1; }
  }
  return V3(qx - measured[0], qy - measured[1], vd[2] - measured[2]);
}

}  // namespace gpo
