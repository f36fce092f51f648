// gtsam_adapter.h — SURVEY.md §8f rank 1: run a REAL gtsam::NonlinearFactorGraph built from gpslam's own factor classes on the
// B200 engine.  Header-only; compiled only where GTSAM (>= 4.0) and gpslam are installed:
//
//     #include <gpslam_b200/gtsam_adapter.h>
//     gtsam::Values result = gpslam_b200::optimizeOnB200(graph, init_values, /*use_lm=*/true);   // == LevenbergMarquardtOptimizer(graph, init).optimize()
//
// STATUS — read before relying on it.  GTSAM, Eigen and Boost do not exist in the container this repository is built and tested in
// (SURVEY.md §8c), so this file has never been compiled against the real headers.  What IS checked: it compiles, warning-free,
// against tests/cpp/gtsam_stub/ (declarations of the few GTSAM / gpslam entry points used below, written from the GTSAM 4.0 API the
// reference's sources call: tests/test_cpp_facade.py::test_gtsam_adapter_compiles_against_api_stubs).  Everything below the
// lowering is the C ABI of include/gpb.h, which is what the GPU parity tests exercise.
//
// Two of gpslam's private members have no accessor: delta_t_ of the GP priors and (delta_t_, tau_) inside the interpolated
// factors' GPbase_.  The prior's delta_t and Qc are recovered from its public noise model (Gaussian::Covariance(calcQ(Qc, dt)),
// gp/GaussianProcessPriorPose3.h:46: Sigma = [[dt^3/3 Qc, dt^2/2 Qc],[dt^2/2 Qc, dt Qc]]).  The interpolated factors are read
// through their own boost::serialization (slam/GPInterpolatedRangeFactorPose3.h:129-135, gp/GaussianProcessInterpolatorPose3.h:
// 152-159): gtsam::serializeXML(factor) carries <delta_t_> and <tau_>.  A maintainer who adds `double delta_t() const` /
// `double tau() const` to those classes can define GPSLAM_HAS_DT_ACCESSORS and skip the XML round trip.
//
// Supported: GaussianProcessPrior{Pose3,Pose2,Rot3}, GPInterpolatedRangeFactorPose{3,2} (without body_P_sensor),
// PriorFactor<Pose3|Pose2|Rot3|Vector6|Vector3|Point3|Point2>, BetweenFactor<Pose3|Pose2|Rot3>; keys Symbol('x'|'v'|'l', i) with
// consecutive state indices (how every call site of the reference names them, matlab/PlazaPose2.m:183-202).  Any other factor:
// std::runtime_error naming its type (the caller can then fall back to GTSAM).
#pragma once
#if defined(__has_include)
#if __has_include(<gtsam/nonlinear/NonlinearFactorGraph.h>) && __has_include(<gpslam/gp/GaussianProcessPriorPose3.h>)
#define GPSLAM_B200_HAVE_GTSAM 1
#endif
#endif

#ifdef GPSLAM_B200_HAVE_GTSAM
#include <gtsam/base/serialization.h>
#include <gtsam/geometry/Pose2.h>
#include <gtsam/geometry/Pose3.h>
#include <gtsam/inference/Symbol.h>
#include <gtsam/linear/NoiseModel.h>
#include <gtsam/nonlinear/NonlinearFactorGraph.h>
#include <gtsam/nonlinear/Values.h>
#include <gtsam/slam/BetweenFactor.h>
#include <gtsam/slam/PriorFactor.h>

#include <gpslam/gp/GaussianProcessPriorPose2.h>
#include <gpslam/gp/GaussianProcessPriorPose3.h>
#include <gpslam/gp/GaussianProcessPriorRot3.h>
#include <gpslam/slam/GPInterpolatedRangeFactorPose2.h>
#include <gpslam/slam/GPInterpolatedRangeFactorPose3.h>

#include <cstdlib>
#include <algorithm>
#include <cmath>
#include <map>
#include <stdexcept>
#include <string>
#include <typeinfo>
#include <vector>

#include "../gpb.h"

namespace gpslam_b200 {
namespace adapter {

inline void check(int rc) { if (rc < 0) throw std::runtime_error(std::string("gpslam_b200: ") + gpb_last_error()); }

// wire layouts of include/gpb.h
inline void wire(const gtsam::Pose3& T, double* p) {
  const gtsam::Matrix3 R = T.rotation().matrix();
  for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) p[r + 3 * c] = R(r, c);
  p[9] = T.x(); p[10] = T.y(); p[11] = T.z();
}
inline void wire(const gtsam::Rot3& Rt, double* p) { const gtsam::Matrix3 R = Rt.matrix(); for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) p[r + 3 * c] = R(r, c); }
inline void wire(const gtsam::Pose2& T, double* p) { p[0] = T.x(); p[1] = T.y(); p[2] = T.theta(); }
template <class V> inline void wireVec(const V& v, int n, double* p) { for (int k = 0; k < n; k++) p[k] = v(k); }

// upper-triangular sqrt information R (m x m, column-major) of a factor's Gaussian noise model
inline std::vector<double> sqrtInfo(const gtsam::SharedNoiseModel& model, int m) {
  auto g = boost::dynamic_pointer_cast<gtsam::noiseModel::Gaussian>(model);
  if (!g) throw std::runtime_error("gpslam_b200: noise model is not Gaussian (robust / constrained models are not lowered)");
  const gtsam::Matrix R = g->R();
  if (R.rows() != m || R.cols() != m) throw std::runtime_error("gpslam_b200: noise model dimension does not match the factor");
  std::vector<double> out(static_cast<size_t>(m) * m);
  for (int c = 0; c < m; c++) for (int r = 0; r < m; r++) out[r + static_cast<size_t>(c) * m] = R(r, c);
  return out;
}

// (delta_t, Qc) of a GP prior from its noise model Sigma = calcQ(Qc, dt) (gp/GPutils.h:24-30): dt = 2 Sigma_pv / Sigma_vv, Qc = Sigma_vv / dt
inline double priorDeltaT(const gtsam::SharedNoiseModel& model, int D, std::vector<double>& Qc) {
  auto g = boost::dynamic_pointer_cast<gtsam::noiseModel::Gaussian>(model);
  if (!g) throw std::runtime_error("gpslam_b200: GP prior without a Gaussian noise model");
  const gtsam::Matrix S = g->covariance();
  if (S.rows() != 2 * D) throw std::runtime_error("gpslam_b200: GP prior noise model has the wrong dimension");
  const double dt = 2.0 * S(0, D) / S(D, D);
  Qc.assign(static_cast<size_t>(D) * D, 0.0);
  for (int c = 0; c < D; c++) for (int r = 0; r < D; r++) Qc[r + static_cast<size_t>(c) * D] = S(D + r, D + c) / dt;
  // covariance() is (R^T R)^-1 recomputed by GTSAM: a diagonal Qc comes back with off-diagonals of rounding size.  Snap them to
  // zero (relative 1e-12) so the engine keeps its diagonal-Qc linearise kernel, and symmetrise what remains.
  for (int c = 0; c < D; c++) for (int r = 0; r < c; r++) {
    double& a = Qc[r + static_cast<size_t>(c) * D]; double& b = Qc[c + static_cast<size_t>(r) * D];
    const double scale = std::sqrt(std::fabs(Qc[r + static_cast<size_t>(r) * D] * Qc[c + static_cast<size_t>(c) * D]));
    const double m = 0.5 * (a + b);
    a = b = (std::fabs(m) <= 1e-12 * scale) ? 0.0 : m;
  }
  return dt;
}

// <tag>value</tag> of the factor's XML archive
inline double xmlNumber(const std::string& xml, const std::string& tag) {
  const std::string open = "<" + tag + ">";
  const size_t a = xml.find(open);
  if (a == std::string::npos) throw std::runtime_error("gpslam_b200: <" + tag + "> not found in the factor's serialisation");
  return std::strtod(xml.c_str() + a + open.size(), nullptr);
}

struct Lowering {
  gpb_graph* g = nullptr;
  int group = GPB_POSE3, D = 6, PS = 12, DL = 3;
  std::map<gtsam::Key, int> state, land;       // 'x' / 'v' key -> chain index, 'l' key -> landmark index
  std::vector<std::vector<double>> qcs;        // Qc models registered so far
  int qcId(const std::vector<double>& Qc) {
    // models recovered from different factors' covariances agree to rounding only: de-duplicate with a relative tolerance, or every
    // distinct delta_t would register its own Qc
    for (size_t k = 0; k < qcs.size(); k++) {
      bool same = qcs[k].size() == Qc.size();
      double scale = 0.0;
      for (double v : qcs[k]) scale = std::max(scale, std::fabs(v));
      for (size_t t = 0; same && t < Qc.size(); t++) same = std::fabs(qcs[k][t] - Qc[t]) <= 1e-9 * scale;
      if (same) return static_cast<int>(k);
    }
    const int id = gpb_add_qc_model(g, Qc.data());
    check(id);
    qcs.push_back(Qc);
    return id;
  }
  int stateOf(gtsam::Key k) const { auto it = state.find(k); if (it == state.end()) throw std::runtime_error("gpslam_b200: factor on a key that is not in Values"); return it->second; }
  int landOf(gtsam::Key k) const { auto it = land.find(k); if (it == land.end()) throw std::runtime_error("gpslam_b200: factor on a landmark key that is not in Values"); return it->second; }

  template <class PRIOR> bool gpPrior(const gtsam::NonlinearFactor::shared_ptr& f) {
    auto p = boost::dynamic_pointer_cast<PRIOR>(f);
    if (!p) return false;
    const int i = stateOf(p->keys()[0]);
    if (stateOf(p->keys()[2]) != i + 1) throw std::runtime_error("gpslam_b200: GP prior must join consecutive states");
    if (stateOf(p->keys()[1]) != i || stateOf(p->keys()[3]) != i + 1) throw std::runtime_error("gpslam_b200: GP prior velocity keys do not belong to its pose keys");
    std::vector<double> Qc;
    const double dt = priorDeltaT(p->noiseModel(), D, Qc);
    check(gpb_add_gp_prior(g, 1, &i, &dt, qcId(Qc)));
    return true;
  }
  template <class RANGE> bool interpRange(const gtsam::NonlinearFactor::shared_ptr& f) {
    auto p = boost::dynamic_pointer_cast<RANGE>(f);
    if (!p) return false;
    const int i = stateOf(p->keys()[0]), l = landOf(p->keys()[4]);
#ifdef GPSLAM_HAS_DT_ACCESSORS
    const double dt = p->delta_t(), tau = p->tau();
#else
    const std::string xml = gtsam::serializeXML(*p);
    if (xml.find("<body_P_sensor_") != std::string::npos && xml.find("<initialized>1</initialized>") != std::string::npos)
      throw std::runtime_error("gpslam_b200: interpolated range factor with body_P_sensor needs the accessor build (GPSLAM_HAS_DT_ACCESSORS)");
    const double dt = xmlNumber(xml, "delta_t_"), tau = xmlNumber(xml, "tau_");
#endif
    const double z = p->measured(), sigma = 1.0 / sqrtInfo(p->noiseModel(), 1)[0];
    check(gpb_add_interp_range(g, 1, &i, &l, &z, &sigma, &dt, &tau, 0, nullptr));  // Lambda / Psi do not depend on Qc: the model id is not used
    return true;
  }
  template <class T, class WIRE> bool prior(const gtsam::NonlinearFactor::shared_ptr& f, int m, char kind, WIRE&& put) {
    auto p = boost::dynamic_pointer_cast<gtsam::PriorFactor<T>>(f);
    if (!p) return false;
    double v[12];
    put(p->prior(), v);
    const std::vector<double> R = sqrtInfo(p->noiseModel(), m);
    const gtsam::Key k = p->keys()[0];
    if (kind == 'l') check(gpb_add_prior_landmark(g, landOf(k), v, R.data()));
    else if (gtsam::Symbol(k).chr() == 'v') check(gpb_add_prior_vel(g, stateOf(k), v, R.data()));
    else check(gpb_add_prior_pose(g, stateOf(k), v, R.data()));
    return true;
  }
  template <class T> bool between(const gtsam::NonlinearFactor::shared_ptr& f) {
    auto p = boost::dynamic_pointer_cast<gtsam::BetweenFactor<T>>(f);
    if (!p) return false;
    double v[12];
    wire(p->measured(), v);
    check(gpb_add_between(g, stateOf(p->keys()[0]), stateOf(p->keys()[1]), v, sqrtInfo(p->noiseModel(), D).data()));
    return true;
  }
};

}  // namespace adapter

/// LevenbergMarquardtOptimizer(graph, initial).optimize() (use_lm) or GaussNewtonOptimizer(graph, initial).optimize() on one B200.
inline gtsam::Values optimizeOnB200(const gtsam::NonlinearFactorGraph& graph, const gtsam::Values& initial, bool use_lm = true, int device = 0,
                                    gpb_stats* stats_out = nullptr) {
  using namespace adapter;
  Lowering L;
  // ---- variables: 'x' i (poses, consecutive), 'v' i (velocities), 'l' j (landmarks)
  std::map<std::uint64_t, gtsam::Key> xs, ls;
  for (const gtsam::Key k : initial.keys()) {
    const gtsam::Symbol s(k);
    if (s.chr() == 'x') xs[s.index()] = k; else if (s.chr() == 'l') ls[s.index()] = k;
    else if (s.chr() != 'v') throw std::runtime_error("gpslam_b200: only 'x', 'v', 'l' keys are lowered");
  }
  if (xs.size() < 2) throw std::runtime_error("gpslam_b200: need at least two states");
  const gtsam::Key x0 = xs.begin()->second;
  if (initial.exists<gtsam::Pose3>(x0)) { L.group = GPB_POSE3; L.D = 6; L.PS = 12; L.DL = 3; }
  else if (initial.exists<gtsam::Pose2>(x0)) { L.group = GPB_POSE2; L.D = 3; L.PS = 3; L.DL = 2; }
  else if (initial.exists<gtsam::Rot3>(x0)) { L.group = GPB_ROT3; L.D = 3; L.PS = 9; L.DL = 0; }
  else throw std::runtime_error("gpslam_b200: states must be Pose3, Pose2 or Rot3 (Vector3 '2DLinear' graphs: use the facade)");
  std::vector<gtsam::Key> xkeys, vkeys, lkeys;
  std::uint64_t prev = 0;
  for (const auto& kv : xs) {
    if (!xkeys.empty() && kv.first != prev + 1) throw std::runtime_error("gpslam_b200: state indices must be consecutive");
    prev = kv.first;
    const int i = static_cast<int>(xkeys.size());
    xkeys.push_back(kv.second); vkeys.push_back(gtsam::Symbol('v', kv.first));
    L.state[xkeys.back()] = i; L.state[vkeys.back()] = i;
  }
  for (const auto& kv : ls) { L.land[kv.second] = static_cast<int>(lkeys.size()); lkeys.push_back(kv.second); }
  const int N = static_cast<int>(xkeys.size()), NL = static_cast<int>(lkeys.size());
  L.g = gpb_graph_create(L.group, 3, N, NL);
  if (!L.g) throw std::runtime_error(std::string("gpslam_b200: ") + gpb_last_error());
  struct Guard { gpb_graph* g; ~Guard() { gpb_graph_destroy(g); } } guard{L.g};
  // ---- factors
  for (const auto& f : graph) {
    if (!f) continue;
    bool ok = false;
    if (L.group == GPB_POSE3) {
      ok = L.gpPrior<gpslam::GaussianProcessPriorPose3>(f) || L.interpRange<gpslam::GPInterpolatedRangeFactorPose3>(f) ||
           L.prior<gtsam::Pose3>(f, 6, 'x', [](const gtsam::Pose3& T, double* p) { wire(T, p); }) ||
           L.prior<gtsam::Vector6>(f, 6, 'v', [](const gtsam::Vector6& v, double* p) { wireVec(v, 6, p); }) ||
           L.prior<gtsam::Point3>(f, 3, 'l', [](const gtsam::Point3& q, double* p) { p[0] = q.x(); p[1] = q.y(); p[2] = q.z(); }) || L.between<gtsam::Pose3>(f);
    } else if (L.group == GPB_POSE2) {
      ok = L.gpPrior<gpslam::GaussianProcessPriorPose2>(f) || L.interpRange<gpslam::GPInterpolatedRangeFactorPose2>(f) ||
           L.prior<gtsam::Pose2>(f, 3, 'x', [](const gtsam::Pose2& T, double* p) { wire(T, p); }) ||
           L.prior<gtsam::Vector3>(f, 3, 'v', [](const gtsam::Vector3& v, double* p) { wireVec(v, 3, p); }) ||
           L.prior<gtsam::Point2>(f, 2, 'l', [](const gtsam::Point2& q, double* p) { p[0] = q.x(); p[1] = q.y(); }) || L.between<gtsam::Pose2>(f);
    } else {
      ok = L.gpPrior<gpslam::GaussianProcessPriorRot3>(f) || L.prior<gtsam::Rot3>(f, 3, 'x', [](const gtsam::Rot3& R, double* p) { wire(R, p); }) ||
           L.prior<gtsam::Vector3>(f, 3, 'v', [](const gtsam::Vector3& v, double* p) { wireVec(v, 3, p); }) || L.between<gtsam::Rot3>(f);
    }
    if (!ok) throw std::runtime_error(std::string("gpslam_b200: factor type not lowered to the B200 engine: ") + typeid(*f).name());
  }
  // ---- values in, optimise, values out
  std::vector<double> P(static_cast<size_t>(N) * L.PS), V(static_cast<size_t>(N) * L.D), Lm(static_cast<size_t>(NL ? NL * L.DL : 1));
  for (int i = 0; i < N; i++) {
    if (L.group == GPB_POSE3) { wire(initial.at<gtsam::Pose3>(xkeys[i]), &P[static_cast<size_t>(i) * 12]); wireVec(initial.at<gtsam::Vector6>(vkeys[i]), 6, &V[static_cast<size_t>(i) * 6]); }
    else if (L.group == GPB_POSE2) { wire(initial.at<gtsam::Pose2>(xkeys[i]), &P[static_cast<size_t>(i) * 3]); wireVec(initial.at<gtsam::Vector3>(vkeys[i]), 3, &V[static_cast<size_t>(i) * 3]); }
    else { wire(initial.at<gtsam::Rot3>(xkeys[i]), &P[static_cast<size_t>(i) * 9]); wireVec(initial.at<gtsam::Vector3>(vkeys[i]), 3, &V[static_cast<size_t>(i) * 3]); }
  }
  for (int l = 0; l < NL; l++) {
    if (L.DL == 3) { const gtsam::Point3 q = initial.at<gtsam::Point3>(lkeys[l]); Lm[3 * l] = q.x(); Lm[3 * l + 1] = q.y(); Lm[3 * l + 2] = q.z(); }
    else if (L.DL == 2) { const gtsam::Point2 q = initial.at<gtsam::Point2>(lkeys[l]); Lm[2 * l] = q.x(); Lm[2 * l + 1] = q.y(); }
  }
  check(gpb_set_values(L.g, P.data(), V.data(), NL ? Lm.data() : nullptr));
  check(gpb_graph_finalize(L.g, device));
  gpb_params prm; gpb_default_params(&prm, use_lm ? 1 : 0);
  gpb_stats st;
  check(gpb_optimize(L.g, &prm, 0, &st));
  if (stats_out) *stats_out = st;
  check(gpb_get_values(L.g, P.data(), V.data(), NL ? Lm.data() : nullptr));
  gtsam::Values result;
  for (int i = 0; i < N; i++) {
    const double* p = &P[static_cast<size_t>(i) * L.PS];
    const double* v = &V[static_cast<size_t>(i) * L.D];
    if (L.group == GPB_POSE3) {
      gtsam::Matrix3 R; for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) R(r, c) = p[r + 3 * c];
      result.insert(xkeys[i], gtsam::Pose3(gtsam::Rot3(R), gtsam::Point3(p[9], p[10], p[11])));
      gtsam::Vector6 vv; for (int k = 0; k < 6; k++) vv(k) = v[k];
      result.insert(vkeys[i], vv);
    } else if (L.group == GPB_POSE2) {
      result.insert(xkeys[i], gtsam::Pose2(p[0], p[1], p[2]));
      gtsam::Vector3 vv; for (int k = 0; k < 3; k++) vv(k) = v[k];
      result.insert(vkeys[i], vv);
    } else {
      gtsam::Matrix3 R; for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) R(r, c) = p[r + 3 * c];
      result.insert(xkeys[i], gtsam::Rot3(R));
      gtsam::Vector3 vv; for (int k = 0; k < 3; k++) vv(k) = v[k];
      result.insert(vkeys[i], vv);
    }
  }
  for (int l = 0; l < NL; l++) {
    if (L.DL == 3) result.insert(lkeys[l], gtsam::Point3(Lm[3 * l], Lm[3 * l + 1], Lm[3 * l + 2]));
    else if (L.DL == 2) result.insert(lkeys[l], gtsam::Point2(Lm[2 * l], Lm[2 * l + 1]));
  }
  return result;
}

}  // namespace gpslam_b200
#endif  // GPSLAM_B200_HAVE_GTSAM
