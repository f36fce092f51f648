// gpslam.h — header-only C++ host facade over the C ABI (include/gpb.h): the reference's factor classes, with the
// reference's names, constructor argument order and evaluateError contract, plus the few GTSAM container / optimiser
// types its call sites use (matlab/PlazaPose2.m:51-230, the "Optimization" unit tests).  Graphs written against
// gpslam + GTSAM compile against this header by switching the include and the namespace alias.
//
// GTSAM itself is not available where this was built, so the handful of value types the interface needs live in
// namespace gpslam_b200::gtsam (Key/Symbol, Pose3, Rot3, Pose2, Point2/3, Vector, Matrix, noiseModel).  Differences from
// the reference signatures, all forced by the absence of Boost/GTSAM: optional Jacobians are `Matrix*` (nullptr = not
// requested) instead of boost::optional<Matrix&>; shared_ptr is std::shared_ptr.
//
// No arithmetic lives here: evaluateError forwards to gpb_eval_factor (the same device code as the batched path) and the
// optimisers lower the graph to a gpb_graph and call gpb_optimize.  Errors become std::runtime_error carrying gpb_last_error().
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <fstream>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../gpb.h"
#include "archive.h"

namespace gpslam_b200 {
namespace gtsam {

using Key = std::uint64_t;
inline Key Symbol(char c, std::uint64_t j) { return (static_cast<Key>(static_cast<unsigned char>(c)) << 56) | j; }
inline char symbolChr(Key k) { return static_cast<char>(k >> 56); }
inline std::uint64_t symbolIndex(Key k) { return k & ((Key(1) << 56) - 1); }
using KeyFormatter = std::function<std::string(Key)>;
inline std::string DefaultKeyFormatter(Key k) {  // gtsam's default: Symbol keys print as <chr><index>, plain integers as the number
  const char c = symbolChr(k);
  return (c >= 33 && c < 127) ? std::string(1, c) + std::to_string(symbolIndex(k)) : std::to_string(k);
}

using Vector = std::vector<double>;
struct Matrix {  // dynamic, column-major (Eigen's default)
  int rows = 0, cols = 0;
  std::vector<double> a;
  Matrix() {}
  Matrix(int r, int c) : rows(r), cols(c), a(static_cast<size_t>(r) * c, 0.0) {}
  double& operator()(int r, int c) { return a[r + static_cast<size_t>(c) * rows]; }
  double operator()(int r, int c) const { return a[r + static_cast<size_t>(c) * rows]; }
  static Matrix Identity(int n, int m) { Matrix I(n, m); for (int k = 0; k < (n < m ? n : m); k++) I(k, k) = 1.0; return I; }
  template <class ARCHIVE> void serialize(ARCHIVE& ar, const unsigned int /*version*/) {
    ar & GPSLAM_B200_NVP(rows); ar & GPSLAM_B200_NVP(cols); ar & make_nvp("data", a);
    if (rows < 0 || cols < 0 || a.size() != static_cast<size_t>(rows) * cols) throw std::runtime_error("gpslam_b200 archive: matrix data does not match its shape");
  }
};
inline Matrix operator*(double s, Matrix m) { for (double& v : m.a) v *= s; return m; }

template <int N> using VectorN = std::array<double, N>;
using Vector3 = VectorN<3>;
using Vector6 = VectorN<6>;
struct Point2 {
  double x = 0, y = 0;
  Point2() {}
  Point2(double x_, double y_) : x(x_), y(y_) {}
  template <class ARCHIVE> void serialize(ARCHIVE& ar, const unsigned int) { ar & GPSLAM_B200_NVP(x); ar & GPSLAM_B200_NVP(y); }
};
struct Point3 {
  double x = 0, y = 0, z = 0;
  Point3() {}
  Point3(double x_, double y_, double z_) : x(x_), y(y_), z(z_) {}
  template <class ARCHIVE> void serialize(ARCHIVE& ar, const unsigned int) { ar & GPSLAM_B200_NVP(x); ar & GPSLAM_B200_NVP(y); ar & GPSLAM_B200_NVP(z); }
};
struct Unit3 {
  double x = 0, y = 0, z = 1;
  Unit3() {}
  Unit3(double x_, double y_, double z_) { const double n = std::sqrt(x_ * x_ + y_ * y_ + z_ * z_); x = x_ / n; y = y_ / n; z = z_ / n; }
  template <class ARCHIVE> void serialize(ARCHIVE& ar, const unsigned int) { ar & GPSLAM_B200_NVP(x); ar & GPSLAM_B200_NVP(y); ar & GPSLAM_B200_NVP(z); }
};

struct Rot3 {
  double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};  // column-major
  static Rot3 Ypr(double y, double p, double r) {  // Rz(y) Ry(p) Rx(r)  (gp/tests/testPose3Utils.cpp:126-133)
    const double cy = std::cos(y), sy = std::sin(y), cp = std::cos(p), sp = std::sin(p), cr = std::cos(r), sr = std::sin(r);
    Rot3 o;
    const double m[3][3] = {{cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr}, {sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr}, {-sp, cp * sr, cp * cr}};
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) o.R[i + 3 * j] = m[i][j];
    return o;
  }
  template <class ARCHIVE> void serialize(ARCHIVE& ar, const unsigned int) {
    ar.array("R", R, 9, [&](size_t n) { if (n != 9) throw std::runtime_error("gpslam_b200 archive: a rotation has nine entries"); return R; });
  }
};
struct Pose3 {
  Rot3 r; Point3 t;
  Pose3() {}
  Pose3(const Rot3& r_, const Point3& t_) : r(r_), t(t_) {}
  void wire(double* p) const { for (int k = 0; k < 9; k++) p[k] = r.R[k]; p[9] = t.x; p[10] = t.y; p[11] = t.z; }
  static Pose3 fromWire(const double* p) { Pose3 T; for (int k = 0; k < 9; k++) T.r.R[k] = p[k]; T.t = Point3(p[9], p[10], p[11]); return T; }
  template <class ARCHIVE> void serialize(ARCHIVE& ar, const unsigned int) { ar & make_nvp("R_", r); ar & make_nvp("t_", t); }
};
struct Pose2 {
  double x = 0, y = 0, theta = 0;
  Pose2() {}
  Pose2(double x_, double y_, double th) : x(x_), y(y_), theta(th) {}
  template <class ARCHIVE> void serialize(ARCHIVE& ar, const unsigned int) { ar & GPSLAM_B200_NVP(x); ar & GPSLAM_B200_NVP(y); ar & GPSLAM_B200_NVP(theta); }
};

namespace noiseModel {
// Every model the gpslam call sites use reduces to a Gaussian with upper-triangular square-root information R.
struct Gaussian {
  int dim = 0;
  Matrix cov;  // covariance (kept for getQc: gp/GPutils.cpp:16-20)
  Matrix R;    // upper-triangular sqrt information
  using shared_ptr = std::shared_ptr<Gaussian>;
  static shared_ptr Covariance(const Matrix& c) {
    auto m = std::make_shared<Gaussian>(); m->dim = c.rows; m->cov = c;
    // R = chol_upper(cov^-1); only needed for measurement models, where the call sites use diagonal covariances
    m->R = Matrix(c.rows, c.cols);
    bool diag = true;
    for (int i = 0; i < c.rows; i++) for (int j = 0; j < c.cols; j++) if (i != j && c(i, j) != 0.0) diag = false;
    if (diag) for (int i = 0; i < c.rows; i++) m->R(i, i) = 1.0 / std::sqrt(c(i, i));
    else m->R = Matrix();  // dense covariance: valid as a Qc model (the engine factors it itself), not as a measurement model
    return m;
  }
  template <class ARCHIVE> void serialize(ARCHIVE& ar, const unsigned int) { ar & GPSLAM_B200_NVP(dim); ar & GPSLAM_B200_NVP(cov); ar & make_nvp("sqrt_information", R); }
};
struct Isotropic { static Gaussian::shared_ptr Sigma(int dim, double sigma) { return Gaussian::Covariance((sigma * sigma) * Matrix::Identity(dim, dim)); } };
struct Unit { static Gaussian::shared_ptr Create(int dim) { return Isotropic::Sigma(dim, 1.0); } };
struct Diagonal {
  static Gaussian::shared_ptr Sigmas(const Vector& s) { Matrix c(static_cast<int>(s.size()), static_cast<int>(s.size())); for (size_t k = 0; k < s.size(); k++) c(static_cast<int>(k), static_cast<int>(k)) = s[k] * s[k]; return Gaussian::Covariance(c); }
};
}  // namespace noiseModel
using SharedNoiseModel = noiseModel::Gaussian::shared_ptr;

/// gtsam::traits<T> for the Testable types of this header (gp/GaussianProcessPriorPose3.h:134-137): Equals / Print forward to the members
template <class T> struct traits {
  static bool Equals(const T& a, const T& b, double tol = 1e-8) { return a.equals(b, tol); }
  static void Print(const T& a, const std::string& s = "") { a.print(s); }
};

}  // namespace gtsam

namespace detail {
inline void check(int rc) { if (rc < 0) throw std::runtime_error(std::string("gpslam_b200: ") + gpb_last_error()); }
inline const gtsam::Matrix& sqrtInfo(const gtsam::SharedNoiseModel& m) {
  if (!m || m->R.rows == 0) throw std::runtime_error("gpslam_b200: measurement noise model must be Gaussian with a diagonal covariance");
  return m->R;
}
inline void wire(const gtsam::Pose3& v, double* p) { v.wire(p); }
inline void wire(const gtsam::Rot3& v, double* p) { for (int k = 0; k < 9; k++) p[k] = v.R[k]; }
inline void wire(const gtsam::Pose2& v, double* p) { p[0] = v.x; p[1] = v.y; p[2] = v.theta; }
inline void wire(const gtsam::Vector3& v, double* p) { for (int k = 0; k < 3; k++) p[k] = v[k]; }
inline void wire(const gtsam::Vector6& v, double* p) { for (int k = 0; k < 6; k++) p[k] = v[k]; }
inline void wire(const gtsam::Point3& v, double* p) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }
inline void wire(const gtsam::Point2& v, double* p) { p[0] = v.x; p[1] = v.y; }
template <class T> struct GroupOf;
template <> struct GroupOf<gtsam::Pose3> { static constexpr int group = GPB_POSE3, D = 6, PS = 12, DL = 3; using Vel = gtsam::Vector6; using Land = gtsam::Point3; static const char* name() { return "Pose3"; } };
template <> struct GroupOf<gtsam::Pose2> { static constexpr int group = GPB_POSE2, D = 3, PS = 3, DL = 2; using Vel = gtsam::Vector3; using Land = gtsam::Point2; static const char* name() { return "Pose2"; } };
template <> struct GroupOf<gtsam::Rot3> { static constexpr int group = GPB_ROT3, D = 3, PS = 9, DL = 0; using Vel = gtsam::Vector3; using Land = gtsam::Point2; static const char* name() { return "Rot3"; } };
template <> struct GroupOf<gtsam::Vector3> { static constexpr int group = GPB_LINEAR, D = 3, PS = 3, DL = 2; using Vel = gtsam::Vector3; using Land = gtsam::Point2; static const char* name() { return "Linear<3>"; } };

// evaluateError through the C ABI: returns e, fills the requested Jacobians (in the factor's variable order)
inline gtsam::Vector eval(int group, int kind, const double* x1, const double* v1, const double* x2, const double* v2, const double* land, const double* prm,
                          std::initializer_list<gtsam::Matrix*> H) {
  double e[12], Hbuf[12 * 6 * 5];
  int dims[5];
  bool want = false;
  for (gtsam::Matrix* h : H) want |= (h != nullptr);
  const int m = gpb_eval_factor(group, kind, x1, v1, x2, v2, land, prm, e, want ? Hbuf : nullptr, dims);
  check(m);
  if (want) {
    int o = 0, v = 0;
    for (gtsam::Matrix* h : H) {
      if (h) { *h = gtsam::Matrix(m, dims[v]); for (int k = 0; k < m * dims[v]; k++) h->a[k] = Hbuf[o + k]; }
      o += m * dims[v]; v++;
    }
  }
  return gtsam::Vector(e, e + m);
}
}  // namespace detail

// ================================================================================== factors
class NonlinearFactor {
 public:
  virtual ~NonlinearFactor() {}
  virtual const std::vector<gtsam::Key>& keys() const = 0;
  virtual size_t size() const { return keys().size(); }
  using shared_ptr = std::shared_ptr<NonlinearFactor>;
  /// deep copy (gp/GaussianProcessPriorPose3.h:55-57)
  virtual shared_ptr clone() const = 0;
  /// print contents (gp/GaussianProcessPriorPose3.h:112-115): the factor's description, then its keys
  virtual void print(const std::string& s = "", const gtsam::KeyFormatter& keyFormatter = gtsam::DefaultKeyFormatter) const {
    std::cout << s << describe() << "\n  keys = {";
    for (gtsam::Key k : keys()) std::cout << " " << keyFormatter(k);
    std::cout << " }" << std::endl;
  }
  virtual std::string describe() const { return "NonlinearFactor"; }
  /// same class, same keys, members equal up to tol (gp/GaussianProcessPriorPose3.h:106-109 and its siblings)
  virtual bool equals(const NonlinearFactor& expected, double tol = 1e-9) const = 0;
  /// dimension of the residual (gtsam::NoiseModelFactor::dim(): the noise model's)
  virtual size_t dim() const = 0;
  /// a GP prior ties (pose, velocity[, angular velocity]) of one state to the next: the optimiser reads the chain order from these links
  struct ChainLink { gtsam::Key x1, v1, w1, x2, v2, w2; bool vw; };
  virtual bool chainLink(ChainLink&) const { return false; }
  // lowering hook used by the optimisers: add this factor to g; idx maps a state key to its chain index, lidx a landmark key
  virtual void lower(gpb_graph* g, int qc_of(void*, const gtsam::Matrix&), void* ctx, const std::map<gtsam::Key, int>& sidx,
                     const std::map<gtsam::Key, int>& lidx) const = 0;
  /// archives (archive.h): the type name a graph archive stores in front of the factor, and the factor's own serialize() reached
  /// through the base pointer (what boost's export / void_cast registration does for the reference's classes)
  virtual std::string archiveTag() const = 0;
  virtual void save(OArchive& ar) const = 0;
  virtual void load(IArchive& ar) = 0;
};
#define GPSLAM_B200_FACTOR(CLASS, TEXT, TAG)                                                              \
  NonlinearFactor::shared_ptr clone() const override { return std::make_shared<CLASS>(*this); }           \
  std::string describe() const override { return TEXT; }                                                  \
  bool equals(const NonlinearFactor& expected, double tol = 1e-9) const override {                        \
    const CLASS* e = dynamic_cast<const CLASS*>(&expected);                                               \
    return e != nullptr && keys() == e->keys() && sameMembers(*e, tol);                                   \
  }                                                                                                       \
  std::string archiveTag() const override { return TAG; }                                                 \
  void save(OArchive& ar) const override { archiveIO(ar, "factor", const_cast<CLASS&>(*this)); }          \
  void load(IArchive& ar) override { archiveIO(ar, "factor", *this); }

/// type name -> default-constructed factor, for loading a graph archive; every factor class of this header is registered, user
/// classes join with FactorRegistry::add<F>()
class FactorRegistry {
 public:
  using Maker = std::function<NonlinearFactor::shared_ptr()>;
  static std::map<std::string, Maker>& table() { static std::map<std::string, Maker> t; return t; }
  template <class F> static void add() { table()[F().archiveTag()] = [] { return NonlinearFactor::shared_ptr(std::make_shared<F>()); }; }
  static inline NonlinearFactor::shared_ptr make(const std::string& tag);
};

namespace detail {
inline int stateOf(const std::map<gtsam::Key, int>& m, gtsam::Key k) {
  auto it = m.find(k);
  if (it == m.end()) throw std::runtime_error("gpslam_b200: factor refers to a key that is not in Values");
  return it->second;
}
// member comparison for equals(): numbers, value types through their wire layout, noise models through their covariance
inline bool same(double a, double b, double tol) { return std::fabs(a - b) <= tol; }
inline bool same(const std::vector<double>& a, const std::vector<double>& b, double tol) {
  if (a.size() != b.size()) return false;
  for (size_t k = 0; k < a.size(); k++) if (!same(a[k], b[k], tol)) return false;
  return true;
}
template <class T> bool sameValue(const T& a, const T& b, double tol) {
  double wa[12] = {0}, wb[12] = {0};
  wire(a, wa); wire(b, wb);
  for (int k = 0; k < 12; k++) if (!same(wa[k], wb[k], tol)) return false;
  return true;
}
inline bool sameModel(const gtsam::SharedNoiseModel& a, const gtsam::SharedNoiseModel& b, double tol) {
  if (!a || !b) return !a && !b;
  return a->dim == b->dim && same(a->cov.a, b->cov.a, tol);
}
inline size_t modelDim(const gtsam::SharedNoiseModel& m) { return m ? static_cast<size_t>(m->dim) : 0; }
// archive pieces shared by the factor classes: the NoiseModelFactorN base (keys and, for measurement factors, the noise model),
// the interpolator a GPInterpolated* factor holds (GPbase_), and an optional sensor pose (boost::optional<POSE> in the reference)
template <class AR> void ioBase(AR& ar, std::vector<gtsam::Key>& keys, size_t nkeys, gtsam::SharedNoiseModel* model) {
  ar.begin("Base");
  ar & make_nvp("keys_", keys);
  if (keys.size() != nkeys) throw std::runtime_error("gpslam_b200 archive: wrong number of keys for this factor type");
  if (model) ar & make_nvp("noiseModel_", *model);
  ar.end();
}
template <class AR> void ioGPbase(AR& ar, double& delta_t, double& tau, gtsam::SharedNoiseModel& Qc) {
  ar.begin("GPbase_");
  ar & make_nvp("delta_t_", delta_t); ar & make_nvp("tau_", tau); ar & make_nvp("Qc", Qc);
  ar.end();
}
template <class AR, class P> void ioOptional(AR& ar, const char* name, bool& has, P& value) {
  ar.begin(name);
  ar & make_nvp("initialized", has);
  if (has) ar & make_nvp("value", value);
  ar.end();
}
template <class T> struct TypeName;
template <> struct TypeName<gtsam::Pose3> { static const char* name() { return "Pose3"; } };
template <> struct TypeName<gtsam::Pose2> { static const char* name() { return "Pose2"; } };
template <> struct TypeName<gtsam::Rot3> { static const char* name() { return "Rot3"; } };
template <> struct TypeName<gtsam::Vector3> { static const char* name() { return "Vector3"; } };
template <> struct TypeName<gtsam::Vector6> { static const char* name() { return "Vector6"; } };
template <> struct TypeName<gtsam::Point3> { static const char* name() { return "Point3"; } };
template <> struct TypeName<gtsam::Point2> { static const char* name() { return "Point2"; } };
}  // namespace detail

/// 4-way GP prior factors — gp/GaussianProcessPrior{Pose3,Pose2,Rot3,Linear}.h (constructor: :43-49)
template <class POSE>
class GaussianProcessPriorT : public NonlinearFactor {
  using G = detail::GroupOf<POSE>;
  std::vector<gtsam::Key> keys_;
  double delta_t_ = 0;
  gtsam::SharedNoiseModel Qc_;

 public:
  GaussianProcessPriorT() {}  ///< default constructor, for loading from an archive only (gp/GaussianProcessPriorPose3.h:40)
  GaussianProcessPriorT(gtsam::Key poseKey1, gtsam::Key velKey1, gtsam::Key poseKey2, gtsam::Key velKey2, double delta_t, const gtsam::SharedNoiseModel& Qc_model)
      : keys_{poseKey1, velKey1, poseKey2, velKey2}, delta_t_(delta_t), Qc_(Qc_model) {
    if (!Qc_model) throw std::runtime_error("gpslam_b200: Qc model is not Gaussian");  // getQc dereferences a failed dynamic_cast in the reference (gp/GPutils.cpp:17-19)
  }
  const std::vector<gtsam::Key>& keys() const override { return keys_; }
  GPSLAM_B200_FACTOR(GaussianProcessPriorT, std::string("4-way Gaussian Process Factor ") + G::name(), std::string("GaussianProcessPrior") + detail::TypeName<POSE>::name())
  /// gp/GaussianProcessPriorPose3.h:118-125 (Base, delta_t_); the base's noise model there is Q(delta_t, Qc) - here the Qc model itself
  template <class ARCHIVE> void serialize(ARCHIVE& ar, const unsigned int /*version*/) {
    detail::ioBase(ar, keys_, 4, nullptr);
    ar & GPSLAM_B200_NVP(delta_t_);
    ar & GPSLAM_B200_NVP(Qc_);
    if (!Qc_) throw std::runtime_error("gpslam_b200 archive: GP prior without a Qc model");
  }
  size_t size() const override { return 4; }
  double delta_t() const { return delta_t_; }
  bool chainLink(ChainLink& c) const override { c = ChainLink{keys_[0], keys_[1], 0, keys_[2], keys_[3], 0, false}; return true; }
  /// factor error function (gp/GaussianProcessPriorPose3.h:60-98)
  gtsam::Vector evaluateError(const POSE& pose1, const typename G::Vel& vel1, const POSE& pose2, const typename G::Vel& vel2, gtsam::Matrix* H1 = nullptr,
                              gtsam::Matrix* H2 = nullptr, gtsam::Matrix* H3 = nullptr, gtsam::Matrix* H4 = nullptr) const {
    double x1[12], x2[12], v1[6], v2[6], prm[20] = {0};
    detail::wire(pose1, x1); detail::wire(pose2, x2); detail::wire(vel1, v1); detail::wire(vel2, v2);
    prm[0] = delta_t_;
    return detail::eval(G::group, GPB_F_GP_PRIOR, x1, v1, x2, v2, nullptr, prm, {H1, H2, H3, H4});
  }
  bool sameMembers(const GaussianProcessPriorT& e, double tol) const { return detail::same(delta_t_, e.delta_t_, tol) && detail::sameModel(Qc_, e.Qc_, tol); }
  size_t dim() const override { return 2 * G::D; }
  void lower(gpb_graph* g, int qc_of(void*, const gtsam::Matrix&), void* ctx, const std::map<gtsam::Key, int>& sidx, const std::map<gtsam::Key, int>&) const override {
    const int i = detail::stateOf(sidx, keys_[0]), j = detail::stateOf(sidx, keys_[2]);
    if (j != i + 1 || detail::stateOf(sidx, keys_[1]) != i || detail::stateOf(sidx, keys_[3]) != j) throw std::runtime_error("gpslam_b200: GP prior must join consecutive states");
    detail::check(gpb_add_gp_prior(g, 1, &i, &delta_t_, qc_of(ctx, Qc_->cov)));
  }
};
using GaussianProcessPriorPose3 = GaussianProcessPriorT<gtsam::Pose3>;
using GaussianProcessPriorPose2 = GaussianProcessPriorT<gtsam::Pose2>;
using GaussianProcessPriorRot3 = GaussianProcessPriorT<gtsam::Rot3>;
template <int Dim> using GaussianProcessPriorLinear = GaussianProcessPriorT<gtsam::VectorN<Dim>>;

/// GP interpolators as value types — gp/GaussianProcessInterpolator{Pose3,Pose2,Rot3,Linear}.h (constructor :43-50, interpolatePose
/// :57-105).  interpolatePose runs on the device through gpb_interpolate_poses (one query; the batched forms are
/// gpb_interpolate_poses with n > 1 and gpb_graph_interpolate on an optimised graph).
namespace detail {
inline void unwire(const double* p, gtsam::Pose3& o) { o = gtsam::Pose3::fromWire(p); }
inline void unwire(const double* p, gtsam::Rot3& o) { for (int k = 0; k < 9; k++) o.R[k] = p[k]; }
inline void unwire(const double* p, gtsam::Pose2& o) { o = gtsam::Pose2(p[0], p[1], p[2]); }
inline void unwire(const double* p, gtsam::Vector3& o) { o = gtsam::Vector3{p[0], p[1], p[2]}; }
inline void unwire(const double* p, gtsam::Vector6& o) { o = gtsam::Vector6{p[0], p[1], p[2], p[3], p[4], p[5]}; }
// one interpolatePose query; Hs: up to four D x D Jacobians (nullptr = not requested)
inline void interpolate(int group, int D, const double* x1, const double* v1, const double* x2, const double* v2, double delta_t, double tau, double* pose_out,
                        std::initializer_list<gtsam::Matrix*> Hs) {
  bool want = false;
  for (gtsam::Matrix* h : Hs) want |= (h != nullptr);
  double Hbuf[4 * 36];
  check(gpb_interpolate_poses(group, 0, 1, x1, v1, x2, v2, &delta_t, &tau, pose_out, want ? Hbuf : nullptr));
  int v = 0;
  for (gtsam::Matrix* h : Hs) {
    if (h) { *h = gtsam::Matrix(D, D); for (int k = 0; k < D * D; k++) h->a[k] = Hbuf[v * D * D + k]; }
    v++;
  }
}
}  // namespace detail

template <class POSE>
class GaussianProcessInterpolatorT {
  using G = detail::GroupOf<POSE>;
  double delta_t_ = 0, tau_ = 0;
  gtsam::SharedNoiseModel Qc_;

 public:
  GaussianProcessInterpolatorT() {}
  GaussianProcessInterpolatorT(const gtsam::SharedNoiseModel& Qc_model, double delta_t, double tau) : delta_t_(delta_t), tau_(tau), Qc_(Qc_model) {
    if (!Qc_model) throw std::runtime_error("gpslam_b200: Qc model is not Gaussian");
  }
  double delta_t() const { return delta_t_; }
  double tau() const { return tau_; }
  /// interpolate pose with Jacobians (gp/GaussianProcessInterpolatorPose3.h:57-105)
  POSE interpolatePose(const POSE& pose1, const typename G::Vel& vel1, const POSE& pose2, const typename G::Vel& vel2, gtsam::Matrix* H1 = nullptr,
                       gtsam::Matrix* H2 = nullptr, gtsam::Matrix* H3 = nullptr, gtsam::Matrix* H4 = nullptr) const {
    double x1[12], x2[12], v1[6], v2[6], out[12];
    detail::wire(pose1, x1); detail::wire(pose2, x2); detail::wire(vel1, v1); detail::wire(vel2, v2);
    detail::interpolate(G::group, G::D, x1, v1, x2, v2, delta_t_, tau_, out, {H1, H2, H3, H4});
    POSE p; detail::unwire(out, p);
    return p;
  }
  /// interpolate velocity with Jacobians (gp/GaussianProcessInterpolatorLinear.h:106-126).  Defined for the Linear interpolator only,
  /// as in the reference (its Lie-group interpolators declare interpolateVelocity and never define it): other groups throw.
  typename G::Vel interpolateVelocity(const POSE& pose1, const typename G::Vel& vel1, const POSE& pose2, const typename G::Vel& vel2, gtsam::Matrix* H1 = nullptr,
                                      gtsam::Matrix* H2 = nullptr, gtsam::Matrix* H3 = nullptr, gtsam::Matrix* H4 = nullptr) const {
    double x1[12], x2[12], v1[6], v2[6], out[6], h[4];
    detail::wire(pose1, x1); detail::wire(pose2, x2); detail::wire(vel1, v1); detail::wire(vel2, v2);
    detail::check(gpb_interpolate_velocities(G::group, 0, 1, G::D, x1, v1, x2, v2, &delta_t_, &tau_, out, h));
    gtsam::Matrix* Hs[4] = {H1, H2, H3, H4};
    for (int k = 0; k < 4; k++) if (Hs[k]) { *Hs[k] = gtsam::Matrix(G::D, G::D); for (int d = 0; d < G::D; d++) (*Hs[k])(d, d) = h[k]; }
    typename G::Vel v; detail::unwire(out, v);
    return v;
  }
  void print(const std::string& s = "") const { std::cout << s << "GaussianProcessInterpolator" << G::name() << std::endl; }
  /// gp/GaussianProcessInterpolatorPose3.h:148-159 (delta_t_, tau_, Qc; Lambda and Psi are functions of these three and are
  /// evaluated on the device at every query, so they are not stored)
  template <class ARCHIVE> void serialize(ARCHIVE& ar, const unsigned int /*version*/) { ar & GPSLAM_B200_NVP(delta_t_); ar & GPSLAM_B200_NVP(tau_); ar & make_nvp("Qc", Qc_); }
  bool equals(const GaussianProcessInterpolatorT& e, double tol = 1e-9) const {
    return detail::same(delta_t_, e.delta_t_, tol) && detail::same(tau_, e.tau_, tol) && detail::sameModel(Qc_, e.Qc_, tol);
  }
};
using GaussianProcessInterpolatorPose3 = GaussianProcessInterpolatorT<gtsam::Pose3>;
using GaussianProcessInterpolatorPose2 = GaussianProcessInterpolatorT<gtsam::Pose2>;
using GaussianProcessInterpolatorRot3 = GaussianProcessInterpolatorT<gtsam::Rot3>;
template <int Dim> using GaussianProcessInterpolatorLinear = GaussianProcessInterpolatorT<gtsam::VectorN<Dim>>;

/// 5-way interpolated range factors — slam/GPInterpolatedRangeFactorPose3.h:46-54, ...Pose2.h
template <class POSE>
class GPInterpolatedRangeFactorT : public NonlinearFactor {
  using G = detail::GroupOf<POSE>;
  std::vector<gtsam::Key> keys_;
  double measured_ = 0, delta_t_ = 0, tau_ = 0;
  gtsam::SharedNoiseModel meas_, Qc_;
  bool has_sensor_ = false;
  POSE body_P_sensor_{};

 public:
  GPInterpolatedRangeFactorT() {}  ///< for loading from an archive only (slam/GPInterpolatedRangeFactorPose3.h:43)
  GPInterpolatedRangeFactorT(double measured, const gtsam::SharedNoiseModel& meas_model, const gtsam::SharedNoiseModel& Qc_model, gtsam::Key poseKey1,
                             gtsam::Key velKey1, gtsam::Key poseKey2, gtsam::Key velKey2, gtsam::Key pointKey, double delta_t, double tau,
                             const POSE* body_P_sensor = nullptr)
      : keys_{poseKey1, velKey1, poseKey2, velKey2, pointKey}, measured_(measured), delta_t_(delta_t), tau_(tau), meas_(meas_model), Qc_(Qc_model) {
    if (body_P_sensor) { has_sensor_ = true; body_P_sensor_ = *body_P_sensor; }
  }
  const std::vector<gtsam::Key>& keys() const override { return keys_; }
  GPSLAM_B200_FACTOR(GPInterpolatedRangeFactorT, "RangeFactor, range = " + std::to_string(measured_), std::string("GPInterpolatedRangeFactor") + detail::TypeName<POSE>::name())
  /// slam/GPInterpolatedRangeFactorPose3.h:125-135 (Base, GPbase_, measured_, body_P_sensor_)
  template <class ARCHIVE> void serialize(ARCHIVE& ar, const unsigned int /*version*/) {
    detail::ioBase(ar, keys_, 5, &meas_);
    detail::ioGPbase(ar, delta_t_, tau_, Qc_);
    ar & GPSLAM_B200_NVP(measured_);
    detail::ioOptional(ar, "body_P_sensor_", has_sensor_, body_P_sensor_);
  }
  double measured() const { return measured_; }
  /// slam/GPInterpolatedRangeFactorPose3.h:107-113 (Base, measured_, body_P_sensor_) + the interpolator's members
  bool sameMembers(const GPInterpolatedRangeFactorT& e, double tol) const {
    return detail::same(measured_, e.measured_, tol) && detail::same(delta_t_, e.delta_t_, tol) && detail::same(tau_, e.tau_, tol) && detail::sameModel(meas_, e.meas_, tol) &&
           detail::sameModel(Qc_, e.Qc_, tol) && has_sensor_ == e.has_sensor_ && (!has_sensor_ || detail::sameValue(body_P_sensor_, e.body_P_sensor_, tol));
  }
  size_t dim() const override { return 1; }
  gtsam::Vector evaluateError(const POSE& pose1, const typename G::Vel& vel1, const POSE& pose2, const typename G::Vel& vel2, const typename G::Land& point,
                              gtsam::Matrix* H1 = nullptr, gtsam::Matrix* H2 = nullptr, gtsam::Matrix* H3 = nullptr, gtsam::Matrix* H4 = nullptr,
                              gtsam::Matrix* H5 = nullptr) const {
    double x1[12], x2[12], v1[6], v2[6], l[3], prm[20] = {0};
    detail::wire(pose1, x1); detail::wire(pose2, x2); detail::wire(vel1, v1); detail::wire(vel2, v2); detail::wire(point, l);
    prm[0] = delta_t_; prm[1] = tau_; prm[2] = measured_;
    if (has_sensor_) { detail::wire(body_P_sensor_, prm + 4); prm[16] = 1.0; }
    return detail::eval(G::group, GPB_F_INTERP_RANGE, x1, v1, x2, v2, l, prm, {H1, H2, H3, H4, H5});
  }
  void lower(gpb_graph* g, int qc_of(void*, const gtsam::Matrix&), void* ctx, const std::map<gtsam::Key, int>& sidx, const std::map<gtsam::Key, int>& lidx) const override {
    const int i = detail::stateOf(sidx, keys_[0]), l = detail::stateOf(lidx, keys_[4]);
    const double sigma = 1.0 / detail::sqrtInfo(meas_)(0, 0);
    double bps[12];
    if (has_sensor_) detail::wire(body_P_sensor_, bps);
    detail::check(gpb_add_interp_range(g, 1, &i, &l, &measured_, &sigma, &delta_t_, &tau_, qc_of(ctx, Qc_->cov), has_sensor_ ? bps : nullptr));
  }
};
using GPInterpolatedRangeFactorPose3 = GPInterpolatedRangeFactorT<gtsam::Pose3>;
using GPInterpolatedRangeFactorPose2 = GPInterpolatedRangeFactorT<gtsam::Pose2>;

/// slam/GPInterpolatedRangeFactor2DLinear.h:42-50 — note the reference's different argument order
class GPInterpolatedRangeFactor2DLinear : public GPInterpolatedRangeFactorT<gtsam::Vector3> {
 public:
  GPInterpolatedRangeFactor2DLinear() {}
  GPInterpolatedRangeFactor2DLinear(double measured, gtsam::Key pose1Key, gtsam::Key vel1Key, gtsam::Key pose2Key, gtsam::Key vel2Key, gtsam::Key pointKey,
                                    const gtsam::SharedNoiseModel& meas_model, const gtsam::SharedNoiseModel& Qc_model, double delta_t, double tau)
      : GPInterpolatedRangeFactorT<gtsam::Vector3>(measured, meas_model, Qc_model, pose1Key, vel1Key, pose2Key, vel2Key, pointKey, delta_t, tau) {}
  GPSLAM_B200_FACTOR(GPInterpolatedRangeFactor2DLinear, "RangeFactor, range = " + std::to_string(measured()), "GPInterpolatedRangeFactor2DLinear")
};

/// gtsam::Cal3_S2 (fx, fy, s, u0, v0) — the calibration the reference's projection-factor tests use
namespace gtsam {
struct Cal3_S2 {
  double fx = 1, fy = 1, s = 0, u0 = 0, v0 = 0;
  Cal3_S2() {}
  Cal3_S2(double fx_, double fy_, double s_, double u0_, double v0_) : fx(fx_), fy(fy_), s(s_), u0(u0_), v0(v0_) {}
  template <class ARCHIVE> void serialize(ARCHIVE& ar, const unsigned int) {
    ar & make_nvp("fx_", fx); ar & make_nvp("fy_", fy); ar & make_nvp("s_", s); ar & make_nvp("u0_", u0); ar & make_nvp("v0_", v0);
  }
};
}  // namespace gtsam

/// slam/GPInterpolatedGPSFactorPose3.h:47-56
class GPInterpolatedGPSFactorPose3 : public NonlinearFactor {
  std::vector<gtsam::Key> keys_;
  gtsam::Point3 measured_;
  double delta_t_ = 0, tau_ = 0;
  gtsam::SharedNoiseModel meas_, Qc_;
  bool has_sensor_ = false;
  gtsam::Pose3 body_P_sensor_;

 public:
  GPInterpolatedGPSFactorPose3() {}  ///< for loading from an archive only
  GPInterpolatedGPSFactorPose3(const gtsam::Point3& measured_point3, const gtsam::SharedNoiseModel& meas_model, const gtsam::SharedNoiseModel& Qc_model,
                               gtsam::Key poseKey1, gtsam::Key velKey1, gtsam::Key poseKey2, gtsam::Key velKey2, double delta_t, double tau,
                               const gtsam::Pose3* body_P_sensor = nullptr)
      : keys_{poseKey1, velKey1, poseKey2, velKey2}, measured_(measured_point3), delta_t_(delta_t), tau_(tau), meas_(meas_model), Qc_(Qc_model) {
    if (body_P_sensor) { has_sensor_ = true; body_P_sensor_ = *body_P_sensor; }
  }
  const std::vector<gtsam::Key>& keys() const override { return keys_; }
  GPSLAM_B200_FACTOR(GPInterpolatedGPSFactorPose3, "GPSFactor, point = (" + std::to_string(measured_.x) + ", " + std::to_string(measured_.y) + ", " + std::to_string(measured_.z) + ")", "GPInterpolatedGPSFactorPose3")
  /// slam/GPInterpolatedGPSFactorPose3.h:123-130 (Base, GPbase_, measured_, body_P_sensor_)
  template <class ARCHIVE> void serialize(ARCHIVE& ar, const unsigned int /*version*/) {
    detail::ioBase(ar, keys_, 4, &meas_);
    detail::ioGPbase(ar, delta_t_, tau_, Qc_);
    ar & GPSLAM_B200_NVP(measured_);
    detail::ioOptional(ar, "body_P_sensor_", has_sensor_, body_P_sensor_);
  }
  gtsam::Point3 measured() const { return measured_; }
  bool sameMembers(const GPInterpolatedGPSFactorPose3& e, double tol) const {
    return detail::sameValue(measured_, e.measured_, tol) && detail::same(delta_t_, e.delta_t_, tol) && detail::same(tau_, e.tau_, tol) && detail::sameModel(meas_, e.meas_, tol) &&
           detail::sameModel(Qc_, e.Qc_, tol) && has_sensor_ == e.has_sensor_ && (!has_sensor_ || detail::sameValue(body_P_sensor_, e.body_P_sensor_, tol));
  }
  size_t dim() const override { return 3; }
  gtsam::Vector evaluateError(const gtsam::Pose3& pose1, const gtsam::Vector6& vel1, const gtsam::Pose3& pose2, const gtsam::Vector6& vel2,
                              gtsam::Matrix* H1 = nullptr, gtsam::Matrix* H2 = nullptr, gtsam::Matrix* H3 = nullptr, gtsam::Matrix* H4 = nullptr) const {
    double x1[12], x2[12], v1[6], v2[6], prm[48] = {0};
    detail::wire(pose1, x1); detail::wire(pose2, x2); detail::wire(vel1, v1); detail::wire(vel2, v2);
    prm[0] = delta_t_; prm[1] = tau_; detail::wire(measured_, prm + 40);
    if (has_sensor_) { detail::wire(body_P_sensor_, prm + 4); prm[16] = 1.0; }
    return detail::eval(GPB_POSE3, GPB_F_INTERP_GPS, x1, v1, x2, v2, nullptr, prm, {H1, H2, H3, H4});
  }
  void lower(gpb_graph* g, int qc_of(void*, const gtsam::Matrix&), void* ctx, const std::map<gtsam::Key, int>& sidx, const std::map<gtsam::Key, int>&) const override {
    const int i = detail::stateOf(sidx, keys_[0]);
    double m[3], bps[12];
    detail::wire(measured_, m);
    if (has_sensor_) detail::wire(body_P_sensor_, bps);
    detail::check(gpb_add_interp_gps(g, 1, &i, m, detail::sqrtInfo(meas_).a.data(), &delta_t_, &tau_, qc_of(ctx, Qc_->cov), has_sensor_ ? bps : nullptr));
  }
};

/// slam/GPInterpolatedProjectionFactorPose3.h:60-75 (CALIBRATION = Cal3_S2; the no-throw cheirality path of :123-138)
template <class CALIBRATION = gtsam::Cal3_S2>
class GPInterpolatedProjectionFactorPose3 : public NonlinearFactor {
  std::vector<gtsam::Key> keys_;
  gtsam::Point2 measured_;
  double delta_t_ = 0, tau_ = 0;
  gtsam::SharedNoiseModel meas_, Qc_;
  std::shared_ptr<CALIBRATION> K_;
  bool has_sensor_ = false;
  gtsam::Pose3 body_P_sensor_;
  bool throwCheirality_ = false, verboseCheirality_ = false;  // kept for the archive's member list: the device path is the no-throw one
  void calib(double* k) const { k[0] = K_->fx; k[1] = K_->fy; k[2] = K_->s; k[3] = K_->u0; k[4] = K_->v0; }

 public:
  GPInterpolatedProjectionFactorPose3() {}  ///< for loading from an archive only
  GPInterpolatedProjectionFactorPose3(const gtsam::Point2& measured, const gtsam::SharedNoiseModel& cam_model, const gtsam::SharedNoiseModel& Qc_model,
                                      gtsam::Key poseKey1, gtsam::Key velKey1, gtsam::Key poseKey2, gtsam::Key velKey2, gtsam::Key pointKey, double delta_t,
                                      double tau, const std::shared_ptr<CALIBRATION>& K, const gtsam::Pose3* body_P_sensor = nullptr)
      : keys_{poseKey1, velKey1, poseKey2, velKey2, pointKey}, measured_(measured), delta_t_(delta_t), tau_(tau), meas_(cam_model), Qc_(Qc_model), K_(K) {
    if (!K_) throw std::runtime_error("gpslam_b200: projection factor needs a calibration");
    if (body_P_sensor) { has_sensor_ = true; body_P_sensor_ = *body_P_sensor; }
  }
  const std::vector<gtsam::Key>& keys() const override { return keys_; }
  GPSLAM_B200_FACTOR(GPInterpolatedProjectionFactorPose3, "GPInterpolatedProjectionFactor, z = (" + std::to_string(measured_.x) + ", " + std::to_string(measured_.y) + ")", "GPInterpolatedProjectionFactorPose3")
  /// slam/GPInterpolatedProjectionFactorPose3.h:186-195 (Base, GPbase_, measured_, K_, throwCheirality_, verboseCheirality_) + the sensor pose
  template <class ARCHIVE> void serialize(ARCHIVE& ar, const unsigned int /*version*/) {
    detail::ioBase(ar, keys_, 5, &meas_);
    detail::ioGPbase(ar, delta_t_, tau_, Qc_);
    ar & GPSLAM_B200_NVP(measured_);
    ar & GPSLAM_B200_NVP(K_);
    ar & GPSLAM_B200_NVP(throwCheirality_);
    ar & GPSLAM_B200_NVP(verboseCheirality_);
    detail::ioOptional(ar, "body_P_sensor_", has_sensor_, body_P_sensor_);
    if (!K_) throw std::runtime_error("gpslam_b200 archive: projection factor without a calibration");
  }
  const gtsam::Point2& measured() const { return measured_; }
  const std::shared_ptr<CALIBRATION> calibration() const { return K_; }
  bool sameMembers(const GPInterpolatedProjectionFactorPose3& e, double tol) const {
    double ka[5], kb[5];
    if (!K_ || !e.K_) return false;
    calib(ka); e.calib(kb);
    for (int k = 0; k < 5; k++) if (!detail::same(ka[k], kb[k], tol)) return false;
    return detail::sameValue(measured_, e.measured_, tol) && detail::same(delta_t_, e.delta_t_, tol) && detail::same(tau_, e.tau_, tol) && detail::sameModel(meas_, e.meas_, tol) &&
           detail::sameModel(Qc_, e.Qc_, tol) && has_sensor_ == e.has_sensor_ && (!has_sensor_ || detail::sameValue(body_P_sensor_, e.body_P_sensor_, tol));
  }
  size_t dim() const override { return 2; }
  gtsam::Vector evaluateError(const gtsam::Pose3& pose1, const gtsam::Vector6& vel1, const gtsam::Pose3& pose2, const gtsam::Vector6& vel2, const gtsam::Point3& point,
                              gtsam::Matrix* H1 = nullptr, gtsam::Matrix* H2 = nullptr, gtsam::Matrix* H3 = nullptr, gtsam::Matrix* H4 = nullptr,
                              gtsam::Matrix* H5 = nullptr) const {
    double x1[12], x2[12], v1[6], v2[6], l[3], prm[48] = {0};
    detail::wire(pose1, x1); detail::wire(pose2, x2); detail::wire(vel1, v1); detail::wire(vel2, v2); detail::wire(point, l);
    prm[0] = delta_t_; prm[1] = tau_; detail::wire(measured_, prm + 40); calib(prm + 43);
    if (has_sensor_) { detail::wire(body_P_sensor_, prm + 4); prm[16] = 1.0; }
    return detail::eval(GPB_POSE3, GPB_F_INTERP_PROJECTION, x1, v1, x2, v2, l, prm, {H1, H2, H3, H4, H5});
  }
  void lower(gpb_graph* g, int qc_of(void*, const gtsam::Matrix&), void* ctx, const std::map<gtsam::Key, int>& sidx, const std::map<gtsam::Key, int>& lidx) const override {
    const int i = detail::stateOf(sidx, keys_[0]), l = detail::stateOf(lidx, keys_[4]);
    double m[2], k[5], bps[12];
    detail::wire(measured_, m); calib(k);
    if (has_sensor_) detail::wire(body_P_sensor_, bps);
    detail::check(gpb_add_interp_projection(g, 1, &i, &l, m, detail::sqrtInfo(meas_).a.data(), &delta_t_, &tau_, qc_of(ctx, Qc_->cov), k, has_sensor_ ? bps : nullptr));
  }
};

// ---------------------------------------------------------------------------------- Pose3 "VW" family
// States (Pose3 'x', Vector3 'v', Vector3 'w'): linear and angular velocity in the world frame (gp/Pose3utils.cpp:27-64).  The
// C ABI carries them as one graph group (GPB_POSE3VW, velocity wire [v | w]); the 12x6 / 3x6 velocity blocks it returns are
// split back into the reference's separate H2|H3 and H5|H6 here.
namespace detail {
inline void splitVW(const gtsam::Matrix& H, gtsam::Matrix* Hv, gtsam::Matrix* Hw) {
  for (int half = 0; half < 2; half++) {
    gtsam::Matrix* o = half ? Hw : Hv;
    if (!o) continue;
    *o = gtsam::Matrix(H.rows, 3);
    for (int c = 0; c < 3; c++) for (int r = 0; r < H.rows; r++) (*o)(r, c) = H(r, 3 * half + c);
  }
}
inline gtsam::Vector evalVW(int kind, const gtsam::Pose3& pose1, const gtsam::Vector3& vel1, const gtsam::Vector3& omega1, const gtsam::Pose3& pose2,
                            const gtsam::Vector3& vel2, const gtsam::Vector3& omega2, const double* prm, gtsam::Matrix* H1, gtsam::Matrix* H2, gtsam::Matrix* H3,
                            gtsam::Matrix* H4, gtsam::Matrix* H5, gtsam::Matrix* H6) {
  double x1[12], x2[12], v1[6], v2[6];
  wire(pose1, x1); wire(pose2, x2); wire(vel1, v1); wire(omega1, v1 + 3); wire(vel2, v2); wire(omega2, v2 + 3);
  gtsam::Matrix Hvw1, Hvw2;
  const gtsam::Vector e = eval(GPB_POSE3VW, kind, x1, v1, x2, v2, nullptr, prm, {H1, (H2 || H3) ? &Hvw1 : nullptr, H4, (H5 || H6) ? &Hvw2 : nullptr});
  if (H2 || H3) splitVW(Hvw1, H2, H3);
  if (H5 || H6) splitVW(Hvw2, H5, H6);
  return e;
}
inline void checkVWKeys(const std::map<gtsam::Key, int>& sidx, const std::vector<gtsam::Key>& k) {
  const int i = stateOf(sidx, k[0]), j = stateOf(sidx, k[3]);
  if (j != i + 1 || stateOf(sidx, k[1]) != i || stateOf(sidx, k[2]) != i || stateOf(sidx, k[4]) != j || stateOf(sidx, k[5]) != j)
    throw std::runtime_error("gpslam_b200: a Pose3VW factor must join consecutive states (x_i, v_i, w_i, x_{i+1}, v_{i+1}, w_{i+1})");
}
}  // namespace detail

/// gp/GaussianProcessPriorPose3VW.h:43-51 (6-way factor; evaluateError :62-117)
class GaussianProcessPriorPose3VW : public NonlinearFactor {
  std::vector<gtsam::Key> keys_;
  double delta_t_ = 0;
  gtsam::SharedNoiseModel Qc_;

 public:
  GaussianProcessPriorPose3VW() {}  ///< for loading from an archive only
  GaussianProcessPriorPose3VW(gtsam::Key poseKey1, gtsam::Key velKey1, gtsam::Key omegaKey1, gtsam::Key poseKey2, gtsam::Key velKey2, gtsam::Key omegaKey2, double delta_t,
                              const gtsam::SharedNoiseModel& Qc_model)
      : keys_{poseKey1, velKey1, omegaKey1, poseKey2, velKey2, omegaKey2}, delta_t_(delta_t), Qc_(Qc_model) {
    if (!Qc_model) throw std::runtime_error("gpslam_b200: Qc model is not Gaussian");
  }
  const std::vector<gtsam::Key>& keys() const override { return keys_; }
  GPSLAM_B200_FACTOR(GaussianProcessPriorPose3VW, "4-way Gaussian Process Factor Pose3 VW", "GaussianProcessPriorPose3VW")
  /// gp/GaussianProcessPriorPose3VW.h:139-144 (Base, delta_t_)
  template <class ARCHIVE> void serialize(ARCHIVE& ar, const unsigned int /*version*/) {
    detail::ioBase(ar, keys_, 6, nullptr);
    ar & GPSLAM_B200_NVP(delta_t_);
    ar & GPSLAM_B200_NVP(Qc_);
    if (!Qc_) throw std::runtime_error("gpslam_b200 archive: GP prior without a Qc model");
  }
  bool chainLink(ChainLink& c) const override { c = ChainLink{keys_[0], keys_[1], keys_[2], keys_[3], keys_[4], keys_[5], true}; return true; }
  size_t size() const override { return 6; }
  gtsam::Vector evaluateError(const gtsam::Pose3& pose1, const gtsam::Vector3& vel1, const gtsam::Vector3& omega1, const gtsam::Pose3& pose2, const gtsam::Vector3& vel2,
                              const gtsam::Vector3& omega2, gtsam::Matrix* H1 = nullptr, gtsam::Matrix* H2 = nullptr, gtsam::Matrix* H3 = nullptr, gtsam::Matrix* H4 = nullptr,
                              gtsam::Matrix* H5 = nullptr, gtsam::Matrix* H6 = nullptr) const {
    double prm[48] = {0};
    prm[0] = delta_t_;
    return detail::evalVW(GPB_F_GP_PRIOR, pose1, vel1, omega1, pose2, vel2, omega2, prm, H1, H2, H3, H4, H5, H6);
  }
  bool sameMembers(const GaussianProcessPriorPose3VW& e, double tol) const { return detail::same(delta_t_, e.delta_t_, tol) && detail::sameModel(Qc_, e.Qc_, tol); }
  size_t dim() const override { return 12; }
  void lower(gpb_graph* g, int qc_of(void*, const gtsam::Matrix&), void* ctx, const std::map<gtsam::Key, int>& sidx, const std::map<gtsam::Key, int>&) const override {
    if (gpb_graph_group(g) != GPB_POSE3VW) throw std::runtime_error("gpslam_b200: GaussianProcessPriorPose3VW needs an optimiser created with group GPB_POSE3VW");
    detail::checkVWKeys(sidx, keys_);
    const int i = detail::stateOf(sidx, keys_[0]);
    detail::check(gpb_add_gp_prior(g, 1, &i, &delta_t_, qc_of(ctx, Qc_->cov)));
  }
};

/// gp/GaussianProcessInterpolatorPose3VW.h:43-124 as a value type (interpolatePose :58-108 through gpb_interpolate_poses)
class GaussianProcessInterpolatorPose3VW {
  double delta_t_ = 0, tau_ = 0;
  gtsam::SharedNoiseModel Qc_;

 public:
  GaussianProcessInterpolatorPose3VW() {}
  GaussianProcessInterpolatorPose3VW(const gtsam::SharedNoiseModel& Qc_model, double delta_t, double tau) : delta_t_(delta_t), tau_(tau), Qc_(Qc_model) {}
  double delta_t() const { return delta_t_; }
  double tau() const { return tau_; }
  const gtsam::SharedNoiseModel& Qc() const { return Qc_; }
  gtsam::Pose3 interpolatePose(const gtsam::Pose3& pose1, const gtsam::Vector3& v1, const gtsam::Vector3& omega1, const gtsam::Pose3& pose2, const gtsam::Vector3& v2,
                               const gtsam::Vector3& omega2, gtsam::Matrix* H1 = nullptr, gtsam::Matrix* H2 = nullptr, gtsam::Matrix* H3 = nullptr,
                               gtsam::Matrix* H4 = nullptr, gtsam::Matrix* H5 = nullptr, gtsam::Matrix* H6 = nullptr) const {
    double x1[12], x2[12], vw1[6], vw2[6], out[12];
    detail::wire(pose1, x1); detail::wire(pose2, x2); detail::wire(v1, vw1); detail::wire(omega1, vw1 + 3); detail::wire(v2, vw2); detail::wire(omega2, vw2 + 3);
    gtsam::Matrix Hvw1, Hvw2;
    detail::interpolate(GPB_POSE3VW, 6, x1, vw1, x2, vw2, delta_t_, tau_, out, {H1, (H2 || H3) ? &Hvw1 : nullptr, H4, (H5 || H6) ? &Hvw2 : nullptr});
    if (H2 || H3) detail::splitVW(Hvw1, H2, H3);
    if (H5 || H6) detail::splitVW(Hvw2, H5, H6);
    return gtsam::Pose3::fromWire(out);
  }
  /// gp/GaussianProcessInterpolatorPose3VW.h:175-184 (delta_t_, tau_, Qc)
  template <class ARCHIVE> void serialize(ARCHIVE& ar, const unsigned int /*version*/) { ar & GPSLAM_B200_NVP(delta_t_); ar & GPSLAM_B200_NVP(tau_); ar & make_nvp("Qc", Qc_); }
  bool equals(const GaussianProcessInterpolatorPose3VW& e, double tol = 1e-9) const {
    return detail::same(delta_t_, e.delta_t_, tol) && detail::same(tau_, e.tau_, tol) && detail::sameModel(Qc_, e.Qc_, tol);
  }
  void print(const std::string& s = "") const { std::cout << s << "GaussianProcessInterpolatorPose3VW" << std::endl; }
};

/// slam/GPInterpolatedGPSFactorPose3VW.h:50-61 (evaluateError :71-106)
class GPInterpolatedGPSFactorPose3VW : public NonlinearFactor {
  std::vector<gtsam::Key> keys_;
  gtsam::Point3 measured_;
  GaussianProcessInterpolatorPose3VW GPbase_;
  gtsam::SharedNoiseModel meas_;
  bool has_sensor_ = false;
  gtsam::Pose3 body_P_sensor_;

 public:
  GPInterpolatedGPSFactorPose3VW() {}  ///< for loading from an archive only
  GPInterpolatedGPSFactorPose3VW(const gtsam::Point3& measured_point3, const gtsam::SharedNoiseModel& meas_model, const gtsam::SharedNoiseModel& Qc_model,
                                 gtsam::Key poseKey1, gtsam::Key velKey1, gtsam::Key omegaKey1, gtsam::Key poseKey2, gtsam::Key velKey2, gtsam::Key omegaKey2, double delta_t,
                                 double tau, const gtsam::Pose3* body_P_sensor = nullptr)
      : keys_{poseKey1, velKey1, omegaKey1, poseKey2, velKey2, omegaKey2}, measured_(measured_point3), GPbase_(Qc_model, delta_t, tau), meas_(meas_model) {
    if (body_P_sensor) { has_sensor_ = true; body_P_sensor_ = *body_P_sensor; }
  }
  const std::vector<gtsam::Key>& keys() const override { return keys_; }
  GPSLAM_B200_FACTOR(GPInterpolatedGPSFactorPose3VW, "GPSFactor, point = (" + std::to_string(measured_.x) + ", " + std::to_string(measured_.y) + ", " + std::to_string(measured_.z) + ")", "GPInterpolatedGPSFactorPose3VW")
  /// slam/GPInterpolatedGPSFactorPose3VW.h:132-139 (Base, GPbase_, measured_, body_P_sensor_)
  template <class ARCHIVE> void serialize(ARCHIVE& ar, const unsigned int /*version*/) {
    detail::ioBase(ar, keys_, 6, &meas_);
    ar & GPSLAM_B200_NVP(GPbase_);
    ar & GPSLAM_B200_NVP(measured_);
    detail::ioOptional(ar, "body_P_sensor_", has_sensor_, body_P_sensor_);
  }
  size_t size() const override { return 6; }
  gtsam::Point3 measured() const { return measured_; }
  bool sameMembers(const GPInterpolatedGPSFactorPose3VW& e, double tol) const {
    return detail::sameValue(measured_, e.measured_, tol) && GPbase_.equals(e.GPbase_, tol) && detail::sameModel(meas_, e.meas_, tol) && has_sensor_ == e.has_sensor_ &&
           (!has_sensor_ || detail::sameValue(body_P_sensor_, e.body_P_sensor_, tol));
  }
  size_t dim() const override { return 3; }
  gtsam::Vector evaluateError(const gtsam::Pose3& pose1, const gtsam::Vector3& vel1, const gtsam::Vector3& omega1, const gtsam::Pose3& pose2, const gtsam::Vector3& vel2,
                              const gtsam::Vector3& omega2, gtsam::Matrix* H1 = nullptr, gtsam::Matrix* H2 = nullptr, gtsam::Matrix* H3 = nullptr, gtsam::Matrix* H4 = nullptr,
                              gtsam::Matrix* H5 = nullptr, gtsam::Matrix* H6 = nullptr) const {
    double prm[48] = {0};
    prm[0] = GPbase_.delta_t(); prm[1] = GPbase_.tau(); detail::wire(measured_, prm + 40);
    if (has_sensor_) { detail::wire(body_P_sensor_, prm + 4); prm[16] = 1.0; }
    return detail::evalVW(GPB_F_INTERP_GPS, pose1, vel1, omega1, pose2, vel2, omega2, prm, H1, H2, H3, H4, H5, H6);
  }
  void lower(gpb_graph* g, int qc_of(void*, const gtsam::Matrix&), void* ctx, const std::map<gtsam::Key, int>& sidx, const std::map<gtsam::Key, int>&) const override {
    if (gpb_graph_group(g) != GPB_POSE3VW) throw std::runtime_error("gpslam_b200: GPInterpolatedGPSFactorPose3VW needs an optimiser created with group GPB_POSE3VW");
    detail::checkVWKeys(sidx, keys_);
    const int i = detail::stateOf(sidx, keys_[0]);
    double m[3], bps[12];
    detail::wire(measured_, m);
    if (has_sensor_) detail::wire(body_P_sensor_, bps);
    const double dt = GPbase_.delta_t(), tau = GPbase_.tau();
    detail::check(gpb_add_interp_gps(g, 1, &i, m, detail::sqrtInfo(meas_).a.data(), &dt, &tau, qc_of(ctx, GPbase_.Qc()->cov), has_sensor_ ? bps : nullptr));
  }
};

/// slam/GPInterpolatedAttitudeFactorRot3.h:44-51
class GPInterpolatedAttitudeFactorRot3 : public NonlinearFactor {
  std::vector<gtsam::Key> keys_;
  double delta_t_ = 0, tau_ = 0;
  gtsam::SharedNoiseModel Qc_, meas_;
  gtsam::Unit3 nZ_, bRef_;

 public:
  GPInterpolatedAttitudeFactorRot3() {}  ///< for loading from an archive only
  GPInterpolatedAttitudeFactorRot3(gtsam::Key poseKey1, gtsam::Key velKey1, gtsam::Key poseKey2, gtsam::Key velKey2, double delta_t, double tau,
                                   const gtsam::SharedNoiseModel& Qc_model, const gtsam::SharedNoiseModel& meas_model, const gtsam::Unit3& nZ,
                                   const gtsam::Unit3& bRef = gtsam::Unit3(0, 0, 1))
      : keys_{poseKey1, velKey1, poseKey2, velKey2}, delta_t_(delta_t), tau_(tau), Qc_(Qc_model), meas_(meas_model), nZ_(nZ), bRef_(bRef) {}
  const std::vector<gtsam::Key>& keys() const override { return keys_; }
  GPSLAM_B200_FACTOR(GPInterpolatedAttitudeFactorRot3, "GP Interpolated AttitudeFactor", "GPInterpolatedAttitudeFactorRot3")
  /// slam/GPInterpolatedAttitudeFactorRot3.h:103-112 (NoiseModelFactor4, AttitudeFactor {nZ_, bRef_}, GPbase_)
  template <class ARCHIVE> void serialize(ARCHIVE& ar, const unsigned int /*version*/) {
    detail::ioBase(ar, keys_, 4, &meas_);
    ar.begin("AttitudeFactor"); ar & GPSLAM_B200_NVP(nZ_); ar & GPSLAM_B200_NVP(bRef_); ar.end();
    detail::ioGPbase(ar, delta_t_, tau_, Qc_);
  }
  bool sameMembers(const GPInterpolatedAttitudeFactorRot3& e, double tol) const {
    return detail::same(delta_t_, e.delta_t_, tol) && detail::same(tau_, e.tau_, tol) && detail::sameModel(Qc_, e.Qc_, tol) && detail::sameModel(meas_, e.meas_, tol) &&
           detail::same(nZ_.x, e.nZ_.x, tol) && detail::same(nZ_.y, e.nZ_.y, tol) && detail::same(nZ_.z, e.nZ_.z, tol) && detail::same(bRef_.x, e.bRef_.x, tol) &&
           detail::same(bRef_.y, e.bRef_.y, tol) && detail::same(bRef_.z, e.bRef_.z, tol);
  }
  size_t dim() const override { return 2; }
  gtsam::Vector evaluateError(const gtsam::Rot3& pose1, const gtsam::Vector3& vel1, const gtsam::Rot3& pose2, const gtsam::Vector3& vel2, gtsam::Matrix* H1 = nullptr,
                              gtsam::Matrix* H2 = nullptr, gtsam::Matrix* H3 = nullptr, gtsam::Matrix* H4 = nullptr) const {
    double x1[9], x2[9], prm[20] = {0};
    detail::wire(pose1, x1); detail::wire(pose2, x2);
    prm[0] = delta_t_; prm[1] = tau_; prm[4] = nZ_.x; prm[5] = nZ_.y; prm[6] = nZ_.z; prm[7] = bRef_.x; prm[8] = bRef_.y; prm[9] = bRef_.z;
    return detail::eval(GPB_ROT3, GPB_F_INTERP_ATTITUDE, x1, vel1.data(), x2, vel2.data(), nullptr, prm, {H1, H2, H3, H4});
  }
  void lower(gpb_graph* g, int qc_of(void*, const gtsam::Matrix&), void* ctx, const std::map<gtsam::Key, int>& sidx, const std::map<gtsam::Key, int>&) const override {
    const int i = detail::stateOf(sidx, keys_[0]);
    const double sigma = 1.0 / detail::sqrtInfo(meas_)(0, 0), nz[3] = {nZ_.x, nZ_.y, nZ_.z}, br[3] = {bRef_.x, bRef_.y, bRef_.z};
    detail::check(gpb_add_interp_attitude(g, 1, &i, &delta_t_, &tau_, qc_of(ctx, Qc_->cov), nz, br, &sigma));
  }
};

// ---------------------------------------------------------------------------------- plain 2-way factors of gpslam/slam
/// gtsam::Rot2 as the reference's bearing measurement uses it (an angle)
namespace gtsam {
struct Rot2 {
  double theta_ = 0;
  Rot2() {}
  Rot2(double theta) : theta_(theta) {}  // implicit, as GTSAM's (slam/tests/testSerializationSLAM.cpp:52 passes a double for the bearing)
  static Rot2 fromAngle(double theta) { return Rot2(theta); }
  double theta() const { return theta_; }
  template <class ARCHIVE> void serialize(ARCHIVE& ar, const unsigned int) { ar & GPSLAM_B200_NVP(theta_); }
};
}  // namespace gtsam

/// slam/RangeFactor2DLinear.h:30-34 (Vector3 "linear Pose2" state, Point2 landmark; evaluateError :43-56) and
/// slam/RangeFactorPose2.h:15 (= gtsam::RangeFactor<Pose2, Point2>): one class template over the pose type
template <class POSE>
class RangeFactor2DT : public NonlinearFactor {
  using G = detail::GroupOf<POSE>;
  std::vector<gtsam::Key> keys_;
  double measured_ = 0;
  gtsam::SharedNoiseModel model_;

 public:
  RangeFactor2DT() {}  ///< for loading from an archive only
  RangeFactor2DT(gtsam::Key poseKey, gtsam::Key pointKey, double measured, const gtsam::SharedNoiseModel& model)
      : keys_{poseKey, pointKey}, measured_(measured), model_(model) {}
  const std::vector<gtsam::Key>& keys() const override { return keys_; }
  GPSLAM_B200_FACTOR(RangeFactor2DT, "RangeFactor, range = " + std::to_string(measured_), std::string("RangeFactor2D") + detail::TypeName<POSE>::name())
  /// slam/RangeFactor2DLinear.h:79-84 (Base, measured_)
  template <class ARCHIVE> void serialize(ARCHIVE& ar, const unsigned int /*version*/) { detail::ioBase(ar, keys_, 2, &model_); ar & GPSLAM_B200_NVP(measured_); }
  double measured() const { return measured_; }
  bool sameMembers(const RangeFactor2DT& e, double tol) const { return detail::same(measured_, e.measured_, tol) && detail::sameModel(model_, e.model_, tol); }
  size_t dim() const override { return 1; }
  gtsam::Vector evaluateError(const POSE& pose, const gtsam::Point2& point, gtsam::Matrix* H1 = nullptr, gtsam::Matrix* H2 = nullptr) const {
    double x[3], l[2], prm[48] = {0};
    detail::wire(pose, x); detail::wire(point, l);
    prm[2] = measured_;
    return detail::eval(G::group, GPB_F_RANGE_2D, x, nullptr, nullptr, nullptr, l, prm, {H1, H2});
  }
  void lower(gpb_graph* g, int (*)(void*, const gtsam::Matrix&), void*, const std::map<gtsam::Key, int>& sidx, const std::map<gtsam::Key, int>& lidx) const override {
    detail::check(gpb_add_range_2d(g, detail::stateOf(sidx, keys_[0]), detail::stateOf(lidx, keys_[1]), measured_, 1.0 / detail::sqrtInfo(model_)(0, 0)));
  }
};
using RangeFactor2DLinear = RangeFactor2DT<gtsam::Vector3>;
using RangeFactorPose2 = RangeFactor2DT<gtsam::Pose2>;

/// slam/RangeBearingFactor2DLinear.h:33-38 (evaluateError :47-84): residual (bearing, range)
class RangeBearingFactor2DLinear : public NonlinearFactor {
  std::vector<gtsam::Key> keys_;
  double range_ = 0;
  gtsam::Rot2 bearing_;
  gtsam::SharedNoiseModel model_;

 public:
  RangeBearingFactor2DLinear() {}  ///< for loading from an archive only
  RangeBearingFactor2DLinear(gtsam::Key poseKey, gtsam::Key pointKey, double range, const gtsam::Rot2& bearing, const gtsam::SharedNoiseModel& model)
      : keys_{poseKey, pointKey}, range_(range), bearing_(bearing), model_(model) {}
  const std::vector<gtsam::Key>& keys() const override { return keys_; }
  GPSLAM_B200_FACTOR(RangeBearingFactor2DLinear, "RangeBearingFactor, range = " + std::to_string(range_), "RangeBearingFactor2DLinear")
  /// slam/RangeBearingFactor2DLinear.h:112-118 (Base, range_, bearing_)
  template <class ARCHIVE> void serialize(ARCHIVE& ar, const unsigned int /*version*/) { detail::ioBase(ar, keys_, 2, &model_); ar & GPSLAM_B200_NVP(range_); ar & GPSLAM_B200_NVP(bearing_); }
  bool sameMembers(const RangeBearingFactor2DLinear& e, double tol) const {
    return detail::same(range_, e.range_, tol) && detail::same(bearing_.theta(), e.bearing_.theta(), tol) && detail::sameModel(model_, e.model_, tol);
  }
  size_t dim() const override { return 2; }
  gtsam::Vector evaluateError(const gtsam::Vector3& pose, const gtsam::Point2& point, gtsam::Matrix* H1 = nullptr, gtsam::Matrix* H2 = nullptr) const {
    double x[3], l[2], prm[48] = {0};
    detail::wire(pose, x); detail::wire(point, l);
    prm[2] = range_; prm[3] = bearing_.theta();
    return detail::eval(GPB_LINEAR, GPB_F_RANGE_BEARING_2D, x, nullptr, nullptr, nullptr, l, prm, {H1, H2});
  }
  void lower(gpb_graph* g, int (*)(void*, const gtsam::Matrix&), void*, const std::map<gtsam::Key, int>& sidx, const std::map<gtsam::Key, int>& lidx) const override {
    detail::check(gpb_add_range_bearing_2d(g, detail::stateOf(sidx, keys_[0]), detail::stateOf(lidx, keys_[1]), range_, bearing_.theta(), detail::sqrtInfo(model_).a.data()));
  }
};

/// slam/OdometryFactor2DLinear.h:36-40 (evaluateError :50-75): body-frame odometry (dx, dy, dtheta) between two Vector3 states
class OdometryFactor2DLinear : public NonlinearFactor {
  std::vector<gtsam::Key> keys_;
  gtsam::Vector3 measured_{};
  gtsam::SharedNoiseModel model_;

 public:
  OdometryFactor2DLinear() {}  ///< for loading from an archive only
  OdometryFactor2DLinear(gtsam::Key pose1Key, gtsam::Key pose2Key, const gtsam::Vector3& betweenMeasured, const gtsam::SharedNoiseModel& model)
      : keys_{pose1Key, pose2Key}, measured_(betweenMeasured), model_(model) {}
  const std::vector<gtsam::Key>& keys() const override { return keys_; }
  GPSLAM_B200_FACTOR(OdometryFactor2DLinear, "2-way projected odometry factor", "OdometryFactor2DLinear")
  /// slam/OdometryFactor2DLinear.h:104-110 (NoiseModelFactor2, measured_)
  template <class ARCHIVE> void serialize(ARCHIVE& ar, const unsigned int /*version*/) { detail::ioBase(ar, keys_, 2, &model_); ar & GPSLAM_B200_NVP(measured_); }
  bool sameMembers(const OdometryFactor2DLinear& e, double tol) const { return detail::sameValue(measured_, e.measured_, tol) && detail::sameModel(model_, e.model_, tol); }
  size_t dim() const override { return 3; }
  gtsam::Vector evaluateError(const gtsam::Vector3& pose1, const gtsam::Vector3& pose2, gtsam::Matrix* H1 = nullptr, gtsam::Matrix* H2 = nullptr) const {
    double x1[3], x2[3], prm[48] = {0};
    detail::wire(pose1, x1); detail::wire(pose2, x2); detail::wire(measured_, prm + 4);
    return detail::eval(GPB_LINEAR, GPB_F_ODOMETRY_2D, x1, nullptr, x2, nullptr, nullptr, prm, {H1, H2});
  }
  void lower(gpb_graph* g, int (*)(void*, const gtsam::Matrix&), void*, const std::map<gtsam::Key, int>& sidx, const std::map<gtsam::Key, int>&) const override {
    double m[3];
    detail::wire(measured_, m);
    detail::check(gpb_add_odometry_2d(g, detail::stateOf(sidx, keys_[0]), detail::stateOf(sidx, keys_[1]), m, detail::sqrtInfo(model_).a.data()));
  }
};

/// gtsam::PriorFactor<T> on a pose ('x'), velocity ('v') or landmark ('l') key
template <class T>
class PriorFactor : public NonlinearFactor {
  std::vector<gtsam::Key> keys_;
  T prior_{};
  gtsam::SharedNoiseModel model_;

 public:
  PriorFactor() {}  ///< for loading from an archive only
  PriorFactor(gtsam::Key key, const T& prior, const gtsam::SharedNoiseModel& model) : keys_{key}, prior_(prior), model_(model) {}
  const std::vector<gtsam::Key>& keys() const override { return keys_; }
  GPSLAM_B200_FACTOR(PriorFactor, "PriorFactor", std::string("PriorFactor") + detail::TypeName<T>::name())
  /// gtsam/slam/PriorFactor.h (Base, prior_)
  template <class ARCHIVE> void serialize(ARCHIVE& ar, const unsigned int /*version*/) { detail::ioBase(ar, keys_, 1, &model_); ar & GPSLAM_B200_NVP(prior_); }
  const T& prior() const { return prior_; }
  bool sameMembers(const PriorFactor& e, double tol) const { return detail::sameValue(prior_, e.prior_, tol) && detail::sameModel(model_, e.model_, tol); }
  size_t dim() const override { return detail::modelDim(model_); }
  void lower(gpb_graph* g, int (*)(void*, const gtsam::Matrix&), void*, const std::map<gtsam::Key, int>& sidx, const std::map<gtsam::Key, int>& lidx) const override {
    double v[12];
    detail::wire(prior_, v);
    const gtsam::Matrix& R = detail::sqrtInfo(model_);
    // the role of the key comes from the optimiser's maps, not from its name: lidx holds landmarks (index >= 0) and tags
    // velocity keys with -1, angular-velocity keys of a VW graph with -2; any other key is a pose
    const auto lit = lidx.find(keys_[0]);
    const char c = lit == lidx.end() ? 'x' : (lit->second >= 0 ? 'l' : (lit->second == -2 ? 'w' : 'v'));
    if (c == 'l') detail::check(gpb_add_prior_landmark(g, detail::stateOf(lidx, keys_[0]), v, R.a.data()));
    else if ((c == 'v' || c == 'w') && gpb_graph_group(g) == GPB_POSE3VW) {
      // PriorFactor<Vector3> on the linear ('v') or angular ('w') velocity of a VW state: a 6x6 sqrt information over [v | w] whose
      // other half is zero (zero rows add nothing to the error or the normal equations)
      if (R.rows != 3) throw std::runtime_error("gpslam_b200: PriorFactor on a VW velocity key needs a 3-dimensional model");
      const int o = c == 'w' ? 3 : 0;
      double v6[6] = {0}, R6[36] = {0};
      for (int k = 0; k < 3; k++) { v6[o + k] = v[k]; for (int r = 0; r < 3; r++) R6[(o + r) + 6 * (o + k)] = R(r, k); }
      detail::check(gpb_add_prior_vel(g, detail::stateOf(sidx, keys_[0]), v6, R6));
    }
    else if (c == 'v') detail::check(gpb_add_prior_vel(g, detail::stateOf(sidx, keys_[0]), v, R.a.data()));
    else detail::check(gpb_add_prior_pose(g, detail::stateOf(sidx, keys_[0]), v, R.a.data()));
  }
};

/// gtsam::BetweenFactor<POSE> between consecutive poses (odometry)
template <class POSE>
class BetweenFactor : public NonlinearFactor {
  std::vector<gtsam::Key> keys_;
  POSE measured_{};
  gtsam::SharedNoiseModel model_;

 public:
  BetweenFactor() {}  ///< for loading from an archive only
  BetweenFactor(gtsam::Key key1, gtsam::Key key2, const POSE& measured, const gtsam::SharedNoiseModel& model) : keys_{key1, key2}, measured_(measured), model_(model) {}
  const std::vector<gtsam::Key>& keys() const override { return keys_; }
  GPSLAM_B200_FACTOR(BetweenFactor, "BetweenFactor", std::string("BetweenFactor") + detail::TypeName<POSE>::name())
  /// gtsam/slam/BetweenFactor.h (Base, measured_)
  template <class ARCHIVE> void serialize(ARCHIVE& ar, const unsigned int /*version*/) { detail::ioBase(ar, keys_, 2, &model_); ar & GPSLAM_B200_NVP(measured_); }
  const POSE& measured() const { return measured_; }
  bool sameMembers(const BetweenFactor& e, double tol) const { return detail::sameValue(measured_, e.measured_, tol) && detail::sameModel(model_, e.model_, tol); }
  size_t dim() const override { return detail::modelDim(model_); }
  void lower(gpb_graph* g, int (*)(void*, const gtsam::Matrix&), void*, const std::map<gtsam::Key, int>& sidx, const std::map<gtsam::Key, int>&) const override {
    double v[12];
    detail::wire(measured_, v);
    detail::check(gpb_add_between(g, detail::stateOf(sidx, keys_[0]), detail::stateOf(sidx, keys_[1]), v, detail::sqrtInfo(model_).a.data()));
  }
};

// ================================================================================== containers and optimisers
class NonlinearFactorGraph {
  std::vector<NonlinearFactor::shared_ptr> factors_;

 public:
  template <class F> void add(const F& f) { factors_.push_back(std::make_shared<F>(f)); }
  void push_back(const NonlinearFactor::shared_ptr& f) { factors_.push_back(f); }
  size_t size() const { return factors_.size(); }
  const std::vector<NonlinearFactor::shared_ptr>& factors() const { return factors_; }
  /// every factor behind its type name (FactorRegistry restores the class on loading); noise models and calibrations shared between
  /// factors are written once and stay shared after loading
  template <class ARCHIVE> void serialize(ARCHIVE& ar, const unsigned int /*version*/) {
    std::uint64_t size = factors_.size();
    ar & GPSLAM_B200_NVP(size);
    if constexpr (ARCHIVE::is_loading) {
      if (size > (std::uint64_t(1) << 32)) throw std::runtime_error("gpslam_b200 archive: implausible factor count");
      factors_.assign(static_cast<size_t>(size), nullptr);
    }
    for (auto& f : factors_) {
      std::string type = f ? f->archiveTag() : std::string();
      ar & GPSLAM_B200_NVP(type);
      if constexpr (ARCHIVE::is_loading) { f = FactorRegistry::make(type); f->load(ar); }
      else { if (!f) throw std::runtime_error("gpslam_b200 archive: null factor in the graph"); f->save(ar); }
    }
  }
};

/// gtsam::Values restricted to the variable types of a gpslam trajectory: poses 'x', velocities 'v', landmarks 'l'
class Values {
  friend class NonlinearOptimizer;
  std::map<gtsam::Key, std::vector<double>> v_;  // wire layout per key

 public:
  template <class T> void insert(gtsam::Key k, const T& value) {
    if (v_.count(k)) throw std::runtime_error("gpslam_b200: Values::insert: key already exists");
    double w[12]; detail::wire(value, w);
    int n = 3;
    if (std::is_same<T, gtsam::Pose3>::value) n = 12; else if (std::is_same<T, gtsam::Rot3>::value) n = 9; else if (std::is_same<T, gtsam::Vector6>::value) n = 6;
    else if (std::is_same<T, gtsam::Point2>::value) n = 2;
    v_[k] = std::vector<double>(w, w + n);
  }
  bool exists(gtsam::Key k) const { return v_.count(k) != 0; }
  size_t size() const { return v_.size(); }
  const std::vector<double>& wire(gtsam::Key k) const { auto it = v_.find(k); if (it == v_.end()) throw std::runtime_error("gpslam_b200: Values::at: key not found"); return it->second; }
  template <class T> T at(gtsam::Key k) const;
  const std::map<gtsam::Key, std::vector<double>>& all() const { return v_; }
  std::map<gtsam::Key, std::vector<double>>& all() { return v_; }
  /// key -> value in wire layout (the value's type follows from its length and the trajectory group, as everywhere in this class)
  template <class ARCHIVE> void serialize(ARCHIVE& ar, const unsigned int /*version*/) {
    std::uint64_t size = v_.size();
    ar & GPSLAM_B200_NVP(size);
    if constexpr (ARCHIVE::is_loading) {
      v_.clear();
      for (std::uint64_t k = 0; k < size; k++) {
        gtsam::Key key = 0; std::vector<double> value;
        ar & GPSLAM_B200_NVP(key); ar & GPSLAM_B200_NVP(value);
        if (value.empty() || value.size() > 12 || !v_.emplace(key, std::move(value)).second) throw std::runtime_error("gpslam_b200 archive: bad or duplicate entry in Values");
      }
    } else {
      for (auto& kv : v_) { gtsam::Key key = kv.first; ar & GPSLAM_B200_NVP(key); ar & make_nvp("value", kv.second); }
    }
  }
};
template <> inline gtsam::Pose3 Values::at<gtsam::Pose3>(gtsam::Key k) const { return gtsam::Pose3::fromWire(wire(k).data()); }
template <> inline gtsam::Rot3 Values::at<gtsam::Rot3>(gtsam::Key k) const { gtsam::Rot3 r; for (int i = 0; i < 9; i++) r.R[i] = wire(k)[i]; return r; }
template <> inline gtsam::Pose2 Values::at<gtsam::Pose2>(gtsam::Key k) const { const auto& w = wire(k); return gtsam::Pose2(w[0], w[1], w[2]); }
template <> inline gtsam::Vector3 Values::at<gtsam::Vector3>(gtsam::Key k) const { const auto& w = wire(k); return gtsam::Vector3{w[0], w[1], w[2]}; }
template <> inline gtsam::Vector6 Values::at<gtsam::Vector6>(gtsam::Key k) const { const auto& w = wire(k); return gtsam::Vector6{w[0], w[1], w[2], w[3], w[4], w[5]}; }
template <> inline gtsam::Point3 Values::at<gtsam::Point3>(gtsam::Key k) const { const auto& w = wire(k); return gtsam::Point3(w[0], w[1], w[2]); }
template <> inline gtsam::Point2 Values::at<gtsam::Point2>(gtsam::Key k) const { const auto& w = wire(k); return gtsam::Point2(w[0], w[1]); }

struct GaussNewtonParams { int maxIterations = 100; double relativeErrorTol = 1e-5, absoluteErrorTol = 1e-5, errorTol = 0.0; void setVerbosity(const std::string&) {} };
struct LevenbergMarquardtParams : GaussNewtonParams { double lambdaInitial = 1e-5, lambdaFactor = 10.0, lambdaUpperBound = 1e5, lambdaLowerBound = 0.0, minModelFidelity = 1e-3; };

/// Lowers (graph, values) onto one B200 and drives gpb_optimize; base of GaussNewtonOptimizer / LevenbergMarquardtOptimizer.
class NonlinearOptimizer {
 protected:
  gpb_graph* g_ = nullptr;
  Values values_;
  std::vector<gtsam::Key> xkeys_, vkeys_, wkeys_, lkeys_;  // wkeys_: angular-velocity keys of a GPB_POSE3VW graph
  bool vw_ = false;
  int PS_ = 0, D_ = 0, DL_ = 0, iterations_ = 0;
  double error_ = 0.0;
  gpb_params params_;
  std::vector<gtsam::Matrix> qc_models_;

  static int qcOf(void* self, const gtsam::Matrix& cov) {
    auto* o = static_cast<NonlinearOptimizer*>(self);
    for (size_t k = 0; k < o->qc_models_.size(); k++) if (o->qc_models_[k].a == cov.a) return static_cast<int>(k);
    const int id = gpb_add_qc_model(o->g_, cov.a.data());
    detail::check(id);
    o->qc_models_.push_back(cov);
    return id;
  }
  void pull() {
    std::vector<double> P(xkeys_.size() * PS_), V(xkeys_.size() * D_), L(lkeys_.size() * (DL_ ? DL_ : 1));
    detail::check(gpb_get_values(g_, P.data(), V.data(), lkeys_.empty() ? nullptr : L.data()));
    for (size_t i = 0; i < xkeys_.size(); i++) {
      values_.all()[xkeys_[i]].assign(P.begin() + i * PS_, P.begin() + (i + 1) * PS_);
      if (vw_) { values_.all()[vkeys_[i]].assign(V.begin() + i * 6, V.begin() + i * 6 + 3); values_.all()[wkeys_[i]].assign(V.begin() + i * 6 + 3, V.begin() + i * 6 + 6); }
      else values_.all()[vkeys_[i]].assign(V.begin() + i * D_, V.begin() + (i + 1) * D_);
    }
    for (size_t l = 0; l < lkeys_.size(); l++) values_.all()[lkeys_[l]].assign(L.begin() + l * DL_, L.begin() + (l + 1) * DL_);
  }

 public:
  NonlinearOptimizer(const NonlinearFactorGraph& graph, const Values& initial, int group, int device = 0) : values_(initial), vw_(group == GPB_POSE3VW) {
    // The trajectory order comes from the graph, not from key names: every GP prior links (pose, velocity[, omega]) of one state to
    // the next, so the chain is the path these links form (any keys, any numbering).  Keys outside the chain are landmarks.
    // Graphs without GP priors fall back to the naming convention 'x' i / 'v' i / 'l' j with consecutive i.
    std::map<gtsam::Key, int> sidx, lidx;
    std::map<gtsam::Key, NonlinearFactor::ChainLink> next;   // by first pose key
    std::map<gtsam::Key, int> has_pred;
    for (const auto& f : graph.factors()) {
      NonlinearFactor::ChainLink c;
      if (!f->chainLink(c)) continue;
      if (c.vw != vw_) throw std::runtime_error("gpslam_b200: GP prior family does not match the trajectory group");
      if (next.count(c.x1)) throw std::runtime_error("gpslam_b200: two GP priors start at the same state");
      next[c.x1] = c; has_pred[c.x2] = 1;
    }
    if (!next.empty()) {
      gtsam::Key start = 0; int nstart = 0;
      for (const auto& kv : next) if (!has_pred.count(kv.first)) { start = kv.first; nstart++; }
      if (nstart != 1) throw std::runtime_error("gpslam_b200: the GP priors must form one chain (found " + std::to_string(nstart) + " chain starts)");
      auto push = [&](gtsam::Key x, gtsam::Key v, gtsam::Key w) {
        const int i = static_cast<int>(xkeys_.size());
        xkeys_.push_back(x); vkeys_.push_back(v); sidx[x] = i; sidx[v] = i;
        if (vw_) { wkeys_.push_back(w); sidx[w] = i; }
      };
      gtsam::Key cur = start;
      while (true) {
        auto it = next.find(cur);
        if (it == next.end()) break;
        const auto& c = it->second;
        if (xkeys_.empty()) push(c.x1, c.v1, c.w1);
        else if (vkeys_.back() != c.v1 || (vw_ && wkeys_.back() != c.w1)) throw std::runtime_error("gpslam_b200: a state's pose key is paired with two different velocity keys");
        push(c.x2, c.v2, c.w2);
        cur = c.x2;
        if (xkeys_.size() > next.size() + 1) throw std::runtime_error("gpslam_b200: the GP priors form a cycle");
      }
      if (xkeys_.size() != next.size() + 1) throw std::runtime_error("gpslam_b200: the GP priors must form one chain");
      for (const auto& kv : initial.all()) if (!sidx.count(kv.first)) { lidx[kv.first] = static_cast<int>(lkeys_.size()); lkeys_.push_back(kv.first); }
    } else {
      std::map<std::uint64_t, gtsam::Key> xs, ls;
      for (const auto& kv : initial.all()) {
        const char c = gtsam::symbolChr(kv.first);
        if (c == 'x') xs[gtsam::symbolIndex(kv.first)] = kv.first; else if (c == 'l') ls[gtsam::symbolIndex(kv.first)] = kv.first;
        else if (c != 'v' && !(c == 'w' && vw_)) throw std::runtime_error("gpslam_b200: a graph without GP priors needs 'x', 'v', 'l' keys ('w' on GPB_POSE3VW graphs)");
      }
      std::uint64_t prev = 0; bool first = true;
      for (const auto& kv : xs) {
        if (!first && kv.first != prev + 1) throw std::runtime_error("gpslam_b200: state indices must be consecutive");
        prev = kv.first; first = false;
        const int i = static_cast<int>(xkeys_.size());
        xkeys_.push_back(kv.second); vkeys_.push_back(gtsam::Symbol('v', kv.first));
        sidx[kv.second] = i; sidx[vkeys_.back()] = i;
        if (vw_) { wkeys_.push_back(gtsam::Symbol('w', kv.first)); sidx[wkeys_.back()] = i; }
      }
      for (const auto& kv : ls) { lidx[kv.second] = static_cast<int>(lkeys_.size()); lkeys_.push_back(kv.second); }
    }
    if (xkeys_.size() < 2) throw std::runtime_error("gpslam_b200: need at least two states");
    for (size_t i = 0; i < xkeys_.size(); i++) { lidx[vkeys_[i]] = -1; if (vw_) lidx[wkeys_[i]] = -2; }   // key roles for PriorFactor (see its lower())
    const bool se3 = group == GPB_POSE3 || vw_;
    PS_ = se3 ? 12 : group == GPB_ROT3 ? 9 : 3; D_ = se3 ? 6 : 3; DL_ = se3 ? 3 : group == GPB_ROT3 ? 0 : 2;
    g_ = gpb_graph_create(group, 3, static_cast<int>(xkeys_.size()), static_cast<int>(lkeys_.size()));
    if (!g_) throw std::runtime_error(std::string("gpslam_b200: ") + gpb_last_error());
    struct Guard { gpb_graph*& g; bool armed = true; ~Guard() { if (armed && g) { gpb_graph_destroy(g); g = nullptr; } } } guard{g_};  // a throwing constructor runs no destructor
    for (const auto& f : graph.factors()) f->lower(g_, &NonlinearOptimizer::qcOf, this, sidx, lidx);
    std::vector<double> P(xkeys_.size() * PS_), V(xkeys_.size() * D_), L(lkeys_.size() * (DL_ ? DL_ : 1));
    for (size_t i = 0; i < xkeys_.size(); i++) {
      const auto& p = initial.wire(xkeys_[i]); const auto& v = initial.wire(vkeys_[i]);
      if (static_cast<int>(p.size()) != PS_ || static_cast<int>(v.size()) != (vw_ ? 3 : D_)) throw std::runtime_error("gpslam_b200: value type does not match the trajectory group");
      std::copy(p.begin(), p.end(), P.begin() + i * PS_); std::copy(v.begin(), v.end(), V.begin() + i * D_);
      if (vw_) {
        const auto& w = initial.wire(wkeys_[i]);
        if (w.size() != 3) throw std::runtime_error("gpslam_b200: 'w' values must be Vector3");
        std::copy(w.begin(), w.end(), V.begin() + i * D_ + 3);
      }
    }
    for (size_t l = 0; l < lkeys_.size(); l++) { const auto& w = initial.wire(lkeys_[l]); std::copy(w.begin(), w.end(), L.begin() + l * DL_); }
    detail::check(gpb_set_values(g_, P.data(), V.data(), lkeys_.empty() ? nullptr : L.data()));
    detail::check(gpb_graph_finalize(g_, device));
    detail::check(gpb_error(g_, &error_));
    guard.armed = false;
  }
  NonlinearOptimizer(const NonlinearOptimizer&) = delete;
  virtual ~NonlinearOptimizer() { if (g_) gpb_graph_destroy(g_); }
  /// one optimiser iteration (matlab/PlazaPose2.m:225)
  /// LM keeps its damping between calls, as GTSAM's iterate() keeps state_.lambda: N calls equal one optimize of N iterations
  void iterate() { gpb_stats st; detail::check(gpb_optimize(g_, &params_, 1, &st)); error_ = st.error_final; iterations_ += st.iterations; if (params_.use_lm) params_.lambda_initial = st.lambda; pull(); }
  double lambda() const { return params_.lambda_initial; }
  /// n iterations in one call (same result as n calls of iterate(); values are pulled once)
  void iterate(int n) { gpb_stats st; detail::check(gpb_optimize(g_, &params_, n, &st)); error_ = st.error_final; iterations_ += st.iterations; if (params_.use_lm) params_.lambda_initial = st.lambda; pull(); }
  /// NonlinearOptimizer::optimize(): iterate to GTSAM's convergence test
  const Values& optimize() { gpb_stats st; detail::check(gpb_optimize(g_, &params_, 0, &st)); error_ = st.error_final; iterations_ += st.iterations; pull(); return values_; }
  const Values& values() const { return values_; }
  double error() const { return error_; }
  int iterations() const { return iterations_; }
};

template <class POSE> int groupOfValues() { return detail::GroupOf<POSE>::group; }

class GaussNewtonOptimizer : public NonlinearOptimizer {
 public:
  GaussNewtonOptimizer(const NonlinearFactorGraph& graph, const Values& initial, const GaussNewtonParams& p = GaussNewtonParams(), int group = GPB_POSE3, int device = 0)
      : NonlinearOptimizer(graph, initial, group, device) {
    gpb_default_params(&params_, 0);
    params_.max_iterations = p.maxIterations; params_.rel_tol = p.relativeErrorTol; params_.abs_tol = p.absoluteErrorTol; params_.err_tol = p.errorTol;
  }
};
class LevenbergMarquardtOptimizer : public NonlinearOptimizer {
 public:
  LevenbergMarquardtOptimizer(const NonlinearFactorGraph& graph, const Values& initial, const LevenbergMarquardtParams& p = LevenbergMarquardtParams(),
                              int group = GPB_POSE3, int device = 0)
      : NonlinearOptimizer(graph, initial, group, device) {
    gpb_default_params(&params_, 1);
    params_.max_iterations = p.maxIterations; params_.rel_tol = p.relativeErrorTol; params_.abs_tol = p.absoluteErrorTol; params_.err_tol = p.errorTol;
    params_.lambda_initial = p.lambdaInitial; params_.lambda_factor = p.lambdaFactor; params_.lambda_upper = p.lambdaUpperBound; params_.lambda_lower = p.lambdaLowerBound;
    params_.min_model_fidelity = p.minModelFidelity;
  }
};

// ================================================================================== archives: registry and the gtsam-style helpers
namespace detail {
inline void registerBuiltinFactors() {
  static const bool once = [] {
    FactorRegistry::add<GaussianProcessPriorPose3>(); FactorRegistry::add<GaussianProcessPriorPose2>(); FactorRegistry::add<GaussianProcessPriorRot3>();
    FactorRegistry::add<GaussianProcessPriorLinear<3>>(); FactorRegistry::add<GaussianProcessPriorPose3VW>();
    FactorRegistry::add<GPInterpolatedRangeFactorPose3>(); FactorRegistry::add<GPInterpolatedRangeFactorPose2>();
    FactorRegistry::add<GPInterpolatedRangeFactorT<gtsam::Vector3>>(); FactorRegistry::add<GPInterpolatedRangeFactor2DLinear>();
    FactorRegistry::add<GPInterpolatedGPSFactorPose3>(); FactorRegistry::add<GPInterpolatedGPSFactorPose3VW>();
    FactorRegistry::add<GPInterpolatedProjectionFactorPose3<gtsam::Cal3_S2>>(); FactorRegistry::add<GPInterpolatedAttitudeFactorRot3>();
    FactorRegistry::add<RangeFactor2DLinear>(); FactorRegistry::add<RangeFactorPose2>(); FactorRegistry::add<RangeBearingFactor2DLinear>();
    FactorRegistry::add<OdometryFactor2DLinear>();
    FactorRegistry::add<PriorFactor<gtsam::Pose3>>(); FactorRegistry::add<PriorFactor<gtsam::Pose2>>(); FactorRegistry::add<PriorFactor<gtsam::Rot3>>();
    FactorRegistry::add<PriorFactor<gtsam::Vector3>>(); FactorRegistry::add<PriorFactor<gtsam::Vector6>>(); FactorRegistry::add<PriorFactor<gtsam::Point3>>();
    FactorRegistry::add<PriorFactor<gtsam::Point2>>();
    FactorRegistry::add<BetweenFactor<gtsam::Pose3>>(); FactorRegistry::add<BetweenFactor<gtsam::Pose2>>(); FactorRegistry::add<BetweenFactor<gtsam::Rot3>>();
    FactorRegistry::add<BetweenFactor<gtsam::Vector3>>();
    return true;
  }();
  (void)once;
}
}  // namespace detail
inline NonlinearFactor::shared_ptr FactorRegistry::make(const std::string& tag) {
  detail::registerBuiltinFactors();
  auto it = table().find(tag);
  if (it == table().end()) throw std::runtime_error("gpslam_b200 archive: unknown factor type '" + tag + "' (FactorRegistry::add<F>() registers a user class)");
  return it->second();
}

/// gtsam/base/serialization.h: serialize / deserialize to a string, serializeToFile / deserializeFromFile
template <class T> std::string serialize(const T& input) {
  std::ostringstream os;
  OArchive ar(os);
  ar & make_nvp("data", input);
  return os.str();
}
template <class T> void deserialize(const std::string& serialized, T& output) {
  std::istringstream is(serialized);
  IArchive ar(is);
  ar & make_nvp("data", output);
}
template <class T> bool serializeToFile(const T& input, const std::string& filename) {
  std::ofstream os(filename.c_str());
  if (!os.is_open()) return false;
  OArchive ar(os);
  ar & make_nvp("data", input);
  os.flush();
  return os.good();
}
template <class T> bool deserializeFromFile(const std::string& filename, T& output) {
  std::ifstream is(filename.c_str());
  if (!is.is_open()) return false;
  IArchive ar(is);
  ar & make_nvp("data", output);
  return true;
}
/// one factor behind a base pointer (what a graph archive does per entry)
inline std::string serializeFactor(const NonlinearFactor& f) {
  std::ostringstream os;
  OArchive ar(os);
  std::string type = f.archiveTag();
  ar & GPSLAM_B200_NVP(type);
  f.save(ar);
  return os.str();
}
inline NonlinearFactor::shared_ptr deserializeFactor(const std::string& serialized) {
  std::istringstream is(serialized);
  IArchive ar(is);
  std::string type;
  ar & GPSLAM_B200_NVP(type);
  NonlinearFactor::shared_ptr f = FactorRegistry::make(type);
  f->load(ar);
  return f;
}

}  // namespace gpslam_b200
