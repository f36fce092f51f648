// TEST HARNESS ONLY — declarations (no behaviour) of the GTSAM 4.0 / gpslam entry points include/gpslam_b200/gtsam_adapter.h calls,
// so that the adapter can at least be type-checked where GTSAM, Eigen and Boost do not exist.  Signatures follow
// the way the reference's own sources use them (file:line next to each).  Nothing here is part of the product, and none of it is
// GTSAM code: bodies are placeholders that only have to link.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace boost {  // the reference uses boost::shared_ptr / dynamic_pointer_cast (slam/GPInterpolatedRangeFactorPose3.h:42)
using std::shared_ptr;
using std::dynamic_pointer_cast;
}  // namespace boost

namespace gtsam {
typedef std::uint64_t Key;
struct Matrix {  // Eigen::MatrixXd subset: rows(), cols(), (r, c)
  int r_ = 0, c_ = 0; std::vector<double> a;
  Matrix() {}
  Matrix(int r, int c) : r_(r), c_(c), a(static_cast<size_t>(r) * c, 0.0) {}
  int rows() const { return r_; }
  int cols() const { return c_; }
  double& operator()(int r, int c) { return a[r + static_cast<size_t>(c) * r_]; }
  double operator()(int r, int c) const { return a[r + static_cast<size_t>(c) * r_]; }
};
template <int N> struct FixedVector { double v[N] = {0}; double& operator()(int k) { return v[k]; } double operator()(int k) const { return v[k]; } };
typedef FixedVector<3> Vector3;
typedef FixedVector<6> Vector6;
struct Matrix3 { double m[9] = {0}; double& operator()(int r, int c) { return m[r + 3 * c]; } double operator()(int r, int c) const { return m[r + 3 * c]; } };
struct Point2 { double x_ = 0, y_ = 0; Point2() {} Point2(double x, double y) : x_(x), y_(y) {} double x() const { return x_; } double y() const { return y_; } };
struct Point3 { double x_ = 0, y_ = 0, z_ = 0; Point3() {} Point3(double x, double y, double z) : x_(x), y_(y), z_(z) {} double x() const { return x_; } double y() const { return y_; } double z() const { return z_; } };
struct Rot3 { Matrix3 R; Rot3() {} explicit Rot3(const Matrix3& R_) : R(R_) {} Matrix3 matrix() const { return R; } };                 // gp/Pose3utils.cpp:32
struct Pose3 {                                                                                                                          // gp/GaussianProcessPriorPose3.h:60
  Rot3 r; Point3 t;
  Pose3() {}
  Pose3(const Rot3& r_, const Point3& t_) : r(r_), t(t_) {}
  const Rot3& rotation() const { return r; }
  double x() const { return t.x(); } double y() const { return t.y(); } double z() const { return t.z(); }
};
struct Pose2 { double x_ = 0, y_ = 0, th_ = 0; Pose2() {} Pose2(double x, double y, double th) : x_(x), y_(y), th_(th) {} double x() const { return x_; } double y() const { return y_; } double theta() const { return th_; } };
class Symbol {                                                                                                                          // gp/tests/testGaussianProcessPriorPose3.cpp:36
  unsigned char c_; std::uint64_t j_;
 public:
  Symbol(unsigned char c, std::uint64_t j) : c_(c), j_(j) {}
  Symbol(Key k) : c_(static_cast<unsigned char>(k >> 56)), j_(k & ((Key(1) << 56) - 1)) {}
  operator Key() const { return (static_cast<Key>(c_) << 56) | j_; }
  unsigned char chr() const { return c_; }
  std::uint64_t index() const { return j_; }
};
namespace noiseModel {
struct Base { virtual ~Base() {} };
struct Gaussian : Base {                                                                                                                // gp/GPutils.cpp:16-20 (R()), gp/GaussianProcessPriorPose3.h:46 (Covariance)
  Matrix R_, cov_;
  virtual Matrix R() const { return R_; }
  virtual Matrix covariance() const { return cov_; }
};
}  // namespace noiseModel
typedef boost::shared_ptr<noiseModel::Base> SharedNoiseModel;
class NonlinearFactor {
 protected:
  std::vector<Key> keys_;
 public:
  typedef boost::shared_ptr<NonlinearFactor> shared_ptr;
  virtual ~NonlinearFactor() {}
  const std::vector<Key>& keys() const { return keys_; }
};
class NoiseModelFactor : public NonlinearFactor {
 protected:
  SharedNoiseModel noiseModel_;
 public:
  const SharedNoiseModel& noiseModel() const { return noiseModel_; }
};
class NonlinearFactorGraph {                                                                                                            // matlab/PlazaPose2.m:51
  std::vector<NonlinearFactor::shared_ptr> f_;
 public:
  typedef std::vector<NonlinearFactor::shared_ptr>::const_iterator const_iterator;
  const_iterator begin() const { return f_.begin(); }
  const_iterator end() const { return f_.end(); }
  void push_back(const NonlinearFactor::shared_ptr& f) { f_.push_back(f); }
};
class Values {                                                                                                                          // matlab/PlazaPose2.m:52,183-202
  std::map<Key, std::shared_ptr<void>> v_;
  std::map<Key, const void*> type_;
  template <class T> static const void* tag() { static const char t = 0; return &t; }
 public:
  std::vector<Key> keys() const { std::vector<Key> k; for (const auto& kv : v_) k.push_back(kv.first); return k; }
  template <class T> void insert(Key k, const T& val) { v_[k] = std::make_shared<T>(val); type_[k] = tag<T>(); }
  template <class T> const T* exists(Key k) const { auto it = type_.find(k); return (it != type_.end() && it->second == tag<T>()) ? static_cast<const T*>(v_.at(k).get()) : nullptr; }
  template <class T> const T& at(Key k) const { return *static_cast<const T*>(v_.at(k).get()); }
};
template <class T> class PriorFactor : public NoiseModelFactor {                                                                        // gp/tests/testGaussianProcessPriorPose3.cpp:170
  T prior_;
 public:
  PriorFactor(Key k, const T& p, const SharedNoiseModel& m) : prior_(p) { keys_ = {k}; noiseModel_ = m; }
  const T& prior() const { return prior_; }
};
template <class T> class BetweenFactor : public NoiseModelFactor {                                                                     // matlab/PlazaPose2.m:124
  T measured_;
 public:
  BetweenFactor(Key k1, Key k2, const T& m, const SharedNoiseModel& nm) : measured_(m) { keys_ = {k1, k2}; noiseModel_ = nm; }
  const T& measured() const { return measured_; }
};
template <class T> std::string serializeXML(const T&) { return std::string("<GPbase_><delta_t_>0.1</delta_t_><tau_>0.05</tau_></GPbase_><body_P_sensor_><initialized>0</initialized></body_P_sensor_>"); }  // gtsam/base/serialization.h
}  // namespace gtsam

namespace gpslam {
#define GPSLAM_STUB_PRIOR(NAME) \
  class NAME : public gtsam::NoiseModelFactor { \
   public: \
    NAME(gtsam::Key x1, gtsam::Key v1, gtsam::Key x2, gtsam::Key v2, double, const gtsam::SharedNoiseModel& Q) { keys_ = {x1, v1, x2, v2}; noiseModel_ = Q; } \
  };
GPSLAM_STUB_PRIOR(GaussianProcessPriorPose3)   // gp/GaussianProcessPriorPose3.h:43-49
GPSLAM_STUB_PRIOR(GaussianProcessPriorPose2)   // gp/GaussianProcessPriorPose2.h:41-47
GPSLAM_STUB_PRIOR(GaussianProcessPriorRot3)    // gp/GaussianProcessPriorRot3.h:41-47
#define GPSLAM_STUB_RANGE(NAME) \
  class NAME : public gtsam::NoiseModelFactor { \
    double z_; \
   public: \
    NAME(double z, const gtsam::SharedNoiseModel& m, const gtsam::SharedNoiseModel&, gtsam::Key x1, gtsam::Key v1, gtsam::Key x2, gtsam::Key v2, gtsam::Key l, double, double) : z_(z) { \
      keys_ = {x1, v1, x2, v2, l}; noiseModel_ = m; } \
    double measured() const { return z_; } \
  };
GPSLAM_STUB_RANGE(GPInterpolatedRangeFactorPose3)   // slam/GPInterpolatedRangeFactorPose3.h:46-54, :101-103
GPSLAM_STUB_RANGE(GPInterpolatedRangeFactorPose2)   // slam/GPInterpolatedRangeFactorPose2.h:46-54
}  // namespace gpslam
