// CPU test of the facade's archives (include/gpslam_b200/archive.h and the serialize() members in gpslam.h), written after the
// reference's / GTSAM's serialization tests (gtsam/base/serializationTestHelpers.h: object -> string -> object -> equals):
//   * every factor class, interpolator, value type, a whole graph and a Values container survive object -> text -> object, the
//     second archive being byte-identical to the first (doubles are written with 17 significant digits: exact round trip);
//   * factors are restored behind the base pointer from their type name; shared noise models / calibrations stay shared;
//   * file round trip (serializeToFile / deserializeFromFile);
//   * damaged input fails loudly: truncated archive, wrong member name, wrong key count, unknown factor type, newer class version;
//   * equals(other, tol) / dim() / traits<T> of every factor class and the interpolators.
// Nothing here touches the device (no evaluateError, no optimiser): it runs on a box without a GPU.  Exit code 0 = all passed.
#include <cmath>
#include <cstdio>
#include <limits>

#include "gpslam_b200/gpslam.h"

using namespace gpslam_b200;
using namespace gpslam_b200::gtsam;

static int failures = 0;
#define EXPECT(cond) do { if (!(cond)) { std::printf("EXPECT failed: %s (line %d)\n", #cond, __LINE__); failures++; } } while (0)
template <class F> static bool throws(F f, const char* what) {
  try { f(); } catch (const std::runtime_error& e) { if (std::string(e.what()).find(what) != std::string::npos) return true; std::printf("  threw '%s', expected '%s'\n", e.what(), what); return false; }
  std::printf("  did not throw (expected '%s')\n", what);
  return false;
}

// object -> text -> fresh object -> text: the two texts must be identical
template <class T> static bool roundTrip(const T& obj, T* out = nullptr) {
  const std::string a = serialize(obj);
  T back;
  deserialize(a, back);
  const std::string b = serialize(back);
  if (out) *out = back;
  if (a != b) std::printf("--- first\n%s--- second\n%s", a.c_str(), b.c_str());
  return a == b;
}
static bool roundTripFactor(const NonlinearFactor& f, NonlinearFactor::shared_ptr* out = nullptr) {
  const std::string a = serializeFactor(f);
  NonlinearFactor::shared_ptr back = deserializeFactor(a);
  const std::string b = serializeFactor(*back);
  if (out) *out = back;
  if (a != b) std::printf("--- first\n%s--- second\n%s", a.c_str(), b.c_str());
  return a == b && back->archiveTag() == f.archiveTag() && back->keys() == f.keys() && back->describe() == f.describe();
}

static Matrix denseQc3() { Matrix Q(3, 3); const double v[9] = {2.0, 0.3, -0.1, 0.3, 1.5, 0.2, -0.1, 0.2, 1.1}; for (int k = 0; k < 9; k++) Q.a[k] = v[k]; return Q; }

static void testValueTypes() {
  Matrix M(2, 3); for (int k = 0; k < 6; k++) M.a[k] = 0.1 * k - 1.0 / 3.0;
  Matrix Mb; EXPECT(roundTrip(M, &Mb)); EXPECT(Mb.rows == 2 && Mb.cols == 3 && Mb.a == M.a);
  Matrix E; EXPECT(roundTrip(E));   // empty matrix (a dense-covariance model has an empty sqrt information)
  Point2 p2(1.5, -2.25), p2b; EXPECT(roundTrip(p2, &p2b)); EXPECT(p2b.x == 1.5 && p2b.y == -2.25);
  Point3 p3(1e-300, -3.0, 7e250), p3b; EXPECT(roundTrip(p3, &p3b)); EXPECT(p3b.x == 1e-300 && p3b.z == 7e250);
  Unit3 u(1, 2, 3), ub; EXPECT(roundTrip(u, &ub)); EXPECT(ub.x == u.x && ub.y == u.y && ub.z == u.z);
  Rot3 R = Rot3::Ypr(0.3, -0.2, 1.1), Rb; EXPECT(roundTrip(R, &Rb)); for (int k = 0; k < 9; k++) EXPECT(R.R[k] == Rb.R[k]);
  Pose3 T(R, Point3(0.1, M_PI, -std::sqrt(2.0))), Tb; EXPECT(roundTrip(T, &Tb)); EXPECT(Tb.t.y == M_PI && Tb.r.R[5] == R.R[5]);
  Pose2 q(1, 2, 0.7), qb; EXPECT(roundTrip(q, &qb)); EXPECT(qb.theta == 0.7);
  Rot2 r2(0.123456789012345678), r2b; EXPECT(roundTrip(r2, &r2b)); EXPECT(r2b.theta() == r2.theta());
  Cal3_S2 K(554.25, 553.0, 0.01, 320.5, 240.5), Kb; EXPECT(roundTrip(K, &Kb)); EXPECT(Kb.fx == 554.25 && Kb.v0 == 240.5);
  // exactness at the edges of the double range, and the non-finite values a diverged run may hold
  Point3 edge(std::numeric_limits<double>::denorm_min(), std::numeric_limits<double>::max(), -0.0), eb;
  EXPECT(roundTrip(edge, &eb)); EXPECT(eb.x == edge.x && eb.y == edge.y && std::signbit(eb.z));
  Point2 nf(std::numeric_limits<double>::infinity(), std::nan("")), nfb;
  deserialize(serialize(nf), nfb); EXPECT(std::isinf(nfb.x) && std::isnan(nfb.y));
}

static void testInterpolators() {
  auto Qc = noiseModel::Gaussian::Covariance(denseQc3());
  GaussianProcessInterpolatorPose2 a(Qc, 0.1, 0.04), ab;
  EXPECT(roundTrip(a, &ab)); EXPECT(ab.equals(a)); EXPECT(ab.delta_t() == 0.1 && ab.tau() == 0.04);
  auto Qc6 = noiseModel::Isotropic::Sigma(6, 0.3);
  GaussianProcessInterpolatorPose3 b(Qc6, 0.5, 0.2), bb; EXPECT(roundTrip(b, &bb)); EXPECT(bb.equals(b));
  GaussianProcessInterpolatorRot3 c(Qc, 0.25, 0.2), cb; EXPECT(roundTrip(c, &cb)); EXPECT(cb.equals(c));
  GaussianProcessInterpolatorLinear<3> d(Qc, 0.25, 0.3), db; EXPECT(roundTrip(d, &db)); EXPECT(db.equals(d));
  GaussianProcessInterpolatorPose3VW e(Qc6, 0.2, 0.1), eb; EXPECT(roundTrip(e, &eb)); EXPECT(eb.equals(e));
  GaussianProcessInterpolatorPose3 empty, emptyb; EXPECT(roundTrip(empty, &emptyb));   // default-constructed: null Qc pointer
}

static NonlinearFactorGraph everyFactor(SharedNoiseModel* sharedQc = nullptr) {
  auto Qc6 = noiseModel::Isotropic::Sigma(6, 0.2), Qc3 = noiseModel::Gaussian::Covariance(denseQc3());
  auto m1 = noiseModel::Isotropic::Sigma(1, 0.05), m2 = noiseModel::Diagonal::Sigmas({0.1, 0.2}), m3 = noiseModel::Isotropic::Sigma(3, 0.3), m6 = noiseModel::Isotropic::Sigma(6, 0.01);
  if (sharedQc) *sharedQc = Qc6;
  const Key x0 = Symbol('x', 0), v0 = Symbol('v', 0), x1 = Symbol('x', 1), v1 = Symbol('v', 1), w0 = Symbol('w', 0), w1 = Symbol('w', 1), l0 = Symbol('l', 0);
  const Pose3 sensor(Rot3::Ypr(0.1, 0.2, 0.3), Point3(0.1, 0.0, -0.2));
  const Pose2 sensor2(0.1, -0.1, 0.05);
  NonlinearFactorGraph g;
  g.add(GaussianProcessPriorPose3(x0, v0, x1, v1, 0.1, Qc6));
  g.add(GaussianProcessPriorPose2(x0, v0, x1, v1, 0.2, Qc3));
  g.add(GaussianProcessPriorRot3(x0, v0, x1, v1, 0.3, Qc3));
  g.add(GaussianProcessPriorLinear<3>(x0, v0, x1, v1, 0.4, Qc3));
  g.add(GaussianProcessPriorPose3VW(x0, v0, w0, x1, v1, w1, 0.5, Qc6));
  g.add(GPInterpolatedRangeFactorPose3(3.25, m1, Qc6, x0, v0, x1, v1, l0, 0.1, 0.04, &sensor));
  g.add(GPInterpolatedRangeFactorPose3(1.0 / 3.0, m1, Qc6, x0, v0, x1, v1, l0, 0.1, 0.06));
  g.add(GPInterpolatedRangeFactorPose2(2.5, m1, Qc3, x0, v0, x1, v1, l0, 0.2, 0.1, &sensor2));
  g.add(GPInterpolatedRangeFactor2DLinear(4.5, x0, v0, x1, v1, l0, m1, Qc3, 0.2, 0.15));
  g.add(GPInterpolatedGPSFactorPose3(Point3(1, 2, 3), m3, Qc6, x0, v0, x1, v1, 0.1, 0.03, &sensor));
  g.add(GPInterpolatedGPSFactorPose3VW(Point3(-1, 0.5, 2), m3, Qc6, x0, v0, w0, x1, v1, w1, 0.5, 0.25));
  auto K = std::make_shared<Cal3_S2>(554.0, 554.0, 0.0, 320.0, 240.0);
  g.add(GPInterpolatedProjectionFactorPose3<Cal3_S2>(Point2(300.5, 200.25), m2, Qc6, x0, v0, x1, v1, l0, 0.1, 0.05, K, &sensor));
  g.add(GPInterpolatedProjectionFactorPose3<Cal3_S2>(Point2(310.5, 210.25), m2, Qc6, x0, v0, x1, v1, l0, 0.1, 0.07, K));
  g.add(GPInterpolatedAttitudeFactorRot3(x0, v0, x1, v1, 0.3, 0.1, Qc3, m2, Unit3(0, 0, -1), Unit3(0.1, 0.2, 1.0)));
  g.add(RangeFactor2DLinear(x0, l0, 2.0, m1));
  g.add(RangeFactorPose2(x1, l0, 2.5, m1));
  g.add(RangeBearingFactor2DLinear(x0, l0, 3.0, Rot2::fromAngle(0.4), m2));
  g.add(OdometryFactor2DLinear(x0, x1, Vector3{0.5, 0.1, 0.02}, m3));
  g.add(PriorFactor<Pose3>(x0, sensor, m6));
  g.add(PriorFactor<Pose2>(x0, sensor2, m3));
  g.add(PriorFactor<Rot3>(x0, Rot3::Ypr(0.3, 0.2, 0.1), m3));
  g.add(PriorFactor<Vector3>(v0, Vector3{0.1, 0.2, 0.3}, m3));
  g.add(PriorFactor<Vector6>(v0, Vector6{0.1, 0.2, 0.3, 0.4, 0.5, 0.6}, m6));
  g.add(PriorFactor<Point3>(l0, Point3(5, 6, 7), m3));
  g.add(PriorFactor<Point2>(l0, Point2(5, 6), m2));
  g.add(BetweenFactor<Pose3>(x0, x1, sensor, m6));
  g.add(BetweenFactor<Pose2>(x0, x1, sensor2, m3));
  return g;
}

static void testEveryFactorClass() {
  const NonlinearFactorGraph g = everyFactor();
  std::map<std::string, int> tags;
  for (const auto& f : g.factors()) {
    NonlinearFactor::shared_ptr back;
    EXPECT(roundTripFactor(*f, &back));
    EXPECT(back.get() != f.get());
    tags[f->archiveTag()]++;
    // the restored object is of the same dynamic class: clone() and print() behave the same
    EXPECT(back->clone()->describe() == f->describe());
    EXPECT(back->size() == f->size());
    NonlinearFactor::ChainLink c1, c2;
    EXPECT(back->chainLink(c1) == f->chainLink(c2));
  }
  EXPECT(tags.size() == 25);   // 27 factors, two classes appear twice (with / without a sensor pose)
  EXPECT(tags.count("GaussianProcessPriorPose3") && tags.count("GPInterpolatedRangeFactor2DLinear") && tags.count("PriorFactorVector6"));
  // typed access: a concrete factor restored as its own class keeps its members
  GPInterpolatedRangeFactorPose3 rf, rf0(3.25, noiseModel::Isotropic::Sigma(1, 0.05), noiseModel::Isotropic::Sigma(6, 0.2), 1, 2, 3, 4, 5, 0.1, 0.04);
  deserialize(serialize(rf0), rf);
  EXPECT(rf.measured() == 3.25 && rf.keys() == rf0.keys());
  GaussianProcessPriorPose3 gp, gp0(1, 2, 3, 4, 0.125, noiseModel::Isotropic::Sigma(6, 0.2));
  deserialize(serialize(gp0), gp);
  EXPECT(gp.equals(gp0) && gp.delta_t() == 0.125);
}

static void testGraphAndValues() {
  SharedNoiseModel Qc6;
  const NonlinearFactorGraph g = everyFactor(&Qc6);
  const std::string a = serialize(g);
  NonlinearFactorGraph gb;
  deserialize(a, gb);
  EXPECT(gb.size() == g.size());
  EXPECT(serialize(gb) == a);
  for (size_t k = 0; k < g.size(); k++) EXPECT(gb.factors()[k]->archiveTag() == g.factors()[k]->archiveTag() && gb.factors()[k]->keys() == g.factors()[k]->keys());
  // shared objects are written once: the 6-d Qc model is used by 10 factors but its covariance appears a single time
  size_t bodies = 0, pos = 0;
  const std::string needle = "dim 6";
  while ((pos = a.find(needle, pos)) != std::string::npos) { bodies++; pos += needle.size(); }
  EXPECT(bodies == 2);   // Qc6 and the 6-d measurement model m6, once each
  Values v;
  v.insert(Symbol('x', 0), Pose3(Rot3::Ypr(0.1, 0.2, 0.3), Point3(1, 2, 3)));
  v.insert(Symbol('v', 0), Vector6{1, 2, 3, 4, 5, 6});
  v.insert(Symbol('l', 7), Point3(0.5, 0.25, 0.125));
  v.insert(12345, Point2(1, 2));   // a plain integer key
  Values vb;
  EXPECT(roundTrip(v, &vb));
  EXPECT(vb.size() == 4 && vb.at<Point3>(Symbol('l', 7)).y == 0.25 && vb.at<Vector6>(Symbol('v', 0))[5] == 6 && vb.exists(12345));
  EXPECT(vb.at<Pose3>(Symbol('x', 0)).r.R[3] == v.at<Pose3>(Symbol('x', 0)).r.R[3]);
  // files
  const std::string path = "/tmp/gpslam_b200_test_archive.txt";
  EXPECT(serializeToFile(g, path));
  NonlinearFactorGraph gf;
  EXPECT(deserializeFromFile(path, gf));
  EXPECT(serialize(gf) == a);
  EXPECT(!deserializeFromFile("/nonexistent/dir/file.txt", gf));
  EXPECT(!serializeToFile(g, "/nonexistent/dir/file.txt"));
  std::remove(path.c_str());
}

// gp/tests/testSerializationGP.cpp:38-92 and slam/tests/testSerializationSLAM.cpp:33-83 with their objects and constructor
// arguments (the GP test's commented-out objects included; Linear<6> is not a group of this engine: Linear<3> stands in).
// equalsObj (gtsam/base/serializationTestHelpers.h) = write, read into a default-constructed object, compare with equals().
template <class T> static bool equalsObj(const T& input) {
  T output;
  deserialize(serialize(input), output);
  return input.equals(output) && output.equals(input);
}
static void testReferenceSerializationTests() {
  SharedNoiseModel Qcmodel_6 = noiseModel::Isotropic::Sigma(6, 0.1), Qcmodel_3 = noiseModel::Isotropic::Sigma(3, 0.1);
  // bases (the reference's GaussianProcessFactorBase* are today's GaussianProcessInterpolator*)
  EXPECT(equalsObj(GaussianProcessInterpolatorLinear<3>(Qcmodel_3, 0.1, 0.04)));
  EXPECT(equalsObj(GaussianProcessInterpolatorPose3(Qcmodel_6, 0.1, 0.04)));
  EXPECT(equalsObj(GaussianProcessInterpolatorPose3VW(Qcmodel_6, 0.1, 0.04)));
  EXPECT(equalsObj(GaussianProcessInterpolatorPose2(Qcmodel_3, 0.1, 0.04)));
  EXPECT(equalsObj(GaussianProcessInterpolatorRot3(Qcmodel_3, 0.1, 0.04)));
  // factors
  EXPECT(equalsObj(GaussianProcessPriorLinear<3>(1, 2, 3, 4, 0.1, Qcmodel_3)));
  EXPECT(equalsObj(GaussianProcessPriorPose3(1, 2, 3, 4, 0.1, Qcmodel_6)));
  EXPECT(equalsObj(GaussianProcessPriorPose3VW(1, 2, 3, 4, 5, 6, 0.1, Qcmodel_6)));
  EXPECT(equalsObj(GaussianProcessPriorPose2(1, 2, 3, 4, 0.1, Qcmodel_3)));
  EXPECT(equalsObj(GaussianProcessPriorRot3(1, 2, 3, 4, 0.1, Qcmodel_3)));
  // slam
  SharedNoiseModel unit1 = noiseModel::Unit::Create(1), unit2 = noiseModel::Unit::Create(2), unit3 = noiseModel::Unit::Create(3);
  std::shared_ptr<Cal3_S2> K(new Cal3_S2());
  EXPECT(equalsObj(GPInterpolatedGPSFactorPose3(Point3(0.3, 0.6, 0.9), unit2, Qcmodel_6, 1, 2, 3, 4, 0.1, 0.04)));
  EXPECT(equalsObj(GPInterpolatedGPSFactorPose3VW(Point3(0.3, 0.6, 0.9), unit2, Qcmodel_6, 1, 2, 3, 4, 5, 6, 0.1, 0.04)));
  EXPECT(equalsObj(GPInterpolatedProjectionFactorPose3<Cal3_S2>(Point2(10, 20), unit2, Qcmodel_6, 1, 2, 3, 4, 5, 0.1, 0.04, K)));
  EXPECT(equalsObj(GPInterpolatedRangeFactorPose3(10.0, unit1, Qcmodel_6, 1, 2, 3, 4, 5, 0.1, 0.04)));
  EXPECT(equalsObj(OdometryFactor2DLinear(1, 2, Vector3{0.1, 0.2, 3.0}, unit3)));
  EXPECT(equalsObj(RangeFactor2DLinear(1, 2, 10.0, unit1)));
  EXPECT(equalsObj(RangeBearingFactor2DLinear(1, 2, 0.1, 10.0, unit2)));
}

// a user-defined factor joins through FactorRegistry::add
class MyFactor : public NonlinearFactor {
  std::vector<Key> keys_;
  double weight_ = 0;

 public:
  MyFactor() {}
  MyFactor(Key k, double w) : keys_{k}, weight_(w) {}
  const std::vector<Key>& keys() const override { return keys_; }
  GPSLAM_B200_FACTOR(MyFactor, "MyFactor", "test::MyFactor")
  double weight() const { return weight_; }
  bool sameMembers(const MyFactor& e, double tol) const { return std::fabs(weight_ - e.weight_) <= tol; }
  size_t dim() const override { return 1; }
  template <class ARCHIVE> void serialize(ARCHIVE& ar, const unsigned int) { detail::ioBase(ar, keys_, 1, nullptr); ar & GPSLAM_B200_NVP(weight_); }
  void lower(gpb_graph*, int (*)(void*, const Matrix&), void*, const std::map<Key, int>&, const std::map<Key, int>&) const override {}
};

static void testUserFactorAndStrings() {
  MyFactor f(Symbol('x', 3), 2.5);
  const std::string a = serializeFactor(f);
  EXPECT(throws([&] { deserializeFactor(a); }, "unknown factor type 'test::MyFactor'"));
  FactorRegistry::add<MyFactor>();
  auto back = deserializeFactor(a);
  EXPECT(std::dynamic_pointer_cast<MyFactor>(back) && std::dynamic_pointer_cast<MyFactor>(back)->weight() == 2.5);
  // strings: spaces, percent signs, control bytes, the empty string
  for (const std::string& s : {std::string("plain"), std::string("with space and %"), std::string(""), std::string("tab\tnewline\n\x01"), std::string("100%%")}) {
    std::ostringstream os; { OArchive ar(os); std::string t = s; ar & make_nvp("text", t); }
    std::istringstream is(os.str()); IArchive ar(is); std::string t; ar & make_nvp("text", t);
    EXPECT(t == s);
  }
}

// equals / dim / traits (SURVEY §8b "other virtuals to honour": gp/GaussianProcessPriorPose3.h:101-115,134-137)
static void testEqualsDimTraits() {
  const NonlinearFactorGraph g = everyFactor();
  const size_t dims[] = {12, 6, 6, 6, 12, 1, 1, 1, 1, 3, 3, 2, 2, 2, 1, 1, 2, 3, 6, 3, 3, 3, 6, 3, 2, 6, 3};
  EXPECT(g.size() == sizeof(dims) / sizeof(dims[0]));
  for (size_t k = 0; k < g.size(); k++) {
    const auto& f = g.factors()[k];
    EXPECT(f->dim() == dims[k]);
    EXPECT(f->equals(*f));
    EXPECT(f->equals(*f->clone()));
    EXPECT(deserializeFactor(serializeFactor(*f))->equals(*f, 0.0));     // archives are exact
    for (size_t j = 0; j < g.size(); j++) if (j != k) EXPECT(!f->equals(*g.factors()[j]));   // another class, other keys or other members
  }
  // a member off by more than the tolerance, the same member inside it, other keys
  auto Qc = noiseModel::Isotropic::Sigma(6, 0.2), Qc2 = noiseModel::Isotropic::Sigma(6, 0.21), m1 = noiseModel::Isotropic::Sigma(1, 0.05);
  const GaussianProcessPriorPose3 a(1, 2, 3, 4, 0.1, Qc), b(1, 2, 3, 4, 0.1 + 1e-12, Qc), c(1, 2, 3, 4, 0.1 + 1e-6, Qc), d(1, 2, 3, 5, 0.1, Qc), e(1, 2, 3, 4, 0.1, Qc2);
  EXPECT(a.equals(b) && !a.equals(c) && a.equals(c, 1e-5) && !a.equals(d) && !a.equals(e));
  const Pose3 s1(Rot3::Ypr(0.1, 0.2, 0.3), Point3(1, 2, 3)), s2(Rot3::Ypr(0.1, 0.2, 0.3), Point3(1, 2, 3.001));
  const GPInterpolatedRangeFactorPose3 r0(2.0, m1, Qc, 1, 2, 3, 4, 5, 0.1, 0.04), r1(2.0, m1, Qc, 1, 2, 3, 4, 5, 0.1, 0.04, &s1), r2(2.0, m1, Qc, 1, 2, 3, 4, 5, 0.1, 0.04, &s2),
      r3(2.0, m1, Qc, 1, 2, 3, 4, 5, 0.1, 0.05), r4(2.5, m1, Qc, 1, 2, 3, 4, 5, 0.1, 0.04);
  EXPECT(r0.equals(r0) && !r0.equals(r1) && !r1.equals(r2) && r1.equals(r2, 0.01) && !r0.equals(r3) && !r0.equals(r4) && !r0.equals(a));
  // the 2-D linear range factor is its own class although it shares its members with the Vector3 instantiation of the template
  auto Qc3 = noiseModel::Isotropic::Sigma(3, 0.2);
  const GPInterpolatedRangeFactor2DLinear l1(4.5, 1, 2, 3, 4, 5, m1, Qc3, 0.2, 0.15);
  const GPInterpolatedRangeFactorT<Vector3> l2(4.5, m1, Qc3, 1, 2, 3, 4, 5, 0.2, 0.15);
  EXPECT(l1.equals(l1) && !l1.equals(l2) && l2.equals(l1));   // as with any derived class: the base accepts the derived object, not the reverse
  // traits forward to the members; interpolators are Testable value types
  EXPECT(traits<GaussianProcessPriorPose3>::Equals(a, b) && !traits<GaussianProcessPriorPose3>::Equals(a, c));
  const GaussianProcessInterpolatorPose3 i1(Qc, 0.1, 0.04), i2(Qc, 0.1, 0.04), i3(Qc, 0.1, 0.05), i4(Qc2, 0.1, 0.04);
  EXPECT(traits<GaussianProcessInterpolatorPose3>::Equals(i1, i2) && !traits<GaussianProcessInterpolatorPose3>::Equals(i1, i3) && !i1.equals(i4));
  {  // gp/tests/testGaussianProcessInterpolatorLinear.cpp:25-45, the reference's own equals test
    typedef GaussianProcessInterpolatorLinear<3> GPBase3;
    auto Qc_model = noiseModel::Gaussian::Covariance(0.01 * Matrix::Identity(3, 3)), Qc2_model = noiseModel::Gaussian::Covariance(0.02 * Matrix::Identity(3, 3));
    double dt = 0.1, tau = 0.03;
    GPBase3 base1(Qc_model, dt, tau), base2(Qc_model, dt, tau), base3(Qc2_model, dt, tau), base4(Qc_model, 0.2, 0.03), base5(Qc_model, 0.1, 0.06);
    EXPECT(base1.equals(base2, 1e-9));
    EXPECT(!base1.equals(base3, 1e-9));
    EXPECT(!base1.equals(base4, 1e-9));
    EXPECT(!base1.equals(base5, 1e-9));
  }
  const GaussianProcessInterpolatorPose3VW w1(Qc, 0.1, 0.04), w2(Qc, 0.1, 0.04), w3(Qc, 0.2, 0.04);
  EXPECT(traits<GaussianProcessInterpolatorPose3VW>::Equals(w1, w2) && !w1.equals(w3));
}

static void testDamagedInput() {
  const GaussianProcessPriorPose3 gp(1, 2, 3, 4, 0.125, noiseModel::Isotropic::Sigma(6, 0.2));
  const std::string a = serialize(gp);
  GaussianProcessPriorPose3 out;
  EXPECT(throws([&] { deserialize(a.substr(0, a.size() / 2), out); }, "archive"));                       // truncated (mid-token or missing rest)
  EXPECT(throws([&] { deserialize(std::string("gpslam_b200::archive 1\n"), out); }, "truncated"));
  EXPECT(throws([&] { deserialize(std::string("something else"), out); }, "expected 'gpslam_b200::archive'"));
  EXPECT(throws([&] { deserialize(std::string("gpslam_b200::archive 2\n") + a.substr(a.find('\n') + 1), out); }, "unsupported format version"));
  std::string b = a; b.replace(b.find("delta_t_"), 8, "delta_x_");
  EXPECT(throws([&] { deserialize(b, out); }, "expected 'delta_t_', found 'delta_x_'"));
  std::string c = a; c.replace(c.find("[ 4 1 2 3 4 ]"), 13, "[ 3 1 2 3 ]");
  EXPECT(throws([&] { deserialize(c, out); }, "wrong number of keys"));
  std::string d = a; d.replace(d.find("0.125"), 5, "0.12x");
  EXPECT(throws([&] { deserialize(d, out); }, "is not a number"));
  std::string e = a; e.replace(e.find("version 0"), 9, "version 7");
  EXPECT(throws([&] { deserialize(e, out); }, "newer version"));
  std::string f = a; f.replace(f.find("Qc_ @1 {"), 8, "Qc_ @0");   // a null Qc where the class needs one (the body that follows is then unexpected, too)
  EXPECT(throws([&] { deserialize(f, out); }, "archive"));
  // a reference to a shared object that was never defined
  NonlinearFactorGraph g2; g2.add(gp); g2.add(gp);
  std::string h = serialize(g2);
  const size_t second = h.rfind("Qc_ @1");
  h.replace(second, 6, "Qc_ @9");
  NonlinearFactorGraph gout;
  EXPECT(throws([&] { deserialize(h, gout); }, "archive"));
  // matrix whose data does not match its shape
  Matrix M(2, 2); std::string m = serialize(M); m.replace(m.find("cols 2"), 6, "cols 3");
  Matrix Mo; EXPECT(throws([&] { deserialize(m, Mo); }, "does not match its shape"));
  // Values: duplicate key
  Values v; v.insert(1, Point2(1, 2)); v.insert(2, Point2(3, 4));
  std::string vs = serialize(v); vs.replace(vs.find("key 2"), 5, "key 1");
  Values vo; EXPECT(throws([&] { deserialize(vs, vo); }, "duplicate"));
}

int main() {
  testValueTypes();
  testInterpolators();
  testEveryFactorClass();
  testGraphAndValues();
  testReferenceSerializationTests();
  testUserFactorAndStrings();
  testEqualsDimTraits();
  testDamagedInput();
  if (failures) { std::printf("%d EXPECT(s) failed\n", failures); return 1; }
  std::printf("archive tests passed\n");
  return 0;
}
