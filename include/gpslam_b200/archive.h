// Text archives for the host facade (gpslam.h): every factor, interpolator, graph and Values container can be written to and
// read back from a stream, the way the reference's classes go through boost::serialization
// (gp/GaussianProcessPriorPose3.h:118-125, gp/GaussianProcessInterpolatorPose3.h:148-159, slam/GPInterpolatedRangeFactorPose3.h:125-135,
// and gtsam/base/serialization.h for the serialize / deserialize / ...ToFile helpers mirrored at the end of gpslam.h).
//
// Boost is not a dependency of this package, and the byte layout of a boost archive of GTSAM objects (class-id tables, tracked
// pointers of GTSAM's noise-model hierarchy) cannot be reproduced or checked without it - so the format here is this package's
// own: a line-oriented, self-describing text archive.  What is kept from the reference is the interface (a
// `template <class ARCHIVE> void serialize(ARCHIVE& ar, const unsigned int version)` member per class, `ar & NVP(member)`, default
// constructors for loading, per-class versions) and the member lists, in the reference's order and under its names.
//
//   gpslam_b200::archive 1
//   <name> <scalar>                         numbers (doubles as %.17g: exact round trip), bools as 0 / 1, strings percent-encoded
//   <name> [ <n> v0 v1 ... ]                arrays of doubles / keys
//   <name> {  ... }                         an object: its `version`, then its members
//   <name> @<id> {  ... }   /  <name> @<id>      a shared object (noise model, calibration): written once, referred to afterwards
//
// Reading checks every name against the one the class expects and throws std::runtime_error on the first mismatch, on a
// truncated archive and on an unknown factor type - an archive is never half-loaded silently.
#pragma once
#include <array>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <istream>
#include <map>
#include <memory>
#include <ostream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

namespace gpslam_b200 {

template <class T> struct NVP { const char* name; T& value; };
template <class T> NVP<T> make_nvp(const char* name, T& value) { return NVP<T>{name, value}; }
#define GPSLAM_B200_NVP(member) ::gpslam_b200::make_nvp(#member, member)

/// per-class archive version (boost's BOOST_CLASS_VERSION): specialise to bump
template <class T> struct ArchiveVersion { static constexpr unsigned value = 0; };

namespace detail {
inline std::string encodeToken(const std::string& s) {  // whitespace-free token: bytes outside '!'..'~' and '%' become %XX; "" becomes %
  if (s.empty()) return "%";
  std::string o;
  char b[4];
  for (unsigned char c : s) {
    if (c > 32 && c < 127 && c != '%') o.push_back(static_cast<char>(c));
    else { std::snprintf(b, sizeof b, "%%%02X", c); o += b; }
  }
  return o;
}
inline std::string decodeToken(const std::string& t) {
  if (t == "%") return "";
  std::string o;
  for (size_t i = 0; i < t.size(); i++) {
    if (t[i] != '%') { o.push_back(t[i]); continue; }
    if (i + 3 > t.size()) throw std::runtime_error("gpslam_b200 archive: bad escape in string token");
    o.push_back(static_cast<char>(std::strtol(t.substr(i + 1, 2).c_str(), nullptr, 16)));
    i += 2;
  }
  return o;
}
}  // namespace detail

/// output archive
class OArchive {
  std::ostream& os_;
  int depth_ = 0;
  std::map<const void*, int> ids_;
  void indent() { for (int k = 0; k < depth_; k++) os_ << "  "; }

 public:
  static constexpr bool is_saving = true, is_loading = false;
  explicit OArchive(std::ostream& os) : os_(os) { os_ << "gpslam_b200::archive 1\n"; }
  void scalar(const char* n, const double& v) { char b[40]; std::snprintf(b, sizeof b, "%.17g", v); indent(); os_ << n << ' ' << b << '\n'; }
  void scalar(const char* n, const int& v) { indent(); os_ << n << ' ' << v << '\n'; }
  void scalar(const char* n, const unsigned& v) { indent(); os_ << n << ' ' << v << '\n'; }
  void scalar(const char* n, const std::uint64_t& v) { indent(); os_ << n << ' ' << v << '\n'; }
  void scalar(const char* n, const bool& v) { indent(); os_ << n << ' ' << (v ? 1 : 0) << '\n'; }
  void scalar(const char* n, const std::string& v) { indent(); os_ << n << ' ' << detail::encodeToken(v) << '\n'; }
  /// array of n doubles; on loading the caller's resize(n) runs before the values are read
  template <class RESIZE> void array(const char* name, const double* p, size_t n, RESIZE) {
    indent(); os_ << name << " [ " << n;
    char b[40];
    for (size_t k = 0; k < n; k++) { std::snprintf(b, sizeof b, " %.17g", p[k]); os_ << b; }
    os_ << " ]\n";
  }
  template <class RESIZE> void keys(const char* name, const std::uint64_t* p, size_t n, RESIZE) {
    indent(); os_ << name << " [ " << n;
    for (size_t k = 0; k < n; k++) os_ << ' ' << p[k];
    os_ << " ]\n";
  }
  void begin(const char* n) { indent(); os_ << n << " {\n"; depth_++; }
  void end() { depth_--; indent(); os_ << "}\n"; }
  /// shared object: true = first occurrence, the caller writes the body between here and end(); false = a reference was written
  bool beginShared(const char* n, const void* p) {
    indent();
    if (!p) { os_ << n << " @0\n"; return false; }
    auto it = ids_.find(p);
    if (it != ids_.end()) { os_ << n << " @" << it->second << '\n'; return false; }
    const int id = static_cast<int>(ids_.size()) + 1;
    ids_[p] = id;
    os_ << n << " @" << id << " {\n"; depth_++;
    return true;
  }
};

/// input archive
class IArchive {
  std::istream& is_;
  std::map<int, std::shared_ptr<void>> objs_;
  std::string token() { std::string t; if (!(is_ >> t)) throw std::runtime_error("gpslam_b200 archive: truncated"); return t; }
  void expect(const char* n) { const std::string t = token(); if (t != n) throw std::runtime_error(std::string("gpslam_b200 archive: expected '") + n + "', found '" + t + "'"); }
  double number() { const std::string t = token(); char* e = nullptr; const double v = std::strtod(t.c_str(), &e); if (e == t.c_str() || *e) throw std::runtime_error("gpslam_b200 archive: '" + t + "' is not a number"); return v; }
  std::uint64_t unsignedNumber() { const std::string t = token(); char* e = nullptr; const unsigned long long v = std::strtoull(t.c_str(), &e, 10); if (e == t.c_str() || *e) throw std::runtime_error("gpslam_b200 archive: '" + t + "' is not an unsigned integer"); return v; }
  size_t count() { expect("["); const std::uint64_t n = unsignedNumber(); if (n > (std::uint64_t(1) << 40)) throw std::runtime_error("gpslam_b200 archive: implausible array length"); return static_cast<size_t>(n); }

 public:
  static constexpr bool is_saving = false, is_loading = true;
  explicit IArchive(std::istream& is) : is_(is) {
    expect("gpslam_b200::archive");
    if (unsignedNumber() != 1) throw std::runtime_error("gpslam_b200 archive: unsupported format version");
  }
  void scalar(const char* n, double& v) { expect(n); v = number(); }
  void scalar(const char* n, int& v) { expect(n); v = static_cast<int>(number()); }
  void scalar(const char* n, unsigned& v) { expect(n); v = static_cast<unsigned>(unsignedNumber()); }
  void scalar(const char* n, std::uint64_t& v) { expect(n); v = unsignedNumber(); }
  void scalar(const char* n, bool& v) { expect(n); v = unsignedNumber() != 0; }
  void scalar(const char* n, std::string& v) { expect(n); v = detail::decodeToken(token()); }
  template <class RESIZE> void array(const char* name, double*, size_t, RESIZE resize) {
    expect(name);
    const size_t n = count();
    double* p = resize(n);
    for (size_t k = 0; k < n; k++) p[k] = number();
    expect("]");
  }
  template <class RESIZE> void keys(const char* name, std::uint64_t*, size_t, RESIZE resize) {
    expect(name);
    const size_t n = count();
    std::uint64_t* p = resize(n);
    for (size_t k = 0; k < n; k++) p[k] = unsignedNumber();
    expect("]");
  }
  void begin(const char* n) { expect(n); expect("{"); }
  void end() { expect("}"); }
  /// shared object: returns its id (0 = null pointer) and whether a body follows (first occurrence)
  int beginShared(const char* n, bool& body) {
    expect(n);
    const std::string t = token();
    if (t.size() < 2 || t[0] != '@') throw std::runtime_error("gpslam_b200 archive: expected a shared-object id, found '" + t + "'");
    const int id = std::atoi(t.c_str() + 1);
    body = false;
    if (id == 0) return 0;
    if (objs_.count(id)) return id;
    expect("{");
    body = true;
    return id;
  }
  void remember(int id, const std::shared_ptr<void>& p) { objs_[id] = p; }
  std::shared_ptr<void> recall(int id) const { auto it = objs_.find(id); if (it == objs_.end()) throw std::runtime_error("gpslam_b200 archive: reference to an unknown shared object"); return it->second; }
};

// ---------------------------------------------------------------------------------- what `ar & nvp` does, per kind of member
namespace detail {
template <class T> struct IsArchive : std::false_type {};
template <> struct IsArchive<OArchive> : std::true_type {};
template <> struct IsArchive<IArchive> : std::true_type {};
template <class T> using Plain = typename std::remove_const<T>::type;
template <class AR, class T> auto hasSerialize(int) -> decltype(std::declval<T&>().serialize(std::declval<AR&>(), 0u), std::true_type{});
template <class AR, class T> std::false_type hasSerialize(...);
}  // namespace detail

// scalars
template <class AR, class T>
typename std::enable_if<std::is_arithmetic<T>::value || std::is_same<T, std::string>::value>::type archiveIO(AR& ar, const char* n, T& v) { ar.scalar(n, v); }
// fixed and dynamic arrays of doubles, key lists
template <class AR, size_t N> void archiveIO(AR& ar, const char* n, std::array<double, N>& v) {
  ar.array(n, v.data(), N, [&](size_t m) { if (m != N) throw std::runtime_error(std::string("gpslam_b200 archive: '") + n + "' has the wrong length"); return v.data(); });
}
template <class AR> void archiveIO(AR& ar, const char* n, std::vector<double>& v) { ar.array(n, v.data(), v.size(), [&](size_t m) { v.resize(m); return v.data(); }); }
template <class AR> void archiveIO(AR& ar, const char* n, std::vector<std::uint64_t>& v) { ar.keys(n, v.data(), v.size(), [&](size_t m) { v.resize(m); return v.data(); }); }
// any class with a serialize member: an object scope carrying the class version
template <class AR, class T>
typename std::enable_if<decltype(detail::hasSerialize<AR, T>(0))::value>::type archiveIO(AR& ar, const char* n, T& v) {
  ar.begin(n);
  unsigned version = ArchiveVersion<T>::value;
  ar.scalar("version", version);
  if (version > ArchiveVersion<T>::value) throw std::runtime_error(std::string("gpslam_b200 archive: '") + n + "' was written by a newer version of its class");
  v.serialize(ar, version);
  ar.end();
}
// shared_ptr to a class with a serialize member: written once per archive, pointer identity restored on loading
template <class T> void archiveIO(OArchive& ar, const char* n, std::shared_ptr<T>& p) {
  if (ar.beginShared(n, p.get())) {
    unsigned version = ArchiveVersion<T>::value;
    ar.scalar("version", version);
    p->serialize(ar, version);
    ar.end();
  }
}
template <class T> void archiveIO(IArchive& ar, const char* n, std::shared_ptr<T>& p) {
  bool body = false;
  const int id = ar.beginShared(n, body);
  if (id == 0) { p.reset(); return; }
  if (!body) { p = std::static_pointer_cast<T>(ar.recall(id)); return; }
  p = std::make_shared<T>();
  unsigned version = 0;
  ar.scalar("version", version);
  if (version > ArchiveVersion<T>::value) throw std::runtime_error(std::string("gpslam_b200 archive: '") + n + "' was written by a newer version of its class");
  p->serialize(ar, version);
  ar.end();
  ar.remember(id, p);
}

/// `ar & GPSLAM_B200_NVP(member)` for both archive kinds; saving never modifies the member (the const is cast away only to share one
/// serialize() between the two directions, as boost does)
template <class AR, class T>
typename std::enable_if<detail::IsArchive<AR>::value, AR&>::type operator&(AR& ar, const NVP<T>& p) {
  archiveIO(ar, p.name, const_cast<detail::Plain<T>&>(p.value));
  return ar;
}

}  // namespace gpslam_b200
