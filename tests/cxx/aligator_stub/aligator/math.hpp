// TEST INFRASTRUCTURE -- stand-ins for the few Eigen / aligator types the adapter
// (integration/aligator/gar/b200-riccati.hpp) touches, with the same member surface
// (.data() .size() .rows() .cols() .resize() .setZero(), implicit Ref conversions), so that the adapter
// compiles and runs in an image without Eigen.  Nothing here is part of the product.
#pragma once
#include <cstddef>
#include <stdexcept>
#include <string>
#include <vector>

namespace aligator {
using uint = unsigned int;
struct RuntimeError : std::runtime_error {
  using std::runtime_error::runtime_error;
};
namespace stub {
struct VectorXs {
  std::vector<double> v;
  VectorXs() = default;
  explicit VectorXs(long n) : v((size_t)n, 0.) {}
  double *data() { return v.data(); }
  const double *data() const { return v.data(); }
  long size() const { return (long)v.size(); }
  long rows() const { return (long)v.size(); }
  long cols() const { return 1; }
  void resize(long n) { v.assign((size_t)n, 0.); }
  void setZero() { std::fill(v.begin(), v.end(), 0.); }
  double &operator[](long i) { return v[(size_t)i]; }
  const double &operator[](long i) const { return v[(size_t)i]; }
};
template <bool RowMajor> struct Mat {
  std::vector<double> v;
  long r = 0, c = 0;
  Mat() = default;
  Mat(long r_, long c_) : v((size_t)(r_ * c_), 0.), r(r_), c(c_) {}
  double *data() { return v.data(); }
  const double *data() const { return v.data(); }
  long size() const { return r * c; }
  long rows() const { return r; }
  long cols() const { return c; }
  void resize(long r_, long c_) {
    r = r_;
    c = c_;
    v.assign((size_t)(r * c), 0.);
  }
  void setZero() { std::fill(v.begin(), v.end(), 0.); }
  double &operator()(long i, long j) { return v[(size_t)(RowMajor ? i * c + j : i + j * r)]; }
  const double &operator()(long i, long j) const { return v[(size_t)(RowMajor ? i * c + j : i + j * r)]; }
};
using MatrixXs = Mat<false>;
using RowMatrixXs = Mat<true>;
struct VectorRef { // Eigen::Ref<VectorXs>
  double *p;
  long n;
  VectorRef(VectorXs &x) : p(x.data()), n(x.size()) {}
  double *data() const { return p; }
  long size() const { return n; }
};
struct ConstVectorRef { // Eigen::Ref<const VectorXs>
  const double *p;
  long n;
  ConstVectorRef(const VectorXs &x) : p(x.data()), n(x.size()) {}
  const double *data() const { return p; }
  long size() const { return n; }
};
struct RowMatrixRef { // Eigen::Ref<RowMatrixXs>
  double *p;
  long r, c;
  RowMatrixRef(RowMatrixXs &m) : p(m.data()), r(m.rows()), c(m.cols()) {}
  double *data() const { return p; }
  long rows() const { return r; }
  long cols() const { return c; }
};
} // namespace stub
} // namespace aligator

#define ALIGATOR_DYNAMIC_TYPEDEFS_WITH_ROW_TYPES(Scalar)                                                     \
  using VectorXs = ::aligator::stub::VectorXs;                                                               \
  using MatrixXs = ::aligator::stub::MatrixXs;                                                               \
  using RowMatrixXs = ::aligator::stub::RowMatrixXs;                                                         \
  using VectorRef = ::aligator::stub::VectorRef;                                                             \
  using ConstVectorRef = ::aligator::stub::ConstVectorRef;                                                   \
  using RowMatrixRef = ::aligator::stub::RowMatrixRef
#define ALIGATOR_DYNAMIC_TYPEDEFS(Scalar) ALIGATOR_DYNAMIC_TYPEDEFS_WITH_ROW_TYPES(Scalar)
// (the reference formats with fmt; the stand-in keeps the format string)
#define ALIGATOR_RUNTIME_ERROR(...) throw ::aligator::RuntimeError(::aligator::stub_message(__VA_ARGS__))
namespace aligator {
template <class... A> inline std::string stub_message(const char *fmt, const A &...) { return std::string(fmt); }
inline std::string stub_message(const char *fmt, const char *arg) { return std::string(fmt) + " [" + arg + "]"; }
} // namespace aligator
