// TEST INFRASTRUCTURE -- every stand-in the adapter test needs, in one header: minimal matrix / vector types with the
// member surface the adapter uses, the knot / problem containers and the solver interface it implements.  The headers
// under aligator/ only forward here, so that `#include "aligator/gar/..."` in the adapter resolves without aligator.
#pragma once
#include <algorithm>
#include <optional>
// ---- matrix / vector stand-ins ----
// stand-ins for the few Eigen / aligator types the adapter
// (integration/aligator/gar/b200-riccati.hpp) touches, with the same member surface
// (.data() .size() .rows() .cols() .resize() .setZero(), implicit Ref conversions), so that the adapter
// compiles and runs in an image without Eigen.  Nothing here is part of the product.
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

// ---- knot and problem containers (field names and meaning of gar/lqr-problem.hpp:49-210) ----
namespace aligator {
namespace gar {
template <typename Scalar> struct LqrKnotTpl {
  ALIGATOR_DYNAMIC_TYPEDEFS(Scalar);
  uint nx = 0, nu = 0, nc = 0, nx2 = 0, nth = 0;
  MatrixXs Q, S, R;
  VectorXs q, r;
  MatrixXs A, B;
  VectorXs f;
  MatrixXs C, D;
  VectorXs d;
  MatrixXs Gth, Gx, Gu, Gv;
  VectorXs gamma;
  LqrKnotTpl() = default;
  LqrKnotTpl(uint nx_, uint nu_, uint nc_, uint nx2_, uint nth_ = 0)
      : nx(nx_), nu(nu_), nc(nc_), nx2(nx2_), nth(nth_), Q(nx_, nx_), S(nx_, nu_), R(nu_, nu_), q(nx_), r(nu_), A(nx2_, nx_),
        B(nx2_, nu_), f(nx2_), C(nc_, nx_), D(nc_, nu_), d(nc_), Gth(nth_, nth_), Gx(nx_, nth_), Gu(nu_, nth_), Gv(nc_, nth_),
        gamma(nth_) {}
};
template <typename Scalar> struct LqrProblemTpl {
  ALIGATOR_DYNAMIC_TYPEDEFS(Scalar);
  using KnotType = LqrKnotTpl<Scalar>;
  MatrixXs G0;
  VectorXs g0;
  std::vector<KnotType> stages;
  int horizon() const noexcept { return (int)stages.size() - 1; }
  uint nc0() const noexcept { return (uint)g0.rows(); }
  LqrProblemTpl(std::vector<KnotType> knots, long nc0_) : G0(nc0_, knots.empty() ? 0 : (long)knots[0].nx), g0(nc0_), stages(std::move(knots)) {}
};
} // namespace gar
} // namespace aligator

// ---- the interface the adapter implements (the six virtuals of gar/riccati-base.hpp:13-37) ----
namespace aligator {
namespace gar {
template <typename Scalar> struct LqrKnotTpl;
template <typename _Scalar> class RiccatiSolverBase {
public:
  using Scalar = _Scalar;
  using LqrKnot = LqrKnotTpl<Scalar>;
  ALIGATOR_DYNAMIC_TYPEDEFS_WITH_ROW_TYPES(Scalar);
  virtual bool backward(const Scalar mueq) = 0;
  virtual bool forward(std::vector<VectorXs> &xs, std::vector<VectorXs> &us, std::vector<VectorXs> &vs,
                       std::vector<VectorXs> &lbdas, const std::optional<ConstVectorRef> &theta_ = std::nullopt) const = 0;
  virtual void cycleAppend(const LqrKnot &knot) = 0;
  virtual void collapseFeedback() {}
  virtual VectorRef getFeedforward(size_t) = 0;
  virtual RowMatrixRef getFeedback(size_t) = 0;
  virtual ~RiccatiSolverBase() = default;
};
} // namespace gar
} // namespace aligator