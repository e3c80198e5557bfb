// TEST INFRASTRUCTURE -- stand-in for gar/lqr-problem.hpp:49-210 (same field names and meaning).
#pragma once
#include "aligator/math.hpp"
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
