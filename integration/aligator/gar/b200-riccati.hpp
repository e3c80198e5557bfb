/// @file b200-riccati.hpp
/// The reference-side binding of the B200 sweep: what a maintainer of aligator adds as
/// `include/aligator/gar/b200-riccati.hpp` (INTEGRATION.md sections 1-2).  A `gar::RiccatiSolverBase<double>`
/// (gar/riccati-base.hpp:13-37) whose six virtuals forward to the C ABI of libaligator_b200_gar.so
/// (include/aligator_b200/gar.h).  Complete, no elided bodies.  It touches Eigen objects only through
/// `.data() / .size() / .rows() / .cols() / .resize() / .setZero()` and the implicit `Ref` conversions, so
/// this repository -- which has no Eigen -- compiles and runs it against stand-in types with that surface
/// (tests/cxx/aligator_stub, tests/cxx/adapter_test.cpp); with the real headers nothing changes.
#pragma once

#include "aligator/gar/riccati-base.hpp"
#include "aligator/gar/lqr-problem.hpp"

#include <aligator_b200/gar.h>

#include <algorithm>
#include <optional>
#include <string>
#include <vector>

namespace aligator {
namespace gar {

/// Serial (`num_legs = 0`) or parallel-in-time (`num_legs >= 2`, the drop-in for ParallelRiccatiSolver,
/// gar/parallel-solver.hpp:21-113) Riccati solver on one B200.  Non-owning pointer to the problem, re-read
/// at every backward() like the reference (proximal-riccati.hpp:46).
template <typename _Scalar> class B200RiccatiSolver : public RiccatiSolverBase<_Scalar> {
public:
  using Scalar = _Scalar;
  static_assert(sizeof(Scalar) == sizeof(double), "the B200 path is fp64 (context.hpp:9)");
  using Base = RiccatiSolverBase<Scalar>;
  ALIGATOR_DYNAMIC_TYPEDEFS_WITH_ROW_TYPES(Scalar);
  using KnotType = LqrKnotTpl<Scalar>;
  using ProblemType = LqrProblemTpl<Scalar>;

  explicit B200RiccatiSolver(const ProblemType &problem, int num_legs = 0, int device = 0)
      : problem_(&problem), num_legs_(num_legs) {
    const int N = problem.horizon();
    if (N < 0)
      ALIGATOR_RUNTIME_ERROR("empty problem");
    const KnotType &kt = problem.stages[(size_t)N];
    if (kt.nu != 0)
      ALIGATOR_RUNTIME_ERROR("B200RiccatiSolver: the terminal knot must have nu = 0 (pad it as aligator_b200's "
                             "CudaRiccatiSolver does, DESIGN.md section 4)");
    dims_.nx = (int)kt.nx;
    dims_.nu = N > 0 ? (int)problem.stages[0].nu : 2;
    dims_.nc = N > 0 ? (int)problem.stages[0].nc : 0;
    dims_.nct = (int)kt.nc;
    dims_.nc0 = (int)problem.nc0();
    dims_.horizon = N;
    dims_.batch = 1;
    dims_.device = device;
    nth_ = N > 0 ? (int)problem.stages[0].nth : (int)kt.nth;
    if (num_legs_ >= 2)
      check(ab2_gar_create_parallel(&dims_, num_legs_, &h_)); // throws for < 2 like parallel-solver.hxx:42-46
    else if (num_legs_ != 0)
      ALIGATOR_RUNTIME_ERROR("numThreads ({}) should be greater than or equal to 2.", num_legs_);
    else if (nth_ > 0)
      check(ab2_gar_create_parametric(&dims_, nth_, &h_));
    else
      check(ab2_gar_create(&dims_, &h_));
    srec_ = num_legs_ >= 2 || nth_ == 0 ? ab2_gar_stage_record_doubles(dims_.nx, dims_.nu, dims_.nc)
                                         : ab2_gar_stage_record_doubles_th(dims_.nx, dims_.nu, dims_.nc, nth_);
    trec_ = num_legs_ >= 2 || nth_ == 0 ? ab2_gar_term_record_doubles(dims_.nx, dims_.nct)
                                         : ab2_gar_term_record_doubles_th(dims_.nx, dims_.nct, nth_);
    const size_t nr = (size_t)dims_.nu + dims_.nc + dims_.nx;
    stage_.assign((size_t)N * srec_, 0.);
    term_.assign(trec_, 0.);
    ffbuf_.assign((size_t)N * nr, 0.);
    fbbuf_.assign((size_t)N * nr * dims_.nx, 0.);
    ff_.resize((size_t)N + 1);
    fb_.resize((size_t)N + 1);
    for (int t = 0; t < N; ++t) {
      ff_[(size_t)t].resize((long)nr);
      fb_[(size_t)t].resize((long)nr, (long)dims_.nx);
    }
    ff_[(size_t)N].resize((long)dims_.nct); // terminal knot: [z] / [Z] only (riccati-kernel.hxx:146-149)
    fb_[(size_t)N].resize((long)dims_.nct, (long)dims_.nx);
  }
  ~B200RiccatiSolver() override { ab2_gar_destroy(h_); }
  B200RiccatiSolver(const B200RiccatiSolver &) = delete;
  B200RiccatiSolver &operator=(const B200RiccatiSolver &) = delete;

  /// riccati-base.hpp:19
  bool backward(const Scalar mueq) override {
    pack();
    check(ab2_gar_set_problem(h_, stage_.data(), term_.data(), problem_->G0.data(), problem_->g0.data(), AB2_HOST, nullptr));
    check(ab2_gar_backward(h_, mueq, nullptr));
    fetch_gains();
    int st = 0;
    check(ab2_gar_status(h_, &st, AB2_HOST, nullptr));
    check(ab2_gar_synchronize(h_, nullptr));
    if (st & 1)
      ALIGATOR_RUNTIME_ERROR("Failed stage LDL factorization"); // riccati-kernel.hxx:239-241
    return true;
  }

  /// riccati-base.hpp:21-24
  bool forward(std::vector<VectorXs> &xs, std::vector<VectorXs> &us, std::vector<VectorXs> &vs,
               std::vector<VectorXs> &lbdas, const std::optional<ConstVectorRef> &theta_ = std::nullopt) const override {
    const int N = dims_.horizon, nx = dims_.nx, nu = dims_.nu, nc = dims_.nc, nct = dims_.nct, nc0 = dims_.nc0;
    if (theta_.has_value() && nth_ > 0 && num_legs_ < 2) // the parallel solver ignores theta (parallel-solver.hxx:211)
      check(ab2_gar_forward_theta(h_, theta_->data(), AB2_HOST, nullptr));
    else
      check(ab2_gar_forward(h_, nullptr));
    traj_.resize((size_t)(N + 1) * nx + (size_t)N * (nu + nc + nx) + nct + nc0);
    double *X = traj_.data(), *U = X + (size_t)(N + 1) * nx, *V = U + (size_t)N * nu, *VT = V + (size_t)N * nc,
           *L0 = VT + nct, *L = L0 + nc0;
    auto get = [&](int what, double *dst, size_t n) {
      if (n)
        check(ab2_gar_get(h_, what, dst, AB2_HOST, nullptr));
    };
    get(AB2_OUT_XS, X, (size_t)(N + 1) * nx);
    get(AB2_OUT_US, U, (size_t)N * nu);
    get(AB2_OUT_VS, V, (size_t)N * nc);
    get(AB2_OUT_VST, VT, (size_t)nct);
    get(AB2_OUT_LBD0, L0, (size_t)nc0);
    get(AB2_OUT_LBDAS, L, (size_t)N * nx);
    check(ab2_gar_synchronize(h_, nullptr));
    for (int t = 0; t <= N; ++t)
      std::copy_n(X + (size_t)t * nx, nx, xs[(size_t)t].data());
    for (int t = 0; t < N; ++t) {
      std::copy_n(U + (size_t)t * nu, nu, us[(size_t)t].data());
      std::copy_n(V + (size_t)t * nc, nc, vs[(size_t)t].data());
      std::copy_n(L + (size_t)t * nx, nx, lbdas[(size_t)t + 1].data());
    }
    std::copy_n(VT, nct, vs[(size_t)N].data());
    std::copy_n(L0, nc0, lbdas[0].data());
    return true;
  }

  /// proximal-riccati.hxx:79-86 / parallel-solver.hxx:246-258
  void cycleAppend(const KnotType &knot) override {
    std::vector<double> rec(srec_, 0.);
    pack_stage(knot, rec.data());
    check(ab2_gar_cycle_append(h_, rec.data(), AB2_HOST, nullptr));
    fetch_gains();
    check(ab2_gar_synchronize(h_, nullptr));
  }

  /// parallel-solver.hpp:41-51 (no-op for the serial solver, riccati-base.hpp:32)
  void collapseFeedback() override {
    if (num_legs_ < 2)
      return;
    check(ab2_gar_collapse_feedback(h_, nullptr));
    fetch_gains();
    check(ab2_gar_synchronize(h_, nullptr));
  }

  VectorRef getFeedforward(size_t i) override { return ff_[i]; }
  RowMatrixRef getFeedback(size_t i) override { return fb_[i]; }

  /// datas[i].vm.Vxx (column-major nx x nx) -- what tests and bindings read (proximal-riccati.hpp:40-43)
  MatrixXs Vxx(size_t i) const {
    MatrixXs out;
    out.resize((long)dims_.nx, (long)dims_.nx);
    check(ab2_gar_get_range(h_, AB2_OUT_VXX, 0, 1, (int)i, 1, out.data(), AB2_HOST, nullptr));
    check(ab2_gar_synchronize(h_, nullptr));
    return out;
  }
  ab2_gar_solver *handle() const { return h_; }

private:
  static void check(int rc) {
    if (rc != AB2_OK)
      ALIGATOR_RUNTIME_ERROR("aligator_b200: {}", ab2_gar_last_error());
  }
  template <class M> static double *put(double *dst, const M &m, size_t n, const char *what) {
    if ((size_t)m.size() != n)
      ALIGATOR_RUNTIME_ERROR("knot field {} has the wrong size (uniform dims, nx2 = nx)", what);
    std::copy_n(m.data(), n, dst); // Eigen's default storage is column-major, like the record
    return dst + n;
  }
  /// LqrKnot -> [A | B | f | Q | S | R | q | r | C | D | d (| Gx | Gu | Gv | Gth | gamma)]: one memcpy per field
  void pack_stage(const KnotType &k, double *o) const {
    const size_t nx = (size_t)dims_.nx, nu = (size_t)dims_.nu, nc = (size_t)dims_.nc;
    if (k.nx != nx || k.nu != nu || k.nc != nc || k.nx2 != nx)
      ALIGATOR_RUNTIME_ERROR("stage knot dims differ from the solver's (uniform dims, nx2 = nx)");
    o = put(o, k.A, nx * nx, "A");
    o = put(o, k.B, nx * nu, "B");
    o = put(o, k.f, nx, "f");
    o = put(o, k.Q, nx * nx, "Q");
    o = put(o, k.S, nx * nu, "S");
    o = put(o, k.R, nu * nu, "R");
    o = put(o, k.q, nx, "q");
    o = put(o, k.r, nu, "r");
    o = put(o, k.C, nc * nx, "C");
    o = put(o, k.D, nc * nu, "D");
    o = put(o, k.d, nc, "d");
    if (nth_ > 0 && num_legs_ < 2) {
      const size_t nth = (size_t)nth_;
      o = put(o, k.Gx, nx * nth, "Gx");
      o = put(o, k.Gu, nu * nth, "Gu");
      o = put(o, k.Gv, nc * nth, "Gv");
      o = put(o, k.Gth, nth * nth, "Gth");
      o = put(o, k.gamma, nth, "gamma");
    }
  }
  void pack() {
    const int N = dims_.horizon;
    const size_t nx = (size_t)dims_.nx, nct = (size_t)dims_.nct;
    if (problem_->horizon() != N || (int)problem_->nc0() != dims_.nc0)
      ALIGATOR_RUNTIME_ERROR("the problem changed its horizon or nc0 since the solver was built");
    for (int t = 0; t < N; ++t)
      pack_stage(problem_->stages[(size_t)t], stage_.data() + (size_t)t * srec_);
    const KnotType &k = problem_->stages[(size_t)N];
    if (k.nx != nx || k.nu != 0 || k.nc != nct)
      ALIGATOR_RUNTIME_ERROR("terminal knot dims differ from the solver's");
    double *o = term_.data();
    o = put(o, k.Q, nx * nx, "Q");
    o = put(o, k.q, nx, "q");
    o = put(o, k.C, nct * nx, "C");
    o = put(o, k.d, nct, "d");
    if (nth_ > 0 && num_legs_ < 2) {
      const size_t nth = (size_t)nth_;
      o = put(o, k.Gx, nx * nth, "Gx");
      o = put(o, k.Gv, nct * nth, "Gv");
      o = put(o, k.Gth, nth * nth, "Gth");
      o = put(o, k.gamma, nth, "gamma");
    }
  }
  /// ff / fb of every knot into the Eigen objects getFeedforward / getFeedback hand out (fb is row-major on both sides)
  void fetch_gains() {
    const int N = dims_.horizon;
    const size_t nr = (size_t)dims_.nu + dims_.nc + dims_.nx, nx = (size_t)dims_.nx;
    if (N > 0) {
      check(ab2_gar_get(h_, AB2_OUT_FF, ffbuf_.data(), AB2_HOST, nullptr));
      check(ab2_gar_get(h_, AB2_OUT_FB, fbbuf_.data(), AB2_HOST, nullptr));
    }
    if (dims_.nct > 0) {
      check(ab2_gar_get(h_, AB2_OUT_FFT, ff_[(size_t)N].data(), AB2_HOST, nullptr));
      check(ab2_gar_get(h_, AB2_OUT_FBT, fb_[(size_t)N].data(), AB2_HOST, nullptr));
    }
    check(ab2_gar_synchronize(h_, nullptr));
    for (int t = 0; t < N; ++t) {
      std::copy_n(ffbuf_.data() + (size_t)t * nr, nr, ff_[(size_t)t].data());
      std::copy_n(fbbuf_.data() + (size_t)t * nr * nx, nr * nx, fb_[(size_t)t].data());
    }
  }

  const ProblemType *problem_;
  int num_legs_ = 0, nth_ = 0;
  ab2_gar_dims dims_{};
  ab2_gar_solver *h_ = nullptr;
  size_t srec_ = 0, trec_ = 0;
  std::vector<double> stage_, term_, ffbuf_, fbbuf_;
  mutable std::vector<double> traj_;
  std::vector<VectorXs> ff_;
  std::vector<RowMatrixXs> fb_;
};

} // namespace gar
} // namespace aligator
