// aligator_b200/riccati_solver.hpp -- C++ host side above the C ABI (gar.h).
//
// Mirrors the reference's operator interface for this path, name for name:
//   aligator::gar::LqrKnotTpl<double>        gar/lqr-problem.hpp:49-118
//   aligator::gar::LqrProblemTpl<double>     gar/lqr-problem.hpp:120-210
//   aligator::gar::RiccatiSolverBase<double> gar/riccati-base.hpp:13-37
//   aligator::gar::ProximalRiccatiSolver     gar/proximal-riccati.hpp:12-47
// so that a caller written against the reference (bench/gar-riccati.cpp:42-50,
// SolverProxDDPTpl::innerLoop, solver-proxddp.hxx:605-632) reads the same.
//
// Header-only, C++17, no Eigen (Eigen is not available in this build image); the
// Eigen-typed adapter a maintainer would use inside aligator is shown in
// INTEGRATION.md.  Matrices are column-major std::vector<double>, fb is row-major,
// exactly the reference's storage orders.  Errors: hard failures throw
// aligator_b200::RuntimeError like the reference throws aligator::RuntimeError
// (utils/exceptions.hpp:8-10; "Failed stage LDL factorization",
// riccati-kernel.hxx:239-241).  No CPU fallback exists.
#pragma once

#include <algorithm>
#include <cstddef>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "gar.h"

namespace aligator_b200 {

struct RuntimeError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

namespace gar {

using uint = unsigned int;
using VectorXs = std::vector<double>;

/// One stage of the LQ problem; zero-initialised like lqr-problem.hxx:29-72.
struct LqrKnot {
  uint nx = 0, nu = 0, nc = 0, nx2 = 0, nth = 0;
  std::vector<double> Q, S, R, q, r; // Q nx*nx, S nx*nu, R nu*nu (column-major)
  std::vector<double> A, B, f;       // A nx2*nx, B nx2*nu
  std::vector<double> C, D, d;       // C nc*nx, D nc*nu
  LqrKnot() = default;
  LqrKnot(uint nx_, uint nu_, uint nc_, uint nx2_, uint nth_ = 0)
      : nx(nx_), nu(nu_), nc(nc_), nx2(nx2_), nth(nth_), Q((size_t)nx_ * nx_), S((size_t)nx_ * nu_),
        R((size_t)nu_ * nu_), q(nx_), r(nu_), A((size_t)nx2_ * nx_), B((size_t)nx2_ * nu_), f(nx2_),
        C((size_t)nc_ * nx_), D((size_t)nc_ * nu_), d(nc_) {
    if (nth_ != 0)
      throw RuntimeError("parameterised knots (nth > 0) are not supported by the CUDA path");
  }
  LqrKnot(uint nx_, uint nu_, uint nc_) : LqrKnot(nx_, nu_, nc_, nx_, 0) {}
};

struct LqrProblem {
  std::vector<LqrKnot> stages;
  std::vector<double> G0; // nc0 x nx0 column-major
  std::vector<double> g0;
  LqrProblem() = default;
  LqrProblem(std::vector<LqrKnot> knots, long nc0)
      : stages(std::move(knots)), G0((size_t)nc0 * (stages.empty() ? 0 : stages[0].nx)), g0((size_t)nc0) {}
  int horizon() const noexcept { return (int)stages.size() - 1; }
  uint nc0() const noexcept { return (uint)g0.size(); }
};

/// Row-major matrix view (the reference's RowMatrixRef).
struct RowMatrixRef {
  double *data;
  int rows, cols;
  double &operator()(int i, int j) const { return data[(size_t)i * cols + j]; }
};
struct VectorRef {
  double *data;
  int size;
  double &operator[](int i) const { return data[i]; }
};

/// gar/riccati-base.hpp:13-37, same six virtuals.
class RiccatiSolverBase {
public:
  virtual bool backward(const double mueq) = 0;
  virtual bool forward(std::vector<VectorXs> &xs, std::vector<VectorXs> &us, std::vector<VectorXs> &vs,
                       std::vector<VectorXs> &lbdas,
                       const std::optional<VectorXs> &theta = std::nullopt) const = 0;
  virtual void cycleAppend(const LqrKnot &knot) = 0;
  virtual void collapseFeedback() {}
  virtual VectorRef getFeedforward(size_t i) = 0;
  virtual RowMatrixRef getFeedback(size_t i) = 0;
  virtual ~RiccatiSolverBase() = default;
};

/// gar/utils.hpp:114-142
inline void lqrInitializeSolution(const LqrProblem &p, std::vector<VectorXs> &xs, std::vector<VectorXs> &us,
                                  std::vector<VectorXs> &vs, std::vector<VectorXs> &lbdas) {
  const int N = p.horizon();
  xs.assign(N + 1, {});
  us.assign(N + 1, {});
  vs.assign(N + 1, {});
  lbdas.assign(N + 1, {});
  lbdas[0].assign(p.nc0(), 0.);
  for (int i = 0; i <= N; ++i) {
    const LqrKnot &k = p.stages[i];
    xs[i].assign(k.nx, 0.);
    us[i].assign(k.nu, 0.);
    vs[i].assign(k.nc, 0.);
    if (i == N)
      break;
    lbdas[i + 1].assign(k.nx2, 0.);
  }
  if (p.stages.back().nu == 0)
    us.pop_back();
}

/// Page-locked host buffer of doubles (ab2_gar_pinned_alloc): copies to / from it are truly
/// asynchronous, so backward() overlaps its uploads and downloads on the stream.
class PinnedBuf {
public:
  PinnedBuf() = default;
  PinnedBuf(const PinnedBuf &) = delete;
  PinnedBuf &operator=(const PinnedBuf &) = delete;
  ~PinnedBuf() { ab2_gar_pinned_free(p_); }
  void assign(size_t n, double v) {
    if (n != n_) {
      ab2_gar_pinned_free(p_);
      p_ = nullptr;
      n_ = 0;
      if (n) {
        void *q = nullptr;
        if (ab2_gar_pinned_alloc(n * sizeof(double), &q) != AB2_OK)
          throw RuntimeError(std::string("aligator_b200: ") + ab2_gar_last_error());
        p_ = static_cast<double *>(q);
        n_ = n;
      }
    }
    for (size_t i = 0; i < n_; ++i)
      p_[i] = v;
  }
  void resize(size_t n) {
    if (n != n_)
      assign(n, 0.0);
  }
  double *data() const { return p_; }
  double *begin() const { return p_; }
  size_t size() const { return n_; }
  bool empty() const { return n_ == 0; }

private:
  double *p_ = nullptr;
  size_t n_ = 0;
};

/// Drop-in for gar::ProximalRiccatiSolver backed by the B200 sweep.  Constructed from
/// one problem (batch = 1, the reference's use) or from `batch` problems of identical
/// dimensions.  Keeps NON-owning pointers to the problems and re-reads them at every
/// backward(), as the reference does (proximal-riccati.hpp:46).
class CudaRiccatiSolver : public RiccatiSolverBase {
public:
  explicit CudaRiccatiSolver(const LqrProblem &problem, int device = 0)
      : CudaRiccatiSolver(std::vector<const LqrProblem *>{&problem}, device) {}

  explicit CudaRiccatiSolver(std::vector<const LqrProblem *> problems, int device = 0)
      : CudaRiccatiSolver(std::move(problems), device, 0) {}

protected:
  /// num_legs >= 2: the parallel-in-time solver (ab2_gar_create_parallel)
  CudaRiccatiSolver(std::vector<const LqrProblem *> problems, int device, int num_legs)
      : problems_(std::move(problems)) {
    if (problems_.empty() || problems_[0]->stages.empty())
      throw RuntimeError("empty problem");
    const LqrProblem &p0 = *problems_[0];
    int N = p0.horizon();
    const LqrKnot &kt = p0.stages[N];
    // A terminal knot WITH controls (terminalSolve's nu > 0 branch, riccati-kernel.hxx:150-173) is
    // solved as one more stage knot with A = B = f = 0 before a null terminal knot: the stage step
    // from the zero value function is that branch exactly (tests/test_terminal_controls.py).
    term_controls_ = kt.nu != 0;
    if (term_controls_ && num_legs)
      throw RuntimeError("the parallel solver needs a terminal knot with nu = 0");
    if (term_controls_)
      ++N; // the library's horizon: the caller's knots 0..N are its stage knots
    // Stage knots of different (nu, nc) (gar/lqr-problem.hpp:49-118 lets every knot have its own) are
    // padded to the largest with decoupled controls (R = I on their diagonal, zero S / B / r columns)
    // and null constraint rows: the KKT matrix is block diagonal with the caller's block first, so its
    // factorisation is the reference's; the padding rows of the gains are exact zeros and are dropped.
    size_t numax = 0, ncmax = 0;
    for (int t = 0; t < N; ++t) {
      knot_dims_.emplace_back((int)p0.stages[t].nu, (int)p0.stages[t].nc);
      numax = std::max<size_t>(numax, p0.stages[t].nu);
      ncmax = std::max<size_t>(ncmax, p0.stages[t].nc);
      ragged_ = ragged_ || knot_dims_[t] != knot_dims_[0];
    }
    if (ragged_ && num_legs)
      throw RuntimeError("the parallel solver needs uniform stage dims");
    dims_.nx = (int)kt.nx;
    dims_.nu = N > 0 ? (int)numax : 2;
    dims_.nc = N > 0 ? (int)ncmax : 0;
    dims_.nct = term_controls_ ? 0 : (int)kt.nc;
    dims_.nc0 = (int)p0.nc0();
    dims_.horizon = N;
    dims_.batch = (int)problems_.size();
    dims_.device = device;
    check(num_legs ? ab2_gar_create_parallel(&dims_, num_legs, &h_) : ab2_gar_create(&dims_, &h_));
    srec_ = ab2_gar_stage_record_doubles(dims_.nx, dims_.nu, dims_.nc);
    trec_ = ab2_gar_term_record_doubles(dims_.nx, dims_.nct);
    const int nr = dims_.nu + dims_.nc + dims_.nx;
    stage_.assign((size_t)dims_.batch * N * srec_, 0.);
    term_.assign((size_t)dims_.batch * trec_, 0.);
    G0_.assign((size_t)dims_.batch * dims_.nc0 * dims_.nx, 0.);
    g0_.assign((size_t)dims_.batch * dims_.nc0, 0.);
    ff_.assign((size_t)dims_.batch * N * nr, 0.);
    fb_.assign((size_t)dims_.batch * N * nr * dims_.nx, 0.);
    ffT_.assign((size_t)dims_.batch * dims_.nct, 0.);
    fbT_.assign((size_t)dims_.batch * dims_.nct * dims_.nx, 0.);
  }

public:
  ~CudaRiccatiSolver() override { ab2_gar_destroy(h_); }
  CudaRiccatiSolver(const CudaRiccatiSolver &) = delete;
  CudaRiccatiSolver &operator=(const CudaRiccatiSolver &) = delete;

  /// riccati-base.hpp:19
  bool backward(const double mueq) override {
    pack();
    check(ab2_gar_set_problem(h_, stage_.data(), term_.data(), G0_.data(), g0_.data(), AB2_HOST, nullptr));
    check(ab2_gar_backward(h_, mueq, nullptr));
    std::vector<int> st(dims_.batch);
    check(ab2_gar_status(h_, st.data(), AB2_HOST, nullptr));
    check(ab2_gar_get(h_, AB2_OUT_FF, ff_.data(), AB2_HOST, nullptr));
    check(ab2_gar_get(h_, AB2_OUT_FB, fb_.data(), AB2_HOST, nullptr));
    if (dims_.nct > 0) {
      check(ab2_gar_get(h_, AB2_OUT_FFT, ffT_.data(), AB2_HOST, nullptr));
      check(ab2_gar_get(h_, AB2_OUT_FBT, fbT_.data(), AB2_HOST, nullptr));
    }
    check(ab2_gar_synchronize(h_, nullptr));
    compact_gains();
    for (int b = 0; b < dims_.batch; ++b)
      if (st[b] & 1)
        throw RuntimeError("Failed stage LDL factorization (instance " + std::to_string(b) + ")");
    return true;
  }

  /// riccati-base.hpp:21-24 (batch = 1)
  bool forward(std::vector<VectorXs> &xs, std::vector<VectorXs> &us, std::vector<VectorXs> &vs,
               std::vector<VectorXs> &lbdas, const std::optional<VectorXs> &theta = std::nullopt) const override {
    if (theta.has_value())
      throw RuntimeError("theta is not supported (nth = 0 only)");
    run_forward();
    scatter(0, xs, us, vs, lbdas);
    return true;
  }
  /// batched variant: one solution set per problem
  bool forward(std::vector<std::vector<VectorXs>> &xs, std::vector<std::vector<VectorXs>> &us,
               std::vector<std::vector<VectorXs>> &vs, std::vector<std::vector<VectorXs>> &lbdas) const {
    run_forward();
    for (int b = 0; b < dims_.batch; ++b)
      scatter(b, xs[b], us[b], vs[b], lbdas[b]);
    return true;
  }

  /// proximal-riccati.hxx:79-86 (same knot appended to every instance)
  void cycleAppend(const LqrKnot &knot) override {
    if (ragged_ || term_controls_)
      throw RuntimeError("cycleAppend needs uniform stage dims and a terminal knot with nu = 0");
    std::vector<double> rec((size_t)dims_.batch * srec_, 0.);
    for (int b = 0; b < dims_.batch; ++b)
      pack_stage(knot, rec.data() + (size_t)b * srec_);
    check(ab2_gar_cycle_append(h_, rec.data(), AB2_HOST, nullptr));
    check(ab2_gar_get(h_, AB2_OUT_FF, ff_.data(), AB2_HOST, nullptr));
    check(ab2_gar_get(h_, AB2_OUT_FB, fb_.data(), AB2_HOST, nullptr));
    check(ab2_gar_synchronize(h_, nullptr));
  }

  VectorRef getFeedforward(size_t i) override { return getFeedforward(i, 0); }
  RowMatrixRef getFeedback(size_t i) override { return getFeedback(i, 0); }
  VectorRef getFeedforward(size_t i, int b) {
    const int N = dims_.horizon, nr = dims_.nu + dims_.nc + dims_.nx;
    if ((int)i == N)
      return {ffT_.data() + (size_t)b * dims_.nct, dims_.nct};
    if (ragged_) // compacted in place by compact_gains(): [k; z; a] of the caller's dims
      return {ff_.data() + ((size_t)b * N + i) * nr, knot_dims_[i].first + knot_dims_[i].second + dims_.nx};
    return {ff_.data() + ((size_t)b * N + i) * nr, nr};
  }
  RowMatrixRef getFeedback(size_t i, int b) {
    const int N = dims_.horizon, nr = dims_.nu + dims_.nc + dims_.nx;
    if ((int)i == N)
      return {fbT_.data() + (size_t)b * dims_.nct * dims_.nx, dims_.nct, dims_.nx};
    if (ragged_)
      return {fb_.data() + ((size_t)b * N + i) * nr * dims_.nx, knot_dims_[i].first + knot_dims_[i].second + dims_.nx,
              dims_.nx};
    return {fb_.data() + ((size_t)b * N + i) * nr * dims_.nx, nr, dims_.nx};
  }
  /// datas[i].vm.Vxx (column-major nx*nx) / vm.vx, fetched on demand
  std::vector<double> Vxx(size_t i, int b = 0) const {
    std::vector<double> out((size_t)dims_.nx * dims_.nx);
    check(ab2_gar_get_range(h_, AB2_OUT_VXX, b, 1, (int)i, 1, out.data(), AB2_HOST, nullptr));
    check(ab2_gar_synchronize(h_, nullptr));
    return out;
  }
  std::vector<double> vx(size_t i, int b = 0) const {
    std::vector<double> out((size_t)dims_.nx);
    check(ab2_gar_get_range(h_, AB2_OUT_VX, b, 1, (int)i, 1, out.data(), AB2_HOST, nullptr));
    check(ab2_gar_synchronize(h_, nullptr));
    return out;
  }
  ab2_gar_solver *handle() const { return h_; }
  const ab2_gar_dims &dims() const { return dims_; }

private:
  static void check(int rc) {
    if (rc != AB2_OK)
      throw RuntimeError(std::string("aligator_b200: ") + ab2_gar_last_error());
  }
  template <class Vec> static double *put(double *dst, const Vec &src, size_t n, const char *name) {
    if (src.size() != n)
      throw RuntimeError(std::string("knot field has the wrong size: ") + name);
    for (size_t i = 0; i < n; ++i)
      dst[i] = src[i];
    return dst + n;
  }
  /// column-major rows x cols block `src` into a block of leading dimension ld (>= rows); the rest stays as it is
  template <class Vec>
  static void put_block(double *dst, size_t ld, const Vec &src, size_t rows, size_t cols, const char *name) {
    if (src.size() != rows * cols)
      throw RuntimeError(std::string("knot field has the wrong size: ") + name);
    for (size_t j = 0; j < cols; ++j)
      for (size_t i = 0; i < rows; ++i)
        dst[i + j * ld] = src[i + j * rows];
  }
  void pack_stage(const LqrKnot &k, double *o, bool last_with_controls = false) const {
    const size_t nx = dims_.nx, nu = dims_.nu, nc = dims_.nc;
    if (k.nx != nx || k.nu > nu || k.nc > nc || (k.nx2 != nx && !last_with_controls) || k.nth != 0)
      throw RuntimeError("stage knot dims differ from the solver's (nx2 = nx, nth = 0)");
    for (size_t i = 0; i < srec_; ++i)
      o[i] = 0.;
    double *A = o, *B = A + nx * nx, *f = B + nx * nu, *Q = f + nx, *S = Q + nx * nx, *R = S + nx * nu,
           *q = R + nu * nu, *r = q + nx, *C = r + nu, *D = C + nc * nx, *d = D + nc * nu;
    if (!last_with_controls) { // (a terminal knot with controls has no successor: A = B = f = 0)
      put_block(A, nx, k.A, nx, nx, "A");
      put_block(B, nx, k.B, nx, k.nu, "B");
      put_block(f, nx, k.f, nx, 1, "f");
    }
    put_block(Q, nx, k.Q, nx, nx, "Q");
    put_block(S, nx, k.S, nx, k.nu, "S");
    put_block(R, nu, k.R, k.nu, k.nu, "R");
    for (size_t c = k.nu; c < nu; ++c)
      R[c + c * nu] = 1.0; // padding controls: decoupled, gain rows exactly zero
    put_block(q, nx, k.q, nx, 1, "q");
    put_block(r, nu, k.r, k.nu, 1, "r");
    put_block(C, nc, k.C, k.nc, nx, "C");
    put_block(D, nc, k.D, k.nc, k.nu, "D");
    put_block(d, nc, k.d, k.nc, 1, "d");
  }
  /// ragged problems: drop the padding rows of every knot's [k; z; a] / [K; Z; Ahat] in the staging buffers
  void compact_gains() {
    if (!ragged_)
      return;
    const int N = dims_.horizon, nx = dims_.nx, nu = dims_.nu, nc = dims_.nc, nr = nu + nc + nx;
    for (int b = 0; b < dims_.batch; ++b)
      for (int t = 0; t < N; ++t) {
        const int nut = knot_dims_[t].first, nct = knot_dims_[t].second;
        double *ff = ff_.data() + ((size_t)b * N + t) * nr, *fb = fb_.data() + ((size_t)b * N + t) * nr * nx;
        int dst = nut;
        auto move = [&](int src0, int n) {
          for (int i = 0; i < n; ++i, ++dst) {
            ff[dst] = ff[src0 + i];
            for (int j = 0; j < nx; ++j)
              fb[(size_t)dst * nx + j] = fb[(size_t)(src0 + i) * nx + j];
          }
        };
        move(nu, nct);
        move(nu + nc, nx);
      }
  }
  void pack() {
    const int N = dims_.horizon;
    const size_t nx = dims_.nx, nct = dims_.nct;
    for (int b = 0; b < dims_.batch; ++b) {
      const LqrProblem &p = *problems_[b];
      if ((int)p.horizon() + (term_controls_ ? 1 : 0) != N || (int)p.nc0() != dims_.nc0)
        throw RuntimeError("problems of a batch must share horizon and nc0");
      for (int t = 0; t < N; ++t) {
        if ((int)p.stages[t].nu != knot_dims_[t].first || (int)p.stages[t].nc != knot_dims_[t].second)
          throw RuntimeError("problems of a batch must share the per-knot dims");
        pack_stage(p.stages[t], stage_.data() + ((size_t)b * N + t) * srec_, term_controls_ && t == N - 1);
      }
      if (!term_controls_) { // (with terminal controls the library's terminal knot is null: term_ stays zero)
        const LqrKnot &k = p.stages[N];
        if (k.nx != nx || k.nu != 0 || k.nc != nct)
          throw RuntimeError("terminal knot dims differ from the solver's");
        double *o = term_.data() + (size_t)b * trec_;
        o = put(o, k.Q, nx * nx, "Q");
        o = put(o, k.q, nx, "q");
        o = put(o, k.C, nct * nx, "C");
        o = put(o, k.d, nct, "d");
      }
      put(G0_.data() + (size_t)b * dims_.nc0 * nx, p.G0, (size_t)dims_.nc0 * nx, "G0");
      put(g0_.data() + (size_t)b * dims_.nc0, p.g0, (size_t)dims_.nc0, "g0");
    }
  }
  void run_forward() const {
    check(ab2_gar_forward(h_, nullptr));
    const int N = dims_.horizon, B = dims_.batch;
    xs_.resize((size_t)B * (N + 1) * dims_.nx);
    us_.resize((size_t)B * N * dims_.nu);
    vs_.resize((size_t)B * N * dims_.nc);
    vsT_.resize((size_t)B * dims_.nct);
    l0_.resize((size_t)B * dims_.nc0);
    ls_.resize((size_t)B * N * dims_.nx);
    auto get = [&](int what, PinnedBuf &v) {
      if (!v.empty())
        check(ab2_gar_get(h_, what, v.data(), AB2_HOST, nullptr));
    };
    get(AB2_OUT_XS, xs_);
    get(AB2_OUT_US, us_);
    get(AB2_OUT_VS, vs_);
    get(AB2_OUT_VST, vsT_);
    get(AB2_OUT_LBD0, l0_);
    get(AB2_OUT_LBDAS, ls_);
    check(ab2_gar_synchronize(h_, nullptr));
  }
  void scatter(int b, std::vector<VectorXs> &xs, std::vector<VectorXs> &us, std::vector<VectorXs> &vs,
               std::vector<VectorXs> &lbdas) const {
    const int N = dims_.horizon, nx = dims_.nx, nu = dims_.nu, nc = dims_.nc;
    if (term_controls_) { // the caller's horizon is N-1: N states, N controls, N multipliers; the null knot is dropped
      for (int t = 0; t < N; ++t) {
        xs[t].assign(xs_.begin() + ((size_t)b * (N + 1) + t) * nx, xs_.begin() + ((size_t)b * (N + 1) + t + 1) * nx);
        us[t].assign(us_.begin() + ((size_t)b * N + t) * nu, us_.begin() + ((size_t)b * N + t) * nu + knot_dims_[t].first);
        vs[t].assign(vs_.begin() + ((size_t)b * N + t) * nc, vs_.begin() + ((size_t)b * N + t) * nc + knot_dims_[t].second);
        if (t + 1 < N)
          lbdas[t + 1].assign(ls_.begin() + ((size_t)b * N + t) * nx, ls_.begin() + ((size_t)b * N + t + 1) * nx);
      }
      lbdas[0].assign(l0_.begin() + (size_t)b * dims_.nc0, l0_.begin() + (size_t)(b + 1) * dims_.nc0);
      return;
    }
    for (int t = 0; t <= N; ++t)
      xs[t].assign(xs_.begin() + ((size_t)b * (N + 1) + t) * nx, xs_.begin() + ((size_t)b * (N + 1) + t + 1) * nx);
    for (int t = 0; t < N; ++t) {
      us[t].assign(us_.begin() + ((size_t)b * N + t) * nu, us_.begin() + ((size_t)b * N + t) * nu + knot_dims_[t].first);
      vs[t].assign(vs_.begin() + ((size_t)b * N + t) * nc, vs_.begin() + ((size_t)b * N + t) * nc + knot_dims_[t].second);
      lbdas[t + 1].assign(ls_.begin() + ((size_t)b * N + t) * nx, ls_.begin() + ((size_t)b * N + t + 1) * nx);
    }
    vs[N].assign(vsT_.begin() + (size_t)b * dims_.nct, vsT_.begin() + (size_t)(b + 1) * dims_.nct);
    lbdas[0].assign(l0_.begin() + (size_t)b * dims_.nc0, l0_.begin() + (size_t)(b + 1) * dims_.nc0);
  }

  std::vector<const LqrProblem *> problems_;
  ab2_gar_dims dims_{};
  ab2_gar_solver *h_ = nullptr;
  size_t srec_ = 0, trec_ = 0;
  bool term_controls_ = false, ragged_ = false;
  std::vector<std::pair<int, int>> knot_dims_; // (nu, nc) of the caller's stage knots
  // pinned staging: uploads and downloads are asynchronous on the stream, one synchronisation per call
  PinnedBuf stage_, term_, G0_, g0_;
  PinnedBuf ff_, fb_, ffT_, fbT_;
  mutable PinnedBuf xs_, us_, vs_, vsT_, l0_, ls_;

protected:
  void refetch_gains() {
    check(ab2_gar_get(h_, AB2_OUT_FF, ff_.data(), AB2_HOST, nullptr));
    check(ab2_gar_get(h_, AB2_OUT_FB, fb_.data(), AB2_HOST, nullptr));
    check(ab2_gar_synchronize(h_, nullptr));
    compact_gains();
  }
};

/// Drop-in for gar::ParallelRiccatiSolver (gar/parallel-solver.hpp:21-113): same constructor
/// (problem, num_threads), same six virtuals.  The legs of the horizon run as work items of one
/// launch on the device, the condensed block-tridiagonal system is solved there too.  Throws like
/// the reference for num_threads < 2 (parallel-solver.hxx:42-46).  Unlike the reference it does
/// not re-parameterise the caller's problem in place.
class CudaParallelRiccatiSolver : public CudaRiccatiSolver {
public:
  CudaParallelRiccatiSolver(const LqrProblem &problem, const uint num_threads, int device = 0)
      : CudaRiccatiSolver(std::vector<const LqrProblem *>{&problem}, device, check_threads(num_threads)),
        numThreads_(num_threads) {}
  CudaParallelRiccatiSolver(std::vector<const LqrProblem *> problems, const uint num_threads, int device = 0)
      : CudaRiccatiSolver(std::move(problems), device, check_threads(num_threads)), numThreads_(num_threads) {}
  uint getNumThreads() const noexcept { return numThreads_; }
  /// parallel-solver.hpp:41-51
  void collapseFeedback() override {
    if (ab2_gar_collapse_feedback(handle(), nullptr) != AB2_OK)
      throw RuntimeError(std::string("aligator_b200: ") + ab2_gar_last_error());
    refetch_gains();
  }

private:
  static int check_threads(uint n) {
    if (n < 2)
      throw RuntimeError("(CudaParallelRiccatiSolver) numThreads (" + std::to_string(n) +
                         ") should be greater than or equal to 2.");
    return (int)n;
  }
  uint numThreads_;
};

} // namespace gar
} // namespace aligator_b200
