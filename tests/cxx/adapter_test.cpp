// The reference-side adapter (integration/aligator/gar/b200-riccati.hpp) compiled against stand-in Eigen /
// aligator types (tests/cxx/aligator_stub) and run like the reference's tests/gar/riccati.cpp: build a problem,
// construct the solver through the RiccatiSolverBase interface, backward(mueq), forward(...), compare the gains
// and the trajectory with the CPU oracle.  Needs a GPU to run; compiling and linking it is a CPU test.
#include <cmath>
#include <cstdio>
#include <memory>
#include <random>

#include "aligator/gar/b200-riccati.hpp"
#include "../../oracle/gar_oracle.hpp"

namespace ag = aligator::gar;
namespace orc = gar_oracle;
using Knot = ag::LqrKnotTpl<double>;
using Problem = ag::LqrProblemTpl<double>;
using Vec = aligator::stub::VectorXs;

static Knot random_knot(std::mt19937 &rng, unsigned nx, unsigned nu, unsigned nc) {
  std::normal_distribution<double> nrm;
  std::uniform_real_distribution<double> uni(-1, 1);
  Knot k(nx, nu, nc, nx);
  const unsigned n = nx + nu;
  std::vector<double> W((size_t)n * (n + 1));
  for (auto &w : W) w = nrm(rng);
  auto H = [&](unsigned i, unsigned j) {
    double s = 0;
    for (unsigned c = 0; c <= n; ++c) s += W[i + (size_t)c * n] * W[j + (size_t)c * n];
    return s / std::max(nx, std::max(nu, 1u));
  };
  for (unsigned j = 0; j < nx; ++j)
    for (unsigned i = 0; i < nx; ++i) k.Q(i, j) = H(i, j);
  for (unsigned j = 0; j < nu; ++j) {
    for (unsigned i = 0; i < nx; ++i) k.S(i, j) = H(i, nx + j);
    for (unsigned i = 0; i < nu; ++i) k.R(i, j) = H(nx + i, nx + j) * (i == j ? 1 + 1e-6 : 1);
  }
  for (unsigned j = 0; j < nx; ++j)
    for (unsigned i = 0; i < nx; ++i) k.A(i, j) = (i == j) + 0.1 * nrm(rng) / std::sqrt((double)nx);
  for (unsigned j = 0; j < nu; ++j)
    for (unsigned i = 0; i < nx; ++i) k.B(i, j) = uni(rng);
  for (unsigned i = 0; i < nx; ++i) { k.f[i] = nrm(rng); k.q[i] = uni(rng); }
  for (unsigned i = 0; i < nu; ++i) k.r[i] = uni(rng);
  for (unsigned m = 0; m < nc && m < nu; ++m)
    if (uni(rng) > 0) { k.D(m, m) = 1.0; k.d[m] = uni(rng); }
  return k;
}

static orc::Problem to_oracle(const Problem &p) {
  orc::Problem o;
  for (const auto &k : p.stages) {
    orc::Knot q(k.nx, k.nu, k.nc, k.nx2, 0);
    q.Q = k.Q.v; q.S = k.S.v; q.R = k.R.v; q.q = k.q.v; q.r = k.r.v;
    q.A = k.A.v; q.B = k.B.v; q.f = k.f.v; q.C = k.C.v; q.D = k.D.v; q.d = k.d.v;
    o.stages.push_back(q);
  }
  o.nc0 = p.nc0();
  o.G0 = p.G0.v;
  o.g0 = p.g0.v;
  return o;
}

static double rel_fro(const double *a, const double *b, size_t n) {
  double num = 0, den = 0;
  for (size_t i = 0; i < n; ++i) { num += (a[i] - b[i]) * (a[i] - b[i]); den += b[i] * b[i]; }
  return den > 0 ? std::sqrt(num / den) : std::sqrt(num);
}

static int check(unsigned nx, unsigned nu, unsigned nc, int N, double mueq, int legs, unsigned seed) {
  std::mt19937 rng(seed);
  std::normal_distribution<double> nrm;
  std::vector<Knot> knots;
  for (int t = 0; t <= N; ++t) knots.push_back(random_knot(rng, nx, t < N ? nu : 0, t < N ? nc : 0));
  Problem prob(std::move(knots), nx);
  for (unsigned i = 0; i < nx; ++i) { prob.G0(i, i) = -1.0; prob.g0[i] = nrm(rng); }
  // through the base-class pointer, like SolverProxDDPTpl::linear_solver_ (solver-proxddp.hpp:181)
  std::unique_ptr<ag::RiccatiSolverBase<double>> solver = std::make_unique<ag::B200RiccatiSolver<double>>(prob, legs);
  if (!solver->backward(mueq)) return 1;
  std::vector<Vec> xs, us, vs, lbdas;
  for (int t = 0; t <= N; ++t) { xs.emplace_back(nx); vs.emplace_back(t < N ? nc : 0); lbdas.emplace_back(nx); }
  for (int t = 0; t < N; ++t) us.emplace_back(nu);
  if (!solver->forward(xs, us, vs, lbdas)) return 2;
  solver->collapseFeedback();

  orc::Problem op = to_oracle(prob);
  orc::ProximalRiccatiSolver ref(op);
  ref.backward(mueq);
  orc::Solution sol = orc::lqrInitializeSolution(op);
  ref.forward(sol);
  double worst = 0, wgain = 0;
  for (int t = 0; t <= N; ++t) {
    worst = std::max(worst, rel_fro(xs[t].data(), sol.xs[t].data(), nx));
    worst = std::max(worst, rel_fro(lbdas[t].data(), sol.lbdas[t].data(), nx));
    if (t < N) worst = std::max(worst, rel_fro(us[t].data(), sol.us[t].data(), nu));
  }
  if (legs == 0)
    for (int t = 0; t < N; ++t) {
      auto fb = solver->getFeedback((size_t)t);
      auto ff = solver->getFeedforward((size_t)t);
      wgain = std::max(wgain, rel_fro(fb.data(), ref.datas[t].fb.data(), (size_t)fb.rows() * fb.cols()));
      wgain = std::max(wgain, rel_fro(ff.data(), ref.datas[t].ff.data(), (size_t)ff.size()));
    }
  std::printf("adapter nx=%u nu=%u nc=%u N=%d legs=%d: trajectory %.2e gains %.2e (rel-Frobenius vs oracle)\n", nx, nu, nc, N,
              legs, worst, wgain);
  const double tol = legs ? 1e-7 : 1e-10; // tests/gar/parallel.cpp:193 for the parallel solver
  return (worst <= tol && wgain <= 1e-10) ? 0 : 3;
}

int main() {
  int rc = 0;
  rc |= check(12, 6, 0, 50, 1e-8, 0, 1);
  rc |= check(4, 2, 2, 20, 1e-3, 0, 2);
  rc |= check(14, 7, 0, 60, 1e-9, 4, 3);
  try { // num_threads = 1 throws like parallel-solver.hxx:42-46
    std::mt19937 rng(4);
    std::vector<Knot> knots{random_knot(rng, 4, 2, 0), random_knot(rng, 4, 0, 0)};
    Problem prob(std::move(knots), 4);
    ag::B200RiccatiSolver<double> bad(prob, 1);
    rc |= 16;
  } catch (const aligator::RuntimeError &e) {
    std::printf("num_legs = 1 -> RuntimeError: %s\n", e.what());
  }
  std::printf(rc == 0 ? "ADAPTER OK\n" : "ADAPTER FAILED rc=%d\n", rc);
  return rc;
}
