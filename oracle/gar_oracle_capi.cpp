// gar_oracle_capi.cpp -- extern "C" surface of the CPU oracle, for ctypes.
// TEST INFRASTRUCTURE ONLY (see gar_oracle.hpp header).  Build: oracle/Makefile.
#include "gar_oracle.hpp"
#include "gar_oracle_dense.hpp"

#include <chrono>
#include <memory>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace gar_oracle;

namespace {

// Flat knot record (doubles), in this order:
//   Q S R q r | A B f | C D d | Gth Gx Gu Gv gamma        (all column-major)
size_t knot_size(uint nx, uint nu, uint nc, uint nx2, uint nth) {
  return (size_t)nx * nx + (size_t)nx * nu + (size_t)nu * nu + nx + nu +
         (size_t)nx2 * nx + (size_t)nx2 * nu + nx2 + (size_t)nc * nx +
         (size_t)nc * nu + nc + (size_t)nth * nth + (size_t)nx * nth +
         (size_t)nu * nth + (size_t)nc * nth + nth;
}

const double *take(vecd &dst, const double *src) {
  std::copy(src, src + dst.size(), dst.begin());
  return src + dst.size();
}

const double *load_knot(Knot &k, const double *p) {
  p = take(k.Q, p);
  p = take(k.S, p);
  p = take(k.R, p);
  p = take(k.q, p);
  p = take(k.r, p);
  p = take(k.A, p);
  p = take(k.B, p);
  p = take(k.f, p);
  p = take(k.C, p);
  p = take(k.D, p);
  p = take(k.d, p);
  p = take(k.Gth, p);
  p = take(k.Gx, p);
  p = take(k.Gu, p);
  p = take(k.Gv, p);
  p = take(k.gamma, p);
  return p;
}

struct SolHolder {
  Solution s;
};

void flatten(const std::vector<vecd> &vv, double *out) {
  for (const auto &v : vv) {
    std::copy(v.begin(), v.end(), out);
    out += v.size();
  }
}
void unflatten(std::vector<vecd> &vv, const double *in) {
  for (auto &v : vv) {
    std::copy(in, in + v.size(), v.begin());
    in += v.size();
  }
}

enum Field {
  F_FF = 0,
  F_FB = 1,
  F_FTH = 2,
  F_VXX = 3,
  F_VX = 4,
  F_VXT = 5,
  F_VTT = 6,
  F_VT = 7,
  F_KKTMAT = 8,
  F_BK_MAT = 9,
  F_BK_SUBDIAG = 10,
  F_BK_PIV = 11, // written as doubles
  F_QHAT = 12,
  F_RHAT = 13,
  F_SHAT = 14,
  F_QVEC = 15,
  F_RVEC = 16,
};

int get_field(const StageFactor &d, int field, double *out) {
  const vecd *src = nullptr;
  switch (field) {
  case F_FF: src = &d.ff; break;
  case F_FB: src = &d.fb; break;
  case F_FTH: src = &d.fth; break;
  case F_VXX: src = &d.vm.Vxx; break;
  case F_VX: src = &d.vm.vx; break;
  case F_VXT: src = &d.vm.Vxt; break;
  case F_VTT: src = &d.vm.Vtt; break;
  case F_VT: src = &d.vm.vt; break;
  case F_KKTMAT: src = &d.kktMat; break;
  case F_BK_MAT: src = &d.kktChol.mat; break;
  case F_BK_SUBDIAG: src = &d.kktChol.subdiag; break;
  case F_QHAT: src = &d.Qhat; break;
  case F_RHAT: src = &d.Rhat; break;
  case F_SHAT: src = &d.Shat; break;
  case F_QVEC: src = &d.qhat; break;
  case F_RVEC: src = &d.rhat; break;
  case F_BK_PIV:
    for (size_t i = 0; i < d.kktChol.piv.size(); ++i)
      out[i] = (double)d.kktChol.piv[i];
    return (int)d.kktChol.piv.size();
  default: return -1;
  }
  std::copy(src->begin(), src->end(), out);
  return (int)src->size();
}

} // namespace

extern "C" {

long gar_oracle_knot_size(int nx, int nu, int nc, int nx2, int nth) {
  return (long)knot_size(nx, nu, nc, nx2, nth);
}

int gar_oracle_num_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// ---- problem ---------------------------------------------------------------
// dims: n_knots x 5 ints (nx,nu,nc,nx2,nth); data: knot records, then G0
// (nc0 x nx0 col-major), then g0.
void *gar_oracle_problem_create(int n_knots, const int *dims, int nc0,
                                const double *data) {
  auto *p = new Problem();
  p->stages.reserve(n_knots);
  const double *cur = data;
  for (int t = 0; t < n_knots; ++t) {
    const int *dm = dims + 5 * t;
    p->stages.emplace_back(dm[0], dm[1], dm[2], dm[3], dm[4]);
    cur = load_knot(p->stages.back(), cur);
  }
  p->nc0 = nc0;
  p->G0.assign((size_t)nc0 * p->stages[0].nx, 0.);
  p->g0.assign(nc0, 0.);
  cur = take(p->G0, cur);
  cur = take(p->g0, cur);
  return p;
}
void gar_oracle_problem_destroy(void *p) { delete (Problem *)p; }
// overwrite the numerical data of an existing problem (same dims)
void gar_oracle_problem_update(void *pv, const double *data) {
  auto *p = (Problem *)pv;
  const double *cur = data;
  for (auto &k : p->stages)
    cur = load_knot(k, cur);
  cur = take(p->G0, cur);
  cur = take(p->g0, cur);
}
// read back dims (the parallel solver re-parameterises the caller's problem)
void gar_oracle_problem_dims(void *pv, int *dims) {
  auto *p = (Problem *)pv;
  for (size_t t = 0; t < p->stages.size(); ++t) {
    const Knot &k = p->stages[t];
    dims[5 * t + 0] = k.nx;
    dims[5 * t + 1] = k.nu;
    dims[5 * t + 2] = k.nc;
    dims[5 * t + 3] = k.nx2;
    dims[5 * t + 4] = k.nth;
  }
}

// ---- solution --------------------------------------------------------------
void *gar_oracle_solution_create(void *pv) {
  auto *h = new SolHolder();
  h->s = lqrInitializeSolution(*(Problem *)pv);
  return h;
}
void gar_oracle_solution_destroy(void *h) { delete (SolHolder *)h; }
// sizes[4] = total doubles in xs, us, vs, lbdas; counts[4] = number of vectors
void gar_oracle_solution_sizes(void *hv, long *sizes, long *counts) {
  auto *h = (SolHolder *)hv;
  const std::vector<vecd> *all[4] = {&h->s.xs, &h->s.us, &h->s.vs, &h->s.lbdas};
  for (int i = 0; i < 4; ++i) {
    long n = 0;
    for (auto &v : *all[i])
      n += (long)v.size();
    sizes[i] = n;
    counts[i] = (long)all[i]->size();
  }
}
void gar_oracle_solution_get(void *hv, double *xs, double *us, double *vs,
                             double *lbdas) {
  auto *h = (SolHolder *)hv;
  flatten(h->s.xs, xs);
  flatten(h->s.us, us);
  flatten(h->s.vs, vs);
  flatten(h->s.lbdas, lbdas);
}
void gar_oracle_solution_set(void *hv, const double *xs, const double *us,
                             const double *vs, const double *lbdas) {
  auto *h = (SolHolder *)hv;
  unflatten(h->s.xs, xs);
  unflatten(h->s.us, us);
  unflatten(h->s.vs, vs);
  unflatten(h->s.lbdas, lbdas);
}

void gar_oracle_kkt_error(void *pv, void *hv, double mueq, const double *theta,
                          double *out3) {
  KktError e =
      lqrComputeKktError(*(Problem *)pv, ((SolHolder *)hv)->s, mueq, theta);
  out3[0] = e.dyn;
  out3[1] = e.cstr;
  out3[2] = e.dual;
}

// ---- serial solver -----------------------------------------------------------
void *gar_oracle_serial_create(void *pv) {
  return new ProximalRiccatiSolver(*(Problem *)pv);
}
void gar_oracle_serial_destroy(void *s) { delete (ProximalRiccatiSolver *)s; }
int gar_oracle_serial_backward(void *s, double mueq) {
  return ((ProximalRiccatiSolver *)s)->backward(mueq) ? 1 : 0;
}
int gar_oracle_serial_forward(void *s, void *hv, const double *theta) {
  return ((ProximalRiccatiSolver *)s)->forward(((SolHolder *)hv)->s, theta) ? 1
                                                                            : 0;
}
int gar_oracle_serial_get(void *s, int t, int field, double *out) {
  return get_field(((ProximalRiccatiSolver *)s)->datas[t], field, out);
}
// kkt0: which = 0 ff, 1 fth, 2 mat, 3 thGrad, 4 thHess
int gar_oracle_serial_get_kkt0(void *sv, int which, double *out) {
  auto *s = (ProximalRiccatiSolver *)sv;
  const vecd *src = nullptr;
  switch (which) {
  case 0: src = &s->kkt0.ff; break;
  case 1: src = &s->kkt0.fth; break;
  case 2: src = &s->kkt0.mat; break;
  case 3: src = &s->thGrad; break;
  case 4: src = &s->thHess; break;
  default: return -1;
  }
  std::copy(src->begin(), src->end(), out);
  return (int)src->size();
}
// cycleAppend with a knot given by dims + flat record
void gar_oracle_serial_cycle_append(void *sv, const int *dm,
                                    const double *rec) {
  Knot k(dm[0], dm[1], dm[2], dm[3], dm[4]);
  load_knot(k, rec);
  ((ProximalRiccatiSolver *)sv)->cycleAppend(k);
}

// ---- stage-dense solver (gar/dense-riccati.hxx) --------------------------------
void *gar_oracle_dense_create(void *pv) { return new RiccatiSolverDense(*(Problem *)pv); }
void gar_oracle_dense_destroy(void *s) { delete (RiccatiSolverDense *)s; }
int gar_oracle_dense_backward(void *s, double mueq) {
  return ((RiccatiSolverDense *)s)->backward(mueq) ? 1 : 0;
}
int gar_oracle_dense_forward(void *s, void *hv, const double *theta) {
  return ((RiccatiSolverDense *)s)->forward(((SolHolder *)hv)->s, theta) ? 1 : 0;
}
// field: 0 ff [k;z;l;y], 1 fb row-major [K;Z;L;Y], 2 ft, 3 Pxx, 4 px, 5 Pxt, 6 Ptt, 7 pt
int gar_oracle_dense_get(void *sv, int t, int field, double *out) {
  auto *s = (RiccatiSolverDense *)sv;
  const vecd *src = nullptr;
  switch (field) {
  case 0: src = &s->stage_factors[t].ff; break;
  case 1: src = &s->stage_factors[t].fb; break;
  case 2: src = &s->stage_factors[t].ft; break;
  case 3: src = &s->P[t].Pxx; break;
  case 4: src = &s->P[t].px; break;
  case 5: src = &s->P[t].Pxt; break;
  case 6: src = &s->P[t].Ptt; break;
  case 7: src = &s->P[t].pt; break;
  default: return -1;
  }
  std::copy(src->begin(), src->end(), out);
  return (int)src->size();
}
int gar_oracle_dense_get_kkt0(void *sv, int which, double *out) {
  auto *s = (RiccatiSolverDense *)sv;
  const vecd *src = nullptr;
  switch (which) {
  case 0: src = &s->kkt0.ff; break;
  case 1: src = &s->kkt0.fth; break;
  case 2: src = &s->kkt0.mat; break;
  case 3: src = &s->thGrad; break;
  case 4: src = &s->thHess; break;
  default: return -1;
  }
  std::copy(src->begin(), src->end(), out);
  return (int)src->size();
}

// ---- parallel solver ---------------------------------------------------------
void *gar_oracle_parallel_create(void *pv, int nthreads, int threaded) {
  auto *s = new ParallelRiccatiSolver(*(Problem *)pv, (uint)nthreads);
  s->threaded = threaded != 0;
  return s;
}
void gar_oracle_parallel_destroy(void *s) { delete (ParallelRiccatiSolver *)s; }
void gar_oracle_parallel_set_refinement(void *s, int steps, double thr) {
  ((ParallelRiccatiSolver *)s)->maxRefinementSteps = (uint)steps;
  ((ParallelRiccatiSolver *)s)->condensedThreshold = thr;
}
int gar_oracle_parallel_backward(void *s, double mueq) {
  return ((ParallelRiccatiSolver *)s)->backward(mueq) ? 1 : 0;
}
int gar_oracle_parallel_forward(void *s, void *hv) {
  return ((ParallelRiccatiSolver *)s)->forward(((SolHolder *)hv)->s) ? 1 : 0;
}
void gar_oracle_parallel_collapse(void *s) {
  ((ParallelRiccatiSolver *)s)->collapseFeedback();
}
int gar_oracle_parallel_get(void *s, int t, int field, double *out) {
  return get_field(((ParallelRiccatiSolver *)s)->datas[t], field, out);
}
int gar_oracle_parallel_ok(void *s) {
  return ((ParallelRiccatiSolver *)s)->ok_ ? 1 : 0;
}

// ---- Bunch-Kaufman alone -----------------------------------------------------
// a: n x n column-major (lower triangle read).  Outputs: mat (n*n), subdiag (n),
// piv (n).  Returns info.
int gar_oracle_bk_compute(int n, const double *a, double *mat, double *subdiag,
                          int *piv) {
  BunchKaufman bk(n);
  bk.compute(a, n);
  std::copy(bk.mat.begin(), bk.mat.end(), mat);
  std::copy(bk.subdiag.begin(), bk.subdiag.end(), subdiag);
  std::copy(bk.piv.begin(), bk.piv.end(), piv);
  return bk.info;
}
// solve A X = B in place; x is n x nrhs column-major
int gar_oracle_bk_solve(int n, const double *a, int nrhs, double *x) {
  BunchKaufman bk(n);
  bk.compute(a, n);
  if (bk.info != BK_SUCCESS)
    return bk.info;
  bk.solveInPlace(x, nrhs, 1, n);
  return 0;
}

// ---- batched uniform-dims driver (CPU baseline + parity checker) --------------
// Packed device-style layout of the product (see include/aligator_b200/gar.h):
//   stage record  [A | B | f | Q | S | R | q | r | C | D | d]   (nx2 = nx)
//   term  record  [Q | q | C | d]   (nu = 0, nc = nct)
// All instances share dims.  Builds `batch` problems + solvers once; the timed
// region (returned in seconds) is exactly `reps` x { backward; forward } per
// instance, as bench/gar-riccati.cpp:46-49, parallelised over instances.
struct Batched {
  int nx, nu, nc, nct, nc0, N, batch;
  std::vector<std::unique_ptr<Problem>> probs;
  std::vector<std::unique_ptr<ProximalRiccatiSolver>> solvers;
  std::vector<Solution> sols;
};

void *gar_oracle_batched_create(int nx, int nu, int nc, int nct, int nc0, int N,
                                int batch, const double *stage,
                                const double *term, const double *G0,
                                const double *g0) {
  auto *b = new Batched{nx, nu, nc, nct, nc0, N, batch, {}, {}, {}};
  const size_t srec = (size_t)nx * nx + (size_t)nx * nu + nx + (size_t)nx * nx +
                      (size_t)nx * nu + (size_t)nu * nu + nx + nu +
                      (size_t)nc * nx + (size_t)nc * nu + nc;
  const size_t trec = (size_t)nx * nx + nx + (size_t)nct * nx + nct;
  b->probs.resize(batch);
  b->solvers.resize(batch);
  b->sols.resize(batch);
  for (int i = 0; i < batch; ++i) {
    auto p = std::make_unique<Problem>();
    p->stages.reserve(N + 1);
    for (int t = 0; t < N; ++t) {
      p->stages.emplace_back(nx, nu, nc, nx, 0);
      Knot &k = p->stages.back();
      const double *c = stage + ((size_t)i * N + t) * srec;
      c = take(k.A, c);
      c = take(k.B, c);
      c = take(k.f, c);
      c = take(k.Q, c);
      c = take(k.S, c);
      c = take(k.R, c);
      c = take(k.q, c);
      c = take(k.r, c);
      c = take(k.C, c);
      c = take(k.D, c);
      c = take(k.d, c);
    }
    p->stages.emplace_back(nx, 0, nct, nx, 0); // terminal knot, nu = 0
    {
      Knot &k = p->stages.back();
      const double *c = term + (size_t)i * trec;
      c = take(k.Q, c);
      c = take(k.q, c);
      c = take(k.C, c);
      c = take(k.d, c);
    }
    p->nc0 = nc0;
    p->G0.assign(G0 + (size_t)i * nc0 * nx, G0 + (size_t)(i + 1) * nc0 * nx);
    p->g0.assign(g0 + (size_t)i * nc0, g0 + (size_t)(i + 1) * nc0);
    b->solvers[i] = std::make_unique<ProximalRiccatiSolver>(*p);
    b->sols[i] = lqrInitializeSolution(*p);
    b->probs[i] = std::move(p);
  }
  return b;
}
void gar_oracle_batched_destroy(void *b) { delete (Batched *)b; }

// returns elapsed seconds of the timed region; status[i] = 1 ok / 0 failed
double gar_oracle_batched_sweep(void *bv, double mueq, int reps, int nthreads,
                                int *status) {
  auto *b = (Batched *)bv;
#ifdef _OPENMP
  if (nthreads <= 0)
    nthreads = omp_get_max_threads();
#else
  nthreads = 1;
#endif
  // One parallel region around ALL repetitions: instances are independent, so each thread
  // sweeps its own instances `reps` times back to back -- no fork/join or barrier per sweep
  // (with many threads and a passive wait policy that overhead dwarfs the work).
  auto t0 = std::chrono::steady_clock::now();
#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads)
#endif
  {
#ifdef _OPENMP
    const int tid = omp_get_thread_num(), nt = omp_get_num_threads();
#pragma omp barrier
#pragma omp master
    t0 = std::chrono::steady_clock::now();
#pragma omp barrier
#else
    const int tid = 0, nt = 1;
#endif
    for (int r = 0; r < reps; ++r)
      for (int i = tid; i < b->batch; i += nt) {
        bool ok = b->solvers[i]->backward(mueq);
        ok = b->solvers[i]->forward(b->sols[i]) && ok;
        if (status)
          status[i] = ok ? 1 : 0;
      }
  }
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

// Outputs in the product's batched layout (see include/aligator_b200/gar.h):
//   ff   [batch][N][nu+nc+nx]      fb  [batch][N][(nu+nc+nx)*nx] row-major
//   Vxx  [batch][N+1][nx*nx]       vx  [batch][N+1][nx]
//   ffT  [batch][nct]              fbT [batch][nct*nx]
//   xs [batch][N+1][nx]  us [batch][N][nu]  vs [batch][N][nc]  vsT [batch][nct]
//   lbd0 [batch][nc0]    lbdas [batch][N][nx]
// Any pointer may be null.
void gar_oracle_batched_get(void *bv, double *ff, double *fb, double *Vxx,
                            double *vx, double *ffT, double *fbT, double *xs,
                            double *us, double *vs, double *vsT, double *lbd0,
                            double *lbdas) {
  auto *b = (Batched *)bv;
  const int nx = b->nx, nu = b->nu, nc = b->nc, nct = b->nct, nc0 = b->nc0,
            N = b->N;
  const size_t nr = (size_t)nu + nc + nx;
  for (int i = 0; i < b->batch; ++i) {
    const auto &S = *b->solvers[i];
    const auto &sol = b->sols[i];
    for (int t = 0; t <= N; ++t) {
      const StageFactor &d = S.datas[t];
      if (t < N) {
        if (ff)
          std::copy(d.ff.begin(), d.ff.end(), ff + ((size_t)i * N + t) * nr);
        if (fb)
          std::copy(d.fb.begin(), d.fb.end(),
                    fb + ((size_t)i * N + t) * nr * nx);
        if (us)
          std::copy(sol.us[t].begin(), sol.us[t].end(),
                    us + ((size_t)i * N + t) * nu);
        if (vs)
          std::copy(sol.vs[t].begin(), sol.vs[t].end(),
                    vs + ((size_t)i * N + t) * nc);
        if (lbdas)
          std::copy(sol.lbdas[t + 1].begin(), sol.lbdas[t + 1].end(),
                    lbdas + ((size_t)i * N + t) * nx);
      } else {
        if (ffT)
          std::copy(d.ff.begin(), d.ff.begin() + nct, ffT + (size_t)i * nct);
        if (fbT)
          std::copy(d.fb.begin(), d.fb.begin() + (size_t)nct * nx,
                    fbT + (size_t)i * nct * nx);
        if (vsT)
          std::copy(sol.vs[t].begin(), sol.vs[t].end(), vsT + (size_t)i * nct);
      }
      if (Vxx)
        std::copy(d.vm.Vxx.begin(), d.vm.Vxx.end(),
                  Vxx + ((size_t)i * (N + 1) + t) * nx * nx);
      if (vx)
        std::copy(d.vm.vx.begin(), d.vm.vx.end(),
                  vx + ((size_t)i * (N + 1) + t) * nx);
      if (xs)
        std::copy(sol.xs[t].begin(), sol.xs[t].end(),
                  xs + ((size_t)i * (N + 1) + t) * nx);
    }
    if (lbd0)
      std::copy(sol.lbdas[0].begin(), sol.lbdas[0].end(),
                lbd0 + (size_t)i * nc0);
  }
}

} // extern "C"
