// gar_oracle_dense.hpp -- CPU restatement of aligator's STAGE-DENSE Riccati solver
// (gar::DenseKernel, gar/dense-kernel.hpp:13-211; gar::RiccatiSolverDense,
// gar/dense-riccati.hxx:13-146), the reference's second, algorithmically independent
// solver of the same LQ problem: one Bunch-Kaufman factorisation per knot of the
// (nu + nc + 2 nx2)^2 matrix
//     [[R, D^T, B^T, 0], [D, -mu I, 0, 0], [B, 0, 0, -I], [0, 0, -I, P']]
// instead of the proximal kernel's reduced (nu + nc)^2 system.
//
// *** TEST INFRASTRUCTURE ONLY *** (same rule as gar_oracle.hpp).  It exists to
// cross-check the proximal restatement AND the CUDA path against a different algorithm:
// both must agree on K, k, Pxx = Vxx, px = vx and on the primal-dual trajectory to
// rounding.  PARITY UNPINNED like the rest
// of the oracle (no Eigen in the image).
#pragma once

#include "gar_oracle.hpp"

namespace gar_oracle {

struct DenseData { // DenseKernel::Data, dense-kernel.hpp:18-44
  uint nx, nu, nc, nx2, nth;
  int n;       // nu + nc + 2 nx2
  vecd kktMat; // n x n column-major, blocks (nu, nc, nx2, nx2)
  vecd fb;     // row-major n x nx   = [K; Z; L; Y]
  vecd ft;     // row-major n x nth  = [Kth; Zth; Lth; Yth]
  vecd ff;     // n                  = [k; z; l; y]
  BunchKaufman ldl;
  DenseData(uint nx_, uint nu_, uint nc_, uint nx2_, uint nth_)
      : nx(nx_), nu(nu_), nc(nc_), nx2(nx2_), nth(nth_), n((int)(nu_ + nc_ + 2 * nx2_)),
        kktMat((size_t)n * n, 0.), fb((size_t)n * nx_, 0.), ft((size_t)n * nth_, 0.), ff(n, 0.),
        ldl(n) {}
  void setZero() {
    std::fill(kktMat.begin(), kktMat.end(), 0.);
    std::fill(fb.begin(), fb.end(), 0.);
    std::fill(ft.begin(), ft.end(), 0.);
    std::fill(ff.begin(), ff.end(), 0.);
  }
};

struct DenseValue { // DenseKernel::value, dense-kernel.hpp:47-53 (owning here)
  vecd Pxx, Pxt, Ptt, px, pt;
  DenseValue(uint nx, uint nth)
      : Pxx((size_t)nx * nx, 0.), Pxt((size_t)nx * nth, 0.), Ptt((size_t)nth * nth, 0.), px(nx, 0.),
        pt(nth, 0.) {}
};

struct DenseKernel {
  // out (r x c, column-major) = (init ? out : 0) + op(A) * op(B) helpers, written out as the
  // plain triple loops the expressions of dense-kernel.hpp:81-96, 151-178 denote
  // value update shared by terminalSolve (with_next = false, 2 blocks) and stageKernelSolve
  static void update_value(const Knot &k, const DenseData &d, DenseValue &v, const DenseValue *vn,
                           bool stage) {
    const int nx = k.nx, nu = k.nu, nc = k.nc, nx2 = k.nx2, nth = k.nth;
    const double *K = d.fb.data(), *Z = K + (size_t)nu * nx, *L = Z + (size_t)nc * nx,
                 *Y = L + (size_t)nx2 * nx;
    const double *Kth = d.ft.data(), *Zth = Kth + (size_t)nu * nth, *Yth = Zth + (size_t)(nc + nx2) * nth;
    const double *kff = d.ff.data(), *zff = kff + nu, *lff = zff + nc, *yff = lff + nx2;
    // Pxx = Q + S K + C^T Z (+ A^T L)          (:81-82, :151-153)
    for (int j = 0; j < nx; ++j)
      for (int i = 0; i < nx; ++i) {
        double s = k.Q[i + (size_t)j * nx];
        double a = 0.0;
        for (int c = 0; c < nu; ++c)
          a += k.S[i + (size_t)c * nx] * K[(size_t)c * nx + j];
        s += a;
        a = 0.0;
        for (int c = 0; c < nc; ++c)
          a += k.C[c + (size_t)i * nc] * Z[(size_t)c * nx + j];
        s += a;
        if (stage) {
          a = 0.0;
          for (int c = 0; c < nx2; ++c)
            a += k.A[c + (size_t)i * nx2] * L[(size_t)c * nx + j];
          s += a;
        }
        v.Pxx[i + (size_t)j * nx] = s;
      }
    // Pxt = Gx + K^T Gu + Z^T Gv (+ Y^T Pxt')  (:84-85, :155-159)
    for (int j = 0; j < nth; ++j)
      for (int i = 0; i < nx; ++i) {
        double s = k.Gx[i + (size_t)j * nx];
        double a = 0.0;
        for (int c = 0; c < nu; ++c)
          a += K[(size_t)c * nx + i] * k.Gu[c + (size_t)j * nu];
        s += a;
        a = 0.0;
        for (int c = 0; c < nc; ++c)
          a += Z[(size_t)c * nx + i] * k.Gv[c + (size_t)j * nc];
        s += a;
        if (stage && vn) {
          a = 0.0;
          for (int c = 0; c < nx2; ++c)
            a += Y[(size_t)c * nx + i] * vn->Pxt[c + (size_t)j * nx2];
          s += a;
        }
        v.Pxt[i + (size_t)j * nx] = s;
      }
    // Ptt = Gth + Gu^T Kth + Gv^T Zth (+ Yth^T Pxt')   (:87-88, :161-165; the stage kernel
    // writes Kth^T Gu -- the transpose of the same product, Ptt being symmetric in exact
    // arithmetic; restated as written per kernel)
    for (int j = 0; j < nth; ++j)
      for (int i = 0; i < nth; ++i) {
        double s = k.Gth[i + (size_t)j * nth];
        double a = 0.0;
        for (int c = 0; c < nu; ++c)
          a += stage ? Kth[(size_t)c * nth + i] * k.Gu[c + (size_t)j * nu]
                     : k.Gu[c + (size_t)i * nu] * Kth[(size_t)c * nth + j];
        s += a;
        a = 0.0;
        for (int c = 0; c < nc; ++c)
          a += stage ? Zth[(size_t)c * nth + i] * k.Gv[c + (size_t)j * nc]
                     : k.Gv[c + (size_t)i * nc] * Zth[(size_t)c * nth + j];
        s += a;
        if (stage && vn) {
          a = 0.0;
          for (int c = 0; c < nx2; ++c)
            a += Yth[(size_t)c * nth + i] * vn->Pxt[c + (size_t)j * nx2];
          s += a;
        }
        v.Ptt[i + (size_t)j * nth] = s;
      }
    // px = q + S k + C^T z (+ A^T l)           (:90-91, :167-169)
    for (int i = 0; i < nx; ++i) {
      double s = k.q[i];
      double a = 0.0;
      for (int c = 0; c < nu; ++c)
        a += k.S[i + (size_t)c * nx] * kff[c];
      s += a;
      a = 0.0;
      for (int c = 0; c < nc; ++c)
        a += k.C[c + (size_t)i * nc] * zff[c];
      s += a;
      if (stage) {
        a = 0.0;
        for (int c = 0; c < nx2; ++c)
          a += k.A[c + (size_t)i * nx2] * lff[c];
        s += a;
      }
      v.px[i] = s;
    }
    // pt = gamma + Gu^T k + Gv^T z (+ Pxt'^T y) (:93-94, :171-174)
    for (int i = 0; i < nth; ++i) {
      double s = k.gamma[i];
      double a = 0.0;
      for (int c = 0; c < nu; ++c)
        a += k.Gu[c + (size_t)i * nu] * kff[c];
      s += a;
      a = 0.0;
      for (int c = 0; c < nc; ++c)
        a += k.Gv[c + (size_t)i * nc] * zff[c];
      s += a;
      if (stage && vn) {
        a = 0.0;
        for (int c = 0; c < nx2; ++c)
          a += vn->Pxt[c + (size_t)i * nx2] * yff[c];
        s += a;
      }
      v.pt[i] = s;
    }
  }

  // dense-kernel.hpp:55-95.  (The KKT matrix keeps the full 4-block size; the last two
  // blocks stay zero, so the factorisation reports a singular column when nx2 > 0 -- the
  // reference ignores compute()'s status here too; with the ProxDDP terminal knot nx2 = 0.)
  static bool terminalSolve(const Knot &k, DenseData &d, DenseValue &v, double mueq) {
    const int nx = k.nx, nu = k.nu, nc = k.nc, nth = k.nth, n = d.n;
    std::fill(d.kktMat.begin(), d.kktMat.end(), 0.);
    CM M{d.kktMat.data(), n};
    for (int j = 0; j < nu; ++j)
      for (int i = 0; i < nu; ++i)
        M(i, j) = k.R[i + (size_t)j * nu];
    for (int i = 0; i < nc; ++i)
      for (int j = 0; j < nu; ++j) {
        M(nu + i, j) = k.D[i + (size_t)j * nc];
        M(j, nu + i) = k.D[i + (size_t)j * nc];
      }
    for (int i = 0; i < nc; ++i)
      M(nu + i, nu + i) = -mueq;
    for (int i = 0; i < nu; ++i)
      d.ff[i] = -k.r[i];
    for (int i = 0; i < nc; ++i)
      d.ff[nu + i] = -k.d[i];
    for (int i = 0; i < nu; ++i)
      for (int j = 0; j < nx; ++j)
        d.fb[(size_t)i * nx + j] = -k.S[j + (size_t)i * nx];
    for (int i = 0; i < nc; ++i)
      for (int j = 0; j < nx; ++j)
        d.fb[(size_t)(nu + i) * nx + j] = -k.C[i + (size_t)j * nc];
    d.ldl.compute(d.kktMat.data(), n);
    d.ldl.solveInPlace(d.ff.data(), 1, 1, n);
    d.ldl.solveInPlace(d.fb.data(), nx, nx, 1);
    for (int i = 0; i < nu; ++i)
      for (int j = 0; j < nth; ++j)
        d.ft[(size_t)i * nth + j] = -k.Gu[i + (size_t)j * nu];
    for (int i = 0; i < nc; ++i)
      for (int j = 0; j < nth; ++j)
        d.ft[(size_t)(nu + i) * nth + j] = -k.Gv[i + (size_t)j * nc];
    d.ldl.solveInPlace(d.ft.data(), nth, nth, 1);
    update_value(k, d, v, nullptr, false);
    return d.ldl.info == BK_SUCCESS;
  }

  // dense-kernel.hpp:97-175
  static bool stageKernelSolve(const Knot &k, DenseData &d, DenseValue &v, const DenseValue *vn,
                               double mueq) {
    const int nx = k.nx, nu = k.nu, nc = k.nc, nx2 = k.nx2, nth = k.nth, n = d.n;
    const int o2 = nu + nc, o3 = o2 + nx2;
    std::fill(d.kktMat.begin(), d.kktMat.end(), 0.);
    CM M{d.kktMat.data(), n};
    for (int j = 0; j < nu; ++j)
      for (int i = 0; i < nu; ++i)
        M(i, j) = k.R[i + (size_t)j * nu];
    for (int i = 0; i < nc; ++i)
      for (int j = 0; j < nu; ++j) {
        M(nu + i, j) = k.D[i + (size_t)j * nc];
        M(j, nu + i) = k.D[i + (size_t)j * nc];
      }
    for (int i = 0; i < nc; ++i)
      M(nu + i, nu + i) = -mueq;
    for (int i = 0; i < nx2; ++i)
      for (int j = 0; j < nu; ++j) {
        M(o2 + i, j) = k.B[i + (size_t)j * nx2];
        M(j, o2 + i) = k.B[i + (size_t)j * nx2];
      }
    for (int i = 0; i < nx2; ++i) {
      M(o2 + i, o3 + i) = -1.0;
      M(o3 + i, o2 + i) = -1.0;
    }
    if (vn)
      for (int j = 0; j < nx2; ++j)
        for (int i = 0; i < nx2; ++i)
          M(o3 + i, o3 + j) = vn->Pxx[i + (size_t)j * nx2];
    d.ldl.compute(d.kktMat.data(), n); // (reads the lower triangle, bunchkaufman.hpp:669-670)
    for (int i = 0; i < nu; ++i)
      d.ff[i] = -k.r[i];
    for (int i = 0; i < nc; ++i)
      d.ff[nu + i] = -k.d[i];
    for (int i = 0; i < nx2; ++i)
      d.ff[o2 + i] = -k.f[i];
    if (vn) // (yff keeps its previous content otherwise, :123-125; vn is always given by the solver)
      for (int i = 0; i < nx2; ++i)
        d.ff[o3 + i] = -vn->px[i];
    for (int i = 0; i < nu; ++i)
      for (int j = 0; j < nx; ++j)
        d.fb[(size_t)i * nx + j] = -k.S[j + (size_t)i * nx];
    for (int i = 0; i < nc; ++i)
      for (int j = 0; j < nx; ++j)
        d.fb[(size_t)(nu + i) * nx + j] = -k.C[i + (size_t)j * nc];
    for (int i = 0; i < nx2; ++i)
      for (int j = 0; j < nx; ++j) {
        d.fb[(size_t)(o2 + i) * nx + j] = -k.A[i + (size_t)j * nx2];
        d.fb[(size_t)(o3 + i) * nx + j] = 0.0;
      }
    for (int i = 0; i < nu; ++i)
      for (int j = 0; j < nth; ++j)
        d.ft[(size_t)i * nth + j] = -k.Gu[i + (size_t)j * nu];
    for (int i = 0; i < nc; ++i)
      for (int j = 0; j < nth; ++j)
        d.ft[(size_t)(nu + i) * nth + j] = -k.Gv[i + (size_t)j * nc];
    for (int i = 0; i < nx2; ++i)
      for (int j = 0; j < nth; ++j) {
        d.ft[(size_t)(o2 + i) * nth + j] = 0.0;
        if (vn)
          d.ft[(size_t)(o3 + i) * nth + j] = -vn->Pxt[i + (size_t)j * nx2];
      }
    d.ldl.solveInPlace(d.ff.data(), 1, 1, n);
    d.ldl.solveInPlace(d.fb.data(), nx, nx, 1);
    d.ldl.solveInPlace(d.ft.data(), nth, nth, 1);
    update_value(k, d, v, vn, true);
    return d.ldl.info == BK_SUCCESS;
  }

  // dense-kernel.hpp:177-215
  static void forwardStep(size_t i, bool isTerminal, const Knot &k, const DenseData &d, Solution &s,
                          const double *theta) {
    const int nx = k.nx, nu = k.nu, nc = k.nc, nx2 = k.nx2, nth = k.nth;
    const int o2 = nu + nc, o3 = o2 + nx2;
    const double *x = s.xs[i].data();
    auto row = [&](int r) {
      double a = 0.0;
      for (int c = 0; c < nx; ++c)
        a += d.fb[(size_t)r * nx + c] * x[c];
      double v = d.ff[r] + a;
      if (theta && nth > 0) {
        double b = 0.0;
        for (int c = 0; c < nth; ++c)
          b += d.ft[(size_t)r * nth + c] * theta[c];
        v += b;
      }
      return v;
    };
    if (nu > 0)
      for (int r = 0; r < nu; ++r)
        s.us[i][r] = row(r);
    for (int r = 0; r < nc; ++r)
      s.vs[i][r] = row(nu + r);
    if (isTerminal)
      return;
    for (int r = 0; r < nx2; ++r)
      s.lbdas[i + 1][r] = row(o2 + r);
    for (int r = 0; r < nx2; ++r)
      s.xs[i + 1][r] = row(o3 + r);
  }
};

// gar/dense-riccati.hxx:13-123
struct RiccatiSolverDense {
  const Problem *problem_;
  std::vector<DenseData> stage_factors;
  std::vector<DenseValue> P;
  Kkt0 kkt0;
  vecd thGrad, thHess;

  explicit RiccatiSolverDense(const Problem &p)
      : problem_(&p), kkt0(p.stages[0].nx, p.nc0, p.ntheta()), thGrad(p.ntheta(), 0.),
        thHess((size_t)p.ntheta() * p.ntheta(), 0.) {
    for (const Knot &k : p.stages) {
      P.emplace_back(k.nx, k.nth);
      stage_factors.emplace_back(k.nx, k.nu, k.nc, k.nx2, k.nth);
    }
  }

  bool backward(double mueq) { // :47-99
    const auto &st = problem_->stages;
    const int N = problem_->horizon();
    // (the terminal KKT matrix keeps zero blocks for the unused nx2 rows, so its compute()
    // reports a singular column once the (nu + nc) leading part is factored; like the
    // reference, which discards the status, only the stage factorisations are reported)
    DenseKernel::terminalSolve(st[N], stage_factors[N], P[N], mueq);
    bool ok = true;
    for (int i = N - 1; i >= 0; --i)
      ok = DenseKernel::stageKernelSolve(st[i], stage_factors[i], P[i], &P[i + 1], mueq) && ok;
    const int nx = kkt0.nx, nc0 = kkt0.nc, nth = kkt0.nth, n0 = nx + nc0;
    std::fill(kkt0.mat.begin(), kkt0.mat.end(), 0.);
    CM M{kkt0.mat.data(), n0};
    CCM G0{problem_->G0.data(), nc0};
    for (int j = 0; j < nx; ++j)
      for (int i = 0; i < nx; ++i)
        M(i, j) = P[0].Pxx[i + (size_t)j * nx];
    for (int i = 0; i < nc0; ++i)
      for (int j = 0; j < nx; ++j) {
        M(nx + i, j) = G0(i, j);
        M(j, nx + i) = G0(i, j);
      }
    kkt0.chol.compute(kkt0.mat.data(), n0);
    for (int i = 0; i < nx; ++i)
      kkt0.ff[i] = -P[0].px[i];
    for (int i = 0; i < nc0; ++i)
      kkt0.ff[nx + i] = -problem_->g0[i];
    for (int i = 0; i < nx; ++i)
      for (int j = 0; j < nth; ++j)
        kkt0.fth[(size_t)i * nth + j] = -P[0].Pxt[i + (size_t)j * nx];
    for (int i = 0; i < nc0; ++i)
      for (int j = 0; j < nth; ++j)
        kkt0.fth[(size_t)(nx + i) * nth + j] = 0.0;
    kkt0.chol.solveInPlace(kkt0.ff.data(), 1, 1, n0);
    kkt0.chol.solveInPlace(kkt0.fth.data(), nth, nth, 1);
    for (int i = 0; i < nth; ++i) {
      double s = 0.0;
      for (int c = 0; c < nx; ++c)
        s += P[0].Pxt[c + (size_t)i * nx] * kkt0.ff[c];
      thGrad[i] = P[0].pt[i] + s;
    }
    for (int j = 0; j < nth; ++j)
      for (int i = 0; i < nth; ++i) {
        double s = 0.0;
        for (int c = 0; c < nx; ++c)
          s += P[0].Pxt[c + (size_t)i * nx] * kkt0.fth[(size_t)c * nth + j];
        thHess[i + (size_t)j * nth] = P[0].Ptt[i + (size_t)j * nth] + s;
      }
    return ok && kkt0.chol.info == BK_SUCCESS;
  }

  bool forward(Solution &s, const double *theta = nullptr) const { // :101-123
    Kernel::computeInitial(s.xs[0], s.lbdas[0], kkt0, theta);
    const size_t N = (size_t)problem_->horizon();
    for (size_t i = 0; i <= N; ++i)
      DenseKernel::forwardStep(i, i == N, problem_->stages[i], stage_factors[i], s, theta);
    return true;
  }
};

} // namespace gar_oracle
