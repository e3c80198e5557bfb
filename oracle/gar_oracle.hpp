// gar_oracle.hpp -- CPU restatement of aligator's `gar` Riccati path.
//
// *** TEST INFRASTRUCTURE ONLY. ***  Nothing in the product path
// (aligator_b200/, include/) may include, link or call this file.  Only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs use it, and only as the checker / the timed CPU arm.
//
// PARITY UNPINNED: the reference cannot be compiled in this image (Eigen,
// mimalloc, Boost, fmt absent) and its tests hold no golden vectors for
// K, k, Vxx (only KKT-residual thresholds).  This restatement is therefore
// pinned against (a) the reference's own test thresholds on the reference's
// own problem generators' *shape*, (b) a dense numpy solve of the full KKT
// system, (c) LAPACK dsytf2 pivot sequences.  See DESIGN.md "Oracle".
//
// Every function cites the reference file:line it restates (paths relative
// to the aligator source tree).  Plain C++17, no Eigen.  Storage conventions
// are the reference's: matrices column-major except fb / fth / AtV / BtV
// which are row-major (math.hpp:23-27, riccati-kernel.hpp:86-101).
#pragma once

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <utility>
#include <vector>

namespace gar_oracle {

using uint = unsigned int;
using vecd = std::vector<double>;

// ---------------------------------------------------------------------------
// Tiny dense views (no ownership).
// ---------------------------------------------------------------------------
struct CM { // column-major view
  double *p;
  int ld;
  double &operator()(int i, int j) const { return p[i + (std::ptrdiff_t)j * ld]; }
};
struct CCM {
  const double *p;
  int ld;
  double operator()(int i, int j) const { return p[i + (std::ptrdiff_t)j * ld]; }
};
struct RM { // row-major view
  double *p;
  int ld;
  double &operator()(int i, int j) const { return p[(std::ptrdiff_t)i * ld + j]; }
};
// generic strided view (used for BK right-hand sides of either storage order)
struct SV {
  double *p;
  std::ptrdiff_t rs, cs;
  double &operator()(int i, int j) const { return p[i * rs + j * cs]; }
};

// ---------------------------------------------------------------------------
// Bunch-Kaufman LDL^T, lower, in place.   core/bunchkaufman.hpp
// ---------------------------------------------------------------------------
enum BkInfo { BK_SUCCESS = 0, BK_NUMERICAL_ISSUE = 1, BK_INVALID = 3 };

inline double bk_alpha() { return (1.0 + std::sqrt(17.0)) / 8.0; } // bunchkaufman.hpp:28

// Unblocked factorization of the n x n block `a` (lower part referenced).
// Restates bunch_kaufman_in_place_unblocked, core/bunchkaufman.hpp:22-169.
inline int bk_unblocked(CM a, int n, int *piv, long &pivot_count) {
  const double alpha = bk_alpha();
  pivot_count = 0;
  if (n == 0)
    return BK_SUCCESS;
  if (n == 1) { // :36-42
    if (std::fabs(a(0, 0)) == 0.0)
      return BK_NUMERICAL_ISSUE;
    a(0, 0) = 1.0 / a(0, 0);
    return BK_SUCCESS;
  }
  int k = 0;
  while (k < n) {
    int kstep = 1;
    const double abs_akk = std::fabs(a(k, k));
    int imax = 0;
    double colmax = 0.0;
    if (k + 1 < n) { // first-max index, as Eigen's maxCoeff(&imax)  (:53)
      colmax = std::fabs(a(k + 1, k));
      for (int i = k + 2; i < n; ++i) {
        const double v = std::fabs(a(i, k));
        if (v > colmax) {
          colmax = v;
          imax = i - (k + 1);
        }
      }
    }
    imax += k + 1;
    int kp;
    if (std::max(abs_akk, colmax) == 0.0) // :58-59
      return BK_NUMERICAL_ISSUE;
    if (abs_akk >= colmax * alpha) { // :61
      kp = k;
    } else {
      double rowmax = 0.0; // :64-72
      for (int j = k; j < imax; ++j)
        rowmax = std::max(rowmax, std::fabs(a(imax, j)));
      for (int i = imax + 1; i < n; ++i)
        rowmax = std::max(rowmax, std::fabs(a(i, imax)));
      if (abs_akk >= (alpha * colmax) * (colmax / rowmax)) { // :74
        kp = k;
      } else if (std::fabs(a(imax, imax)) >= alpha * rowmax) { // :76
        kp = imax;
      } else { // :78-80
        kp = imax;
        kstep = 2;
      }
    }
    const int kk = k + kstep - 1;
    if (kp != kk) { // symmetric interchange inside the trailing block, :85-103
      pivot_count += 1;
      for (int i = kp + 1; i < n; ++i)
        std::swap(a(i, kk), a(i, kp));
      for (int j = kk + 1; j < kp; ++j)
        std::swap(a(j, kk), a(kp, j));
      std::swap(a(kk, kk), a(kp, kp));
      if (kstep == 2)
        std::swap(a(k + 1, k), a(kp, k));
    }
    if (kstep == 1) { // :105-124   1x1 pivot, stored inverted
      const double d11 = 1.0 / a(k, k);
      a(k, k) = d11;
      const int m = n - k - 1;
      for (int j = 0; j < m; ++j) {
        const double d11xj = a(k + 1 + j, k) * d11;
        for (int i = j; i < m; ++i)
          a(k + 1 + i, k + 1 + j) -= d11xj * a(k + 1 + i, k);
      }
      for (int i = 0; i < m; ++i)
        a(k + 1 + i, k) *= d11;
    } else { // :125-151   2x2 pivot, inverse block stored
      const double d21_abs = std::fabs(a(k + 1, k));
      const double d21_inv = 1.0 / d21_abs;
      const double d11 = d21_inv * a(k + 1, k + 1);
      const double d22 = d21_inv * a(k, k);
      const double t = 1.0 / ((d11 * d22) - 1.0);
      const double d = t * d21_inv;
      const double d21 = a(k + 1, k) * d21_inv;
      a(k, k) = d11 * d;
      a(k + 1, k) = -d21 * d;
      a(k + 1, k + 1) = d22 * d;
      for (int j = k + 2; j < n; ++j) {
        const double wk = ((a(j, k) * d11) - (a(j, k + 1) * d21)) * d;
        const double wkp1 = ((a(j, k + 1) * d22) - (a(j, k) * d21)) * d;
        for (int i = j; i < n; ++i)
          a(i, j) -= a(i, k) * wk + a(i, k + 1) * wkp1;
        a(j, k) = wk;
        a(j, k + 1) = wkp1;
      }
    }
    if (kstep == 1) { // :158-163
      piv[k] = kp;
    } else {
      piv[k] = -1 - kp;
      piv[k + 1] = -1 - kp;
    }
    k += kstep;
  }
  return BK_SUCCESS;
}

// One panel of the blocked algorithm (LAPACK dlasyf, lower).
// Restates bunch_kaufman_in_place_one_block, core/bunchkaufman.hpp:171-344.
inline int bk_one_block(CM a, int n, CM w, int nb, int *piv, long &pivot_count,
                        int &processed_cols) {
  const double alpha = bk_alpha();
  pivot_count = 0;
  processed_cols = 0;
  if (n == 0)
    return BK_SUCCESS;
  int k = 0;
  while (k < n && k + 1 < nb) {
    // w(k:n,k) = a(k:n,k) - a(k:n,0:k) * w(k,0:k)^T          (:191-197)
    for (int i = k; i < n; ++i) {
      double s = a(i, k);
      for (int c = 0; c < k; ++c)
        s -= a(i, c) * w(k, c);
      w(i, k) = s;
    }
    int kstep = 1;
    const double abs_akk = std::fabs(w(k, k));
    int imax = 0;
    double colmax = 0.0;
    if (k + 1 < n) {
      colmax = std::fabs(w(k + 1, k));
      for (int i = k + 2; i < n; ++i) {
        const double v = std::fabs(w(i, k));
        if (v > colmax) {
          colmax = v;
          imax = i - (k + 1);
        }
      }
    }
    imax += k + 1;
    int kp;
    if (std::max(abs_akk, colmax) == 0.0)
      return BK_NUMERICAL_ISSUE;
    if (abs_akk >= colmax * alpha) {
      kp = k;
    } else {
      // column k+1 of w <- updated column imax of the matrix   (:216-226)
      for (int j = k; j < imax; ++j)
        w(j, k + 1) = a(imax, j);
      for (int i = imax; i < n; ++i)
        w(i, k + 1) = a(i, imax);
      for (int i = k; i < n; ++i) {
        double s = w(i, k + 1);
        for (int c = 0; c < k; ++c)
          s -= a(i, c) * w(imax, c);
        w(i, k + 1) = s;
      }
      double rowmax = 0.0; // :228-237
      for (int i = k; i < imax; ++i)
        rowmax = std::max(rowmax, std::fabs(w(i, k + 1)));
      for (int i = imax + 1; i < n; ++i)
        rowmax = std::max(rowmax, std::fabs(w(i, k + 1)));
      if (abs_akk >= (alpha * colmax) * (colmax / rowmax)) {
        kp = k;
      } else if (std::fabs(w(imax, k + 1)) >= alpha * rowmax) {
        kp = imax;
        for (int i = k; i < n; ++i)
          w(i, k) = w(i, k + 1);
      } else {
        kp = imax;
        kstep = 2;
      }
    }
    const int kk = k + kstep - 1;
    if (kp != kk) { // :253-264
      pivot_count += 1;
      a(kp, kp) = a(kk, kk);
      for (int j = kk + 1; j < kp; ++j)
        a(kp, j) = a(j, kk);
      for (int i = kp + 1; i < n; ++i)
        a(i, kp) = a(i, kk);
      for (int c = 0; c < k; ++c)
        std::swap(a(kk, c), a(kp, c));
      for (int c = 0; c < kk + 1; ++c)
        std::swap(w(kk, c), w(kp, c));
    }
    if (kstep == 1) { // :266-275
      for (int i = k; i < n; ++i)
        a(i, k) = w(i, k);
      const double d11 = 1.0 / w(k, k);
      a(k, k) = d11;
      for (int i = k + 1; i < n; ++i)
        a(i, k) *= d11;
    } else { // :276-303
      const double d21_abs = std::fabs(w(k + 1, k));
      const double d21_inv = 1.0 / d21_abs;
      const double d11 = d21_inv * w(k + 1, k + 1);
      const double d22 = d21_inv * w(k, k);
      const double t = 1.0 / ((d11 * d22) - 1.0);
      const double d21 = w(k + 1, k) * d21_inv;
      const double d = t * d21_inv;
      a(k, k) = d11 * d;
      a(k + 1, k) = -d21 * d;
      a(k + 1, k + 1) = d22 * d;
      for (int j = k + 2; j < n; ++j) {
        const double wk = ((w(j, k) * d11) - (w(j, k + 1) * d21)) * d;
        const double wkp1 = ((w(j, k + 1) * d22) - (w(j, k) * d21)) * d;
        a(j, k) = wk;
        a(j, k + 1) = wkp1;
      }
    }
    if (kstep == 1) {
      piv[k] = kp;
    } else {
      piv[k] = -1 - kp;
      piv[k + 1] = -1 - kp;
    }
    k += kstep;
  }
  // trailing update, lower triangle only                       (:319-323)
  for (int j = k; j < n; ++j)
    for (int i = j; i < n; ++i) {
      double s = a(i, j);
      for (int c = 0; c < k; ++c)
        s -= a(i, c) * w(j, c);
      a(i, j) = s;
    }
  int j = k - 1;
  processed_cols = k;
  while (true) { // apply the panel's interchanges to its own left part (:326-343)
    const int jj = j;
    int jp = piv[j];
    if (jp < 0) {
      jp = -1 - jp;
      j -= 1;
    }
    if (j == 0)
      return BK_SUCCESS;
    j -= 1;
    if (jp != jj)
      for (int c = 0; c < j + 1; ++c)
        std::swap(a(jp, c), a(jj, c));
    if (j == 0)
      return BK_SUCCESS;
  }
}

// Driver.  Restates bunch_kaufman_in_place, core/bunchkaufman.hpp:346-420.
inline int bk_in_place(CM a, int n, double *subdiag, int *piv, CM w,
                       int blocksize, long &pivot_count) {
  int k = 0;
  pivot_count = 0;
  while (k < n) {
    int kb = 0;
    long kpc = 0;
    CM ablk{&a(k, k), a.ld};
    int info;
    if (blocksize != 0 && blocksize < n - k) {
      info = bk_one_block(ablk, n - k, w, blocksize, piv + k, kpc, kb);
    } else {
      info = bk_unblocked(ablk, n - k, piv + k, kpc);
      kb = n - k;
    }
    if (info != BK_SUCCESS)
      return info;
    for (int j = k; j < k + kb; ++j) { // :378-389
      if (piv[j] >= 0)
        piv[j] += k;
      else
        piv[j] -= k;
    }
    pivot_count += kpc;
    k += kb;
  }
  k = 0; // :393-404  2x2 off-diagonals move to subdiag
  while (k < n) {
    if (piv[k] < 0) {
      subdiag[k] = a(k + 1, k);
      subdiag[k + 1] = 0.0;
      a(k + 1, k) = 0.0;
      k += 2;
    } else {
      subdiag[k] = 0.0;
      k += 1;
    }
  }
  k = 0; // :406-417  interchanges applied to the columns left of each pivot
  while (k < n) {
    int p = piv[k];
    if (p < 0) {
      p = -1 - p;
      for (int c = 0; c < k; ++c)
        std::swap(a(k + 1, c), a(p, c));
      k += 2;
    } else {
      for (int c = 0; c < k; ++c)
        std::swap(a(k, c), a(p, c));
      k += 1;
    }
  }
  return BK_SUCCESS;
}

// Restates bunch_kaufman_solve_in_place<false>, core/bunchkaufman.hpp:451-518.
inline void bk_solve_in_place(CCM L, const double *subdiag, const int *piv,
                              int n, SV x, int nrhs) {
  int k = 0;
  while (k < n) { // forward interchanges
    int p = piv[k];
    if (p < 0) {
      p = -1 - p;
      if (p != k + 1)
        for (int j = 0; j < nrhs; ++j)
          std::swap(x(k + 1, j), x(p, j));
      k += 2;
    } else {
      if (p != k)
        for (int j = 0; j < nrhs; ++j)
          std::swap(x(k, j), x(p, j));
      k += 1;
    }
  }
  // unit-lower solve                                           (:472)
  for (int c = 0; c < n; ++c)
    for (int i = c + 1; i < n; ++i) {
      const double l = L(i, c);
      if (l != 0.0)
        for (int j = 0; j < nrhs; ++j)
          x(i, j) -= l * x(c, j);
    }
  k = 0; // D^-1 (inverses are stored)                          (:474-502)
  while (k < n) {
    if (piv[k] < 0) {
      const double akp1k = subdiag[k];
      const double ak = L(k, k);
      const double akp1 = L(k + 1, k + 1);
      for (int j = 0; j < nrhs; ++j) {
        const double xk = x(k, j);
        const double xkp1 = x(k + 1, j);
        x(k, j) = xk * ak + xkp1 * akp1k;
        x(k + 1, j) = xkp1 * akp1 + xk * akp1k;
      }
      k += 2;
    } else {
      const double ak = L(k, k);
      for (int j = 0; j < nrhs; ++j)
        x(k, j) *= ak;
      k += 1;
    }
  }
  // unit-upper solve with L^T                                  (:504)
  for (int c = n - 1; c >= 0; --c)
    for (int i = c + 1; i < n; ++i) {
      const double l = L(i, c);
      if (l != 0.0)
        for (int j = 0; j < nrhs; ++j)
          x(c, j) -= l * x(i, j);
    }
  k = n; // reverse interchanges                                (:506-517)
  while (k > 0) {
    k -= 1;
    int p = piv[k];
    if (p < 0) {
      p = -1 - p;
      if (p != k)
        for (int j = 0; j < nrhs; ++j)
          std::swap(x(k, j), x(p, j));
      k -= 1;
    } else {
      if (p != k)
        for (int j = 0; j < nrhs; ++j)
          std::swap(x(k, j), x(p, j));
    }
  }
}

// Solver object.  Restates Eigen::BunchKaufman<MatrixXd, Lower>,
// core/bunchkaufman.hpp:521-676.
struct BunchKaufman {
  static constexpr int BlockSize = 32; // :531
  int n = 0;
  vecd mat;     // n x n column-major, L below the diagonal, D^-1 on it
  vecd subdiag; // 2x2 off-diagonals of D^-1
  std::vector<int> piv;
  vecd work; // n x blocksize
  int blocksize = 0;
  long pivot_count = 0;
  int info = BK_INVALID;

  BunchKaufman() = default;
  explicit BunchKaufman(int size) { resize(size); }
  void resize(int size) {
    n = size;
    mat.assign((size_t)n * n, 0.0);
    subdiag.assign(n, 0.0);
    piv.assign(n, 0);
    blocksize = n <= BlockSize ? 0 : BlockSize;
    work.assign((size_t)n * blocksize, 0.0);
  }
  // compute(): zero-fill, copy the LOWER triangle only, factor (:654-676)
  void compute(const double *a, int lda) {
    std::fill(mat.begin(), mat.end(), 0.0);
    std::fill(subdiag.begin(), subdiag.end(), 0.0);
    std::fill(piv.begin(), piv.end(), 0);
    std::fill(work.begin(), work.end(), 0.0);
    for (int j = 0; j < n; ++j)
      for (int i = j; i < n; ++i)
        mat[i + (size_t)j * n] = a[i + (size_t)j * lda];
    info = bk_in_place(CM{mat.data(), n}, n, subdiag.data(), piv.data(),
                       CM{work.data(), n}, blocksize, pivot_count);
  }
  // x is n x nrhs with arbitrary strides (row stride rs, column stride cs)
  void solveInPlace(double *x, int nrhs, std::ptrdiff_t rs,
                    std::ptrdiff_t cs) const {
    if (n == 0 || nrhs == 0)
      return;
    bk_solve_in_place(CCM{mat.data(), n}, subdiag.data(), piv.data(), n,
                      SV{x, rs, cs}, nrhs);
  }
};

// ---------------------------------------------------------------------------
// LQ problem types.   gar/lqr-problem.hpp:49-210
// ---------------------------------------------------------------------------
struct Knot { // LqrKnotTpl, lqr-problem.hpp:49-118; zero-initialised (.hxx:29-72)
  uint nx = 0, nu = 0, nc = 0, nx2 = 0, nth = 0;
  vecd Q, S, R, q, r; // Q nx*nx, S nx*nu, R nu*nu (column-major)
  vecd A, B, f;       // A nx2*nx, B nx2*nu
  vecd C, D, d;       // C nc*nx, D nc*nu
  vecd Gth, Gx, Gu, Gv, gamma;

  Knot() = default;
  Knot(uint nx_, uint nu_, uint nc_, uint nx2_, uint nth_ = 0)
      : nx(nx_), nu(nu_), nc(nc_), nx2(nx2_), nth(nth_) {
    Q.assign((size_t)nx * nx, 0.);
    S.assign((size_t)nx * nu, 0.);
    R.assign((size_t)nu * nu, 0.);
    q.assign(nx, 0.);
    r.assign(nu, 0.);
    A.assign((size_t)nx2 * nx, 0.);
    B.assign((size_t)nx2 * nu, 0.);
    f.assign(nx2, 0.);
    C.assign((size_t)nc * nx, 0.);
    D.assign((size_t)nc * nu, 0.);
    d.assign(nc, 0.);
    addParameterization(nth_);
  }
  // lqr-problem.hxx:233-242
  Knot &addParameterization(uint nth_) {
    nth = nth_;
    Gth.assign((size_t)nth * nth, 0.);
    Gx.assign((size_t)nx * nth, 0.);
    Gu.assign((size_t)nu * nth, 0.);
    Gv.assign((size_t)nc * nth, 0.);
    gamma.assign(nth, 0.);
    return *this;
  }
};

struct Problem { // LqrProblemTpl, lqr-problem.hpp:120-210
  std::vector<Knot> stages;
  vecd G0; // nc0 x nx0 column-major
  vecd g0; // nc0
  uint nc0 = 0;
  int horizon() const { return (int)stages.size() - 1; }
  uint ntheta() const { return stages[0].nth; }
  void addParameterization(uint nth) {
    for (auto &s : stages)
      s.addParameterization(nth);
  }
};

// ---------------------------------------------------------------------------
// Per-knot factor data.   gar/riccati-kernel.hpp:30-102, .hxx:12-50
// ---------------------------------------------------------------------------
struct CostToGo {
  vecd Vxx, vx, Vxt, Vtt, vt;
  CostToGo() = default;
  CostToGo(uint nx, uint nth)
      : Vxx((size_t)nx * nx, 0.), vx(nx, 0.), Vxt((size_t)nx * nth, 0.),
        Vtt((size_t)nth * nth, 0.), vt(nth, 0.) {}
};

struct StageFactor {
  uint nx, nu, nc, nx2, nth;
  vecd Qhat, Rhat, Shat, qhat, rhat; // column-major
  vecd AtV, BtV;                     // row-major nx x nx2, nu x nx2
  vecd Gxhat, Guhat;
  vecd ff;     // [k (nu); z (nc); a (nx2)]
  vecd fb;     // row-major (nu+nc+nx2) x nx  = [K; Z; Ahat]
  vecd fth;    // row-major (nu+nc+nx2) x nth
  vecd kktMat; // column-major (nu+nc)^2
  BunchKaufman kktChol;
  CostToGo vm;
  vecd vplus; // scratch (the reference allocates it per call, .hxx:217)

  StageFactor(uint nx_, uint nu_, uint nc_, uint nx2_, uint nth_)
      : nx(nx_), nu(nu_), nc(nc_), nx2(nx2_), nth(nth_),
        Qhat((size_t)nx * nx, 0.), Rhat((size_t)nu * nu, 0.),
        Shat((size_t)nx * nu, 0.), qhat(nx, 0.), rhat(nu, 0.),
        AtV((size_t)nx * nx2, 0.), BtV((size_t)nu * nx2, 0.),
        Gxhat((size_t)nx * nth, 0.), Guhat((size_t)nu * nth, 0.),
        ff(nu + nc + nx2, 0.), fb((size_t)(nu + nc + nx2) * nx, 0.),
        fth((size_t)(nu + nc + nx2) * nth, 0.),
        kktMat((size_t)(nu + nc) * (nu + nc), 0.), kktChol((int)(nu + nc)),
        vm(nx, nth), vplus(nx2, 0.) {}
};

// initial-stage saddle system.  riccati-kernel.hpp:110-120
struct Kkt0 {
  uint nx = 0, nc = 0, nth = 0;
  vecd mat; // (nx+nc)^2 column-major
  vecd ff;  // nx+nc
  vecd fth; // row-major (nx+nc) x nth
  BunchKaufman chol;
  Kkt0() = default;
  Kkt0(uint nx_, uint nc_, uint nth_)
      : nx(nx_), nc(nc_), nth(nth_), mat((size_t)(nx_ + nc_) * (nx_ + nc_), 0.),
        ff(nx_ + nc_, 0.), fth((size_t)(nx_ + nc_) * nth_, 0.),
        chol((int)(nx_ + nc_)) {}
};

// ---------------------------------------------------------------------------
// ProximalRiccatiKernel.   gar/riccati-kernel.hxx:105-377
// ---------------------------------------------------------------------------
struct Kernel {
  // terminalSolve, riccati-kernel.hxx:131-193
  static void terminalSolve(const Knot &m, double mueq, StageFactor &d) {
    const int nx = m.nx, nu = m.nu, nc = m.nc, nth = m.nth;
    const int n = nu + nc;
    CostToGo &vc = d.vm;
    double *kff = d.ff.data();
    double *zff = d.ff.data() + nu;
    RM K{d.fb.data(), nx};
    RM Z{d.fb.data() + (size_t)nu * nx, nx};
    RM Kth{d.fth.data(), nth};
    RM Zth{d.fth.data() + (size_t)nu * nth, nth};
    CCM C{m.C.data(), nc}, D{m.D.data(), nc}, S{m.S.data(), nx},
        R{m.R.data(), nu};
    if (nu == 0) { // :146-149
      for (int i = 0; i < nc; ++i) {
        for (int j = 0; j < nx; ++j)
          Z(i, j) = C(i, j) / mueq;
        zff[i] = m.d[i] / mueq;
        for (int j = 0; j < nth; ++j)
          Zth(i, j) = 0.0;
      }
    } else { // :151-173
      CM kkt{d.kktMat.data(), n};
      for (int j = 0; j < nu; ++j)
        for (int i = 0; i < nu; ++i)
          kkt(i, j) = R(i, j);
      for (int i = 0; i < nc; ++i)
        for (int j = 0; j < nu; ++j) {
          kkt(j, nu + i) = D(i, j);
          kkt(nu + i, j) = D(i, j);
        }
      for (int i = 0; i < nc; ++i)
        kkt(nu + i, nu + i) = -mueq;
      d.kktChol.compute(d.kktMat.data(), n);
      for (int i = 0; i < nu; ++i)
        kff[i] = -m.r[i];
      for (int i = 0; i < nc; ++i)
        zff[i] = -m.d[i];
      for (int i = 0; i < nu; ++i)
        for (int j = 0; j < nx; ++j)
          K(i, j) = -S(j, i);
      for (int i = 0; i < nc; ++i)
        for (int j = 0; j < nx; ++j)
          Z(i, j) = -C(i, j);
      d.kktChol.solveInPlace(d.ff.data(), 1, 1, n);
      d.kktChol.solveInPlace(d.fb.data(), nx, nx, 1);
      if (nth > 0) {
        CCM Gu{m.Gu.data(), nu};
        for (int i = 0; i < nu; ++i)
          for (int j = 0; j < nth; ++j)
            Kth(i, j) = -Gu(i, j);
        for (int i = 0; i < nc; ++i)
          for (int j = 0; j < nth; ++j)
            Zth(i, j) = 0.0;
        d.kktChol.solveInPlace(d.fth.data(), nth, nth, 1);
      }
    }
    // :175-183
    CM Vxx{vc.Vxx.data(), nx};
    CCM Q{m.Q.data(), nx};
    for (int j = 0; j < nx; ++j)
      for (int i = 0; i < nx; ++i) {
        double s = Q(i, j);
        for (int c = 0; c < nc; ++c)
          s += C(c, i) * Z(c, j);
        Vxx(i, j) = s;
      }
    for (int i = 0; i < nx; ++i) {
      double s = m.q[i];
      for (int c = 0; c < nc; ++c)
        s += C(c, i) * zff[c];
      vc.vx[i] = s;
    }
    if (nu > 0) {
      for (int j = 0; j < nx; ++j)
        for (int i = 0; i < nx; ++i) {
          double s = 0.0;
          for (int c = 0; c < nu; ++c)
            s += S(i, c) * K(c, j);
          Vxx(i, j) += s;
        }
      for (int i = 0; i < nx; ++i) {
        double s = 0.0;
        for (int c = 0; c < nu; ++c)
          s += S(i, c) * kff[c];
        vc.vx[i] += s;
      }
    }
    if (nth > 0) { // :185-192
      CCM Gx{m.Gx.data(), nx}, Gu{m.Gu.data(), nu}, Gth{m.Gth.data(), nth};
      CM Vxt{vc.Vxt.data(), nx}, Vtt{vc.Vtt.data(), nth};
      for (int j = 0; j < nth; ++j)
        for (int i = 0; i < nx; ++i) {
          double s = Gx(i, j);
          for (int c = 0; c < nu; ++c)
            s += K(c, i) * Gu(c, j);
          Vxt(i, j) = s;
        }
      for (int j = 0; j < nth; ++j)
        for (int i = 0; i < nth; ++i) {
          double s = Gth(i, j);
          for (int c = 0; c < nu; ++c)
            s += Gu(c, i) * Kth(c, j);
          Vtt(i, j) = s;
        }
      for (int i = 0; i < nth; ++i) {
        double s = m.gamma[i];
        for (int c = 0; c < nu; ++c)
          s += Gu(c, i) * kff[c];
        vc.vt[i] = s;
      }
    }
  }

  // stageKernelSolve, riccati-kernel.hxx:210-312.  Returns false where the
  // reference throws "Failed stage LDL factorization" (:239-241).
  static bool stageKernelSolve(const Knot &m, StageFactor &d, CostToGo &vn,
                               double mueq) {
    const int nx = m.nx, nu = m.nu, nc = m.nc, nx2 = m.nx2, nth = m.nth;
    const int n = nu + nc;
    // (1) symmetrise the NEXT Vxx from its lower triangle, in place (:216)
    CM Vn{vn.Vxx.data(), nx2};
    for (int j = 0; j < nx2; ++j)
      for (int i = j + 1; i < nx2; ++i)
        Vn(j, i) = Vn(i, j);
    // vplus = vx' + V' f   (:217-218)
    for (int i = 0; i < nx2; ++i) {
      double s = 0.0;
      for (int c = 0; c < nx2; ++c)
        s += Vn(i, c) * m.f[c];
      d.vplus[i] = vn.vx[i] + s;
    }
    CCM A{m.A.data(), nx2}, B{m.B.data(), nx2};
    RM AtV{d.AtV.data(), nx2}, BtV{d.BtV.data(), nx2};
    // (2) AtV = A^T V', BtV = B^T V'   (:220-221)
    for (int i = 0; i < nx; ++i)
      for (int j = 0; j < nx2; ++j) {
        double s = 0.0;
        for (int c = 0; c < nx2; ++c)
          s += A(c, i) * Vn(c, j);
        AtV(i, j) = s;
      }
    for (int i = 0; i < nu; ++i)
      for (int j = 0; j < nx2; ++j) {
        double s = 0.0;
        for (int c = 0; c < nx2; ++c)
          s += B(c, i) * Vn(c, j);
        BtV(i, j) = s;
      }
    // (3) hatted blocks   (:224-228)
    CM Qh{d.Qhat.data(), nx}, Rh{d.Rhat.data(), nu}, Sh{d.Shat.data(), nx};
    CCM Q{m.Q.data(), nx}, R{m.R.data(), nu}, S{m.S.data(), nx};
    for (int j = 0; j < nx; ++j)
      for (int i = 0; i < nx; ++i) {
        double s = 0.0;
        for (int c = 0; c < nx2; ++c)
          s += AtV(i, c) * A(c, j);
        Qh(i, j) = Q(i, j) + s;
      }
    for (int j = 0; j < nu; ++j)
      for (int i = 0; i < nu; ++i) {
        double s = 0.0;
        for (int c = 0; c < nx2; ++c)
          s += BtV(i, c) * B(c, j);
        Rh(i, j) = R(i, j) + s;
      }
    for (int j = 0; j < nu; ++j)
      for (int i = 0; i < nx; ++i) {
        double s = 0.0;
        for (int c = 0; c < nx2; ++c)
          s += AtV(i, c) * B(c, j);
        Sh(i, j) = S(i, j) + s;
      }
    for (int i = 0; i < nx; ++i) {
      double s = 0.0;
      for (int c = 0; c < nx2; ++c)
        s += A(c, i) * d.vplus[c];
      d.qhat[i] = m.q[i] + s;
    }
    for (int i = 0; i < nu; ++i) {
      double s = 0.0;
      for (int c = 0; c < nx2; ++c)
        s += B(c, i) * d.vplus[c];
      d.rhat[i] = m.r[i] + s;
    }
    // (4) reduced KKT matrix, symmetrised from lower, factored (:232-241)
    CM kkt{d.kktMat.data(), n};
    CCM C{m.C.data(), nc}, D{m.D.data(), nc};
    for (int j = 0; j < nu; ++j)
      for (int i = 0; i < nu; ++i)
        kkt(i, j) = Rh(i, j);
    for (int i = 0; i < nc; ++i)
      for (int j = 0; j < nu; ++j) {
        kkt(j, nu + i) = D(i, j);
        kkt(nu + i, j) = D(i, j);
      }
    for (int i = 0; i < nc; ++i)
      kkt(nu + i, nu + i) = -mueq;
    for (int j = 0; j < n; ++j)
      for (int i = j + 1; i < n; ++i)
        kkt(j, i) = kkt(i, j);
    d.kktChol.compute(d.kktMat.data(), n);
    if (d.kktChol.info != BK_SUCCESS)
      return false;
    // (5) right-hand sides and solve   (:243-262)
    double *kff = d.ff.data();
    double *zff = d.ff.data() + nu;
    double *yff = d.ff.data() + nu + nc;
    RM K{d.fb.data(), nx};
    RM Z{d.fb.data() + (size_t)nu * nx, nx};
    RM Aff{d.fb.data() + (size_t)(nu + nc) * nx, nx};
    for (int i = 0; i < nu; ++i)
      kff[i] = -d.rhat[i];
    for (int i = 0; i < nc; ++i)
      zff[i] = -m.d[i];
    for (int i = 0; i < nu; ++i)
      for (int j = 0; j < nx; ++j)
        K(i, j) = -Sh(j, i);
    for (int i = 0; i < nc; ++i)
      for (int j = 0; j < nx; ++j)
        Z(i, j) = -C(i, j);
    d.kktChol.solveInPlace(d.ff.data(), 1, 1, n);
    d.kktChol.solveInPlace(d.fb.data(), nx, nx, 1);
    // (6) closed-loop dynamics   (:266-267)
    for (int i = 0; i < nx2; ++i) {
      double s = 0.0;
      for (int c = 0; c < nu; ++c)
        s += B(i, c) * kff[c];
      yff[i] = m.f[i] + s;
    }
    for (int i = 0; i < nx2; ++i)
      for (int j = 0; j < nx; ++j) {
        double s = 0.0;
        for (int c = 0; c < nu; ++c)
          s += B(i, c) * K(c, j);
        Aff(i, j) = A(i, j) + s;
      }
    // (7) cost-to-go   (:270-277)
    CostToGo &vc = d.vm;
    CM Vxx{vc.Vxx.data(), nx};
    for (int j = 0; j < nx; ++j)
      for (int i = 0; i < nx; ++i) {
        double s = 0.0;
        for (int c = 0; c < nu; ++c)
          s += Sh(i, c) * K(c, j);
        double s2 = 0.0;
        for (int c = 0; c < nc; ++c)
          s2 += C(c, i) * Z(c, j);
        Vxx(i, j) = (Qh(i, j) + s) + s2;
      }
    for (int i = 0; i < nx; ++i) {
      double s = 0.0;
      for (int c = 0; c < nu; ++c)
        s += Sh(i, c) * kff[c];
      double s2 = 0.0;
      for (int c = 0; c < nc; ++c)
        s2 += C(c, i) * zff[c];
      vc.vx[i] = (d.qhat[i] + s) + s2;
    }
    if (nth > 0) { // (8) parametric terms   (:278-311)
      RM Kth{d.fth.data(), nth};
      RM Zth{d.fth.data() + (size_t)nu * nth, nth};
      RM Yth{d.fth.data() + (size_t)(nu + nc) * nth, nth};
      CCM Gx{m.Gx.data(), nx}, Gu{m.Gu.data(), nu}, Gv{m.Gv.data(), nc},
          Gth{m.Gth.data(), nth};
      CCM Vxtn{vn.Vxt.data(), nx2}, Vttn{vn.Vtt.data(), nth};
      CM Gxh{d.Gxhat.data(), nx}, Guh{d.Guhat.data(), nu};
      for (int j = 0; j < nth; ++j) {
        for (int i = 0; i < nx; ++i) {
          double s = 0.0;
          for (int c = 0; c < nx2; ++c)
            s += A(c, i) * Vxtn(c, j);
          Gxh(i, j) = Gx(i, j) + s;
        }
        for (int i = 0; i < nu; ++i) {
          double s = 0.0;
          for (int c = 0; c < nx2; ++c)
            s += B(c, i) * Vxtn(c, j);
          Guh(i, j) = Gu(i, j) + s;
        }
      }
      for (int i = 0; i < nu; ++i)
        for (int j = 0; j < nth; ++j)
          Kth(i, j) = -Guh(i, j);
      for (int i = 0; i < nc; ++i)
        for (int j = 0; j < nth; ++j)
          Zth(i, j) = -Gv(i, j);
      d.kktChol.solveInPlace(d.fth.data(), nth, nth, 1);
      for (int i = 0; i < nx2; ++i)
        for (int j = 0; j < nth; ++j) {
          double s = 0.0;
          for (int c = 0; c < nu; ++c)
            s += B(i, c) * Kth(c, j);
          Yth(i, j) = s;
        }
      for (int i = 0; i < nth; ++i) { // vt
        double s = m.gamma[i] + vn.vt[i];
        double s1 = 0.0;
        for (int c = 0; c < nu; ++c)
          s1 += Gu(c, i) * kff[c];
        double s2 = 0.0;
        for (int c = 0; c < nx2; ++c)
          s2 += Vxtn(c, i) * yff[c];
        vc.vt[i] = (s + s1) + s2;
      }
      CM Vxt{vc.Vxt.data(), nx}, Vtt{vc.Vtt.data(), nth};
      for (int j = 0; j < nth; ++j)
        for (int i = 0; i < nx; ++i) {
          double s1 = 0.0;
          for (int c = 0; c < nu; ++c)
            s1 += K(c, i) * Gu(c, j);
          double s2 = 0.0;
          for (int c = 0; c < nx2; ++c)
            s2 += Aff(c, i) * Vxtn(c, j);
          Vxt(i, j) = (Gx(i, j) + s1) + s2;
        }
      for (int j = 0; j < nth; ++j)
        for (int i = 0; i < nth; ++i) {
          double s1 = 0.0;
          for (int c = 0; c < nu; ++c)
            s1 += Gu(c, i) * Kth(c, j);
          double s2 = 0.0;
          for (int c = 0; c < nx2; ++c)
            s2 += Vxtn(c, i) * Yth(c, j);
          Vtt(i, j) = ((Gth(i, j) + Vttn(i, j)) + s1) + s2;
        }
    }
    return true;
  }

  // backwardImpl, riccati-kernel.hxx:105-129 (on the span [beg, end))
  static bool backwardImpl(const std::vector<Knot> &stages, size_t beg,
                           size_t end, double mueq,
                           std::vector<StageFactor> &datas) {
    if (end == beg)
      return true;
    const size_t N = end - beg - 1;
    terminalSolve(stages[beg + N], mueq, datas[beg + N]);
    if (N == 0)
      return true;
    bool ok = true;
    size_t t = N - 1;
    while (true) {
      ok &= stageKernelSolve(stages[beg + t], datas[beg + t],
                             datas[beg + t + 1].vm, mueq);
      if (t == 0)
        break;
      --t;
    }
    return ok;
  }

  // computeInitial, riccati-kernel.hxx:196-207
  static void computeInitial(vecd &x0, vecd &lbd0, const Kkt0 &kkt0,
                             const double *theta) {
    for (uint i = 0; i < kkt0.nx; ++i)
      x0[i] = kkt0.ff[i];
    for (uint i = 0; i < kkt0.nc; ++i)
      lbd0[i] = kkt0.ff[kkt0.nx + i];
    if (theta) {
      const int nth = kkt0.nth;
      for (uint i = 0; i < kkt0.nx; ++i) {
        double s = 0.0;
        for (int c = 0; c < nth; ++c)
          s += kkt0.fth[(size_t)i * nth + c] * theta[c];
        x0[i] += s;
      }
      for (uint i = 0; i < kkt0.nc; ++i) {
        double s = 0.0;
        for (int c = 0; c < nth; ++c)
          s += kkt0.fth[(size_t)(kkt0.nx + i) * nth + c] * theta[c];
        lbd0[i] += s;
      }
    }
  }

  // forwardImpl, riccati-kernel.hxx:315-377 (on the span [beg, end))
  static bool forwardImpl(const std::vector<Knot> &stages,
                          const std::vector<StageFactor> &datas, size_t beg,
                          size_t end, std::vector<vecd> &xs,
                          std::vector<vecd> &us, std::vector<vecd> &vs,
                          std::vector<vecd> &lbdas, const double *theta) {
    const size_t N = end - beg - 1;
    for (size_t tt = 0; tt <= N; ++tt) {
      const size_t t = beg + tt;
      const StageFactor &d = datas[t];
      const Knot &m = stages[t];
      const int nx = m.nx, nu = m.nu, nc = m.nc, nx2 = m.nx2, nth = m.nth;
      const double *x = xs[t].data();
      const double *fb = d.fb.data();
      const double *fth = d.fth.data();
      if (nu > 0) {
        for (int i = 0; i < nu; ++i) {
          double s = 0.0;
          for (int c = 0; c < nx; ++c)
            s += fb[(size_t)i * nx + c] * x[c];
          us[t][i] = d.ff[i] + s;
        }
      }
      for (int i = 0; i < nc; ++i) {
        double s = 0.0;
        for (int c = 0; c < nx; ++c)
          s += fb[(size_t)(nu + i) * nx + c] * x[c];
        vs[t][i] = d.ff[nu + i] + s;
      }
      if (nth > 0 && theta) {
        if (nu > 0)
          for (int i = 0; i < nu; ++i) {
            double s = 0.0;
            for (int c = 0; c < nth; ++c)
              s += fth[(size_t)i * nth + c] * theta[c];
            us[t][i] += s;
          }
        for (int i = 0; i < nc; ++i) {
          double s = 0.0;
          for (int c = 0; c < nth; ++c)
            s += fth[(size_t)(nu + i) * nth + c] * theta[c];
          vs[t][i] += s;
        }
      }
      if (tt == N)
        break;
      double *xn = xs[t + 1].data();
      for (int i = 0; i < nx2; ++i) {
        double s = 0.0;
        for (int c = 0; c < nx; ++c)
          s += fb[(size_t)(nu + nc + i) * nx + c] * x[c];
        xn[i] = d.ff[nu + nc + i] + s;
      }
      if (nth > 0 && theta)
        for (int i = 0; i < nx2; ++i) {
          double s = 0.0;
          for (int c = 0; c < nth; ++c)
            s += fth[(size_t)(nu + nc + i) * nth + c] * theta[c];
          xn[i] += s;
        }
      const CostToGo &vn = datas[t + 1].vm;
      for (int i = 0; i < nx2; ++i) {
        double s = 0.0;
        for (int c = 0; c < nx2; ++c)
          s += vn.Vxx[i + (size_t)c * nx2] * xn[c];
        lbdas[t + 1][i] = vn.vx[i] + s;
      }
      if (nth > 0 && theta)
        for (int i = 0; i < nx2; ++i) {
          double s = 0.0;
          for (int c = 0; c < nth; ++c)
            s += vn.Vxt[i + (size_t)c * nx2] * theta[c];
          lbdas[t + 1][i] += s;
        }
    }
    return true;
  }
};

// ---------------------------------------------------------------------------
// lqrInitializeSolution, gar/utils.hpp:114-142
// ---------------------------------------------------------------------------
struct Solution {
  std::vector<vecd> xs, us, vs, lbdas;
};
inline Solution lqrInitializeSolution(const Problem &p) {
  Solution s;
  const uint N = (uint)p.horizon();
  s.xs.resize(N + 1);
  s.us.resize(N + 1);
  s.vs.resize(N + 1);
  s.lbdas.resize(N + 1);
  s.lbdas[0].assign(p.nc0, 0.);
  for (uint i = 0; i <= N; ++i) {
    const Knot &kn = p.stages[i];
    s.xs[i].assign(kn.nx, 0.);
    s.us[i].assign(kn.nu, 0.);
    s.vs[i].assign(kn.nc, 0.);
    if (i == N)
      break;
    s.lbdas[i + 1].assign(kn.nx2, 0.);
  }
  if (p.stages.back().nu == 0)
    s.us.pop_back();
  return s;
}

// ---------------------------------------------------------------------------
// ProximalRiccatiSolver (serial).   gar/proximal-riccati.hxx:13-86
// ---------------------------------------------------------------------------
struct ProximalRiccatiSolver {
  const Problem *problem_;
  std::vector<StageFactor> datas;
  Kkt0 kkt0;
  vecd thGrad, thHess;

  explicit ProximalRiccatiSolver(const Problem &p)
      : problem_(&p), kkt0(p.stages[0].nx, p.nc0, p.ntheta()),
        thGrad(p.ntheta(), 0.), thHess((size_t)p.ntheta() * p.ntheta(), 0.) {
    const uint N = (uint)p.horizon();
    datas.reserve(N + 1);
    for (uint t = 0; t <= N; ++t) {
      const Knot &k = p.stages[t];
      datas.emplace_back(k.nx, k.nu, k.nc, k.nx2, k.nth);
    }
  }

  // proximal-riccati.hxx:34-62
  bool backward(double mueq) {
    bool ret =
        Kernel::backwardImpl(problem_->stages, 0, datas.size(), mueq, datas);
    StageFactor &d0 = datas[0];
    CostToGo &vinit = d0.vm;
    const int nx = kkt0.nx, nc0 = kkt0.nc, nth = kkt0.nth, n0 = nx + nc0;
    CM M{kkt0.mat.data(), n0};
    CCM G0{problem_->G0.data(), nc0};
    for (int j = 0; j < nx; ++j)
      for (int i = 0; i < nx; ++i)
        M(i, j) = vinit.Vxx[i + (size_t)j * nx];
    for (int i = 0; i < nc0; ++i)
      for (int j = 0; j < nx; ++j) {
        M(nx + i, j) = G0(i, j);
        M(j, nx + i) = G0(i, j);
      }
    for (int j = 0; j < nc0; ++j)
      for (int i = 0; i < nc0; ++i)
        M(nx + i, nx + j) = 0.0;
    kkt0.chol.compute(kkt0.mat.data(), n0);
    if (kkt0.chol.info != BK_SUCCESS)
      ret = false;
    for (int i = 0; i < nx; ++i)
      kkt0.ff[i] = -vinit.vx[i];
    for (int i = 0; i < nc0; ++i)
      kkt0.ff[nx + i] = -problem_->g0[i];
    kkt0.chol.solveInPlace(kkt0.ff.data(), 1, 1, n0);
    for (int i = 0; i < nx; ++i)
      for (int j = 0; j < nth; ++j)
        kkt0.fth[(size_t)i * nth + j] = -vinit.Vxt[i + (size_t)j * nx];
    for (int i = 0; i < nc0; ++i)
      for (int j = 0; j < nth; ++j)
        kkt0.fth[(size_t)(nx + i) * nth + j] = 0.0;
    kkt0.chol.solveInPlace(kkt0.fth.data(), nth, nth, 1);
    for (int i = 0; i < nth; ++i) { // :57-59
      double s = 0.0;
      for (int c = 0; c < nx; ++c)
        s += vinit.Vxt[c + (size_t)i * nx] * kkt0.ff[c];
      thGrad[i] = vinit.vt[i] + s;
    }
    for (int j = 0; j < nth; ++j)
      for (int i = 0; i < nth; ++i) {
        double s = 0.0;
        for (int c = 0; c < nx; ++c)
          s += vinit.Vxt[c + (size_t)i * nx] * kkt0.fth[(size_t)c * nth + j];
        thHess[i + (size_t)j * nth] = vinit.Vtt[i + (size_t)j * nth] + s;
      }
    return ret;
  }

  // proximal-riccati.hxx:65-76
  bool forward(Solution &s, const double *theta = nullptr) const {
    Kernel::computeInitial(s.xs[0], s.lbdas[0], kkt0, theta);
    return Kernel::forwardImpl(problem_->stages, datas, 0, datas.size(), s.xs,
                               s.us, s.vs, s.lbdas, theta);
  }

  // proximal-riccati.hxx:79-86 (rotate_vec_left(datas,0,1), utils/mpc-util.hpp:16-22)
  void cycleAppend(const Knot &knot) {
    std::rotate(datas.begin(), datas.begin() + 1, datas.end() - 1);
    const uint N = (uint)problem_->horizon() - 1;
    datas[N] = StageFactor(knot.nx, knot.nu, knot.nc, knot.nx2, knot.nth);
    std::fill(thGrad.begin(), thGrad.end(), 0.);
    std::fill(thHess.begin(), thHess.end(), 0.);
    std::fill(kkt0.mat.begin(), kkt0.mat.end(), 0.);
  }
};

// ---------------------------------------------------------------------------
// Symmetric block-tridiagonal solver.   gar/block-tridiagonal.hpp:52-182
// Blocks are column-major; dims[i] is the size of diagonal block i.
// ---------------------------------------------------------------------------
struct BlockTridiag {
  std::vector<int> dims;
  std::vector<vecd> sub, diag, super; // sub[i]: dims[i+1] x dims[i]
  std::vector<vecd> diagFacs, upFacs;
  std::vector<BunchKaufman> ldlt;
};

// c <- beta c + A b.   block-tridiagonal.hpp:52-75
inline void blockTridiagMatMul(const std::vector<int> &dims,
                               const std::vector<vecd> &sub,
                               const std::vector<vecd> &diag,
                               const std::vector<vecd> &super,
                               const std::vector<vecd> &b, std::vector<vecd> &c,
                               double beta) {
  const size_t N = super.size();
  auto gemv = [](const vecd &M, int r, int cdim, const vecd &x, vecd &y) {
    for (int i = 0; i < r; ++i) {
      double s = 0.0;
      for (int j = 0; j < cdim; ++j)
        s += M[i + (size_t)j * r] * x[j];
      y[i] += s;
    }
  };
  for (auto &ci : c)
    for (auto &v : ci)
      v *= beta;
  gemv(diag[0], dims[0], dims[0], b[0], c[0]);
  gemv(super[0], dims[0], dims[1], b[1], c[0]);
  for (size_t i = 1; i < N; ++i) {
    gemv(sub[i - 1], dims[i], dims[i - 1], b[i - 1], c[i]);
    gemv(diag[i], dims[i], dims[i], b[i], c[i]);
    gemv(super[i], dims[i], dims[i + 1], b[i + 1], c[i]);
  }
  gemv(sub[N - 1], dims[N], dims[N - 1], b[N - 1], c[N]);
  gemv(diag[N], dims[N], dims[N], b[N], c[N]);
}

// block-tridiagonal.hpp:82-138 (backward-looking U D U^T)
inline bool symmetricBlockTridiagSolve(const std::vector<int> &dims,
                                       std::vector<vecd> &sub,
                                       std::vector<vecd> &diag,
                                       const std::vector<vecd> &super,
                                       std::vector<vecd> &rhs,
                                       std::vector<BunchKaufman> &facs) {
  if (sub.size() != super.size() || diag.size() != super.size() + 1 ||
      rhs.size() != diag.size())
    return false;
  const size_t N = super.size();
  size_t i = N - 1;
  while (true) {
    BunchKaufman &ldl = facs[i + 1];
    const int di = dims[i], dn = dims[i + 1];
    ldl.compute(diag[i + 1].data(), dn);
    if (ldl.info != BK_SUCCESS)
      return false;
    ldl.solveInPlace(rhs[i + 1].data(), 1, 1, dn);
    const vecd &B = super[i]; // di x dn
    vecd &Cm = sub[i];        // dn x di
    for (int r = 0; r < di; ++r) {
      double s = 0.0;
      for (int c = 0; c < dn; ++c)
        s += B[r + (size_t)c * di] * rhs[i + 1][c];
      rhs[i][r] -= s;
    }
    ldl.solveInPlace(Cm.data(), di, 1, dn);
    for (int c = 0; c < di; ++c)
      for (int r = 0; r < di; ++r) {
        double s = 0.0;
        for (int m = 0; m < dn; ++m)
          s += B[r + (size_t)m * di] * Cm[m + (size_t)c * dn];
        diag[i][r + (size_t)c * di] -= s;
      }
    if (i == 0)
      break;
    i--;
  }
  {
    BunchKaufman &ldl = facs[0];
    ldl.compute(diag[0].data(), dims[0]);
    if (ldl.info != BK_SUCCESS)
      return false;
    ldl.solveInPlace(rhs[0].data(), 1, 1, dims[0]);
  }
  for (size_t k = 0; k < N; ++k) {
    const vecd &U = sub[k]; // dims[k+1] x dims[k]
    const int dn = dims[k + 1], di = dims[k];
    for (int r = 0; r < dn; ++r) {
      double s = 0.0;
      for (int c = 0; c < di; ++c)
        s += U[r + (size_t)c * dn] * rhs[k][c];
      rhs[k + 1][r] -= s;
    }
  }
  return true;
}

// block-tridiagonal.hpp:147-182
inline bool blockTridiagRefinementStep(const std::vector<int> &dims,
                                       const std::vector<vecd> &upT,
                                       const std::vector<vecd> &super,
                                       const std::vector<BunchKaufman> &facs,
                                       std::vector<vecd> &rhs) {
  const size_t N = super.size();
  size_t i = N - 1;
  while (true) {
    const int di = dims[i], dn = dims[i + 1];
    facs[i + 1].solveInPlace(rhs[i + 1].data(), 1, 1, dn);
    const vecd &B = super[i];
    for (int r = 0; r < di; ++r) {
      double s = 0.0;
      for (int c = 0; c < dn; ++c)
        s += B[r + (size_t)c * di] * rhs[i + 1][c];
      rhs[i][r] -= s;
    }
    if (i == 0)
      break;
    i--;
  }
  facs[0].solveInPlace(rhs[0].data(), 1, 1, dims[0]);
  for (size_t k = 0; k < N; ++k) {
    const vecd &U = upT[k];
    const int dn = dims[k + 1], di = dims[k];
    for (int r = 0; r < dn; ++r) {
      double s = 0.0;
      for (int c = 0; c < di; ++c)
        s += U[r + (size_t)c * dn] * rhs[k][c];
      rhs[k + 1][r] -= s;
    }
  }
  return true;
}

// ---------------------------------------------------------------------------
// ParallelRiccatiSolver.   gar/parallel-solver.hxx:23-258, .hpp:41-51
// `threaded` only selects whether legs run under OpenMP; the arithmetic is
// identical either way.
// ---------------------------------------------------------------------------
struct WorkRange {
  uint beg, end;
};
inline WorkRange get_work(uint horz, uint tid, uint nthreads) { // :23-28
  return {tid * (horz + 1) / nthreads, (tid + 1) * (horz + 1) / nthreads};
}

struct ParallelRiccatiSolver {
  Problem *problem_;
  uint numThreads_;
  std::vector<StageFactor> datas;
  BlockTridiag cond;
  std::vector<vecd> condRhs, condSol, condErr; // blocked by cond.dims
  double condensedThreshold = 1e-10;
  uint maxRefinementSteps = 5u;
  bool threaded = true;
  bool ok_ = true;

  ParallelRiccatiSolver(Problem &p, uint num_threads)
      : problem_(&p), numThreads_(num_threads) {
    ok_ = num_threads >= 2; // the reference throws (:42-46)
    if (ok_)
      initialize();
  }

  void initialize() { // :51-82 + initializeTridiagSystem :260-291
    const uint N = (uint)problem_->horizon();
    datas.clear();
    for (uint i = 0; i < numThreads_; ++i) {
      auto [i0, i1] = get_work(N, i, numThreads_);
      const bool last = i == numThreads_ - 1;
      const uint nth = problem_->stages[i1 - 1].nx2;
      for (uint t = i0; t < i1; ++t) {
        Knot &knot = problem_->stages[t];
        if (!last)
          knot.addParameterization(nth);
        datas.emplace_back(knot.nx, knot.nu, knot.nc, knot.nx2, knot.nth);
      }
    }
    cond = BlockTridiag{};
    cond.dims = {(int)problem_->nc0, (int)problem_->stages[0].nx};
    for (uint i = 0; i + 1 < numThreads_; ++i) {
      auto [i0, i1] = get_work(N, i, numThreads_);
      cond.dims.push_back((int)problem_->stages[i0].nx);
      cond.dims.push_back((int)problem_->stages[i1 - 1].nx);
    }
    const auto &dm = cond.dims;
    condRhs.clear();
    for (int dd : dm)
      condRhs.emplace_back(dd, 0.);
    condSol = condRhs;
    condErr = condRhs;
    for (size_t i = 0; i < dm.size(); ++i) {
      cond.diag.emplace_back((size_t)dm[i] * dm[i], 0.);
      cond.diagFacs.emplace_back((size_t)dm[i] * dm[i], 0.);
      cond.ldlt.emplace_back(dm[i]);
      if (i + 1 < dm.size()) {
        cond.super.emplace_back((size_t)dm[i] * dm[i + 1], 0.);
        cond.sub.emplace_back((size_t)dm[i + 1] * dm[i], 0.);
      }
    }
    for (size_t i = 0; i < dm.size(); ++i) // upFacs indexed like diag (:283-285)
      cond.upFacs.emplace_back(
          i + 1 < dm.size() ? (size_t)dm[i + 1] * dm[i] : (size_t)0, 0.);
  }

  void assembleCondensedSystem(double mudyn) { // :85-129
    const uint N = (uint)problem_->horizon();
    auto &dg = cond.diag;
    auto &sp = cond.super;
    auto &sb = cond.sub;
    const auto &dm = cond.dims;
    std::fill(dg[0].begin(), dg[0].end(), 0.);
    for (int i = 0; i < dm[0]; ++i)
      dg[0][i + (size_t)i * dm[0]] = -mudyn;
    sp[0] = problem_->G0;
    dg[1] = datas[0].vm.Vxx;
    if (sp.size() > 1)
      sp[1] = datas[0].vm.Vxt;
    for (uint i = 0; i + 1 < numThreads_; ++i) {
      auto [i0, i1] = get_work(N, i, numThreads_);
      const uint ip1 = i + 1;
      dg[2 * ip1] = datas[i0].vm.Vtt;
      dg[2 * ip1 + 1] = datas[i1].vm.Vxx;
      vecd &I = sp[2 * ip1];
      std::fill(I.begin(), I.end(), 0.);
      for (int c = 0; c < dm[2 * ip1]; ++c)
        I[c + (size_t)c * dm[2 * ip1]] = -1.0;
      if (ip1 + 1 < numThreads_)
        sp[2 * ip1 + 1] = datas[i1].vm.Vxt;
    }
    for (size_t i = 0; i < sb.size(); ++i) { // sub = super^T
      const int r = dm[i], c = dm[i + 1];
      for (int a = 0; a < r; ++a)
        for (int b = 0; b < c; ++b)
          sb[i][b + (size_t)a * c] = sp[i][a + (size_t)b * r];
    }
    for (int i = 0; i < dm[0]; ++i)
      condRhs[0][i] = -problem_->g0[i];
    for (int i = 0; i < dm[1]; ++i)
      condRhs[1][i] = -datas[0].vm.vx[i];
    for (uint i = 0; i + 1 < numThreads_; ++i) {
      auto [i0, i1] = get_work(N, i, numThreads_);
      const uint ip1 = i + 1;
      for (int c = 0; c < dm[2 * ip1]; ++c)
        condRhs[2 * ip1][c] = -datas[i0].vm.vt[c];
      for (int c = 0; c < dm[2 * ip1 + 1]; ++c)
        condRhs[2 * ip1 + 1][c] = -datas[i1].vm.vx[c];
    }
  }

  bool backward(double mueq) { // :132-206
    if (!ok_)
      return false;
    const uint N = (uint)problem_->horizon();
    for (uint i = 0; i + 1 < numThreads_; ++i) { // configure_knot :136-147
      const uint end = get_work(N, i, numThreads_).end;
      Knot &k = problem_->stages[end - 1];
      for (uint a = 0; a < k.nx; ++a)
        for (uint b = 0; b < k.nx2; ++b)
          k.Gx[a + (size_t)b * k.nx] = k.A[b + (size_t)a * k.nx2];
      for (uint a = 0; a < k.nu; ++a)
        for (uint b = 0; b < k.nx2; ++b)
          k.Gu[a + (size_t)b * k.nu] = k.B[b + (size_t)a * k.nx2];
      std::fill(k.Gth.begin(), k.Gth.end(), 0.);
      k.gamma = k.f;
    }
    bool legs_ok = true;
#if defined(_OPENMP)
#pragma omp parallel for num_threads(numThreads_) schedule(static, 1) if (threaded) reduction(&& : legs_ok)
#endif
    for (int i = 0; i < (int)numThreads_; ++i) {
      auto [beg, end] = get_work(N, (uint)i, numThreads_);
      legs_ok = legs_ok &&
                Kernel::backwardImpl(problem_->stages, beg, end, mueq, datas);
    }
    assembleCondensedSystem(0.0);
    condSol = condRhs;
    cond.diagFacs = cond.diag;
    cond.upFacs.assign(cond.sub.begin(), cond.sub.end());
    // factor on the copies, keep the originals (the reference solves on the
    // originals then swaps with the copies, :176-181 -- same end state)
    std::vector<vecd> facDiag = cond.diag, facSub = cond.sub;
    symmetricBlockTridiagSolve(cond.dims, facSub, facDiag, cond.super, condSol,
                               cond.ldlt);
    cond.diagFacs = facDiag;
    cond.upFacs = facSub;
    for (uint it = 0; it < maxRefinementSteps; ++it) { // :185-202
      blockTridiagMatMul(cond.dims, cond.sub, cond.diag, cond.super, condSol,
                         condErr, -1.0);
      double resdl = 0.0;
      for (auto &blk : condErr)
        for (auto &v : blk) {
          v *= -1.0;
          resdl = std::max(resdl, std::fabs(v));
        }
      if (resdl <= condensedThreshold)
        return legs_ok;
      blockTridiagRefinementStep(cond.dims, cond.upFacs, cond.super, cond.ldlt,
                                 condErr);
      for (size_t b = 0; b < condSol.size(); ++b)
        for (size_t c = 0; c < condSol[b].size(); ++c)
          condSol[b][c] += condErr[b][c];
      condErr = condRhs;
    }
    return legs_ok;
  }

  bool forward(Solution &s) const { // :209-243
    const uint N = (uint)problem_->horizon();
    for (uint i = 0; i < numThreads_; ++i) {
      const uint i0 = get_work(N, i, numThreads_).beg;
      s.lbdas[i0] = condSol[2 * i];
      s.xs[i0] = condSol[2 * i + 1];
    }
#if defined(_OPENMP)
#pragma omp parallel for num_threads(numThreads_) schedule(static, 1) if (threaded)
#endif
    for (int i = 0; i < (int)numThreads_; ++i) {
      auto [beg, end] = get_work(N, (uint)i, numThreads_);
      if ((uint)i + 1 < numThreads_)
        Kernel::forwardImpl(problem_->stages, datas, beg, end, s.xs, s.us,
                            s.vs, s.lbdas, s.lbdas[end].data());
      else
        Kernel::forwardImpl(problem_->stages, datas, beg, end, s.xs, s.us,
                            s.vs, s.lbdas, nullptr);
    }
    return true;
  }

  // parallel-solver.hpp:41-51   K0 -= Kth0 * subdiagonal[1].
  // NOTE (reference quirk, restated as written): the header comment there says
  // `subdiagonal` holds the U^T factors, but backward() swaps the factored
  // copies back out (parallel-solver.hxx:180-181), so at this point
  // `subdiagonal[1]` is the ORIGINAL block  Vxt_0^T, not upFacs[1].
  void collapseFeedback() {
    StageFactor &d = datas[0];
    const int nu = d.nu, nx = d.nx, nth = d.nth;
    const vecd &Up1t = cond.sub[1]; // dims[2] x dims[1] = nth x nx
    const int r = cond.dims[2];
    for (int i = 0; i < nu; ++i)
      for (int j = 0; j < nx; ++j) {
        double s = 0.0;
        for (int c = 0; c < nth; ++c)
          s += d.fth[(size_t)i * nth + c] * Up1t[c + (size_t)j * r];
        d.fb[(size_t)i * nx + j] -= s;
      }
  }

  void cycleAppend(const Knot &) { // :246-258
    problem_->addParameterization(0);
    initialize();
  }
};

// ---------------------------------------------------------------------------
// lqrComputeKktError.   gar/utils.hxx:88-182.  Returns {dyn, cstr, dual}.
// ---------------------------------------------------------------------------
struct KktError {
  double dyn, cstr, dual;
  double max() const { return std::max({dyn, cstr, dual}); }
};
inline KktError lqrComputeKktError(const Problem &p, const Solution &s,
                                   double mueq, const double *theta) {
  const uint N = (uint)p.horizon();
  double dynErr = 0., cstErr = 0., dualErr = 0.;
  auto inf = [](const vecd &v) {
    double m = 0.;
    for (double x : v)
      m = std::max(m, std::fabs(x));
    return m;
  };
  {
    const Knot &k0 = p.stages[0];
    vecd dyn(p.nc0);
    for (uint i = 0; i < p.nc0; ++i) {
      double a = p.g0[i];
      for (uint c = 0; c < k0.nx; ++c)
        a += p.G0[i + (size_t)c * p.nc0] * s.xs[0][c];
      dyn[i] = a;
    }
    dynErr = std::max(dynErr, inf(dyn));
  }
  for (uint t = 0; t <= N; ++t) {
    const Knot &k = p.stages[t];
    const int nx = k.nx, nu = k.nu, nc = k.nc, nx2 = k.nx2, nth = k.nth;
    const vecd &x = s.xs[t];
    const vecd &v = s.vs[t];
    const double *u = (nu > 0) ? s.us[t].data() : nullptr;
    vecd gx(nx, 0.), gu(nu, 0.), cst(nc, 0.);
    for (int i = 0; i < nc; ++i) {
      double a = k.d[i] - mueq * v[i];
      for (int c = 0; c < nx; ++c)
        a += k.C[i + (size_t)c * nc] * x[c];
      for (int c = 0; c < nu; ++c)
        a += k.D[i + (size_t)c * nc] * u[c];
      cst[i] = a;
    }
    for (int i = 0; i < nx; ++i) {
      double a = k.q[i];
      for (int c = 0; c < nx; ++c)
        a += k.Q[i + (size_t)c * nx] * x[c];
      for (int c = 0; c < nc; ++c)
        a += k.C[c + (size_t)i * nc] * v[c];
      for (int c = 0; c < nu; ++c)
        a += k.S[i + (size_t)c * nx] * u[c];
      gx[i] = a;
    }
    for (int i = 0; i < nu; ++i) {
      double a = k.r[i];
      for (int c = 0; c < nx; ++c)
        a += k.S[c + (size_t)i * nx] * x[c];
      for (int c = 0; c < nc; ++c)
        a += k.D[c + (size_t)i * nc] * v[c];
      for (int c = 0; c < nu; ++c)
        a += k.R[i + (size_t)c * nu] * u[c];
      gu[i] = a;
    }
    if (t == 0) {
      for (int i = 0; i < nx; ++i)
        for (uint c = 0; c < p.nc0; ++c)
          gx[i] += p.G0[c + (size_t)i * p.nc0] * s.lbdas[0][c];
    } else {
      for (int i = 0; i < nx; ++i)
        gx[i] -= s.lbdas[t][i];
    }
    if (t < N) {
      const vecd &xn = s.xs[t + 1];
      const vecd &ln = s.lbdas[t + 1];
      vecd dyn(nx2);
      for (int i = 0; i < nx2; ++i) {
        double a = k.f[i] - xn[i];
        for (int c = 0; c < nx; ++c)
          a += k.A[i + (size_t)c * nx2] * x[c];
        for (int c = 0; c < nu; ++c)
          a += k.B[i + (size_t)c * nx2] * u[c];
        dyn[i] = a;
      }
      for (int i = 0; i < nx; ++i)
        for (int c = 0; c < nx2; ++c)
          gx[i] += k.A[c + (size_t)i * nx2] * ln[c];
      for (int i = 0; i < nu; ++i)
        for (int c = 0; c < nx2; ++c)
          gu[i] += k.B[c + (size_t)i * nx2] * ln[c];
      dynErr = std::max(dynErr, inf(dyn));
    }
    if (theta && nth > 0) {
      for (int i = 0; i < nx; ++i)
        for (int c = 0; c < nth; ++c)
          gx[i] += k.Gx[i + (size_t)c * nx] * theta[c];
      for (int i = 0; i < nu; ++i)
        for (int c = 0; c < nth; ++c)
          gu[i] += k.Gu[i + (size_t)c * nu] * theta[c];
    }
    dualErr = std::max({dualErr, inf(gx), inf(gu)});
    cstErr = std::max(cstErr, inf(cst));
  }
  return {dynErr, cstErr, dualErr};
}

} // namespace gar_oracle
