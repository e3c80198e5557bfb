// riccati_group.cuh -- the batched Riccati sweep as a *group program*.
//
// One group of G lanes (G = 8, 16 or 32: a warp or a sub-warp) owns one problem
// instance and walks its horizon: terminal knot -> stage knots N-1..0 (backward
// factorisation) -> initial saddle system -> knots 0..N (forward rollout).
//
// What it computes is aligator's gar::ProximalRiccatiSolver::backward/forward
// (gar/proximal-riccati.hxx:34-76, gar/riccati-kernel.hxx:105-377) with the
// Bunch-Kaufman factorisation of core/bunchkaufman.hpp:22-169, 451-518; HOW is
// B200-first and shares nothing with the reference's Eigen code:
//
//  * lane-per-column mapping: lane j owns column j of  M = [A | B | f]
//    (nx x (nx+nu+1)); one knot step is
//        W  = V' M                         (V' broadcast from shared memory)
//        H  = [[Q S q],[S^T R r]] + [A B]^T W     ([A B] broadcast)
//        KKT = [[Rhat, D^T],[D, -mu I]] -> Bunch-Kaufman, cooperative, in smem
//        [K k; Z z] = -KKT^-1 [Shat^T rhat; C d]  (each lane solves its column)
//        [Ahat a] = [A f] + B [K k],   [Vxx vx] = [Qhat qhat] + [Shat C^T][K k; Z z]
//    so every lane keeps its columns in registers and the only shared-memory
//    traffic is warp-wide broadcast reads;
//  * the knot record (one contiguous [A|B|f|Q|S|R|q|r|C|D|d] block in HBM) is
//    staged by two TMA bulk copies (cp.async.bulk + mbarrier) issued one knot
//    ahead; V' never leaves the chip between knots;
//  * fp64 throughout (the reference's Scalar, context.hpp:9).
//
// The same source compiles for the host (g++), where a "group" is G std::threads
// and sync() is a std::barrier: tests/test_group_emulation.py runs the exact
// index arithmetic below on the CPU before any GPU time is spent.
#pragma once

#include <cmath>
#include <cstddef>
#include <cstdint>

#if defined(__CUDACC__)
#define AB2_HD __host__ __device__ __forceinline__
#define AB2_UNROLL _Pragma("unroll")
#else
#define AB2_HD inline
#define AB2_UNROLL
#endif

namespace ab2 {

// status bits (per instance)
enum : int { ST_OK = 0, ST_STAGE_FACTOR_FAILED = 1, ST_INIT_FACTOR_FAILED = 2 };

struct SweepParams {
  int N;     // horizon: N stage knots + 1 terminal knot
  int nct;   // terminal-knot constraint rows
  int nc0;   // initial-condition rows
  int batch; // instances
  double mueq;
  int do_bwd, do_fwd;
  // inputs
  const double *stage; // [batch][N][SREC_PAD]
  const double *term;  // [batch][TREC]   [Q | q | C | d]
  const double *G0;    // [batch][nc0*nx] column-major
  const double *g0;    // [batch][nc0]
  // backward outputs
  double *ff;   // [batch][N][NR]          [k; z; a]
  double *fb;   // [batch][N][NR*NX]       row-major [K; Z; Ahat]
  double *Vxx;  // [batch][N+1][NX*NX]     column-major
  double *vx;   // [batch][N+1][NX]
  double *ffT;  // [batch][nct]            terminal z
  double *fbT;  // [batch][nct*NX]         terminal Z (row-major)
  double *kkt0; // [batch][NX+nc0]         initial-stage solution [x0; lbda0]
  // forward outputs
  double *xs;    // [batch][N+1][NX]
  double *us;    // [batch][N][NU]
  double *vs;    // [batch][N][NC]
  double *vsT;   // [batch][nct]
  double *lbd0;  // [batch][nc0]
  double *lbdas; // [batch][N][NX]         lbda_1..lbda_N
  int *status;   // [batch]
};

// ---------------------------------------------------------------------------
// Compile-time shape of one kernel instantiation.
// ---------------------------------------------------------------------------
template <int NX_, int NU_, int NC_, int G_> struct Cfg {
  static constexpr int NX = NX_, NU = NU_, NC = NC_, G = G_;
  static constexpr int NCOL = NX + NU + 1; // columns of M = [A | B | f]
  static constexpr int NXU = NX + NU;      // rows of H
  static constexpr int NK = NU + NC;       // reduced KKT size
  static constexpr int NR = NU + NC + NX;  // rows of ff / fb
  static constexpr int FCOL = NX + NU;     // the lane that owns f / q,r / ff
  // stage record offsets (doubles) -- the reference's 11 buffers, concatenated
  static constexpr int OFF_A = 0;
  static constexpr int OFF_B = OFF_A + NX * NX;
  static constexpr int OFF_F = OFF_B + NX * NU;
  static constexpr int OFF_Q = OFF_F + NX;
  static constexpr int OFF_S = OFF_Q + NX * NX;
  static constexpr int OFF_R = OFF_S + NX * NU;
  static constexpr int OFF_QV = OFF_R + NU * NU;
  static constexpr int OFF_RV = OFF_QV + NX;
  static constexpr int OFF_C = OFF_RV + NU;
  static constexpr int OFF_D = OFF_C + NC * NX;
  static constexpr int OFF_DV = OFF_D + NC * NU;
  static constexpr int SREC = OFF_DV + NC;
  static constexpr int SREC_PAD = (SREC + 1) & ~1; // 16-byte granularity for TMA
  static constexpr int M_DBL = OFF_Q;              // [A|B|f]
  static constexpr int SPLIT = (M_DBL + 1) & ~1;   // part 0 = [0,SPLIT), part 1 = rest
  // shared-memory layout of one group (doubles)
  static constexpr int S_REC = 0;
  static constexpr int S_VN = S_REC + SREC_PAD;  // V' (symmetric, full) NX*NX
  static constexpr int S_VXN = S_VN + NX * NX;   // vx'
  static constexpr int S_KKT = S_VXN + NX;       // NK*NK column-major
  static constexpr int S_RHS = S_KKT + NK * NK;  // NK x (NX+1) row-major, unsolved
  static constexpr int S_SOL = S_RHS + NK * (NX + 1);
  static constexpr int S_DD = S_SOL + NK * (NX + 1);
  static constexpr int S_SD = S_DD + NK;
  static constexpr int S_X = S_SD + NK;    // forward state x_t (NX) + x_{t+1} (NX)
  static constexpr int S_INT = S_X + 2 * NX; // perm[NK], kind[NK] (ints)
  static constexpr int S_STAGE_END = S_INT + NK + 1;

  static_assert(NU >= 1, "stage knots need nu >= 1");
  static_assert(NCOL <= G, "lane-per-column mapping needs nx+nu+1 <= G");
  static_assert(NK <= G, "cooperative Bunch-Kaufman needs nu+nc <= G");
  static_assert(G == 8 || G == 16 || G == 32, "group size");

  // doubles of shared memory per group for a run with nc0 initial rows
  static AB2_HD int group_doubles(int nc0) {
    const int n0 = NX + nc0;
    const int k0 = n0 * n0 + 6 * n0 + 2; // K0, b, x, dd, sd, out, 2*n0 ints
    int m = S_STAGE_END > k0 ? S_STAGE_END : k0;
    return (m + 1) & ~1;
  }
  static AB2_HD int term_rec(int nct) { return NX * NX + NX + nct * NX + nct; }
};

// ---------------------------------------------------------------------------
// Cooperative Bunch-Kaufman (lower), n <= G, matrix in shared memory.
// Same pivot logic and arithmetic as bunch_kaufman_in_place_unblocked
// (core/bunchkaufman.hpp:46-151), restructured: lane i owns row i; row
// interchanges are applied to whole rows at once (so the "apply to previous
// columns" pass of :406-417 is not needed); D^-1 is kept in dd/sd.
//   kind[k] = 0: 1x1 pivot, 1: first row of a 2x2 pivot, 2: second row.
//   perm[i]  = original index now at position i  (the composed interchanges).
// Returns false where the reference reports NumericalIssue (:58-59).
// ---------------------------------------------------------------------------
template <class Ctx>
AB2_HD bool bk_factor_group(Ctx &ctx, double *a, const int lda, const int n,
                            double *dd, double *sd, int *perm, int *kind) {
  const double alpha = 0.6403882032022076; // (1+sqrt(17))/8
  const int lane = ctx.lane;
#define A_(i, j) a[(i) + (j) * lda]
  if (lane < n)
    perm[lane] = lane;
  ctx.sync();
  bool ok = true;
  int k = 0;
  while (k < n) {
    // ---- pivot search: every lane does it redundantly (broadcast reads) ----
    const double akk = A_(k, k);
    const double abs_akk = fabs(akk);
    int imax = k + 1;
    double colmax = 0.0, cval = 0.0;
    for (int i = k + 1; i < n; ++i) {
      const double v = A_(i, k);
      if (fabs(v) > colmax) {
        colmax = fabs(v);
        cval = v;
        imax = i;
      }
    }
    if (fmax(abs_akk, colmax) == 0.0) { // singular column: flag, neutral fill
      ok = false;
      ctx.sync();
      if (lane >= k && lane < n) {
        dd[lane] = 0.0;
        sd[lane] = 0.0;
        kind[lane] = 0;
        for (int j = k; j < lane; ++j)
          A_(lane, j) = 0.0;
      }
      break;
    }
    int kp = k, kstep = 1;
    double aii = akk;
    if (!(abs_akk >= colmax * alpha)) {
      double rowmax = 0.0;
      for (int j = k; j < imax; ++j)
        rowmax = fmax(rowmax, fabs(A_(imax, j)));
      for (int i = imax + 1; i < n; ++i)
        rowmax = fmax(rowmax, fabs(A_(i, imax)));
      aii = A_(imax, imax);
      if (abs_akk >= (alpha * colmax) * (colmax / rowmax)) {
        kp = k;
      } else if (fabs(aii) >= alpha * rowmax) {
        kp = imax;
      } else {
        kp = imax;
        kstep = 2;
      }
    }
    const int kk = k + kstep - 1;
    if (kp != kk) { // ---- symmetric interchange kk <-> kp, whole rows ----
      ctx.sync();   // everyone finished reading before anyone writes
      const int i = lane;
      if (i < n) {
        if (i > kp) {
          const double t = A_(i, kk);
          A_(i, kk) = A_(i, kp);
          A_(i, kp) = t;
        } else if (i > kk && i < kp) {
          const double t = A_(i, kk);
          A_(i, kk) = A_(kp, i);
          A_(kp, i) = t;
        } else if (i < k) {
          const double t = A_(kk, i);
          A_(kk, i) = A_(kp, i);
          A_(kp, i) = t;
        } else if (i == kk) {
          const double t = A_(kk, kk);
          A_(kk, kk) = A_(kp, kp);
          A_(kp, kp) = t;
          const int tp = perm[kk];
          perm[kk] = perm[kp];
          perm[kp] = tp;
        }
        if (kstep == 2 && i == k) { // column k of the 2x2 block: rows k+1 <-> kp
          const double t = A_(k + 1, k);
          A_(k + 1, k) = A_(kp, k);
          A_(kp, k) = t;
        }
      }
    }
    ctx.sync(); // S1
    if (kstep == 1) {
      // post-interchange pivot is akk (no swap) or aii (swap with imax)
      const double d11 = 1.0 / ((kp == k) ? akk : aii);
      double xi = 0.0;
      if (lane > k && lane < n) {
        xi = A_(lane, k);
        A_(lane, k) = xi * d11;
      }
      if (lane == k) {
        dd[k] = d11;
        sd[k] = 0.0;
        kind[k] = 0;
      }
      ctx.sync(); // S2
      if (lane > k && lane < n)
        for (int j = k + 1; j <= lane; ++j)
          A_(lane, j) -= A_(j, k) * xi;
    } else {
      // 2x2 pivot on (k, k+1): a11 = akk, a22 = aii, a21 = the column-k entry
      // that was at row imax (cval); identical whether or not rows moved.
      const double d21_abs = fabs(cval);
      const double d21_inv = 1.0 / d21_abs;
      const double d11 = d21_inv * aii;
      const double d22 = d21_inv * akk;
      const double t = 1.0 / ((d11 * d22) - 1.0);
      const double d = t * d21_inv;
      const double d21 = cval * d21_inv;
      double x0 = 0.0, x1 = 0.0;
      if (lane > k + 1 && lane < n) {
        x0 = A_(lane, k);
        x1 = A_(lane, k + 1);
        const double wk = ((x0 * d11) - (x1 * d21)) * d;
        const double wkp1 = ((x1 * d22) - (x0 * d21)) * d;
        A_(lane, k) = wk;
        A_(lane, k + 1) = wkp1;
      }
      if (lane == k) {
        dd[k] = d11 * d;
        sd[k] = -d21 * d;
        dd[k + 1] = d22 * d;
        sd[k + 1] = 0.0;
        kind[k] = 1;
        kind[k + 1] = 2;
        A_(k + 1, k) = 0.0;
      }
      ctx.sync(); // S2
      if (lane > k + 1 && lane < n)
        for (int j = k + 2; j <= lane; ++j)
          A_(lane, j) -= x0 * A_(j, k) + x1 * A_(j, k + 1);
    }
    ctx.sync(); // S3: trailing block complete before the next search
    k += kstep;
  }
  ctx.sync();
#undef A_
  return ok;
}

// Per-lane solve of one right-hand-side column with the factor above (n = NK
// compile-time; everything in registers, factor read by broadcast).
// Same sequence as bunch_kaufman_solve_in_place (core/bunchkaufman.hpp:451-518):
// interchanges, unit-lower solve, D^-1, unit-upper solve, inverse interchanges.
template <int NK>
AB2_HD void bk_solve_column(const double *a, const double *dd, const double *sd,
                            const int *perm, const int *kind, const double *rhs,
                            double *sol, const int stride, double (&x)[NK > 0 ? NK : 1]) {
  AB2_UNROLL
  for (int i = 0; i < NK; ++i)
    x[i] = rhs[perm[i] * stride];
  AB2_UNROLL
  for (int c = 0; c < NK; ++c) {
    AB2_UNROLL
    for (int i = c + 1; i < NK; ++i)
      x[i] -= a[i + c * NK] * x[c];
  }
  AB2_UNROLL
  for (int k = 0; k < NK; ++k) {
    const int kd = kind[k];
    if (kd == 0) {
      x[k] *= dd[k];
    } else if (kd == 1) {
      if (k + 1 < NK) {
        const double xk = x[k], xk1 = x[k + 1 < NK ? k + 1 : k];
        const double s = sd[k];
        x[k] = xk * dd[k] + xk1 * s;
        x[k + 1 < NK ? k + 1 : k] = xk1 * dd[k + 1 < NK ? k + 1 : k] + xk * s;
      }
    }
  }
  AB2_UNROLL
  for (int c = NK - 1; c >= 0; --c) {
    AB2_UNROLL
    for (int i = c + 1; i < NK; ++i)
      x[c] -= a[i + c * NK] * x[i];
  }
  AB2_UNROLL
  for (int i = 0; i < NK; ++i)
    sol[perm[i] * stride] = x[i];
  AB2_UNROLL
  for (int i = 0; i < NK; ++i)
    x[i] = sol[i * stride];
}

// Group-cooperative solve of ONE vector (runtime n <= G): lane i owns x[i].
// b: input (n), x: work/output in permuted order, out: un-permuted result.
template <class Ctx>
AB2_HD void bk_solve_vec_group(Ctx &ctx, const double *a, const int lda, const int n,
                               const double *dd, const double *sd, const int *perm,
                               const int *kind, const double *b, double *x, double *out) {
  const int lane = ctx.lane;
  if (lane < n)
    x[lane] = b[perm[lane]];
  for (int c = 0; c < n; ++c) { // column-oriented unit-lower solve
    ctx.sync();
    const double xc = x[c];
    if (lane > c && lane < n)
      x[lane] -= a[lane + c * lda] * xc;
  }
  ctx.sync();
  if (lane < n) {
    const int kd = kind[lane];
    if (kd == 0) {
      x[lane] *= dd[lane];
    } else if (kd == 1) {
      const double xk = x[lane], xk1 = x[lane + 1];
      const double s = sd[lane];
      x[lane] = xk * dd[lane] + xk1 * s;
      x[lane + 1] = xk1 * dd[lane + 1] + xk * s;
    }
  }
  for (int i = n - 1; i >= 1; --i) { // unit-upper solve with L^T
    ctx.sync();
    const double xi = x[i];
    if (lane < i)
      x[lane] -= a[i + lane * lda] * xi;
  }
  ctx.sync();
  if (lane < n)
    out[perm[lane]] = x[lane];
  ctx.sync();
}

// ---------------------------------------------------------------------------
// The sweep of one instance by one group.
// ---------------------------------------------------------------------------
template <class C, class Ctx>
AB2_HD void riccati_group_sweep(Ctx &ctx, const SweepParams &p, const int inst,
                                double *__restrict__ sm) {
  constexpr int NX = C::NX, NU = C::NU, NC = C::NC, NK = C::NK, NR = C::NR;
  constexpr int NXU = C::NXU, NCOL = C::NCOL, FCOL = C::FCOL;
  const int lane = ctx.lane;
  const int N = p.N;
  const int nct = p.nct, nc0 = p.nc0;
  const double mueq = p.mueq;

  double *rec = sm + C::S_REC;
  double *Vn = sm + C::S_VN;
  double *vxn = sm + C::S_VXN;
  double *kkt = sm + C::S_KKT;
  double *rhs0 = sm + C::S_RHS;
  double *sol = sm + C::S_SOL;
  double *dd = sm + C::S_DD;
  double *sd = sm + C::S_SD;
  double *xv = sm + C::S_X;
  int *perm = reinterpret_cast<int *>(sm + C::S_INT);
  int *kind = perm + NK;

  const double *stage_b = p.stage + (size_t)inst * N * C::SREC_PAD;
  double *ff_b = p.ff + (size_t)inst * N * NR;
  double *fb_b = p.fb + (size_t)inst * N * NR * NX;
  double *Vxx_b = p.Vxx + (size_t)inst * (N + 1) * NX * NX;
  double *vx_b = p.vx + (size_t)inst * (N + 1) * NX;
  int st = ST_OK;

  // lane classes
  const bool colA = lane < NX;                 // owns a state column
  const bool colB = lane >= NX && lane < NXU;  // owns a control column
  const bool colF = lane == FCOL;              // owns the affine column
  const bool active = lane < NCOL;
  const int jj = colF ? NX : lane; // column index in rhs0/sol (feedback cols, then ff)

  if (p.do_bwd) {
    // prefetch the last stage knot while the terminal knot is processed
    if (N > 0) {
      const double *src = stage_b + (size_t)(N - 1) * C::SREC_PAD;
      ctx.issue_copy(0, rec, src, C::SPLIT);
      ctx.issue_copy(1, rec + C::SPLIT, src + C::SPLIT, C::SREC_PAD - C::SPLIT);
    }
    // ---------------- terminal knot (nu = 0): riccati-kernel.hxx:146-149,175-183
    {
      const double *tr = p.term + (size_t)inst * C::term_rec(nct);
      const double *Qt = tr;
      const double *qt = tr + NX * NX;
      const double *Ct = qt + NX;            // nct x NX column-major
      const double *dt = Ct + (size_t)nct * NX;
      double *VN = Vxx_b + (size_t)N * NX * NX;
      double tv[NX];
      if (colA) {
        const int j = lane;
        for (int m = 0; m < nct; ++m) // Z = C / mu  (row-major nct x NX)
          p.fbT[(size_t)inst * nct * NX + (size_t)m * NX + j] = Ct[m + (size_t)j * nct] / mueq;
        AB2_UNROLL
        for (int i = 0; i < NX; ++i) {
          double s = Qt[i + j * NX];
          double acc = 0.0;
          for (int m = 0; m < nct; ++m)
            acc += Ct[m + (size_t)i * nct] * (Ct[m + (size_t)j * nct] / mueq);
          s += acc;
          tv[i] = s;
          if (i >= j) { // V' for the next step = lower triangle mirrored (:216)
            Vn[i * NX + j] = s;
            Vn[j * NX + i] = s;
          }
        }
      }
      if (colF) {
        for (int m = 0; m < nct; ++m)
          p.ffT[(size_t)inst * nct + m] = dt[m] / mueq;
        AB2_UNROLL
        for (int i = 0; i < NX; ++i) {
          double acc = 0.0;
          for (int m = 0; m < nct; ++m)
            acc += Ct[m + (size_t)i * nct] * (dt[m] / mueq);
          const double s = qt[i] + acc;
          vx_b[(size_t)N * NX + i] = s;
          vxn[i] = s;
        }
      }
      ctx.sync();
      if (colA) { // the step N-1 of the reference symmetrises datas[N].Vxx in place (A1)
        AB2_UNROLL
        for (int i = 0; i < NX; ++i)
          VN[i + lane * NX] = (N > 0) ? Vn[lane * NX + i] : tv[i];
      }
    }

    // ---------------- stage knots N-1 .. 0: riccati-kernel.hxx:210-277
    for (int t = N - 1; t >= 0; --t) {
      ctx.wait_copy(0);
      ctx.wait_copy(1);
      // own column of M = [A|B|f]
      double mcol[NX];
      AB2_UNROLL
      for (int k = 0; k < NX; ++k)
        mcol[k] = active ? rec[lane * NX + k] : 0.0;
      // (A) w = V' m_j  (+ vx' on the affine column: vplus = vx' + V' f, :217-218)
      double w[NX];
      AB2_UNROLL
      for (int i = 0; i < NX; ++i) {
        double s = 0.0;
        AB2_UNROLL
        for (int k = 0; k < NX; ++k)
          s += Vn[i * NX + k] * mcol[k];
        w[i] = s;
      }
      if (colF) {
        AB2_UNROLL
        for (int i = 0; i < NX; ++i)
          w[i] += vxn[i];
      }
      // (B) h = H0[:,j] + [A B]^T w      (:220-228 in one product)
      double h[NXU];
      {
        // H0 = [[Q S q],[S^T R r]]; per-lane base/stride so the code is uniform
        int base1, base2, stride2;
        if (colA) {
          base1 = C::OFF_Q + lane * NX;
          base2 = C::OFF_S + lane;
          stride2 = NX;
        } else if (colB) {
          base1 = C::OFF_S + (lane - NX) * NX;
          base2 = C::OFF_R + (lane - NX) * NU;
          stride2 = 1;
        } else {
          base1 = C::OFF_QV;
          base2 = C::OFF_RV;
          stride2 = 1;
        }
        AB2_UNROLL
        for (int i = 0; i < NX; ++i)
          h[i] = active ? rec[base1 + i] : 0.0;
        AB2_UNROLL
        for (int i = 0; i < NU; ++i)
          h[NX + i] = active ? rec[base2 + i * stride2] : 0.0;
      }
      AB2_UNROLL
      for (int i = 0; i < NXU; ++i) {
        double s = 0.0;
        AB2_UNROLL
        for (int k = 0; k < NX; ++k)
          s += rec[i * NX + k] * w[k];
        h[i] += s;
      }
      // (C) reduced KKT matrix (lower triangle) and right-hand sides (:232-257)
      if (colB) {
        const int c = lane - NX;
        AB2_UNROLL
        for (int r = 0; r < NU; ++r)
          kkt[r + c * NK] = h[NX + r]; // Rhat[r][c]; only r >= c is read
        AB2_UNROLL
        for (int m = 0; m < NC; ++m)
          kkt[NU + m + c * NK] = rec[C::OFF_D + c * NC + m];
      }
      if (lane < NC) { // (1,1) block: -mu on the diagonal, zeros below
        AB2_UNROLL
        for (int m = 0; m < NC; ++m)
          kkt[NU + m + (NU + lane) * NK] = (m == lane) ? -mueq : 0.0;
      }
      if (colA || colF) {
        AB2_UNROLL
        for (int r = 0; r < NU; ++r)
          rhs0[r * (NX + 1) + jj] = -h[NX + r]; // -Shat^T[:,j] / -rhat
        AB2_UNROLL
        for (int m = 0; m < NC; ++m)
          rhs0[(NU + m) * (NX + 1) + jj] =
              colF ? -rec[C::OFF_DV + m] : -rec[C::OFF_C + lane * NC + m];
      }
      ctx.sync();
      // part 1 of the record (Q..d) is consumed: fetch the next knot's
      if (t > 0)
        ctx.issue_copy(1, rec + C::SPLIT,
                       stage_b + (size_t)(t - 1) * C::SREC_PAD + C::SPLIT,
                       C::SREC_PAD - C::SPLIT);
      if (!bk_factor_group(ctx, kkt, NK, NK, dd, sd, perm, kind))
        st |= ST_STAGE_FACTOR_FAILED;
      // (D) solve, closed loop, cost-to-go (:259-277)
      double kz[NK > 0 ? NK : 1];
      double ahat[NX], vnew[NX];
      if (colA || colF) {
        bk_solve_column<NK>(kkt, dd, sd, perm, kind, rhs0 + jj, sol + jj, NX + 1, kz);
        AB2_UNROLL
        for (int i = 0; i < NX; ++i) { // [Ahat a] = [A f] + B [K k]
          double s = 0.0;
          AB2_UNROLL
          for (int c = 0; c < NU; ++c)
            s += rec[C::OFF_B + c * NX + i] * kz[c];
          ahat[i] = mcol[i] + s;
        }
      }
      ctx.sync();
      // part 0 ([A|B|f]) is consumed
      if (t > 0)
        ctx.issue_copy(0, rec, stage_b + (size_t)(t - 1) * C::SREC_PAD, C::SPLIT);
      if (colA || colF) {
        AB2_UNROLL
        for (int i = 0; i < NX; ++i) { // [Vxx vx] = [Qhat qhat] + [Shat C^T][K k; Z z]
          double s1 = 0.0;
          AB2_UNROLL
          for (int r = 0; r < NU; ++r)
            s1 -= rhs0[r * (NX + 1) + i] * kz[r]; // rhs0 holds -Shat^T
          double s2 = 0.0;
          AB2_UNROLL
          for (int m = 0; m < NC; ++m)
            s2 -= rhs0[(NU + m) * (NX + 1) + i] * kz[NU + m]; // and -C
          vnew[i] = (h[i] + s1) + s2;
        }
        // ---- outputs of knot t ----
        double *fbt = fb_b + (size_t)t * NR * NX;
        double *fft = ff_b + (size_t)t * NR;
        if (colA) {
          AB2_UNROLL
          for (int r = 0; r < NK; ++r)
            fbt[r * NX + lane] = kz[r];
          AB2_UNROLL
          for (int i = 0; i < NX; ++i)
            fbt[(NK + i) * NX + lane] = ahat[i];
          double *Vt = Vxx_b + (size_t)t * NX * NX;
          if (t == 0) { // datas[0].Vxx is left unsymmetrised (A1)
            AB2_UNROLL
            for (int i = 0; i < NX; ++i)
              Vt[i + lane * NX] = vnew[i];
          }
          AB2_UNROLL
          for (int i = 0; i < NX; ++i)
            if (i >= lane) { // V' = lower triangle mirrored (:216 of the next step)
              Vn[i * NX + lane] = vnew[i];
              Vn[lane * NX + i] = vnew[i];
            }
        } else {
          AB2_UNROLL
          for (int r = 0; r < NK; ++r)
            fft[r] = kz[r];
          AB2_UNROLL
          for (int i = 0; i < NX; ++i)
            fft[NK + i] = ahat[i];
          AB2_UNROLL
          for (int i = 0; i < NX; ++i) {
            vx_b[(size_t)t * NX + i] = vnew[i];
            vxn[i] = vnew[i];
          }
        }
      }
      ctx.sync();
      if (t > 0 && colA) { // symmetric Vxx_t, as the next step of the reference leaves it
        double *Vt = Vxx_b + (size_t)t * NX * NX;
        AB2_UNROLL
        for (int i = 0; i < NX; ++i)
          Vt[i + lane * NX] = Vn[lane * NX + i];
      }
    }

    // ---------------- initial stage: proximal-riccati.hxx:42-55 (nth = 0)
    {
      const int n0 = NX + nc0;
      double *K0 = sm;              // n0 x n0 column-major (overlays the stage area)
      double *b0 = K0 + n0 * n0;
      double *x0w = b0 + n0;
      double *dd0 = x0w + n0;
      double *sd0 = dd0 + n0;
      double *o0 = sd0 + n0;
      int *perm0 = reinterpret_cast<int *>(o0 + n0);
      int *kind0 = perm0 + n0;
      // pull what is needed out of the stage area before overwriting it
      double vcol[NX];
      double vx0 = 0.0;
      AB2_UNROLL
      for (int i = 0; i < NX; ++i)
        vcol[i] = colA ? Vn[i * NX + lane] : 0.0;
      if (lane < NX)
        vx0 = vxn[lane];
      ctx.sync();
      const double *G0 = p.G0 + (size_t)inst * nc0 * NX;
      const double *g0 = p.g0 + (size_t)inst * nc0;
      if (colA) {
        AB2_UNROLL
        for (int i = 0; i < NX; ++i)
          if (i >= lane)
            K0[i + lane * n0] = vcol[i]; // lower triangle of Vxx_0
        for (int m = 0; m < nc0; ++m)
          K0[NX + m + lane * n0] = G0[m + (size_t)lane * nc0];
        b0[lane] = -vx0;
      }
      for (int m = lane; m < nc0; m += C::G) {
        for (int m2 = m; m2 < nc0; ++m2)
          K0[NX + m2 + (NX + m) * n0] = 0.0;
        b0[NX + m] = -g0[m];
      }
      ctx.sync();
      if (!bk_factor_group(ctx, K0, n0, n0, dd0, sd0, perm0, kind0))
        st |= ST_INIT_FACTOR_FAILED;
      bk_solve_vec_group(ctx, K0, n0, n0, dd0, sd0, perm0, kind0, b0, x0w, o0);
      for (int i = lane; i < n0; i += C::G)
        p.kkt0[(size_t)inst * n0 + i] = o0[i];
      ctx.sync();
    }
    if (lane == 0)
      p.status[inst] = st;
  }

  // ---------------- forward rollout: riccati-kernel.hxx:196-207, 315-377
  if (p.do_fwd) {
    const int n0 = NX + nc0;
    const double *k0 = p.kkt0 + (size_t)inst * n0;
    double *xs_b = p.xs + (size_t)inst * (N + 1) * NX;
    double *us_b = p.us + (size_t)inst * N * NU;
    double *vs_b = p.vs + (size_t)inst * N * NC;
    double *lb_b = p.lbdas + (size_t)inst * N * NX;
    double *xc = xv;       // x_t
    double *xnx = xv + NX; // x_{t+1}
    ctx.sync();
    if (lane < NX) {
      const double v = k0[lane];
      xc[lane] = v;
      xs_b[lane] = v;
    }
    for (int m = lane; m < nc0; m += C::G)
      p.lbd0[(size_t)inst * nc0 + m] = k0[NX + m];
    ctx.sync();
    constexpr int RPL = (NR + C::G - 1) / C::G; // gain rows per lane
    for (int t = 0; t < N; ++t) {
      const double *fbt = fb_b + (size_t)t * NR * NX;
      const double *fft = ff_b + (size_t)t * NR;
      AB2_UNROLL
      for (int q = 0; q < RPL; ++q) {
        const int r = lane + q * C::G;
        if (r < NR) {
          double s = fft[r];
          AB2_UNROLL
          for (int c = 0; c < NX; ++c)
            s += fbt[r * NX + c] * xc[c];
          if (r < NU)
            us_b[(size_t)t * NU + r] = s;
          else if (r < NK)
            vs_b[(size_t)t * NC + (r - NU)] = s;
          else {
            xnx[r - NK] = s;
            xs_b[(size_t)(t + 1) * NX + (r - NK)] = s;
          }
        }
      }
      ctx.sync();
      if (lane < NX) { // lbda_{t+1} = vx_{t+1} + Vxx_{t+1} x_{t+1}
        const double *Vt = Vxx_b + (size_t)(t + 1) * NX * NX;
        double s = vx_b[(size_t)(t + 1) * NX + lane];
        AB2_UNROLL
        for (int c = 0; c < NX; ++c)
          s += Vt[lane + c * NX] * xnx[c];
        lb_b[(size_t)t * NX + lane] = s;
      }
      double *tmp = xc;
      xc = xnx;
      xnx = tmp;
      ctx.sync();
    }
    // terminal multipliers v_N = z + Z x_N
    for (int m = lane; m < nct; m += C::G) {
      double s = p.ffT[(size_t)inst * nct + m];
      for (int c = 0; c < NX; ++c)
        s += p.fbT[(size_t)inst * nct * NX + (size_t)m * NX + c] * xc[c];
      p.vsT[(size_t)inst * nct + m] = s;
    }
  }
}

} // namespace ab2
