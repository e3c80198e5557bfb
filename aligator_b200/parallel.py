"""Host side of ``gar::ParallelRiccatiSolver`` (gar/parallel-solver.hxx:23-258) on top of the
batched CUDA sweep: the horizon is cut into J+1 legs, every non-final leg becomes a PARAMETRIC
problem (theta = the co-state at the head of the next leg) that the CTA-per-instance kernel
solves -- all legs of all instances in one batch per distinct leg length -- and the condensed
symmetric block-tridiagonal system that couples the leg heads (block-tridiagonal.hpp:52-182) is
solved here, on the host, with the reference's Bunch-Kaufman pivoting (core/bunchkaufman.hpp).

The leg back end is pluggable (``backend(knots_of_one_leg, final, mueq) -> LegResult``): the product
back end is the CUDA batch (``CudaLegBackend``; it needs the parametric kernels, which are not in
this build yet -- it raises instead of falling back); the CPU tests plug the oracle in, which
exercises every line of the orchestration without a GPU.
"""
from __future__ import annotations

import numpy as np

from .lqr import LqrKnot, LqrProblem

ALPHA = (1.0 + np.sqrt(17.0)) / 8.0


# ---------------------------------------------------------------------------------------------
# Bunch-Kaufman (lower, unblocked): core/bunchkaufman.hpp:22-169 (LAPACK dsytf2 pivot logic,
# inverted pivots), solve :451-518.  Small blocks (nx, nc0): plain numpy loops.
# ---------------------------------------------------------------------------------------------
class BunchKaufman:
    def __init__(self, a):
        a = np.array(a, dtype=np.float64, order="F")
        n = a.shape[0]
        self.n, self.ok = n, True
        L = np.tril(a)
        piv = np.zeros(n, dtype=np.int64)
        sub = np.zeros(n)
        k = 0
        while k < n:
            kstep, kp = 1, k
            abs_akk = abs(L[k, k])
            if k + 1 < n:
                col = np.abs(L[k + 1:, k])
                imax = k + 1 + int(np.argmax(col))
                colmax = col[imax - k - 1]
            else:
                imax, colmax = k, 0.0
            if max(abs_akk, colmax) == 0.0:
                self.ok = False
                piv[k] = k
                k += 1
                continue
            if not (abs_akk >= colmax * ALPHA):
                rowmax = 0.0
                if imax > k:
                    rowmax = max(rowmax, np.max(np.abs(L[imax, k:imax])))
                if imax + 1 < n:
                    rowmax = max(rowmax, np.max(np.abs(L[imax + 1:, imax])))
                if abs_akk >= (ALPHA * colmax) * (colmax / rowmax):
                    kp = k
                elif abs(L[imax, imax]) >= ALPHA * rowmax:
                    kp = imax
                else:
                    kp, kstep = imax, 2
            kk = k + kstep - 1
            if kp != kk:  # symmetric interchange kk <-> kp on the trailing part
                t = L[kp + 1:, kk].copy()
                L[kp + 1:, kk] = L[kp + 1:, kp]
                L[kp + 1:, kp] = t
                for i in range(kk + 1, kp):
                    L[i, kk], L[kp, i] = L[kp, i], L[i, kk]
                L[kk, kk], L[kp, kp] = L[kp, kp], L[kk, kk]
                if kstep == 2:
                    L[k + 1, k], L[kp, k] = L[kp, k], L[k + 1, k]
            if kstep == 1:
                d11 = 1.0 / L[k, k]
                x = L[k + 1:, k].copy()
                for j in range(k + 1, n):
                    L[j:, j] -= (x[j - k - 1] * d11) * x[j - k - 1:]
                L[k + 1:, k] = x * d11
                L[k, k] = d11
                piv[k] = kp
            else:
                d21_abs = abs(L[k + 1, k])
                d21_inv = 1.0 / d21_abs
                d11 = d21_inv * L[k + 1, k + 1]
                d22 = d21_inv * L[k, k]
                t = 1.0 / ((d11 * d22) - 1.0)
                d21 = L[k + 1, k] * d21_inv
                dm = t * d21_inv
                L[k, k] = d11 * dm
                sub[k] = -d21 * dm
                L[k + 1, k + 1] = d22 * dm
                L[k + 1, k] = 0.0
                for j in range(k + 2, n):
                    wk = ((L[j, k] * d11) - (L[j, k + 1] * d21)) * dm
                    wkp1 = ((L[j, k + 1] * d22) - (L[j, k] * d21)) * dm
                    L[j:, j] -= L[j:, k] * wk + L[j:, k + 1] * wkp1
                    L[j, k], L[j, k + 1] = wk, wkp1
                piv[k] = -(kp + 1)
                piv[k + 1] = -(kp + 1)
            k += kstep
        k = 0
        while k < n:  # interchanges applied to the columns left of each pivot (:406-417)
            pk = piv[k]
            if pk < 0:
                pk = -1 - pk
                if pk != k + 1:
                    L[[k + 1, pk], :k] = L[[pk, k + 1], :k]
                k += 2
            else:
                if pk != k:
                    L[[k, pk], :k] = L[[pk, k], :k]
                k += 1
        self.L, self.piv, self.sub = L, piv, sub

    def solve(self, b):
        """In-place semantics of solveInPlace: returns A^-1 b (b: vector or matrix of columns)."""
        x = np.array(b, dtype=np.float64).reshape(self.n, -1).copy()
        n, L = self.n, self.L
        k = 0
        while k < n:  # interchanges + unit-lower solve interleaved as in :451-478
            if self.piv[k] >= 0:
                kp = self.piv[k]
                if kp != k:
                    x[[k, kp]] = x[[kp, k]]
                k += 1
            else:
                kp = -self.piv[k] - 1
                if kp != k + 1:
                    x[[k + 1, kp]] = x[[kp, k + 1]]
                k += 2
        for c in range(n):
            x[c + 1:] -= np.outer(L[c + 1:, c], x[c])
        k = 0
        while k < n:
            if self.piv[k] >= 0:
                x[k] *= L[k, k]
                k += 1
            else:
                xk, xk1 = x[k].copy(), x[k + 1].copy()
                x[k] = xk * L[k, k] + xk1 * self.sub[k]
                x[k + 1] = xk1 * L[k + 1, k + 1] + xk * self.sub[k]
                k += 2
        for c in range(n - 1, -1, -1):
            x[c] -= L[c + 1:, c] @ x[c + 1:]
        k = n - 1
        while k >= 0:  # inverse interchanges, last to first
            if self.piv[k] >= 0:
                kp = self.piv[k]
                if kp != k:
                    x[[k, kp]] = x[[kp, k]]
                k -= 1
            else:
                kp = -self.piv[k] - 1
                if kp != k:
                    x[[k, kp]] = x[[kp, k]]
                k -= 2
        return x.reshape(np.shape(b))


# ---------------------------------------------------------------------------------------------
# Symmetric block-tridiagonal system: block-tridiagonal.hpp:52-182
# ---------------------------------------------------------------------------------------------
def block_tridiag_matmul(sub, diag, sup, b):
    """c = A b for A = tridiag(sub, diag, sup) (:52-75)."""
    n = len(diag)
    c = [diag[i] @ b[i] for i in range(n)]
    for i in range(n - 1):
        c[i] = c[i] + sup[i] @ b[i + 1]
        c[i + 1] = c[i + 1] + sub[i] @ b[i]
    return c


def block_tridiag_solve(sub, diag, sup, rhs):
    """Backward-looking block U D U^T (:82-138).  Returns (ok, solution, facs, upT) where upT are
    the factored sub-diagonal blocks the refinement step re-uses."""
    n = len(diag)
    diag = [d.copy() for d in diag]
    upT = [s.copy() for s in sub]
    x = [r.copy() for r in rhs]
    facs = [None] * n
    for i in range(n - 2, -1, -1):
        f = BunchKaufman(diag[i + 1])
        if not f.ok:
            return False, x, facs, upT
        facs[i + 1] = f
        x[i + 1] = f.solve(x[i + 1])
        x[i] = x[i] - sup[i] @ x[i + 1]
        upT[i] = f.solve(upT[i])
        diag[i] = diag[i] - sup[i] @ upT[i]
    f = BunchKaufman(diag[0])
    if not f.ok:
        return False, x, facs, upT
    facs[0] = f
    x[0] = f.solve(x[0])
    for k in range(n - 1):
        x[k + 1] = x[k + 1] - upT[k] @ x[k]
    return True, x, facs, upT


def block_tridiag_refine(upT, sup, facs, rhs):
    """One refinement solve with the existing factors (:147-182)."""
    n = len(facs)
    x = [r.copy() for r in rhs]
    for i in range(n - 2, -1, -1):
        x[i + 1] = facs[i + 1].solve(x[i + 1])
        x[i] = x[i] - sup[i] @ x[i + 1]
    x[0] = facs[0].solve(x[0])
    for k in range(n - 1):
        x[k + 1] = x[k + 1] - upT[k] @ x[k]
    return x


# ---------------------------------------------------------------------------------------------
# The solver
# ---------------------------------------------------------------------------------------------
def get_work(horz, tid, nthreads):
    """Knots [beg, end) of leg `tid` (:23-28)."""
    return tid * (horz + 1) // nthreads, (tid + 1) * (horz + 1) // nthreads


class LegResult:
    """What the condensed system and the rollout need from one leg's backward pass: value
    function at the leg head and the per-knot gains (incl. the parametric ones)."""

    def __init__(self, Vxx, vx, Vxt, Vtt, vt, ff, fb, fth, Vxx_all, vx_all, Vxt_all):
        self.Vxx, self.vx, self.Vxt, self.Vtt, self.vt = Vxx, vx, Vxt, Vtt, vt
        self.ff, self.fb, self.fth = ff, fb, fth  # lists over the leg's knots
        self.Vxx_all, self.vx_all, self.Vxt_all = Vxx_all, vx_all, Vxt_all


class ParallelRiccatiSolver:
    """Mirror of ``gar::ParallelRiccatiSolver`` (parallel-solver.hpp:25-106).  Like the reference it
    re-parameterises the problem in place (initialize(), :51-60) and throws for fewer than two legs."""

    def __init__(self, problem: LqrProblem, num_threads: int, backend):
        if num_threads < 2:
            raise RuntimeError("numThreads should be greater than or equal to 2")  # :42-46
        self.problem, self.J1, self.backend = problem, int(num_threads), backend
        self.condensedThreshold, self.maxRefinementSteps = 1e-10, 5
        N = problem.horizon
        for i in range(self.J1 - 1):  # initialize(): every knot of a non-final leg gets nth = nx2 of its last knot
            i0, i1 = get_work(N, i, self.J1)
            nth = problem.stages[i1 - 1].nx2
            for t in range(i0, i1):
                problem.stages[t].addParameterization(nth)
        self.legs = [None] * self.J1

    def backward(self, mueq):
        p, N, J1 = self.problem, self.problem.horizon, self.J1
        for i in range(J1 - 1):  # configure the last knot of each non-final leg (:136-147)
            k = p.stages[get_work(N, i, J1)[1] - 1]
            k.Gx[...] = k.A.T
            k.Gu[...] = k.B.T
            k.Gth[...] = 0.0
            k.gamma[...] = k.f
        for i in range(J1):  # the legs (the reference: one OpenMP thread each, :150-164)
            i0, i1 = get_work(N, i, J1)
            self.legs[i] = self.backend(p.stages[i0:i1], final=(i == J1 - 1), mueq=mueq)
        # condensed system (:85-129): unknowns [lbda_0, x_0, (lbda_{i1}, x_{i1}) per interior head]
        nc0, nx0 = p.nc0, p.stages[0].nx
        diag = [-0.0 * np.eye(nc0), self.legs[0].Vxx.copy()]
        sup = [np.array(p.G0, dtype=np.float64).reshape(nc0, nx0)]
        rhs = [-np.asarray(p.g0, dtype=np.float64), -self.legs[0].vx]
        for i in range(J1 - 1):
            sup.append(self.legs[i].Vxt.copy())
            diag.append(self.legs[i].Vtt.copy())
            diag.append(self.legs[i + 1].Vxx.copy())
            sup.append(-np.eye(self.legs[i].Vtt.shape[0]))
            rhs.append(-self.legs[i].vt)
            rhs.append(-self.legs[i + 1].vx)
        sub = [s.T.copy() for s in sup]
        self.cond = (sub, diag, sup, rhs)
        ok, x, facs, upT = block_tridiag_solve(sub, diag, sup, rhs)
        for _ in range(self.maxRefinementSteps):  # :185-202
            Ax = block_tridiag_matmul(sub, diag, sup, x)
            err = [r - a for r, a in zip(rhs, Ax)]
            if max((np.max(np.abs(e)) if e.size else 0.0) for e in err) <= self.condensedThreshold:
                break
            dx = block_tridiag_refine(upT, sup, facs, err)
            x = [a + b for a, b in zip(x, dx)]
        self.condSol = x
        return ok

    def forward(self, xs, us, vs, lbdas):
        """Scatter the head states / co-states, then roll every leg out with theta = the co-state at
        the head of the next leg (:209-243)."""
        p, N, J1 = self.problem, self.problem.horizon, self.J1
        for i in range(J1):
            i0 = get_work(N, i, J1)[0]
            lbdas[i0][:] = self.condSol[2 * i]
            xs[i0][:] = self.condSol[2 * i + 1]
        for i in range(J1):
            i0, i1 = get_work(N, i, J1)
            leg = self.legs[i]
            theta = lbdas[i1] if i + 1 < J1 else None
            for t in range(i0, i1):  # forwardImpl on [i0, i1) (riccati-kernel.hxx:315-377)
                k = p.stages[t]
                j = t - i0
                nu, nc = k.nu, k.nc
                full = leg.ff[j] + leg.fb[j] @ xs[t]
                if theta is not None and leg.fth[j].size:
                    full = full + leg.fth[j] @ theta
                if nu:
                    us[t][:] = full[:nu]
                vs[t][:] = full[nu:nu + nc]
                if t == N:
                    break
                if t + 1 < i1:  # (the head of the next leg keeps its consensus values: the reference
                    # overwrites them from both sides, a benign race in its threaded mode)
                    xs[t + 1][:] = full[nu + nc:]
                    lam = leg.vx_all[j + 1] + leg.Vxx_all[j + 1] @ xs[t + 1]
                    if theta is not None and leg.Vxt_all[j + 1].size:
                        lam = lam + leg.Vxt_all[j + 1] @ theta
                    lbdas[t + 1][:] = lam
        return True


class CudaLegBackend:
    """The product leg back end: one leg = one (parametric) problem on the CTA-per-instance kernel.
    A non-final leg gets a dummy terminal knot with a zero value function, which makes the kernel's
    stage step on the leg's last knot identical to the reference's ``terminalSolve`` with controls
    (riccati-kernel.hxx:151-192: Rhat = R, Shat = S, Vxt = Gx + K^T Gu, ...)."""

    def __init__(self, device=0):
        self.device = device

    def __call__(self, stages, final, mueq):
        from . import gar
        if not hasattr(gar.lib(), "ab2_gar_create_parametric"):
            # no CPU fallback: the legs need the parametric kernels (branch `parametric`)
            raise gar.GarError("this build of libaligator_b200_gar has no parametric (nth > 0) kernels")
        stages = list(stages)
        if final:
            knots, term = stages[:-1], stages[-1]
        else:
            last = stages[-1]
            knots, term = stages, LqrKnot(last.nx2, 0, 0, last.nx2, last.nth)  # zeros: V' = 0
        n = len(knots)
        nx = term.nx
        k0 = knots[0] if n else None
        nu, nc = (k0.nu, k0.nc) if n else (1, 0)
        nth = term.nth
        prob = LqrProblem(knots + [term], 0)
        stage, tr, G0, g0 = gar.pack_problems([prob])
        s = gar.CudaRiccatiBatch(nx, nu, nc, term.nc, 0, n, 1, self.device, nth=nth)
        s.set_problem(stage, tr, G0, g0)
        s.backward(mueq)
        FF, FB, V, vx = s.get(gar.OUT_FF)[0], s.get(gar.OUT_FB)[0], s.get(gar.OUT_VXX)[0], s.get(gar.OUT_VX)[0]
        if nth:
            FTH, VXT, VTT, VT = (s.get(gar.OUT_FTH)[0], s.get(gar.OUT_VXT)[0], s.get(gar.OUT_VTT)[0],
                                 s.get(gar.OUT_VT)[0])
        else:
            z = np.zeros
            FTH, VXT, VTT, VT = z((n, nu + nc + nx, 0)), z((n + 1, nx, 0)), z((n + 1, 0, 0)), z((n + 1, 0))
        ff, fb, fth = [FF[t] for t in range(n)], [FB[t] for t in range(n)], [FTH[t] for t in range(n)]
        if final:  # the terminal knot's gains are [z], [Z]
            ff.append(s.get(gar.OUT_FFT)[0])
            fb.append(s.get(gar.OUT_FBT)[0])
            fth.append(np.zeros((term.nc, 0)))
        Vs = [V[t] for t in range(n + 1)]
        # (Vxx_0 is left unsymmetrised by the reference and by the kernel: use its lower triangle)
        Vs[0] = np.tril(Vs[0]) + np.tril(Vs[0], -1).T
        res = LegResult(Vs[0], vx[0], VXT[0], VTT[0], VT[0], ff, fb, fth, Vs, [vx[t] for t in range(n + 1)],
                        [VXT[t] for t in range(n + 1)])
        s.close()
        return res
