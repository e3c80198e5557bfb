"""ctypes wrapper around oracle/libgar_oracle.so -- the CPU restatement of
aligator's gar Riccati path.

*** TEST INFRASTRUCTURE ONLY ***: import from tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs; never from aligator_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libgar_oracle.so")

F_FF, F_FB, F_FTH, F_VXX, F_VX, F_VXT, F_VTT, F_VT, F_KKTMAT = range(9)
F_BK_MAT, F_BK_SUBDIAG, F_BK_PIV, F_QHAT, F_RHAT, F_SHAT, F_QVEC, F_RVEC = range(9, 17)

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lp = C.POINTER(C.c_long)


def build(force=False):
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.gar_oracle_knot_size.restype = C.c_long
        L.gar_oracle_problem_create.restype = C.c_void_p
        L.gar_oracle_solution_create.restype = C.c_void_p
        L.gar_oracle_serial_create.restype = C.c_void_p
        L.gar_oracle_parallel_create.restype = C.c_void_p
        L.gar_oracle_batched_create.restype = C.c_void_p
        L.gar_oracle_batched_sweep.restype = C.c_double
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _flat(a):
    return np.asarray(a, dtype=np.float64).ravel(order="F")


def knot_record(k):
    """Q S R q r | A B f | C D d | Gth Gx Gu Gv gamma, column-major."""
    return np.concatenate([
        _flat(k.Q), _flat(k.S), _flat(k.R), _flat(k.q), _flat(k.r),
        _flat(k.A), _flat(k.B), _flat(k.f), _flat(k.C), _flat(k.D), _flat(k.d),
        _flat(k.Gth), _flat(k.Gx), _flat(k.Gu), _flat(k.Gv), _flat(k.gamma)])


def _problem_data(problem):
    parts = [knot_record(k) for k in problem.stages]
    parts += [_flat(problem.G0), _flat(problem.g0)]
    return np.ascontiguousarray(np.concatenate(parts))


class OracleProblem:
    """Owns a C++ copy of an ``aligator_b200.lqr.LqrProblem``-like object."""

    def __init__(self, problem):
        self.py = problem
        self.nk = len(problem.stages)
        dims = np.array([list(k.dims) for k in problem.stages], dtype=np.int32)
        data = _problem_data(problem)
        self.h = C.c_void_p(lib().gar_oracle_problem_create(
            self.nk, dims.ctypes.data_as(_ip), int(problem.nc0), _d(data)))

    def update(self, problem=None):
        """Re-read numerical data (same dims) from the python problem."""
        data = _problem_data(problem or self.py)
        lib().gar_oracle_problem_update(self.h, _d(data))

    def dims(self):
        out = np.zeros((self.nk, 5), dtype=np.int32)
        lib().gar_oracle_problem_dims(self.h, out.ctypes.data_as(_ip))
        return out

    def __del__(self):
        try:
            lib().gar_oracle_problem_destroy(self.h)
        except Exception:
            pass


class OracleSolution:
    def __init__(self, oprob):
        self.oprob = oprob
        self.h = C.c_void_p(lib().gar_oracle_solution_create(oprob.h))

    def get(self):
        sizes = (C.c_long * 4)()
        counts = (C.c_long * 4)()
        lib().gar_oracle_solution_sizes(self.h, sizes, counts)
        bufs = [np.zeros(max(int(s), 1)) for s in sizes]
        lib().gar_oracle_solution_get(self.h, *[_d(b) for b in bufs])
        dims = self.oprob.dims()
        N = self.oprob.nk - 1
        nus = int(counts[1])
        xs, us, vs, lbdas = [], [], [], []
        o = 0
        for t in range(N + 1):
            xs.append(bufs[0][o:o + dims[t, 0]].copy()); o += dims[t, 0]
        o = 0
        for t in range(nus):
            us.append(bufs[1][o:o + dims[t, 1]].copy()); o += dims[t, 1]
        o = 0
        for t in range(N + 1):
            vs.append(bufs[2][o:o + dims[t, 2]].copy()); o += dims[t, 2]
        nc0 = int(self.oprob.py.nc0)
        lbdas.append(bufs[3][:nc0].copy()); o = nc0
        for t in range(N):
            lbdas.append(bufs[3][o:o + dims[t, 3]].copy()); o += dims[t, 3]
        return xs, us, vs, lbdas

    def set(self, xs, us, vs, lbdas):
        cat = lambda v: np.ascontiguousarray(
            np.concatenate([np.asarray(a, dtype=np.float64).ravel() for a in v] + [np.zeros(1)]))
        a, b, c, d = cat(xs), cat(us), cat(vs), cat(lbdas)
        lib().gar_oracle_solution_set(self.h, _d(a), _d(b), _d(c), _d(d))

    def __del__(self):
        try:
            lib().gar_oracle_solution_destroy(self.h)
        except Exception:
            pass


def kkt_error(oprob, osol, mueq, theta=None):
    """lqrComputeKktError (gar/utils.hxx:88-182) -> (dyn, cstr, dual)."""
    out = np.zeros(3)
    th = None if theta is None else np.ascontiguousarray(theta, dtype=np.float64)
    lib().gar_oracle_kkt_error(oprob.h, osol.h, C.c_double(mueq),
                               _d(th) if th is not None else None, _d(out))
    return tuple(out)


class _FactorAccess:
    def _get(self, fn, t, field, n):
        buf = np.zeros(max(n, 1))
        got = fn(self.h, int(t), int(field), _d(buf))
        assert got == n, (got, n, field)
        return buf[:n]

    def factor(self, t):
        """dict of the public StageFactor members of knot t."""
        nx, nu, nc, nx2, nth = [int(v) for v in self.oprob.dims()[t]]
        nr = nu + nc + nx2
        g = lambda f, n: self._get(self._getfn, t, f, n)
        return dict(
            ff=g(F_FF, nr),
            fb=g(F_FB, nr * nx).reshape(nr, nx),
            fth=g(F_FTH, nr * nth).reshape(nr, nth),
            Vxx=g(F_VXX, nx * nx).reshape(nx, nx, order="F"),
            vx=g(F_VX, nx),
            Vxt=g(F_VXT, nx * nth).reshape(nx, nth, order="F"),
            Vtt=g(F_VTT, nth * nth).reshape(nth, nth, order="F"),
            vt=g(F_VT, nth),
            kktMat=g(F_KKTMAT, (nu + nc) ** 2).reshape(nu + nc, nu + nc, order="F"),
            bk_mat=g(F_BK_MAT, (nu + nc) ** 2).reshape(nu + nc, nu + nc, order="F"),
            bk_subdiag=g(F_BK_SUBDIAG, nu + nc),
            bk_piv=g(F_BK_PIV, nu + nc).astype(np.int64),
            Qhat=g(F_QHAT, nx * nx).reshape(nx, nx, order="F"),
            Rhat=g(F_RHAT, nu * nu).reshape(nu, nu, order="F"),
            Shat=g(F_SHAT, nx * nu).reshape(nx, nu, order="F"),
            qhat=g(F_QVEC, nx), rhat=g(F_RVEC, nu),
            dims=(nx, nu, nc, nx2, nth))


class ProximalRiccatiSolver(_FactorAccess):
    """Oracle counterpart of gar::ProximalRiccatiSolver (proximal-riccati.hxx)."""

    def __init__(self, oprob):
        self.oprob = oprob
        self.h = C.c_void_p(lib().gar_oracle_serial_create(oprob.h))
        self._getfn = lib().gar_oracle_serial_get

    def backward(self, mueq):
        return bool(lib().gar_oracle_serial_backward(self.h, C.c_double(mueq)))

    def forward(self, osol, theta=None):
        th = None if theta is None else np.ascontiguousarray(theta, dtype=np.float64)
        return bool(lib().gar_oracle_serial_forward(
            self.h, osol.h, _d(th) if th is not None else None))

    def kkt0(self):
        p = self.oprob.py
        nx, nc0, nth = p.stages[0].nx, p.nc0, int(self.oprob.dims()[0, 4])
        n0 = nx + nc0
        def g(which, n):
            buf = np.zeros(max(n, 1))
            got = lib().gar_oracle_serial_get_kkt0(self.h, which, _d(buf))
            assert got == n
            return buf[:n]
        return dict(ff=g(0, n0), fth=g(1, n0 * nth).reshape(n0, nth),
                    mat=g(2, n0 * n0).reshape(n0, n0, order="F"),
                    thGrad=g(3, nth), thHess=g(4, nth * nth).reshape(nth, nth, order="F"))

    def cycleAppend(self, knot):
        dm = np.array(list(knot.dims), dtype=np.int32)
        rec = np.ascontiguousarray(knot_record(knot))
        lib().gar_oracle_serial_cycle_append(self.h, dm.ctypes.data_as(_ip), _d(rec))

    def __del__(self):
        try:
            lib().gar_oracle_serial_destroy(self.h)
        except Exception:
            pass


class ParallelRiccatiSolver(_FactorAccess):
    """Oracle counterpart of gar::ParallelRiccatiSolver (parallel-solver.hxx).
    Like the reference it re-parameterises the (C++ copy of the) problem."""

    def __init__(self, oprob, num_threads, threaded=True):
        self.oprob = oprob
        self.h = C.c_void_p(lib().gar_oracle_parallel_create(
            oprob.h, int(num_threads), int(bool(threaded))))
        self._getfn = lib().gar_oracle_parallel_get
        if not lib().gar_oracle_parallel_ok(self.h):
            raise RuntimeError("numThreads should be greater than or equal to 2")

    def set_refinement(self, steps, threshold=1e-10):
        lib().gar_oracle_parallel_set_refinement(self.h, int(steps), C.c_double(threshold))

    def backward(self, mueq):
        return bool(lib().gar_oracle_parallel_backward(self.h, C.c_double(mueq)))

    def forward(self, osol):
        return bool(lib().gar_oracle_parallel_forward(self.h, osol.h))

    def collapseFeedback(self):
        lib().gar_oracle_parallel_collapse(self.h)

    def __del__(self):
        try:
            lib().gar_oracle_parallel_destroy(self.h)
        except Exception:
            pass


class RiccatiSolverDense:
    """Oracle counterpart of gar::RiccatiSolverDense (dense-riccati.hxx, dense-kernel.hpp):
    the reference's second, algorithmically independent solver of the same problem."""

    def __init__(self, oprob):
        self.oprob = oprob
        lib().gar_oracle_dense_create.restype = C.c_void_p
        self.h = C.c_void_p(lib().gar_oracle_dense_create(oprob.h))

    def backward(self, mueq):
        return bool(lib().gar_oracle_dense_backward(self.h, C.c_double(mueq)))

    def forward(self, osol, theta=None):
        th = None if theta is None else np.ascontiguousarray(theta, dtype=np.float64)
        return bool(lib().gar_oracle_dense_forward(self.h, osol.h, _d(th) if th is not None else None))

    def factor(self, t):
        """ff = [k; z; l; y], fb = [K; Z; L; Y] (row-major), value function Pxx, px (+ parametric)."""
        nx, nu, nc, nx2, nth = [int(v) for v in self.oprob.dims()[t]]
        n = nu + nc + 2 * nx2

        def g(field, cnt):
            buf = np.zeros(max(cnt, 1))
            got = lib().gar_oracle_dense_get(self.h, int(t), int(field), _d(buf))
            assert got == cnt, (got, cnt, field)
            return buf[:cnt]
        return dict(ff=g(0, n), fb=g(1, n * nx).reshape(n, nx), ft=g(2, n * nth).reshape(n, nth),
                    Pxx=g(3, nx * nx).reshape(nx, nx, order="F"), px=g(4, nx),
                    Pxt=g(5, nx * nth).reshape(nx, nth, order="F"),
                    Ptt=g(6, nth * nth).reshape(nth, nth, order="F"), pt=g(7, nth),
                    dims=(nx, nu, nc, nx2, nth))

    def kkt0(self):
        nx, nth = int(self.oprob.dims()[0][0]), int(self.oprob.dims()[0][4])
        n0 = nx + int(self.oprob.py.nc0)
        ff = np.zeros(max(n0, 1))
        lib().gar_oracle_dense_get_kkt0(self.h, 0, _d(ff))
        out = dict(ff=ff[:n0])
        if nth:
            g = np.zeros(nth); H = np.zeros(nth * nth)
            lib().gar_oracle_dense_get_kkt0(self.h, 3, _d(g))
            lib().gar_oracle_dense_get_kkt0(self.h, 4, _d(H))
            out.update(thGrad=g, thHess=H.reshape(nth, nth, order="F"))
        return out

    def __del__(self):
        try:
            lib().gar_oracle_dense_destroy(self.h)
        except Exception:
            pass


def bk_compute(a):
    """Eigen::BunchKaufman<MatrixXd, Lower>::compute restated
    (core/bunchkaufman.hpp:654-676).  Returns (info, mat, subdiag, piv)."""
    a = np.asfortranarray(a, dtype=np.float64)
    n = a.shape[0]
    mat = np.zeros((n, n), order="F")
    sub = np.zeros(max(n, 1))
    piv = np.zeros(max(n, 1), dtype=np.int32)
    info = lib().gar_oracle_bk_compute(n, _d(a), _d(mat), _d(sub), piv.ctypes.data_as(_ip))
    return info, mat, sub[:n], piv[:n]


def bk_solve(a, b):
    a = np.asfortranarray(a, dtype=np.float64)
    x = np.asfortranarray(np.array(b, dtype=np.float64).reshape(a.shape[0], -1))
    info = lib().gar_oracle_bk_solve(a.shape[0], _d(a), x.shape[1], _d(x))
    return info, x


class BatchedOracle:
    """Uniform-dims batched driver in the product's packed layout; the timed CPU
    baseline (OpenMP over instances) and the parity checker for the CUDA path."""

    def __init__(self, nx, nu, nc, nct, nc0, N, batch, stage, term, G0, g0):
        self.dims = (nx, nu, nc, nct, nc0, N, batch)
        # the product pads odd-sized stage records to an even number of doubles (16-byte
        # TMA granularity); the C driver reads dense records
        srec = 2 * nx * nx + 2 * nx * nu + nu * nu + 2 * nx + nu + nc * (nx + nu + 1)
        stage = np.asarray(stage, dtype=np.float64)
        if N > 0 and stage.shape[-1] > srec:
            stage = stage.reshape(batch, N, -1)[..., :srec]
        self._keep = [np.ascontiguousarray(a, dtype=np.float64) for a in (stage, term, G0, g0)]
        s, t, g, h = self._keep
        self.h = C.c_void_p(lib().gar_oracle_batched_create(
            nx, nu, nc, nct, nc0, N, batch, _d(s), _d(t), _d(g), _d(h)))

    def sweep(self, mueq, reps=1, nthreads=0):
        batch = self.dims[-1]
        self.status = np.zeros(batch, dtype=np.int32)
        return lib().gar_oracle_batched_sweep(
            self.h, C.c_double(mueq), int(reps), int(nthreads),
            self.status.ctypes.data_as(_ip))

    def get(self):
        nx, nu, nc, nct, nc0, N, B = self.dims
        nr = nu + nc + nx
        o = dict(
            ff=np.zeros((B, N, nr)), fb=np.zeros((B, N, nr, nx)),
            Vxx=np.zeros((B, N + 1, nx * nx)), vx=np.zeros((B, N + 1, nx)),
            ffT=np.zeros((B, max(nct, 1))), fbT=np.zeros((B, max(nct * nx, 1))),
            xs=np.zeros((B, N + 1, nx)), us=np.zeros((B, N, max(nu, 1))),
            vs=np.zeros((B, N, max(nc, 1))), vsT=np.zeros((B, max(nct, 1))),
            lbd0=np.zeros((B, max(nc0, 1))), lbdas=np.zeros((B, N, nx)))
        lib().gar_oracle_batched_get(
            self.h, *[_d(o[k]) for k in
                      ("ff", "fb", "Vxx", "vx", "ffT", "fbT", "xs", "us", "vs", "vsT", "lbd0", "lbdas")])
        # Vxx is column-major per block
        o["Vxx"] = o["Vxx"].reshape(B, N + 1, nx, nx).transpose(0, 1, 3, 2)
        if nu == 0: o["us"] = o["us"][:, :, :0]
        if nc == 0: o["vs"] = o["vs"][:, :, :0]
        o["ffT"] = o["ffT"][:, :nct]
        o["fbT"] = o["fbT"][:, :nct * nx].reshape(B, nct, nx)
        o["vsT"] = o["vsT"][:, :nct]
        o["lbd0"] = o["lbd0"][:, :nc0]
        return o

    def __del__(self):
        try:
            lib().gar_oracle_batched_destroy(self.h)
        except Exception:
            pass


def num_threads():
    return int(lib().gar_oracle_num_threads())
