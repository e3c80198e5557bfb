"""Host-side mirror of aligator's LQ problem types (numpy, fp64).

Mirrors ``gar::LqrKnotTpl`` / ``gar::LqrProblemTpl``
(include/aligator/gar/lqr-problem.hpp:49-118, 120-210): same field names,
same shapes, matrices stored column-major (``order='F'``) like Eigen's default.
These are plain containers; the arithmetic lives in the CUDA library.
"""
from __future__ import annotations

import numpy as np

_MAT_FIELDS = ("Q", "S", "R", "A", "B", "C", "D", "Gth", "Gx", "Gu", "Gv")
_VEC_FIELDS = ("q", "r", "f", "d", "gamma")


def _zeros(r, c=None):
    if c is None:
        return np.zeros(r, dtype=np.float64)
    return np.zeros((r, c), dtype=np.float64, order="F")


class LqrKnot:
    """One stage: cost 1/2 [x;u]^T [[Q,S],[S^T,R]] [x;u] + q^T x + r^T u,
    dynamics x' = A x + B u + f, constraint 0 = C x + D u + d
    (lqr-problem.hpp:14-33).  Zero-initialised like lqr-problem.hxx:29-72."""

    __slots__ = ("nx", "nu", "nc", "nx2", "nth") + _MAT_FIELDS + _VEC_FIELDS

    def __init__(self, nx, nu, nc, nx2=None, nth=0):
        self.nx, self.nu, self.nc = int(nx), int(nu), int(nc)
        self.nx2 = self.nx if nx2 is None else int(nx2)
        nx, nu, nc, nx2 = self.nx, self.nu, self.nc, self.nx2
        self.Q, self.S, self.R = _zeros(nx, nx), _zeros(nx, nu), _zeros(nu, nu)
        self.q, self.r = _zeros(nx), _zeros(nu)
        self.A, self.B, self.f = _zeros(nx2, nx), _zeros(nx2, nu), _zeros(nx2)
        self.C, self.D, self.d = _zeros(nc, nx), _zeros(nc, nu), _zeros(nc)
        self.addParameterization(nth)

    def addParameterization(self, nth):
        """lqr-problem.hxx:233-242: (re)allocates the theta terms as zeros."""
        self.nth = int(nth)
        nth = self.nth
        self.Gth = _zeros(nth, nth)
        self.Gx = _zeros(self.nx, nth)
        self.Gu = _zeros(self.nu, nth)
        self.Gv = _zeros(self.nc, nth)
        self.gamma = _zeros(nth)
        return self

    def copy(self):
        k = LqrKnot(self.nx, self.nu, self.nc, self.nx2, self.nth)
        for n in _MAT_FIELDS:
            setattr(k, n, np.array(getattr(self, n), dtype=np.float64, order="F"))
        for n in _VEC_FIELDS:
            setattr(k, n, np.array(getattr(self, n), dtype=np.float64))
        return k

    @property
    def dims(self):
        return (self.nx, self.nu, self.nc, self.nx2, self.nth)


class LqrProblem:
    """``stages`` = N+1 knots (last = terminal), initial condition
    ``G0 x0 + g0 = 0`` (lqr-problem.hpp:120-137)."""

    def __init__(self, stages, nc0):
        self.stages = list(stages)
        nx0 = self.stages[0].nx if self.stages else 0
        self.G0 = _zeros(int(nc0), nx0)
        self.g0 = _zeros(int(nc0))

    @property
    def horizon(self):
        return len(self.stages) - 1

    @property
    def nc0(self):
        return self.g0.shape[0]

    @property
    def ntheta(self):
        return self.stages[0].nth

    def addParameterization(self, nth):
        for s in self.stages:
            s.addParameterization(nth)

    def copy(self):
        p = LqrProblem([s.copy() for s in self.stages], self.nc0)
        p.G0 = np.array(self.G0, dtype=np.float64, order="F")
        p.g0 = np.array(self.g0, dtype=np.float64)
        return p


def lqr_initialize_solution(problem):
    """gar/utils.hpp:114-142: xs[N+1], us[N] (N+1 if terminal nu>0), vs[N+1],
    lbdas[N+1] with lbdas[0] of size nc0."""
    N = problem.horizon
    xs = [np.zeros(k.nx) for k in problem.stages]
    us = [np.zeros(k.nu) for k in problem.stages]
    vs = [np.zeros(k.nc) for k in problem.stages]
    lbdas = [np.zeros(problem.nc0)] + [
        np.zeros(problem.stages[i].nx2) for i in range(N)
    ]
    if problem.stages[-1].nu == 0:
        us.pop()
    return xs, us, vs, lbdas
