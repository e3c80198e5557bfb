"""Host side of the ParallelRiccatiSolver mirror (aligator_b200/parallel.py): Bunch-Kaufman and the
block-tridiagonal solve against the oracle / dense solves, and the whole orchestration (leg split,
re-parameterisation, condensed system, refinement, rollout) with the ORACLE plugged in as the leg
back end -- every host line runs on the CPU; the CUDA back end replaces only the per-leg sweeps."""
import copy

import numpy as np
import pytest

import gen
from aligator_b200 import parallel as par
from aligator_b200.lqr import LqrProblem
from oracle import gar_oracle as orc


@pytest.mark.parametrize("n,kind", [(1, "spd"), (4, "spd"), (7, "indef"), (12, "saddle"), (9, "zero_diag")])
def test_bunch_kaufman_numpy_matches_oracle(n, kind):
    rng = np.random.default_rng(n)
    a = rng.standard_normal((n, n))
    a = a + a.T
    if kind == "spd":
        a = a @ a.T + np.eye(n)
    elif kind == "saddle":
        m = n // 2
        a[m:, m:] = 0.0
    elif kind == "zero_diag":
        np.fill_diagonal(a, 0.0)
    f = par.BunchKaufman(a)
    info, mat, sub, piv = orc.bk_compute(a)
    assert f.ok == (info == 0)
    assert np.array_equal(f.piv, piv)
    assert np.allclose(np.tril(f.L), np.tril(mat), rtol=1e-12, atol=1e-13)
    b = rng.standard_normal((n, 3))
    x = f.solve(b)
    assert np.allclose(a @ x, b, rtol=1e-9, atol=1e-9)
    _, xo = orc.bk_solve(a, b)
    assert np.allclose(x, xo, rtol=1e-10, atol=1e-12)


def test_block_tridiagonal_solve_and_refinement():
    rng = np.random.default_rng(0)
    dims = [3, 5, 4, 5, 2]
    diag = []
    for d in dims:
        m = rng.standard_normal((d, d))
        diag.append(m @ m.T + 3 * np.eye(d))
    sup = [0.3 * rng.standard_normal((dims[i], dims[i + 1])) for i in range(len(dims) - 1)]
    sub = [s.T.copy() for s in sup]
    rhs = [rng.standard_normal(d) for d in dims]
    ok, x, facs, upT = par.block_tridiag_solve(sub, diag, sup, rhs)
    assert ok
    n = sum(dims)
    A = np.zeros((n, n))
    o = np.cumsum([0] + dims)
    for i, d in enumerate(dims):
        A[o[i]:o[i + 1], o[i]:o[i + 1]] = diag[i]
    for i in range(len(dims) - 1):
        A[o[i]:o[i + 1], o[i + 1]:o[i + 2]] = sup[i]
        A[o[i + 1]:o[i + 2], o[i]:o[i + 1]] = sub[i]
    want = np.linalg.solve(A, np.concatenate(rhs))
    assert np.allclose(np.concatenate(x), want, rtol=1e-10, atol=1e-12)
    Ax = par.block_tridiag_matmul(sub, diag, sup, x)
    assert np.allclose(np.concatenate(Ax), np.concatenate(rhs), rtol=1e-10, atol=1e-12)
    dx = par.block_tridiag_refine(upT, sup, facs, [r - a for r, a in zip(rhs, Ax)])
    assert np.max(np.abs(np.concatenate(dx))) <= 1e-12


def oracle_leg_backend(stages, final, mueq):
    """One leg solved by the oracle (its own terminalSolve handles a last knot with controls)."""
    leg = LqrProblem([copy.deepcopy(k) for k in stages], 0)
    op = orc.OracleProblem(leg)
    s = orc.ProximalRiccatiSolver(op)
    s.backward(mueq)  # (the nc0 = 0 initial stage it also solves is not used)
    f = [s.factor(t) for t in range(len(stages))]
    return par.LegResult(f[0]["Vxx"], f[0]["vx"], f[0]["Vxt"], f[0]["Vtt"], f[0]["vt"],
                         [d["ff"] for d in f], [d["fb"] for d in f], [d["fth"] for d in f],
                         [d["Vxx"] for d in f], [d["vx"] for d in f], [d["Vxt"] for d in f])


@pytest.mark.parametrize("shape", [(4, 2, 0, 0, 11, 3), (6, 3, 0, 0, 20, 4), (4, 2, 2, 0, 13, 2), (3, 2, 0, 2, 9, 6)])
def test_parallel_mirror_matches_oracle(shape):
    nx, nu, nc, nct, N, J1 = shape
    mueq = 1e-3 if (nc or nct) else 1e-8
    p_ref = gen.generate_batch(77, 1, N, nx, nu, nc, nct)[0]
    p_mine = copy.deepcopy(p_ref)
    p_ser = copy.deepcopy(p_ref)
    # oracle: parallel and serial
    op = orc.OracleProblem(p_ref)
    ref = orc.ParallelRiccatiSolver(op, J1, threaded=False)
    assert ref.backward(mueq)
    sol_ref = orc.OracleSolution(op)
    assert ref.forward(sol_ref)
    xr, ur, vr, lr = sol_ref.get()
    ops = orc.OracleProblem(p_ser)
    ser = orc.ProximalRiccatiSolver(ops)
    ser.backward(mueq)
    sol_s = orc.OracleSolution(ops)
    ser.forward(sol_s)
    xs_s, us_s, vs_s, ls_s = sol_s.get()
    # the mirror with the oracle as leg back end
    mine = par.ParallelRiccatiSolver(p_mine, J1, oracle_leg_backend)
    assert mine.backward(mueq)
    xs = [np.zeros(nx) for _ in range(N + 1)]
    us = [np.zeros(nu) for _ in range(N)]
    vs = [np.zeros(nc) for _ in range(N)] + [np.zeros(nct)]
    lb = [np.zeros(p_mine.nc0)] + [np.zeros(nx) for _ in range(N)]
    mine.forward(xs, us, vs, lb)
    tol = 1e-7 if (nc or nct) else 1e-8  # consensus to the condensed threshold (1e-10) times conditioning
    assert gen.rel_fro(np.array(xs), np.array(xs_s)) <= tol
    assert gen.rel_fro(np.array(us), np.array(us_s[:N])) <= tol
    assert gen.rel_fro(np.array(lb[1:]), np.array(ls_s[1:])) <= tol
    assert gen.rel_fro(np.array(xs), np.array(xr)) <= tol
    # the problem was re-parameterised in place like the reference does
    assert p_mine.stages[0].nth == nx and p_mine.stages[N].nth == 0
    with pytest.raises(RuntimeError):
        par.ParallelRiccatiSolver(copy.deepcopy(p_ser), 1, oracle_leg_backend)
