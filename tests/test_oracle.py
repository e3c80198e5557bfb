"""CPU tests that pin the oracle (oracle/gar_oracle.hpp).

The reference holds no golden vectors for this path, so the oracle is
pinned against (a) the reference's own test thresholds, on its own test shapes
(tests/gar/riccati.cpp, tests/gar/parallel.cpp), (b) one dense numpy solve of the
full KKT system (tests/gar/test_util.hpp:91-165 layout), (c) an independent numpy
re-derivation of K, k, Vxx, (d) LAPACK dsytf2 pivot sequences.
"""
import numpy as np
import pytest
import scipy.linalg.lapack as lapack

import gen
from aligator_b200.lqr import LqrKnot, LqrProblem
from oracle import gar_oracle as orc


# ----------------------------------------------------------------------------- BK
def _sym(rng, n, kind):
    if kind == "indef":
        a = rng.standard_normal((n, n))
        return a + a.T
    if kind == "kkt":  # [[R, D^T],[D, -mu I]]
        nu = max(1, n // 2)
        nc = n - nu
        w = rng.standard_normal((nu, nu + 1))
        a = np.zeros((n, n))
        a[:nu, :nu] = w @ w.T
        d = np.eye(nc, nu)
        d[rng.uniform(size=nc) < 0.5] = 0
        a[nu:, :nu] = d
        a[:nu, nu:] = d.T
        a[nu:, nu:] = -1e-8 * np.eye(nc)
        return a
    if kind == "zero_diag":
        a = rng.standard_normal((n, n))
        a = a + a.T
        a[np.diag_indices(n)] = 0.0
        return a
    raise ValueError(kind)


@pytest.mark.parametrize("kind", ["indef", "kkt", "zero_diag"])
@pytest.mark.parametrize("n", [1, 2, 3, 4, 6, 7, 12, 24, 32])
def test_bk_unblocked_matches_lapack_dsytf2(n, kind):
    """core/bunchkaufman.hpp:22-169 is LAPACK dsytf2 (lower): same pivots, and the
    factor solves A x = b."""
    if n == 1 and kind == "zero_diag":
        pytest.skip("1x1 zero matrix is singular by construction")
    rng = np.random.default_rng(100 * n + len(kind))
    for _ in range(5):
        a = _sym(rng, n, kind)
        info, mat, sub, piv = orc.bk_compute(a)
        assert info == 0
        ldu, ipiv, linfo = lapack.dsytf2(a, lower=1)
        assert linfo == 0
        ours = np.where(piv >= 0, piv + 1, piv)  # LAPACK is 1-based for 1x1 pivots
        if n > 1:
            assert np.array_equal(ours, ipiv), (piv, ipiv)
        b = rng.standard_normal((n, 3))
        info, x = orc.bk_solve(a, b)
        assert info == 0
        xr = np.linalg.solve(a, b)
        assert gen.rel_fro(x, xr) <= 1e-9 * max(1.0, np.linalg.cond(a) * 1e-4)


@pytest.mark.parametrize("n", [33, 44, 64, 70])
def test_bk_blocked_solves(n):
    """n > 32 takes the blocked path (core/bunchkaufman.hpp:171-420, BlockSize=32)."""
    rng = np.random.default_rng(n)
    for kind in ("indef", "kkt"):
        a = _sym(rng, n, kind)
        b = rng.standard_normal((n, 4))
        info, x = orc.bk_solve(a, b)
        assert info == 0
        assert np.linalg.norm(a @ x - b) <= 1e-9 * (np.linalg.norm(a) * np.linalg.norm(x) + 1)
        if kind == "indef":
            assert gen.rel_fro(x, np.linalg.solve(a, b)) <= 1e-9


def test_bk_reconstruction():
    """P A P^T = L D L^T with the storage of SURVEY A4 (inverted pivots, subdiag)."""
    rng = np.random.default_rng(7)
    for n in (5, 9, 16, 40, 67):
        a = _sym(rng, n, "indef")
        info, mat, sub, piv = orc.bk_compute(a)
        assert info == 0
        L = np.tril(mat, -1) + np.eye(n)
        Dinv = np.zeros((n, n))
        perm = np.arange(n)
        k = 0
        while k < n:
            if piv[k] < 0:
                p = -1 - piv[k]
                perm[[k + 1, p]] = perm[[p, k + 1]]
                Dinv[k, k], Dinv[k + 1, k + 1] = mat[k, k], mat[k + 1, k + 1]
                Dinv[k + 1, k] = Dinv[k, k + 1] = sub[k]
                k += 2
            else:
                p = piv[k]
                perm[[k, p]] = perm[[p, k]]
                Dinv[k, k] = mat[k, k]
                k += 1
        D = np.linalg.inv(Dinv)
        pa = a[np.ix_(perm, perm)]
        assert gen.rel_fro(L @ D @ L.T, pa) <= 1e-10


def test_bk_zero_matrix_fails():
    info, *_ = orc.bk_compute(np.zeros((3, 3)))
    assert info == 1  # NumericalIssue, bunchkaufman.hpp:58-59


# ----------------------------------------------------------- reference test shapes
def _solve_serial(problem, mueq, theta=None):
    op = orc.OracleProblem(problem)
    solver = orc.ProximalRiccatiSolver(op)
    assert solver.backward(mueq)
    sol = orc.OracleSolution(op)
    assert solver.forward(sol, theta)
    return op, solver, sol


@pytest.mark.parametrize("horz", [4, 8, 16])
def test_riccati_short_horz_pb(horz):
    """tests/gar/riccati.cpp:26-85: nx=nu=2, one knot with nc=2 (D=I, d=0.1)."""
    mueq = 1e-14
    rng = np.random.default_rng(horz)
    nx = nu = 2
    x0, x1 = np.ones(nx), -np.ones(nx)
    Brnd, frnd = rng.uniform(-1, 1, (nx, nu)), rng.uniform(-1, 1, nx)

    def init_knot(nc):
        k = LqrKnot(nx, nu, nc)
        k.A[:] = [[0.1, 0.0], [-0.1, 0.01]]
        k.B[:] = Brnd
        k.f[:] = frnd
        k.Q[:] = 0.01 * np.eye(nx)
        k.R[:] = 0.1 * np.eye(nu)
        return k

    knots = [init_knot(0) for _ in range(horz + 1)]
    knots[4] = init_knot(nu)
    knots[4].D[:] = np.eye(nu)
    knots[4].d[:] = 0.1
    knots[horz] = init_knot(0)
    knots[horz].Q[:] = np.eye(nx)
    knots[horz].q[:] = -x1
    prob = LqrProblem(knots, nx)
    prob.g0[:] = -x0
    prob.G0[:] = np.eye(nx)
    op, solver, sol = _solve_serial(prob, mueq)
    xs, us, vs, lbdas = sol.get()
    assert len(xs) == horz + 1 and len(vs) == horz + 1 and len(lbdas) == horz + 1
    err = orc.kkt_error(op, sol, mueq)
    assert max(err) <= 1e-9


def test_riccati_one_knot_prob():
    """tests/gar/riccati.cpp:87-105: horizon 0, us empty."""
    rng = np.random.default_rng(1)
    prob = gen.generate_lq_problem(rng, np.zeros(2), 0, 2, 2, 0, 0, True)
    op, solver, sol = _solve_serial(prob, 1e-13)
    xs, us, vs, lbdas = sol.get()
    assert len(xs) == 1 and len(us) == 0 and len(lbdas) == 1
    assert max(orc.kkt_error(op, sol, 1e-13)) <= 1e-10


@pytest.mark.parametrize("horz", [20, 100])
def test_riccati_random_large_problem(horz):
    """tests/gar/riccati.cpp:107-139: nx=36, nu=12, singular Q; serial <= 1e-6."""
    rng = np.random.default_rng(horz)
    prob = gen.generate_lq_problem(rng, np.zeros(36), horz, 36, 12, 0, 0, True)
    op, solver, sol = _solve_serial(prob, 1e-14)
    assert max(orc.kkt_error(op, sol, 1e-14)) <= 1e-6


def test_riccati_parametric():
    """tests/gar/riccati.cpp:157-192: nx=10, nu=4, N=100, nth=1."""
    rng = np.random.default_rng(3)
    nx, nu, horz, nth = 10, 4, 100, 1
    x0 = rng.standard_normal(nx)
    prob = gen.generate_lq_problem(rng, x0, horz, nx, nu, nth, 0, True)
    theta = rng.uniform(-1, 1, nth)
    op, solver, sol = _solve_serial(prob, 1e-12, theta)
    assert max(orc.kkt_error(op, sol, 1e-12, theta)) <= 1e-9
    k0 = solver.kkt0()
    for key in ("ff", "fth", "thGrad", "thHess"):
        assert np.all(np.isfinite(k0[key]))
    for t in (0, horz):
        f = solver.factor(t)
        for key in ("vt", "Vxt", "Vtt"):
            assert np.all(np.isfinite(f[key]))


def test_bench_native_shape_blocked_bk():
    """bench/gar-riccati.cpp:19-22: nx=36, nu=12, nc=32 -> KKT n=44 > 32 (blocked BK),
    C=[I 0] on every knot incl. terminal, mueq=1e-11."""
    rng = np.random.default_rng(11)
    nx, nu, nc, horz, mueq = 36, 12, 32, 16, 1e-11
    prob = gen.generate_lq_problem(rng, rng.standard_normal(nx), horz, nx, nu, 0, nc, True)
    op, solver, sol = _solve_serial(prob, mueq)
    err = orc.kkt_error(op, sol, mueq)
    assert np.all(np.isfinite(err))
    # badly scaled by construction (Vxx ~ 1/mu); the constraint/dynamics rows still close
    assert err[0] <= 1e-6 and err[1] <= 1e-6


# ------------------------------------------------------------ independent checks
@pytest.mark.parametrize("shape", [(6, 3, 0, 20), (12, 6, 0, 30), (4, 2, 2, 25), (14, 7, 0, 10)])
def test_serial_matches_dense_kkt_solve(shape):
    """One dense solve of the full KKT matrix (E=-I convention) vs backward+forward."""
    nx, nu, nc, N = shape
    rng = np.random.default_rng(sum(shape))
    mueq = 1e-4
    prob = gen.generate_lq_problem(rng, rng.standard_normal(nx), N, nx, nu, 0, nc, False,
                                   conditioned=True, control_rows=nc > 0, term_nc=0)
    op, solver, sol = _solve_serial(prob, mueq)
    xs, us, vs, lbdas = sol.get()
    dxs, dus, dvs, dl = gen.lqr_dense_solve(prob, mueq)
    cat = lambda v: np.concatenate([np.ravel(a) for a in v])
    assert gen.rel_fro(cat(xs), cat(dxs)) <= 1e-9
    assert gen.rel_fro(cat(us), cat(dus)) <= 1e-9
    assert gen.rel_fro(cat(lbdas), cat(dl)) <= 1e-9
    if nc:
        assert gen.rel_fro(cat(vs), cat(dvs)) <= 1e-8


def _numpy_riccati(prob, mueq):
    """Independent re-derivation of K,k,Vxx,vx with np.linalg.solve on the reduced KKT
    and a different product order (A^T (V A)) than riccati-kernel.hxx."""
    st = prob.stages
    N = prob.horizon
    out = [None] * (N + 1)
    m = st[N]
    assert m.nu == 0
    Z = m.C / mueq if m.nc else np.zeros((0, m.nx))
    z = m.d / mueq if m.nc else np.zeros(0)
    V = m.Q + m.C.T @ Z
    v = m.q + m.C.T @ z
    out[N] = dict(Vxx=V, vx=v)
    for t in range(N - 1, -1, -1):
        m = st[t]
        V = np.tril(V) + np.tril(V, -1).T
        vp = v + V @ m.f
        VA, VB = V @ m.A, V @ m.B
        Qh = m.Q + m.A.T @ VA
        Rh = m.R + m.B.T @ VB
        Sh = m.S + m.A.T @ VB
        qh = m.q + m.A.T @ vp
        rh = m.r + m.B.T @ vp
        kkt = np.block([[Rh, m.D.T], [m.D, -mueq * np.eye(m.nc)]])
        sol = np.linalg.solve(kkt, -np.block([[rh[:, None], Sh.T], [m.d[:, None], m.C]]))
        k, K = sol[:m.nu, 0], sol[:m.nu, 1:]
        z, Z = sol[m.nu:, 0], sol[m.nu:, 1:]
        V = Qh + Sh @ K + m.C.T @ Z
        v = qh + Sh @ k + m.C.T @ z
        out[t] = dict(K=K, k=k, Z=Z, z=z, Vxx=V, vx=v, a=m.f + m.B @ k, Ahat=m.A + m.B @ K)
    return out


@pytest.mark.parametrize("shape,mueq", [((12, 6, 0, 100), 1e-8), ((6, 3, 0, 100), 1e-8),
                                         ((14, 7, 0, 200), 1e-8), ((4, 2, 2, 100), 1e-3)])
def test_gains_match_numpy_rederivation(shape, mueq):
    """K, k, Vxx of the oracle vs an independent numpy recursion: <= 1e-10 rel-Frobenius
    on conditioned problems (SURVEY Appendix C shows >= 20x margin)."""
    nx, nu, nc, N = shape
    rng = np.random.default_rng(N + nx)
    prob = gen.generate_lq_problem(rng, rng.standard_normal(nx), N, nx, nu, 0, nc, False,
                                   conditioned=True, control_rows=nc > 0, term_nc=0)
    op, solver, sol = _solve_serial(prob, mueq)
    ref = _numpy_riccati(prob, mueq)
    worst = 0.0
    for t in range(N):
        f = solver.factor(t)
        worst = max(worst, gen.rel_fro(f["fb"][:nu], ref[t]["K"]),
                    gen.rel_fro(f["ff"][:nu], ref[t]["k"]),
                    gen.rel_fro(np.tril(f["Vxx"]), np.tril(ref[t]["Vxx"])))
    assert worst <= 1e-10, worst


def test_vxx_symmetrisation_semantics():
    """A1: datas[t].Vxx is exactly symmetric for t >= 1, not necessarily for t = 0."""
    rng = np.random.default_rng(5)
    prob = gen.generate_lq_problem(rng, np.zeros(8), 10, 8, 3, 0, 0, False)
    op, solver, sol = _solve_serial(prob, 1e-10)
    for t in range(1, 10):
        V = solver.factor(t)["Vxx"]
        assert np.array_equal(V, V.T)
    V0 = solver.factor(0)["Vxx"]
    assert np.allclose(V0, V0.T, rtol=1e-10, atol=1e-12)


def test_terminal_third_block_stays_zero():
    """A6: a / Ahat of the terminal knot are never written; Z = C/mu."""
    rng = np.random.default_rng(6)
    prob = gen.generate_lq_problem(rng, np.zeros(4), 3, 4, 2, 0, 2, False)
    op, solver, sol = _solve_serial(prob, 1e-6)
    f = solver.factor(3)
    nx, nu, nc, nx2, nth = f["dims"]
    assert np.all(f["ff"][nu + nc:] == 0) and np.all(f["fb"][nu + nc:] == 0)
    assert np.allclose(f["fb"][:nc], prob.stages[3].C / 1e-6)


# ---------------------------------------------------------------- parallel solver
def test_parallel_manual_two_legs():
    """tests/gar/parallel.cpp:63-169: split by hand, consensus on theta = costate."""
    EPS = 1e-9
    rng = np.random.default_rng(42)
    nx = nu = 2
    horizon, mueq = 16, 1e-14
    prob = gen.generate_lq_problem(rng, rng.uniform(-1, 1, nx), horizon, nx, nu)
    op, solver, sol = _solve_serial(prob, mueq)
    assert max(orc.kkt_error(op, sol, mueq)) <= EPS
    xs, us, vs, lbdas = sol.get()
    t0 = horizon // 2
    k1 = [prob.stages[i].copy() for i in range(t0)]
    k2 = [prob.stages[i].copy() for i in range(t0, horizon + 1)]
    last = k1[-1].copy()
    p1 = LqrProblem(k1, prob.nc0)
    p1.G0[:], p1.g0[:] = prob.G0, prob.g0
    p1.addParameterization(nx)
    p1.stages[-1].Gx[:] = last.A.T
    p1.stages[-1].Gu[:] = last.B.T
    p1.stages[-1].gamma[:] = last.f
    p2 = LqrProblem(k2, 0)
    p2.addParameterization(nx)
    p2.stages[0].Gx[:] = -np.eye(nx)
    o1, o2 = orc.OracleProblem(p1), orc.OracleProblem(p2)
    s1, s2 = orc.ProximalRiccatiSolver(o1), orc.ProximalRiccatiSolver(o2)
    assert s1.backward(mueq) and s2.backward(mueq)
    H = s1.kkt0()["thHess"] + s2.kkt0()["thHess"]
    g = s1.kkt0()["thGrad"] + s2.kkt0()["thGrad"]
    th = np.linalg.solve(H, -g)
    sol1, sol2 = orc.OracleSolution(o1), orc.OracleSolution(o2)
    s1.forward(sol1, th)
    s2.forward(sol2, th)
    assert max(orc.kkt_error(o1, sol1, mueq, th)) <= EPS
    assert max(orc.kkt_error(o2, sol2, mueq, th)) <= EPS
    xs1, us1, _, l1 = sol1.get()
    xs2, us2, _, l2 = sol2.get()
    xm = xs1 + xs2
    assert len(xs1) == t0 and len(xs2) == horizon - t0 + 1
    assert max(np.abs(a - b).max() for a, b in zip(xs, xm)) <= 1e-8
    lm = l1 + [th] + l2[1:]
    assert max(np.abs(a - b).max() for a, b in zip(lbdas, lm)) <= 1e-8


@pytest.mark.parametrize("threaded", [False, True])
def test_parallel_solver_class(threaded):
    """tests/gar/parallel.cpp:185-245: nx=32, nu=12, N=50, 6 legs vs serial."""
    rng = np.random.default_rng(9)
    nx, nu, horizon, TOL, mueq = 32, 12, 50, 1e-7, 1e-9
    prob = gen.generate_lq_problem(rng, np.zeros(nx), horizon, nx, nu)
    opr, ref, solr = _solve_serial(prob, mueq)
    assert max(orc.kkt_error(opr, solr, mueq)) <= 1e-8
    xr, ur, vr, lr = solr.get()
    op = orc.OracleProblem(prob)
    par = orc.ParallelRiccatiSolver(op, 6, threaded)
    par.set_refinement(10)
    assert par.backward(mueq)
    sol = orc.OracleSolution(op)
    par.forward(sol)
    assert max(orc.kkt_error(op, sol, mueq)) <= TOL
    xs, us, vs, ls = sol.get()
    assert max(np.abs(a - b).max() for a, b in zip(xs, xr)) <= TOL
    assert max(np.abs(a - b).max() for a, b in zip(ls, lr)) <= TOL
    for i in range(3):  # randomlyModifyProblem, parallel.cpp:173-183
        for j in (0, horizon // 3, horizon // 2, horizon // 2 + 1, horizon // 2 + 2, horizon):
            k = prob.stages[j]
            k.A += 0.1 * rng.standard_normal(k.A.shape)
            k.B += 0.1 * rng.standard_normal(k.B.shape)
            k.q += 0.1 * rng.standard_normal(k.q.shape)
        # keep the parameterisation the parallel solver installed on its copy
        dims = op.dims()
        p2 = prob.copy()
        for t, s in enumerate(p2.stages):
            s.addParameterization(int(dims[t, 4]))
        op.update(p2)
        assert par.backward(mueq)
        par.forward(sol)
        assert max(orc.kkt_error(op, sol, mueq)) <= TOL


def test_parallel_needs_two_threads():
    """parallel-solver.hxx:42-46 throws for < 2 threads."""
    rng = np.random.default_rng(1)
    prob = gen.generate_lq_problem(rng, np.zeros(2), 8, 2, 2)
    with pytest.raises(RuntimeError):
        orc.ParallelRiccatiSolver(orc.OracleProblem(prob), 1)


def test_get_work_ranges():
    """parallel-solver.hxx:23-28 leg ranges partition [0, N]."""
    for N in (7, 50, 100, 199):
        for J in (2, 3, 4, 6):
            r = [(i * (N + 1) // J, (i + 1) * (N + 1) // J) for i in range(J)]
            assert r[0][0] == 0 and r[-1][1] == N + 1
            assert all(r[i][1] == r[i + 1][0] for i in range(J - 1))


# ---------------------------------------------------------------- cycleAppend
def test_cycle_append_rotates_factors():
    """proximal-riccati.hxx:79-86: datas[0..N-1] rotate left, datas[N-1] fresh zeros."""
    rng = np.random.default_rng(2)
    prob = gen.generate_lq_problem(rng, np.zeros(3), 6, 3, 2, 0, 0, False)
    op, solver, sol = _solve_serial(prob, 1e-10)
    before = [solver.factor(t) for t in range(7)]
    solver.cycleAppend(gen.generate_knot(rng, 3, 2, 0))
    after = [solver.factor(t) for t in range(7)]
    for t in range(5):
        assert np.array_equal(after[t]["fb"], before[t + 1]["fb"])
        assert np.array_equal(after[t]["Vxx"], before[t + 1]["Vxx"])
    assert np.all(after[5]["fb"] == 0) and np.all(after[5]["ff"] == 0)
    assert np.array_equal(after[6]["Vxx"], before[6]["Vxx"])
    assert np.all(solver.kkt0()["mat"] == 0)


# ---------------------------------------------------------------- batched driver
def test_batched_driver_matches_serial_class():
    nx, nu, nc, nct, N, B = 5, 2, 2, 0, 12, 6
    probs = gen.generate_batch(3, B, N, nx, nu, nc, nct)
    stage, term, G0, g0 = gen.pack_problems(probs)
    bo = orc.BatchedOracle(nx, nu, nc, nct, nx, N, B, stage, term, G0, g0)
    bo.sweep(1e-5, nthreads=2)
    assert np.all(bo.status == 1)
    out = bo.get()
    for b in (0, B - 1):
        op, solver, sol = _solve_serial(probs[b], 1e-5)
        xs, us, vs, lbdas = sol.get()
        for t in (0, N // 2, N - 1):
            f = solver.factor(t)
            assert np.array_equal(out["fb"][b, t], f["fb"])
            assert np.array_equal(out["ff"][b, t], f["ff"])
            assert np.array_equal(out["Vxx"][b, t], f["Vxx"])
            assert np.array_equal(out["us"][b, t], us[t])
            assert np.array_equal(out["lbdas"][b, t], lbdas[t + 1])
        assert np.array_equal(out["xs"][b, N], xs[N])
        assert np.array_equal(out["lbd0"][b], lbdas[0])


def test_oracle_regression_fixture():
    """tests/golden/oracle_regression.npz (oracle-generated, see the README there): the oracle
    still produces these K, k, Vxx, vx, xs, us, lambdas (1e-12: compiler/FMA freedom only)."""
    import importlib.util
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(here, "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    ref = np.load(os.path.join(here, "golden", "oracle_regression.npz"))
    for case in mg.CASES:
        got = mg.solve(case)
        for k in ("fb", "ff", "Vxx", "vx", "xs", "us", "lbdas"):
            assert gen.rel_fro(got[k], ref["%s/%s" % (case, k)]) <= 1e-12, (case, k)
