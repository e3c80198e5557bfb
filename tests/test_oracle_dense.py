"""CPU tests: the two independent algorithms the reference offers for the same LQ problem
-- the proximal Riccati kernel (gar/riccati-kernel.hxx, reduced (nu+nc)^2 KKT per knot) and
the stage-dense kernel (gar/dense-kernel.hpp:98-113, one (nu+nc+2nx)^2 Bunch-Kaufman per
knot) -- restated in the oracle, must agree on K, k, Z, z, the cost-to-go and the
primal-dual trajectory.  A bug in either restatement shows up here; the reference's own test
for the dense solver holds it to the same KKT thresholds (tests/gar/riccati.cpp:60-105)."""
import numpy as np
import pytest

import gen
from oracle import gar_oracle as orc


def _both(prob, mueq, theta=None):
    op = orc.OracleProblem(prob)
    prox = orc.ProximalRiccatiSolver(op)
    dense = orc.RiccatiSolverDense(op)
    assert prox.backward(mueq)
    assert dense.backward(mueq)
    s1, s2 = orc.OracleSolution(op), orc.OracleSolution(op)
    assert prox.forward(s1, theta)
    assert dense.forward(s2, theta)
    return op, prox, dense, s1, s2


@pytest.mark.parametrize("shape", [(6, 3, 0, 0, 30, 1e-8), (12, 6, 0, 0, 100, 1e-8), (4, 2, 2, 0, 40, 1e-3),
                                   (4, 2, 2, 3, 25, 1e-3), (14, 7, 0, 0, 60, 1e-8), (9, 5, 3, 2, 12, 1e-2),
                                   (12, 6, 6, 0, 30, 1e-3)])
def test_dense_and_proximal_kernels_agree(shape):
    nx, nu, nc, nct, N, mueq = shape
    prob = gen.generate_batch(17, 1, N, nx, nu, nc, nct)[0]
    # ProxDDP's terminal knot has nx2 = 0 (solvers/proxddp/workspace.hxx:54-55).  With nx2 > 0 the
    # dense kernel's terminal KKT matrix carries zero blocks, its Bunch-Kaufman stops at the first
    # of them with the trailing pivots left at 0, and solveInPlace then swaps row 0 of the
    # right-hand side into the dead rows (bunchkaufman.hpp:366-371, 459-470): the reference's dense
    # solver itself is only meaningful with nx2 = 0 there, so that is what is compared.
    from aligator_b200.lqr import LqrKnot
    kt = prob.stages[-1]
    k0 = LqrKnot(nx, 0, nct, 0)
    k0.Q[:], k0.q[:], k0.C[:], k0.d[:] = kt.Q, kt.q, kt.C, kt.d
    prob.stages[-1] = k0
    op, prox, dense, s1, s2 = _both(prob, mueq)
    worst = {}
    for t in range(N):
        a, b = prox.factor(t), dense.factor(t)
        nk = nu + nc
        for key, x, y in (("K", a["fb"][:nu], b["fb"][:nu]), ("Z", a["fb"][nu:nk], b["fb"][nu:nk]),
                          ("k", a["ff"][:nu], b["ff"][:nu]), ("z", a["ff"][nu:nk], b["ff"][nu:nk]),
                          # closed loop: Ahat = Y block, a = y block of the dense solution
                          ("Ahat", a["fb"][nk:], b["fb"][nk + nx:]), ("a", a["ff"][nk:], b["ff"][nk + nx:])):
            if x.size:
                worst[key] = max(worst.get(key, 0.0), gen.rel_fro(x, y))
        # cost-to-go: the dense kernel's Pxx is not symmetrised; compare lower triangles
        il = np.tril_indices(nx)
        worst["Vxx"] = max(worst.get("Vxx", 0.0), gen.rel_fro(a["Vxx"][il], b["Pxx"][il]))
        worst["vx"] = max(worst.get("vx", 0.0), gen.rel_fro(a["vx"], b["px"]))
    tol = 1e-10
    tolk = max(tol, 2.4e-16 / mueq) if nc > 0 else tol  # eps*cond(KKT) on constrained knots (App. C)
    bad = {k: v for k, v in worst.items() if v > (tolk if k in ("K", "k", "Z", "z") else tol)}
    assert not bad, (bad, worst)
    cat = lambda v: np.concatenate([np.ravel(x) for x in v] + [np.zeros(0)])
    for x, y in zip(s1.get(), s2.get()):
        if cat(x).size:
            assert gen.rel_fro(cat(x), cat(y)) <= 1e-9
    for sol in (s1, s2):  # tests/gar/riccati.cpp:84,104
        assert max(orc.kkt_error(op, sol, mueq)) <= 1e-9


def test_dense_kernel_reference_style_problem_meets_reference_threshold():
    """tests/gar/riccati.cpp:60-105 (`riccati_short_horz_pb` also runs the dense solver):
    reference-style generator, KKT residuals <= 1e-9."""
    rng = np.random.default_rng(5)
    nx, nu, N = 2, 2, 8
    prob = gen.generate_lq_problem(rng, rng.standard_normal(nx), N, nx, nu, 0, 0, True)
    op, prox, dense, s1, s2 = _both(prob, 1e-14)
    assert max(orc.kkt_error(op, s2, 1e-14)) <= 1e-9


def test_dense_and_proximal_parametric_agree():
    """Parametric terms (nth > 0): Pxt vs Vxt and the theta-dependent trajectories agree.
    Ptt / pt are NOT compared below the terminal knot: DenseKernel::stageKernelSolve
    (dense-kernel.hpp:161-174) never adds the next knot's Ptt and pt (the proximal kernel does,
    riccati-kernel.hxx:296,309), so the dense solver's Ptt, pt, thGrad, thHess are those of a
    different quantity -- restated as written, and only checked at the terminal knot, where the
    two formulas coincide."""
    rng = np.random.default_rng(8)
    nx, nu, N, nth = 6, 3, 20, 2
    prob = gen.generate_lq_problem(rng, rng.standard_normal(nx), N, nx, nu, nth, 0, False, conditioned=True)
    from aligator_b200.lqr import LqrKnot
    kt = prob.stages[-1]
    k0 = LqrKnot(nx, 0, 0, 0, nth)
    for f in ("Q", "q", "Gx", "Gth", "gamma"):
        getattr(k0, f)[...] = getattr(kt, f)
    prob.stages[-1] = k0
    theta = rng.uniform(-1, 1, nth)
    op, prox, dense, s1, s2 = _both(prob, 1e-8, theta)
    for t in range(N + 1):
        a, b = prox.factor(t), dense.factor(t)
        assert gen.rel_fro(a["Vxt"], b["Pxt"]) <= 1e-10
    a, b = prox.factor(N), dense.factor(N)
    assert gen.rel_fro(a["Vtt"], b["Ptt"]) <= 1e-10
    assert gen.rel_fro(a["vt"], b["pt"]) <= 1e-10
    cat = lambda v: np.concatenate([np.ravel(x) for x in v])
    for x, y in zip(s1.get(), s2.get()):
        assert gen.rel_fro(cat(x), cat(y)) <= 1e-9
