"""The padding that lets the library (terminal knot nu = 0) solve problems whose terminal knot HAS
controls: one more stage knot with A = B = f = 0 before a null terminal knot.  Checked on the CPU
with the oracle on both sides (the reference's terminalSolve nu > 0 branch, riccati-kernel.hxx:150-191,
against the stage step from the null knot's zero value function); the GPU test of the same path is
tests/test_gpu_parity.py::test_terminal_knot_with_controls."""
import numpy as np
import pytest

import gen
from oracle import gar_oracle as orc
from aligator_b200.gar import _pad_terminal_controls


@pytest.mark.parametrize("shape", [(4, 2, 0, 5, 1e-8), (6, 3, 2, 7, 1e-3), (5, 3, 2, 0, 1e-2)])
def test_padding_is_the_terminal_solve_with_controls(shape):
    nx, nu, nc, N, mueq = shape
    rng = np.random.default_rng(7)
    p = gen.generate_lq_problem(rng, rng.standard_normal(nx), N, nx, nu, nc=nc, singular=False, conditioned=True)
    p.stages[N] = gen.generate_knot(rng, nx, nu, nc, 0, False, conditioned=True)
    q = _pad_terminal_controls(p)
    assert q.horizon == N + 1 and q.stages[N + 1].nu == 0
    a, b = orc.ProximalRiccatiSolver(orc.OracleProblem(p)), orc.ProximalRiccatiSolver(orc.OracleProblem(q))
    assert a.backward(mueq) and b.backward(mueq)
    for t in range(N + 1):
        fa, fb = a.factor(t), b.factor(t)
        rows = nu + nc if t == N else nu + nc + nx
        assert gen.rel_fro(fb["fb"][:rows], fa["fb"][:rows]) <= 1e-12
        assert gen.rel_fro(fb["ff"][:rows], fa["ff"][:rows]) <= 1e-12
        assert gen.rel_fro(fb["Vxx"], fa["Vxx"]) <= 1e-12 and gen.rel_fro(fb["vx"], fa["vx"]) <= 1e-12
    assert np.all(b.factor(N)["fb"][nu + nc:] == 0.0) and np.all(b.factor(N)["ff"][nu + nc:] == 0.0)
    sa, sb = orc.OracleSolution(orc.OracleProblem(p)), orc.OracleSolution(orc.OracleProblem(q))
    assert a.forward(sa) and b.forward(sb)
    xa, ua, va, la = sa.get()
    xb, ub, vb, lb = sb.get()
    assert len(ua) == N + 1 and len(ub) == N + 1
    assert gen.rel_fro(np.concatenate(xb[:N + 1]), np.concatenate(xa)) <= 1e-12
    assert gen.rel_fro(np.concatenate(ub), np.concatenate(ua)) <= 1e-12
    assert gen.rel_fro(np.concatenate(vb[:N + 1]), np.concatenate(va)) <= 1e-12
    assert gen.rel_fro(np.concatenate(lb[:N + 1]), np.concatenate(la)) <= 1e-12
    assert np.all(xb[N + 1] == 0.0) and np.all(lb[N + 1] == 0.0)


def test_padding_of_ragged_stage_dims():
    """Stage knots of different (nu, nc) padded to the largest (decoupled controls with R = I, null
    constraint rows): the caller's rows of every factor and the whole solution are the unpadded
    problem's, the padding rows exact zeros -- oracle on both sides."""
    from aligator_b200.gar import _pad_stage_dims
    from aligator_b200.lqr import LqrProblem
    nx, N, mueq = 5, 6, 1e-4
    dims = [(3, 0), (2, 2), (3, 1), (1, 0), (3, 2), (2, 0)]
    rng = np.random.default_rng(3)
    knots = [gen.generate_knot(rng, nx, nu, nc, 0, False, conditioned=True) for nu, nc in dims]
    knots.append(gen.generate_knot(rng, nx, 0, 0, 0, False, conditioned=True))
    p = LqrProblem(knots, nx)
    p.G0[:] = -np.eye(nx)
    p.g0[:] = rng.standard_normal(nx)
    q = _pad_stage_dims(p, 3, 2)
    a, b = orc.ProximalRiccatiSolver(orc.OracleProblem(p)), orc.ProximalRiccatiSolver(orc.OracleProblem(q))
    assert a.backward(mueq) and b.backward(mueq)
    for t, (nu, nc) in enumerate(dims):
        fa, fb = a.factor(t), b.factor(t)
        rows = np.r_[0:nu, 3:3 + nc, 5:5 + nx]
        pad = np.setdiff1d(np.arange(5 + nx), rows)
        assert gen.rel_fro(fb["fb"][rows], fa["fb"]) <= 1e-13 and gen.rel_fro(fb["ff"][rows], fa["ff"]) <= 1e-13
        assert np.all(fb["fb"][pad] == 0.0) and np.all(fb["ff"][pad] == 0.0)
        assert gen.rel_fro(fb["Vxx"], fa["Vxx"]) <= 1e-13
    sa, sb = orc.OracleSolution(orc.OracleProblem(p)), orc.OracleSolution(orc.OracleProblem(q))
    assert a.forward(sa) and b.forward(sb)
    xa, ua, va, la = sa.get()
    xb, ub, vb, lb = sb.get()
    assert gen.rel_fro(np.concatenate(xb), np.concatenate(xa)) <= 1e-13
    for t, (nu, nc) in enumerate(dims):
        assert gen.rel_fro(ub[t][:nu], ua[t]) <= 1e-13 and np.all(ub[t][nu:] == 0.0)
        assert gen.rel_fro(vb[t][:nc], va[t]) <= 1e-13 and np.all(vb[t][nc:] == 0.0)
