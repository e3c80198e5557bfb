"""Parametric problems (nth > 0: riccati-kernel.hxx:185-192, 278-311; proximal-riccati.hxx:50-59;
forward with theta) on the CTA-per-instance program, host emulation vs the oracle."""
import ctypes as C

import numpy as np
import pytest

import gen
import test_group_emulation as tg
from oracle import gar_oracle as orc

F = lambda a: np.asarray(a, dtype=np.float64).ravel(order="F")


def make_problem(seed, N, nx, nu, nc, nct, nth):
    rng = np.random.default_rng(seed)
    x0 = rng.standard_normal(nx)
    p = gen.generate_lq_problem(rng, x0, N, nx, nu, nth, nc, singular=False, conditioned=True,
                                control_rows=nc > 0, term_nc=nct)
    for k in p.stages:  # non-trivial parametric blocks everywhere
        k.Gx[...] = 0.3 * rng.standard_normal(k.Gx.shape)
        k.Gu[...] = 0.3 * rng.standard_normal(k.Gu.shape)
        k.Gv[...] = 0.0  # (the reference's Vxt formula drops Z^T Gv; legs always have Gv = 0)
        g = rng.standard_normal((nth, nth))
        k.Gth[...] = g @ g.T / max(nth, 1) + np.eye(nth)
        k.gamma[...] = rng.standard_normal(nth)
    return p


def run(p, nx, nu, nc, nct, nth, N, mueq, theta, nw):
    lib = tg._block_lib()
    srec = lib.emu_block_stage_record_th(nx, nu, nc, nth)
    stage = np.zeros((1, max(N, 1), srec))
    for t in range(N):
        k = p.stages[t]
        rec = np.concatenate([gen.stage_record(k), F(k.Gx), F(k.Gu), F(k.Gv), F(k.Gth), F(k.gamma)])
        stage[0, t, :rec.size] = rec
    kt = p.stages[N]
    term = np.concatenate([gen.term_record(kt), F(kt.Gx), F(kt.Gv), F(kt.Gth), F(kt.gamma)])[None]
    G0, g0 = F(p.G0)[None], np.asarray(p.g0)[None]
    nr, nc0 = nu + nc + nx, p.nc0
    z = lambda *s: np.full(s if np.prod(s) > 0 else (1,), np.nan)
    out = dict(ff=z(1, N, nr), fb=z(1, N, nr, nx), Vxx=z(1, N + 1, nx * nx), vx=z(1, N + 1, nx), ffT=z(1, nct),
               fbT=z(1, nct, nx), kkt0=z(1, nx + nc0), xs=z(1, N + 1, nx), us=z(1, N, nu), vs=z(1, N, nc),
               vsT=z(1, nct), lbd0=z(1, nc0), lbdas=z(1, N, nx), fth=z(1, N, nr, nth), Vxt=z(1, N + 1, nx * nth),
               Vtt=z(1, N + 1, nth * nth), vt=z(1, N + 1, nth), kkt0fth=z(1, nx + nc0, nth), thGrad=z(1, nth),
               thHess=z(1, nth * nth))
    status = np.full(1, -1, dtype=np.int32)
    sp = tg.SweepParams()
    sp.N, sp.nct, sp.nc0, sp.batch, sp.mueq, sp.do_bwd, sp.do_fwd, sp.nth = N, nct, nc0, 1, mueq, 1, 1, nth
    th = np.ascontiguousarray(theta, dtype=np.float64)
    keep = dict(stage=np.ascontiguousarray(stage), term=np.ascontiguousarray(term), G0=np.ascontiguousarray(G0),
                g0=np.ascontiguousarray(g0), theta=th, **out)
    for k, v in keep.items():
        setattr(sp, k, v.ctypes.data_as(tg._dp))
    sp.status = status.ctypes.data_as(C.POINTER(C.c_int))
    assert lib.emu_block_sweep_th(nx, nu, nc, nth, nw, C.byref(sp)) == 0
    assert status[0] == 0
    return out


@pytest.mark.parametrize("shape", [(5, 2, 0, 0, 3, 6, 1e-8, 1), (4, 3, 2, 0, 2, 5, 1e-3, 1), (7, 3, 0, 2, 7, 4, 1e-2, 2),
                                   (6, 2, 1, 0, 1, 3, 1e-3, 1)])
def test_parametric_block_sweep(shape):
    nx, nu, nc, nct, nth, N, mueq, nw = shape
    p = make_problem(sum(shape[:6]), N, nx, nu, nc, nct, nth)
    theta = np.random.default_rng(3).standard_normal(nth)
    got = run(p, nx, nu, nc, nct, nth, N, mueq, theta, nw)
    op = orc.OracleProblem(p)
    ref = orc.ProximalRiccatiSolver(op)
    assert ref.backward(mueq)
    tol = 1e-9 if (nc or nct) else 1e-10
    for t in range(N):
        f = ref.factor(t)
        for key, a, b in (("fb", got["fb"][0, t], f["fb"]), ("ff", got["ff"][0, t], f["ff"]),
                          ("fth", got["fth"][0, t], f["fth"])):
            assert gen.rel_fro(a, b) <= tol, (t, key)
    for t in range(N + 1):
        f = ref.factor(t)
        assert gen.rel_fro(got["Vxt"][0, t].reshape(nth, nx).T, f["Vxt"]) <= tol, t
        assert gen.rel_fro(got["Vtt"][0, t].reshape(nth, nth).T, f["Vtt"]) <= tol, t
        assert gen.rel_fro(got["vt"][0, t], f["vt"]) <= tol, t
    k0 = ref.kkt0()
    assert gen.rel_fro(got["kkt0"][0], k0["ff"]) <= tol
    assert gen.rel_fro(got["kkt0fth"][0], k0["fth"]) <= tol
    assert gen.rel_fro(got["thGrad"][0], k0["thGrad"]) <= tol
    assert gen.rel_fro(got["thHess"][0].reshape(nth, nth).T, k0["thHess"]) <= tol
    sol = orc.OracleSolution(op)
    assert ref.forward(sol, theta)
    xs, us, vs, lb = sol.get()
    assert gen.rel_fro(got["xs"][0], np.array(xs)) <= tol
    assert gen.rel_fro(got["us"][0], np.array(us[:N])) <= tol
    assert gen.rel_fro(got["lbd0"][0], lb[0]) <= tol
    assert gen.rel_fro(got["lbdas"][0], np.array(lb[1:])) <= tol


def test_parametric_random_shapes():
    """Seeded random sweep: odd sizes, constrained / terminal-constrained, nth from 1 to nx."""
    rng = np.random.default_rng(99)
    done = 0
    while done < 10:
        nx, nu = int(rng.integers(1, 9)), int(rng.integers(1, 5))
        nc = int(rng.integers(0, 3)) if rng.random() < 0.4 else 0
        nct = int(rng.integers(0, 3)) if rng.random() < 0.3 else 0
        nth = int(rng.integers(1, nx + 1))
        N = int(rng.integers(0, 5))
        need = max(nx + 1, nu + nc, 2 * nx, nu + nc + nx, nth)
        nw = (need + 31) // 32
        mueq = 1e-3 if (nc or nct) else 1e-8
        test_parametric_block_sweep((nx, nu, nc, nct, nth, N, mueq, nw))
        done += 1
