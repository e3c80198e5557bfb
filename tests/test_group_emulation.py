"""CPU execution of the CUDA group program (riccati_group.cuh) through the host
emulation in tests/emu/group_emu.cpp (G threads + a barrier per group), compared
with the oracle.  This validates the kernel's arithmetic and indexing without a
GPU; the TMA/mbarrier plumbing is the only part it cannot see."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import gen
from oracle import gar_oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_DIR = os.path.join(HERE, "emu")
EMU_LIB = os.path.join(EMU_DIR, "libgroup_emu.so")
_dp = C.POINTER(C.c_double)


class SweepParams(C.Structure):
    _fields_ = [("N", C.c_int), ("nct", C.c_int), ("nc0", C.c_int), ("batch", C.c_int),
                ("mueq", C.c_double), ("do_bwd", C.c_int), ("do_fwd", C.c_int)] + \
               [(n, _dp) for n in ("stage", "term", "G0", "g0", "ff", "fb", "Vxx", "vx", "ffT",
                                   "fbT", "kkt0", "xs", "us", "vs", "vsT", "lbd0", "lbdas")] + \
               [("status", C.POINTER(C.c_int)), ("pivstat", C.POINTER(C.c_int)), ("stagger_ns", C.c_int),
                ("num_sms", C.c_int),
                ("ctas_per_sm", C.c_int), ("stage_head", C.c_int), ("dbg", C.c_int), ("nth", C.c_int)] + \
               [(n, _dp) for n in ("theta", "fth", "Vxt", "Vtt", "vt", "kkt0fth", "thGrad", "thHess")] + \
               [("legs", C.c_int), ("clk", C.c_void_p), ("cond", _dp),
                ("peer_world", C.c_int), ("peer_dst", C.c_void_p * 8), ("peer_off", C.c_longlong)]


def _lib():
    srcs = [os.path.join(EMU_DIR, "group_emu.cpp"),
            os.path.join(HERE, "..", "aligator_b200", "csrc", "riccati_group.cuh"),
            os.path.join(HERE, "..", "aligator_b200", "csrc", "riccati_configs.h")]
    if (not os.path.exists(EMU_LIB)
            or os.path.getmtime(EMU_LIB) < max(os.path.getmtime(s) for s in srcs)):
        subprocess.check_call(["/usr/bin/g++", "-O1", "-std=c++20", "-fPIC", "-shared", "-pthread",
                               "-o", EMU_LIB, srcs[0]])
    return C.CDLL(EMU_LIB)


BLOCK_LIB = os.path.join(EMU_DIR, "libblock_emu.so")


def _block_lib():
    srcs = [os.path.join(EMU_DIR, "block_emu.cpp"),
            os.path.join(HERE, "..", "aligator_b200", "csrc", "riccati_block.cuh"),
            os.path.join(HERE, "..", "aligator_b200", "csrc", "riccati_group.cuh")]
    if (not os.path.exists(BLOCK_LIB)
            or os.path.getmtime(BLOCK_LIB) < max(os.path.getmtime(s) for s in srcs)):
        subprocess.check_call(["/usr/bin/g++", "-O1", "-std=c++20", "-fPIC", "-shared", "-pthread",
                               "-o", BLOCK_LIB, srcs[0]])
    return C.CDLL(BLOCK_LIB)


def run_emulated(nx, nu, nc, nct, N, probs, mueq, db=0, block=0, peers=None):
    """block = 0: the warp-per-instance group program; block = NW > 0: the
    CTA-per-instance program of riccati_block.cuh with NW emulated warps."""
    lib = _block_lib() if block else _lib()
    B = len(probs)
    nc0 = probs[0].nc0
    stage, term, G0, g0 = gen.pack_problems(probs)
    srec = lib.emu_block_stage_record(nx, nu, nc) if block else lib.emu_stage_record(nx, nu, nc)
    assert srec > 0, "shape not instantiated"
    if N > 0 and stage.shape[-1] != srec:
        stage = np.concatenate([stage, np.zeros(stage.shape[:-1] + (srec - stage.shape[-1],))], -1)
    stage = np.ascontiguousarray(stage)
    nr = nu + nc + nx
    z = lambda *s: np.full(s if np.prod(s) > 0 else (1,), np.nan)
    out = dict(ff=z(B, N, nr), fb=z(B, N, nr, nx), Vxx=z(B, N + 1, nx * nx), vx=z(B, N + 1, nx),
               ffT=z(B, nct), fbT=z(B, nct, nx), kkt0=z(B, nx + nc0), xs=z(B, N + 1, nx),
               us=z(B, N, nu), vs=z(B, N, nc), vsT=z(B, nct), lbd0=z(B, nc0), lbdas=z(B, N, nx))
    status = np.full(B, -1, dtype=np.int32)
    p = SweepParams()
    p.N, p.nct, p.nc0, p.batch, p.mueq, p.do_bwd, p.do_fwd = N, nct, nc0, B, mueq, 1, 1
    keep = dict(stage=stage, term=term, G0=G0, g0=g0, **out)
    for k, v in keep.items():
        setattr(p, k, v.ctypes.data_as(_dp))
    p.status = status.ctypes.data_as(C.POINTER(C.c_int))
    pivstat = np.zeros(B, dtype=np.int32)
    p.pivstat = pivstat.ctypes.data_as(C.POINTER(C.c_int))
    if peers is not None:  # sharded batch: (receive buffers of every rank, offset in doubles) -- SweepParams::peer_*
        bufs, off = peers
        p.peer_world = len(bufs)
        for w, b in enumerate(bufs):
            p.peer_dst[w] = b.ctypes.data
        p.peer_off = off
    if block < 0:  # compile-time specialisation of the block program (StaticBlockDims)
        rc = lib.emu_block_sweep_static(nx, nu, nc, int(-block), C.byref(p))
    elif block:
        rc = lib.emu_block_sweep(nx, nu, nc, int(block), C.byref(p))
    else:
        rc = lib.emu_sweep(nx, nu, nc, int(db), C.byref(p))
    assert rc == 0
    out["Vxx"] = out["Vxx"].reshape(B, N + 1, nx, nx).transpose(0, 1, 3, 2)
    out["status"] = status
    out["pivots_2x2"], out["interchanges"] = pivstat & 0x7fff, (pivstat >> 16) & 0xffff
    out["kkt0_fast"] = ((pivstat >> 15) & 1).astype(bool)
    return out


def check_against_oracle(nx, nu, nc, nct, N, B, mueq, seed, tol=1e-10, style="conditioned", db=0,
                         pivoting=False, block=0):
    probs = gen.generate_batch(seed, B, N, nx, nu, nc, nct, style=style)
    if pivoting:
        gen.make_pivoting(probs)
    got = run_emulated(nx, nu, nc, nct, N, probs, mueq, db, block)
    assert np.all(got["status"] == 0)
    stage, term, G0, g0 = gen.pack_problems(probs)
    bo = orc.BatchedOracle(nx, nu, nc, nct, probs[0].nc0, N, B, stage, term, G0, g0)
    bo.sweep(mueq, nthreads=1)
    ref = bo.get()
    worst = {}
    for b in range(B):
        for t in range(N):
            worst["K"] = max(worst.get("K", 0), gen.rel_fro(got["fb"][b, t, :nu], ref["fb"][b, t, :nu]))
            worst["k"] = max(worst.get("k", 0), gen.rel_fro(got["ff"][b, t, :nu], ref["ff"][b, t, :nu]))
            worst["fb"] = max(worst.get("fb", 0), gen.rel_fro(got["fb"][b, t], ref["fb"][b, t]))
            worst["ff"] = max(worst.get("ff", 0), gen.rel_fro(got["ff"][b, t], ref["ff"][b, t]))
        for t in range(N + 1):
            worst["Vxx"] = max(worst.get("Vxx", 0), gen.rel_fro(got["Vxx"][b, t], ref["Vxx"][b, t]))
            worst["vx"] = max(worst.get("vx", 0), gen.rel_fro(got["vx"][b, t], ref["vx"][b, t]))
        for key in ("xs", "us", "vs", "lbdas", "lbd0", "vsT"):
            if ref[key].size:
                worst[key] = max(worst.get(key, 0), gen.rel_fro(got[key][b], ref[key][b]))
    # K alone on control-constrained knots is limited by eps*cond(KKT) ~ 6e-17/mueq between
    # any two correct fp64 solvers; everything else is gated at `tol`.
    tolk = max(tol, 2.4e-16 / mueq) if nc > 0 else tol
    bad = {k: v for k, v in worst.items() if not v <= (tolk if k in ("K", "k") else tol)}
    assert not bad, (bad, worst)
    # structural semantics (A1): Vxx_t symmetric for t >= 1 when N > 0
    for t in range(1, N + 1):
        assert np.array_equal(got["Vxx"][0, t], got["Vxx"][0, t].T)
    return worst


@pytest.mark.parametrize("db", [0, 1])
@pytest.mark.parametrize("shape", [
    (2, 2, 0, 0, 8), (6, 3, 0, 0, 20), (12, 6, 0, 0, 12), (14, 7, 0, 0, 6), (10, 4, 0, 0, 10),
    (3, 2, 0, 0, 9), (8, 3, 0, 0, 7)])
def test_emulated_kernel_unconstrained(shape, db):
    nx, nu, nc, nct, N = shape
    check_against_oracle(nx, nu, nc, nct, N, B=3, mueq=1e-8, seed=sum(shape), db=db)


@pytest.mark.parametrize("shape,mueq", [
    ((4, 2, 2, 0, 25), 1e-3), ((4, 2, 2, 0, 25), 1e-6), ((2, 2, 2, 0, 8), 1e-4),
    ((5, 2, 2, 0, 10), 1e-3), ((12, 6, 6, 0, 8), 1e-3)])
@pytest.mark.parametrize("db", [0, 1])
def test_emulated_kernel_constrained(shape, mueq, db):
    """Constrained knots: D = I rows with a random half inactive (2x2 pivots and
    interchanges occur); gated at the mueq values of SURVEY §8(d)."""
    nx, nu, nc, nct, N = shape
    check_against_oracle(nx, nu, nc, nct, N, B=4, mueq=mueq, seed=7 + sum(shape), db=db)


def test_emulated_kernel_terminal_constraints():
    """Terminal knot with nct > 0 (Z = C/mu branch, riccati-kernel.hxx:146-149)."""
    check_against_oracle(4, 2, 2, 3, 12, B=2, mueq=1e-3, seed=5, tol=1e-9)
    check_against_oracle(6, 3, 0, 2, 10, B=2, mueq=1e-2, seed=6, tol=1e-9)


def test_emulated_kernel_horizon_zero_and_one():
    check_against_oracle(4, 2, 0, 0, 0, B=2, mueq=1e-8, seed=1)
    check_against_oracle(4, 2, 0, 0, 1, B=2, mueq=1e-8, seed=2)


def test_emulated_kernel_reference_style_unstable_A():
    """A ~ U[-1,1] (reference generator style): still within 1e-10 for nc = 0."""
    check_against_oracle(6, 3, 0, 0, 40, B=2, mueq=1e-8, seed=3, style="reference")


@pytest.mark.parametrize("db", [0, 1])
@pytest.mark.parametrize("shape", [(6, 3, 0, 0, 12), (12, 6, 0, 0, 8), (2, 2, 0, 0, 6)])
def test_emulated_kernel_unconstrained_with_interchanges(shape, db):
    """Rhat needs pivoting although nc = 0: the branch-free fast path must detect it and
    fall back to the general Bunch-Kaufman (same results as the oracle)."""
    nx, nu, nc, nct, N = shape
    probs = gen.make_pivoting(gen.generate_batch(31, 2, N, nx, nu, nc, nct))
    # the oracle really does pivot on these problems
    op = orc.OracleProblem(probs[0])
    ref = orc.ProximalRiccatiSolver(op)
    ref.backward(1e-8)
    piv = np.concatenate([ref.factor(t)["bk_piv"] for t in range(N)])
    ident = np.concatenate([np.arange(nu) for _ in range(N)])
    assert np.any(piv != ident)
    check_against_oracle(nx, nu, nc, nct, N, B=2, mueq=1e-8, seed=31, db=db, pivoting=True, tol=1e-9)


@pytest.mark.parametrize("db", [2, 3])
@pytest.mark.parametrize("shape", [(12, 6, 0, 0, 10), (14, 7, 0, 0, 5), (10, 4, 0, 0, 6)])
def test_emulated_tensor_core_step(shape, db):
    """db=2: the DMMA (mma.sync m8n8k4 f64) formulation of the stage step, with the mma
    emulated by a fragment exchange between the 32 lane-threads; db=3: the same with a
    single record buffer refilled in two parts."""
    nx, nu, nc, nct, N = shape
    check_against_oracle(nx, nu, nc, nct, N, B=2, mueq=1e-8, seed=17 + nx, db=db)


def test_emulated_tensor_core_step_with_interchanges():
    check_against_oracle(12, 6, 0, 0, 8, B=2, mueq=1e-8, seed=31, db=2, pivoting=True, tol=1e-9)


@pytest.mark.parametrize("db", [2, 3])
def test_emulated_tensor_core_step_edge_horizons(db):
    check_against_oracle(12, 6, 0, 0, 0, B=1, mueq=1e-8, seed=3, db=db)
    check_against_oracle(12, 6, 0, 0, 1, B=1, mueq=1e-8, seed=4, db=db)
    check_against_oracle(12, 6, 0, 2, 4, B=1, mueq=1e-2, seed=5, db=db, tol=1e-9)


def test_pivot_statistics_and_initial_fast_path():
    """ab2_gar_pivot_stats semantics in the emulation: well-posed problems take the register fast path
    of the initial saddle system and no 2x2 pivot; make_2x2_pivots() problems take min(nu,nc) 2x2
    pivots per knot; make_pivoting() problems take one interchange per knot."""
    nx, nu, N, B = 12, 6, 6, 3
    got = run_emulated(nx, nu, 0, 0, N, gen.generate_batch(3, B, N, nx, nu, 0, 0), 1e-8, db=2)
    assert np.all(got["kkt0_fast"]) and np.all(got["pivots_2x2"] == 0) and np.all(got["interchanges"] == 0)
    got = run_emulated(nx, nu, 0, 0, N, gen.make_pivoting(gen.generate_batch(3, B, N, nx, nu, 0, 0)), 1e-8, db=2)
    assert np.all(got["interchanges"] >= N)
    nx, nu, nc = 4, 2, 2
    probs = gen.make_2x2_pivots(gen.generate_batch(4, B, N, nx, nu, nc, 0))
    got = run_emulated(nx, nu, nc, 0, N, probs, 1e-3, db=1)
    assert np.all(got["pivots_2x2"] >= N * min(nu, nc))
    stage, term, G0, g0 = gen.pack_problems(probs)
    bo = orc.BatchedOracle(nx, nu, nc, 0, nx, N, B, stage, term, G0, g0)
    bo.sweep(1e-3)
    ref = bo.get()
    for k in ("fb", "ff", "Vxx", "xs", "us", "vs", "lbdas"):
        assert gen.rel_fro(got[k], ref[k]) <= 1e-10, k


@pytest.mark.parametrize("shape,db", [((12, 6, 0, 0, 6), 2), ((6, 3, 0, 0, 5), 0), ((4, 2, 2, 0, 4), 1)])
def test_in_sweep_exchange_of_the_first_step_policy(shape, db):
    """Sharded batch: with SweepParams::peer_* set, the sweep itself stores every instance's
    [K_0 | k_0] (nu x (nx+1), row-major) into slot [rank] of EVERY rank's receive buffer at peer_off --
    here three host buffers stand in for the peer-mapped ones.  Everything else in the buffers stays
    untouched, and the sweep's own outputs are unchanged."""
    nx, nu, nc, nct, N = shape
    B, world, rank, slot = 3, 3, 1, 2
    probs = gen.generate_batch(5, B, N, nx, nu, nc, nct)
    per = nu * (nx + 1)
    total = B * per
    bufs = [np.full(3 * world * total, -7.0) for _ in range(world)]
    off = slot * world * total + rank * total
    out = run_emulated(nx, nu, nc, nct, N, probs, 1e-6, db=db, peers=(bufs, off))
    ref = run_emulated(nx, nu, nc, nct, N, probs, 1e-6, db=db)
    assert np.array_equal(out["fb"], ref["fb"]) and np.array_equal(out["xs"], ref["xs"])
    want = np.concatenate([out["fb"][:, 0, :nu, :], out["ff"][:, 0, :nu, None]], axis=2).ravel()
    for b in bufs:
        assert np.array_equal(b[off:off + total], want)
        assert np.all(b[:off] == -7.0) and np.all(b[off + total:] == -7.0)
