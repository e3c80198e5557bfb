"""ab2_gar_pack_stage_sym (host helper of ab2_gar_sweep_host_sym): the triangle-packed stage record
[A | B | f | Qlow | S | Rlow | q | r | C | D | d] against a numpy restatement.  No GPU needed; the device-side
expansion is covered by tests/test_gpu_parity.py::test_host_sweep_with_triangle_packed_records."""
import ctypes as C

import numpy as np
import pytest

import aligator_b200.gar as gar


@pytest.mark.parametrize("dims", [(12, 6, 0), (4, 2, 2), (5, 3, 1), (1, 1, 0), (3, 0, 0), (57, 28, 0)])
def test_pack_stage_sym_matches_numpy(dims):
    nx, nu, nc = dims
    L = gar.lib()
    pad = int(L.ab2_gar_stage_record_doubles(nx, nu, nc))
    sym = int(L.ab2_gar_stage_record_doubles_sym(nx, nu, nc))
    full_size = 2 * nx * nx + 2 * nx * nu + nu * nu + 2 * nx + nu + nc * (nx + nu + 1)
    assert pad in (full_size, full_size + 1)  # (records are padded to an even number of doubles)
    assert sym == full_size - nx * (nx - 1) // 2 - nu * (nu - 1) // 2
    nrec = 3
    rng = np.random.default_rng(0)
    st = np.zeros((nrec, pad))
    want = []
    for r in range(nrec):
        head = rng.standard_normal(nx * nx + nx * nu + nx)
        Q = rng.standard_normal((nx, nx))
        Q = Q + Q.T
        S = rng.standard_normal(nx * nu)
        R = rng.standard_normal((nu, nu))
        R = R + R.T
        rest = rng.standard_normal(nx + nu + nc * nx + nc * nu + nc)
        full = np.concatenate([head, Q.ravel(order="F"), S, R.ravel(order="F"), rest])
        st[r, :full.size] = full
        tri = lambda M: np.concatenate([M[j:, j] for j in range(M.shape[0])]) if M.shape[0] else np.zeros(0)
        want.append(np.concatenate([head, tri(Q), S, tri(R), rest]))
    out = np.full(nrec * sym, np.nan)
    rc = L.ab2_gar_pack_stage_sym(nx, nu, nc, st.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), C.c_long(nrec))
    assert rc == 0
    assert np.array_equal(out.reshape(nrec, sym), np.stack(want))
