"""Line-search consumers of the LQ step: CPU test of the numpy restatement against a
hand-computed case; GPU test of the batched device kernels (ab2_gar_linear_step, ab2_gar_directional_derivative,
ab2_gar_al_value) against that restatement, the step being the solver's own forward pass."""
import numpy as np
import pytest

import gen
from oracle import linesearch as ols


def test_restatement_hand_case():
    xs = [np.array([1.0, 2.0]), np.array([0.0, 1.0])]
    us = [np.array([3.0])]
    vs = [np.zeros(0), np.array([2.0])]
    lams = [np.array([1.0, 1.0]), np.array([2.0, 0.0])]
    dxs = [np.array([1.0, 0.0]), np.array([0.0, 2.0])]
    dus = [np.array([-1.0])]
    dvs = [np.zeros(0), np.array([4.0])]
    dl = [np.array([0.0, 2.0]), np.array([2.0, 2.0])]
    tx, tu, tv, tl = ols.try_linear_step(xs, us, vs, lams, dxs, dus, dvs, dl, 0.5)
    assert np.array_equal(tx[0], [1.5, 2.0]) and np.array_equal(tx[1], [0.0, 2.0]) and np.array_equal(tu[0], [2.5])
    assert np.array_equal(tv[1], [4.0]) and np.array_equal(tl[0], [1.0, 2.0]) and np.array_equal(tl[1], [3.0, 1.0])
    Lxs = [np.array([2.0, 1.0]), np.array([1.0, 1.0])]
    Lus = [np.array([4.0])]
    assert ols.directional_derivative(Lxs, Lus, dxs, dus) == 2.0 + 2.0 - 4.0
    # cost 1 + 1/2 (mucstr*2 + mudyn*4 + mucstr*0 + mucstr*4) with mudyn 0.1, mucstr 10
    assert ols.al_value(1.0, lams, vs, 0.1, 10.0, True) == pytest.approx(1.0 + 0.5 * (20.0 + 0.4 + 0.0 + 40.0))


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(12, 6, 0, 0, 30, 17), (4, 2, 2, 3, 25, 70), (9, 5, 3, 0, 7, 5)])
def test_device_consumers_match_restatement(shape):
    import torch
    assert torch.cuda.is_available()
    import __graft_entry__ as g
    g.build()
    import aligator_b200.gar as gar
    nx, nu, nc, nct, N, B = shape
    mueq = 1e-3 if (nc or nct) else 1e-8
    probs = gen.generate_batch(21, B, N, nx, nu, nc, nct)
    s = gar.CudaRiccatiBatch(nx, nu, nc, nct, nx, N, B)
    s.set_problem(*gar.pack_problems(probs))
    s.sweep(mueq)
    step = {k: s.get(w) for k, w in dict(xs=gar.OUT_XS, us=gar.OUT_US, vs=gar.OUT_VS, vsT=gar.OUT_VST, lam0=gar.OUT_LBD0,
                                         lams=gar.OUT_LBDAS).items()}
    rng = np.random.default_rng(0)
    dev = torch.device("cuda:0")
    cur_h = {k: rng.standard_normal(v.shape) for k, v in step.items()}
    cur = {k: torch.tensor(v, device=dev) for k, v in cur_h.items()}
    trial = {k: torch.empty_like(v) for k, v in cur.items()}
    alpha = 0.37
    s.linear_step(alpha, cur, trial)
    s.synchronize()
    for k in cur:
        # (the device contracts x + alpha*dx into one FMA: a single rounding instead of two)
        assert gen.rel_fro(trial[k].cpu().numpy(), cur_h[k] + alpha * step[k]) <= 1e-15, k
    Lxs_h, Lus_h = rng.standard_normal((B, N + 1, nx)), rng.standard_normal((B, N, nu))
    d1 = s.directional_derivative(torch.tensor(Lxs_h, device=dev), torch.tensor(Lus_h, device=dev))
    cost_h = rng.standard_normal(B)
    val = s.al_value(cur, torch.tensor(cost_h, device=dev), 0.01, 7.0)
    for b in range(B):
        ref = ols.directional_derivative([Lxs_h[b, t] for t in range(N + 1)], [Lus_h[b, t] for t in range(N)],
                                         [step["xs"][b, t] for t in range(N + 1)], [step["us"][b, t] for t in range(N)])
        assert abs(d1[b] - ref) <= 1e-11 * max(1.0, abs(ref))
        lams_plus = [cur_h["lam0"][b]] + [cur_h["lams"][b, t] for t in range(N)]
        vs_plus = [cur_h["vs"][b, t] for t in range(N)] + [cur_h["vsT"][b]]
        refv = ols.al_value(cost_h[b], lams_plus, vs_plus, 0.01, 7.0, nct > 0)
        assert abs(val[b] - refv) <= 1e-11 * max(1.0, abs(refv))
    s.close()
