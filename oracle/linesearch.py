"""CPU restatement (numpy) of the line-search consumers of the LQ step -- TEST INFRASTRUCTURE, same rule as
oracle/gar_oracle.hpp.  Follows, statement by statement:
  tryLinearStep's vector part     solvers/proxddp/solver-proxddp.hxx:111-155 (vector-space integrate)
  costDirectionalDerivative       solvers/proxddp/merit-function.hxx:13-31
  ALFunction::directionalDerivative  merit-function.hxx:68-104 (given Lxs, Lus)
  ALFunction::evaluate            merit-function.hxx:33-66 (given prob_data.cost_)
Parity unpinned (the reference cannot run here); the arithmetic is sums and axpys."""
import numpy as np


def try_linear_step(xs, us, vs, lams, dxs, dus, dvs, dlams, alpha):
    """lists of vectors in, lists out: (trial_xs, trial_us, trial_vs, trial_lams)."""
    trial_lams = [l + alpha * dl for l, dl in zip(lams, dlams)]          # :121-122 vectorMultiplyAdd
    trial_vs = [v + alpha * dv for v, dv in zip(vs, dvs)]               # :123-124
    trial_xs = [x + (alpha * dx) for x, dx in zip(xs, dxs)]             # :139-150 xspace_->integrate (vector space)
    trial_us = [u + (alpha * du) for u, du in zip(us, dus)]
    return trial_xs, trial_us, trial_vs, trial_lams


def directional_derivative(Lxs, Lus, dxs, dus):
    """:82-101 -- d1 = Lxs[0].dxs[0] + sum_i (Lxs[i+1].dxs[i+1] + Lus[i].dus[i])."""
    d1 = float(np.dot(Lxs[0], dxs[0]))
    for i in range(len(dus)):
        d1 += float(np.dot(Lxs[i + 1], dxs[i + 1]))
        d1 += float(np.dot(Lus[i], dus[i]))
    return d1


def al_value(cost, lams_plus, vs_plus, mudyn, mucstr, has_term_cstr):
    """:41-65 -- cost + 1/2 (mucstr |lam_0|^2 + sum_i mudyn |lam_{i+1}|^2 + mucstr |v_i|^2 (+ mucstr |v_N|^2))."""
    nsteps = len(lams_plus) - 1
    pen = 0.5 * mucstr * float(np.dot(lams_plus[0], lams_plus[0]))
    for i in range(nsteps):
        pen += 0.5 * mudyn * float(np.dot(lams_plus[i + 1], lams_plus[i + 1]))
        pen += 0.5 * mucstr * float(np.dot(vs_plus[i], vs_plus[i]))
    if has_term_cstr:
        pen += 0.5 * mucstr * float(np.dot(vs_plus[nsteps], vs_plus[nsteps]))
    return cost + pen
