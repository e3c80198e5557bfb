"""CPU restatement (numpy) of the LQ assembly -- TEST INFRASTRUCTURE ONLY, never imported by
the product.  Follows, statement by statement,

  computeProjectedJacobians   solvers/proxddp/solver-proxddp.hxx:25-69
  updateLQSubproblem          solvers/proxddp/solver-proxddp.hxx:734-805
  applyNormalConeProjectionJacobian   core/constraint-set.hxx:25-37
  computeActiveSet            equality-constraint.hpp:52-55, negative-orthant.hpp:30-33,
                              box-constraint.hpp:39-43

for ONE problem instance; the per-row bounds (lo, hi) encode the constraint sets: a row is
active iff z > hi or z < lo (equality: lo = +inf; negative orthant: lo = -inf, hi = 0).

PARITY UNPINNED: the reference holds no fixture for this step and cannot be built here
(no Eigen); the restatement is checked against a hand-computed case in tests/.
"""
import numpy as np


def active_set(z, lo, hi):
    return (z > hi) | (z < lo)


def projected_jacobians(cJx, cJu, Lv, shifted, lo, hi, mu_inv):
    """:25-69 for one knot -> (Px_proj, Pu_proj, lx_corr, lu_corr); cJu may be None (terminal)."""
    lv = Lv * mu_inv                                   # :46 / :63
    lx = cJx.T @ lv                                    # :47 / :64
    lu = cJu.T @ lv if cJu is not None else None       # :48
    act = active_set(shifted, lo, hi)
    Px = cJx.copy()
    Px[~act, :] = 0.0                                  # :49-50 / :65-66 (row-wise, inactive rows zeroed)
    lx = lx - Px.T @ lv                                # :51 / :67
    Pu = None
    if cJu is not None:
        Pu = cJu.copy()
        Pu[~act, :] = 0.0
        lu = lu - Pu.T @ lv                            # :52
    return Px, Pu, lx, lu


def assemble_problem(inp, N, nx, nu, nc, nct, nc0):
    """inp: dict of per-instance arrays named like ab2_lq_inputs (stage arrays lead with the
    knot index).  Returns a dict with the stage knots' matrices, the terminal knot and G0, g0
    (the contents of LqrProblemTpl after updateLQSubproblem)."""
    preg, mu_inv = inp["preg"], inp["mu_inv"]
    exact = inp.get("Hxx") is not None
    stages = []
    for t in range(N):
        k = {}
        k["A"] = inp["Jx"][t].copy()                   # :755
        k["B"] = inp["Ju"][t].copy()                   # :756
        k["f"] = inp["slack"][t].copy()                # :757
        k["Q"] = inp["Lxx"][t].copy()                  # :759
        k["S"] = inp["Lxu"][t].copy()
        k["R"] = inp["Luu"][t].copy()
        k["q"] = inp["Lx"][t].copy()                   # :764
        k["r"] = inp["Lu"][t].copy()
        k["Q"][np.diag_indices(nx)] += preg            # :767
        k["R"][np.diag_indices(nu)] += preg
        if exact:                                      # :770-774
            k["Q"] += inp["Hxx"][t]
            k["S"] += inp["Hxu"][t]
            k["R"] += inp["Huu"][t]
        if nc > 0:
            Px, Pu, lx, lu = projected_jacobians(inp["cJx"][t], inp["cJu"][t], inp["Lv"][t], inp["shifted"][t],
                                                 inp["lo"], inp["hi"], mu_inv)
            k["C"], k["D"], k["d"] = Px, Pu, inp["Lv"][t].copy()   # :778-780
            k["q"] += lx                               # :782
            k["r"] += lu                               # :783
        else:
            k["C"], k["D"], k["d"] = np.zeros((0, nx)), np.zeros((0, nu)), np.zeros(0)
        stages.append(k)
    term = {"Q": inp["Lxx_N"].copy(), "q": inp["Lx_N"].copy()}      # :787-790
    term["Q"][np.diag_indices(nx)] += preg
    if nct > 0:
        Px, _, lx, _ = projected_jacobians(inp["cJx_N"], None, inp["Lv_N"], inp["shifted_N"], inp["loN"],
                                           inp["hiN"], mu_inv)
        term["C"], term["d"] = Px, inp["Lv_N"].copy()               # :791-792
        term["q"] += lx                                             # :794
    else:
        term["C"], term["d"] = np.zeros((0, nx)), np.zeros(0)
    if inp.get("Hxx0") is not None:                                 # :803-804 (stages[0] is the terminal knot when N = 0)
        (stages[0] if N > 0 else term)["Q"] += inp["Hxx0"]
    G0 = inp["G0"].copy() if nc0 else np.zeros((0, nx))             # :799-800
    g0 = inp["g0"].copy() if nc0 else np.zeros(0)
    return {"stages": stages, "term": term, "G0": G0, "g0": g0}


def pack(prob, N, nx, nu, nc, nct, srec):
    """The product's packed layout (aligator_b200.gar.pack_stage_knot / pack_term_knot)."""
    F = lambda a: np.asarray(a, dtype=np.float64).ravel(order="F")
    stage = np.zeros((N, srec))
    for t, k in enumerate(prob["stages"]):
        rec = np.concatenate([F(k["A"]), F(k["B"]), F(k["f"]), F(k["Q"]), F(k["S"]), F(k["R"]), F(k["q"]),
                              F(k["r"]), F(k["C"]), F(k["D"]), F(k["d"])])
        stage[t, :rec.size] = rec
    kt = prob["term"]
    term = np.concatenate([F(kt["Q"]), F(kt["q"]), F(kt["C"]), F(kt["d"])])
    return stage, term, F(prob["G0"]), F(prob["g0"])
