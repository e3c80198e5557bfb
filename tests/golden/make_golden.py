#!/usr/bin/env python
"""Writes tests/golden/oracle_regression.npz from the CPU oracle (see README.md in this directory)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))

CASES = {  # name: (nx, nu, nc, nct, N, mueq, seed)
    "unconstrained": (6, 3, 0, 0, 12, 1e-8, 101),
    "constrained": (4, 2, 2, 0, 10, 1e-3, 102),
    "terminal": (5, 2, 1, 2, 6, 1e-2, 103),
}


def solve(case):
    import gen
    from oracle import gar_oracle as orc
    nx, nu, nc, nct, N, mueq, seed = CASES[case]
    probs = gen.generate_batch(seed, 2, N, nx, nu, nc, nct)
    stage, term, G0, g0 = gen.pack_problems(probs)
    bo = orc.BatchedOracle(nx, nu, nc, nct, probs[0].nc0, N, 2, stage, term, G0, g0)
    bo.sweep(mueq, nthreads=1)
    return bo.get()


if __name__ == "__main__":
    out = {}
    for c in CASES:
        r = solve(c)
        for k in ("fb", "ff", "Vxx", "vx", "xs", "us", "lbdas"):
            out["%s/%s" % (c, k)] = r[k]
    np.savez_compressed(os.path.join(HERE, "oracle_regression.npz"), **out)
    print("wrote", len(out), "arrays")
