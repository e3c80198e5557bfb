"""Small sweeps of every kernel family, meant to run under compute-sanitizer
(memcheck / racecheck / synccheck): warp-per-instance DMMA (C2 shape), lane-per-column with
constraints (C3 shape), sub-warp groups (C1 shape), CTA-per-instance (odd run-time shape, the
pipelined host sweep with an odd slice start), LQ assembly, KKT error."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gen  # noqa: E402
import aligator_b200.gar as gar  # noqa: E402

CASES = [  # nx, nu, nc, nct, N, B, mueq, variant
    (12, 6, 0, 0, 9, 5, 1e-8, -1),
    (12, 6, 0, 0, 9, 5, 1e-8, 10),
    (14, 7, 0, 0, 7, 3, 1e-8, -1),
    (4, 2, 2, 0, 9, 9, 1e-3, -1),
    (6, 3, 0, 0, 9, 5, 1e-8, -1),
    (9, 5, 3, 0, 7, 3, 1e-3, -1),   # CTA per instance, odd record sizes
    (12, 6, 0, 0, 5, 3, 1e-8, 9),   # CTA per instance forced
]
only = sys.argv[1:] and [int(a) for a in sys.argv[1:]]
for ci, (nx, nu, nc, nct, N, B, mueq, variant) in enumerate(CASES):
    if only and ci not in only:
        continue
    probs = gen.generate_batch(3 + ci, B, N, nx, nu, nc, nct)
    stage, term, G0, g0 = gar.pack_problems(probs)
    s = gar.CudaRiccatiBatch(nx, nu, nc, nct, probs[0].nc0, N, B, 0, variant)
    s.set_problem(stage, term, G0, g0)
    s.sweep(mueq)
    s.backward(mueq)
    s.forward()
    st = s.status()
    xs = s.get(gar.OUT_XS)
    print("case", ci, (nx, nu, nc, nct, N, B), "variant", variant, "status ok", bool(np.all(st == 0)),
          "finite", bool(np.isfinite(xs).all()), flush=True)
    s.close()
print("sanitize cases done")
