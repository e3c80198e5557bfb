"""Debug helper: run one shape through the CUDA path and print per-knot errors vs the oracle."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "tests"))
import numpy as np
import gen
from oracle import gar_oracle as orc
import aligator_b200.gar as gar

nx, nu, nc, nct, N, B = [int(v) for v in sys.argv[1:7]]
mueq = float(sys.argv[7])
variant = int(sys.argv[8]) if len(sys.argv) > 8 else -1
probs = gen.generate_batch(200 + nx, B, N, nx, nu, nc, nct)
stage, term, G0, g0 = gar.pack_problems(probs)
s = gar.CudaRiccatiBatch(nx, nu, nc, nct, probs[0].nc0, N, B, 0, variant)
s.set_problem(stage, term, G0, g0)
s.sweep(mueq)
fb, ff, V = s.get(gar.OUT_FB), s.get(gar.OUT_FF), s.get(gar.OUT_VXX)
print("status", s.status(), "info", s.kernel_info())
bo = orc.BatchedOracle(nx, nu, nc, nct, probs[0].nc0, N, B, stage, term, G0, g0)
bo.sweep(mueq)
ref = bo.get()
for t in range(N - 1, -1, -1):
    print(t, "K %.2e Z %.2e A %.2e ff %.2e V %.2e" % (
        gen.rel_fro(fb[:, t, :nu], ref["fb"][:, t, :nu]),
        gen.rel_fro(fb[:, t, nu:nu + nc], ref["fb"][:, t, nu:nu + nc]) if nc else 0,
        gen.rel_fro(fb[:, t, nu + nc:], ref["fb"][:, t, nu + nc:]),
        gen.rel_fro(ff[:, t], ref["ff"][:, t]), gen.rel_fro(V[:, t], ref["Vxx"][:, t])))
