"""Cycles per phase of the CTA-per-instance kernel at given dims (AB2_PHASE_CLOCKS=1): nx nu N B [legs]"""
import ctypes as C
import os
import sys
os.environ["AB2_PHASE_CLOCKS"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
import aligator_b200.gar as gar  # noqa: E402
nx, nu, N, B = [int(a) for a in sys.argv[1:5]]
legs = int(sys.argv[5]) if len(sys.argv) > 5 else 0
prob = bench.synth_batch_torch(torch, B, N, nx, nu, "cuda:0", 3)
s = gar.CudaRiccatiBatch(nx, nu, 0, 0, nx, N, B, 0, 9 if not legs else -1, legs=legs)
s.set_problem(*prob, memspace=gar.AB2_DEVICE)
clk = (C.c_longlong * 16)()
gar.lib().ab2_gar_phase_clocks.argtypes = [C.c_void_p, C.c_void_p]
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
s.sweep(1e-9); s.synchronize()
gar._check(gar.lib().ab2_gar_phase_clocks(s.h, clk))
ev[0].record(); s.backward(1e-9); ev[1].record(); s.forward(); ev[2].record(); torch.cuda.synchronize()
gar._check(gar.lib().ab2_gar_phase_clocks(s.h, clk))
names = ["wait_copy", "(1) W=V'M", "(2) H", "(3) X,kkt build", "ldlt", "solves", "gains+(4)(5)", "parametric", "-", "loop tail"]
tot = sum(clk[:10])
nk = N if not legs else (N + 1) // legs
print("nx %d nu %d N %d B %d legs %d: backward %.3f ms forward %.3f ms; instance 0 (leg 0): %d cycles = %.0f per knot (%d knots)"
      % (nx, nu, N, B, legs, ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), tot, tot / max(nk, 1), nk))
for i, nme in enumerate(names):
    print("  %-18s %9d  %5.1f%%  %.0f / knot" % (nme, clk[i], 100.0 * clk[i] / max(tot, 1), clk[i] / max(nk, 1)))
print(s.kernel_info())
os._exit(0)
