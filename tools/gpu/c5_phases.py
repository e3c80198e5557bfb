"""Cycles per phase of the CTA-per-instance kernel at BASELINE config 5 (AB2_PHASE_CLOCKS=1)."""
import ctypes as C
import os
import sys
os.environ["AB2_PHASE_CLOCKS"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
import aligator_b200.gar as gar  # noqa: E402
nx, nu, N, B = 57, 28, 150, 512
prob = bench.synth_batch_torch(torch, B, N, nx, nu, "cuda:0", 3)
s = gar.CudaRiccatiBatch(nx, nu, 0, 0, nx, N, B)
s.set_problem(*prob, memspace=gar.AB2_DEVICE)
s.sweep(1e-9); s.synchronize()
clk = (C.c_longlong * 16)()
gar.lib().ab2_gar_phase_clocks.argtypes = [C.c_void_p, C.c_void_p]
gar._check(gar.lib().ab2_gar_phase_clocks(s.h, clk))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); s.sweep(1e-9); e1.record(); torch.cuda.synchronize()
gar._check(gar.lib().ab2_gar_phase_clocks(s.h, clk))
names = ["wait_copy", "(1) W=V'M", "(2) H", "(3) X,kkt build", "ldlt", "solves", "gains+(4)(5)", "parametric", "-", "loop tail"]
tot = sum(clk[:10])
print("sweep %.3f ms; instance 0 backward: %d cycles = %.0f per knot" % (e0.elapsed_time(e1), tot, tot / N))
for i, nme in enumerate(names):
    print("  %-18s %9d  %5.1f%%  %.0f / knot" % (nme, clk[i], 100.0 * clk[i] / max(tot, 1), clk[i] / N))
