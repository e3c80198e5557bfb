"""Parallel-in-time (leg mode) vs serial sweep at small per-GPU batches: BASELINE config 4's strong-scaling
end point (2048 instances over 8 GPUs = 256 per GPU) and neighbours.  Device-timed, inputs resident."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
import aligator_b200.gar as gar  # noqa: E402

nx, nu, N = 14, 7, 200
rows = []
for B in (16, 64, 256):
    stage, term, G0, g0 = bench.synth_batch_torch(torch, B, N, nx, nu, "cuda:0", 7)
    for legs, variant in ((0, -1), (0, 9), (2, -1), (4, -1), (8, -1), (16, -1)):
        s = gar.CudaRiccatiBatch(nx, nu, 0, 0, nx, N, B, 0, variant, legs=legs)
        s.set_problem(stage, term, G0, g0, memspace=gar.AB2_DEVICE)
        for _ in range(3):
            s.sweep(1e-9)
        s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reps = 10
        for _ in range(reps):
            s.sweep(1e-9, stream=torch.cuda.current_stream().cuda_stream)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        ok = bool((s.status() == 0).all())
        kk = float(s.kkt_error(1e-9).max())
        rows.append(dict(batch=B, legs=legs, variant=variant, ms=ms, knots_per_s=B * (N + 1) / ms * 1e3, ok=ok, kkt=kk))
        print(rows[-1], flush=True)
        s.close()
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "legs_bench.json"), "w"), indent=1)
