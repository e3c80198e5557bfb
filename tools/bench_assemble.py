#!/usr/bin/env python
"""Roofline of the LQ assembly kernel (ab2_gar_assemble = updateLQSubproblem +
computeProjectedJacobians on the device) at a BASELINE shape: achieved HBM GB/s =
(bytes read + bytes written per launch, counted from the arrays) / CUDA-event time."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=12)
    ap.add_argument("--nu", type=int, default=6)
    ap.add_argument("--nc", type=int, default=0)
    ap.add_argument("--horizon", type=int, default=100)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--exact", action="store_true", help="with dynamics Hessians (HessianApprox::EXACT)")
    a = ap.parse_args()
    import torch
    import __graft_entry__ as g
    g.build()
    import aligator_b200.gar as gar
    nx, nu, nc, N, B = a.nx, a.nu, a.nc, a.horizon, a.batch
    dev = torch.device("cuda:0")
    r = lambda *s: torch.randn(*s, dtype=torch.float64, device=dev)
    arr = dict(Jx=r(B, N, nx * nx), Ju=r(B, N, nx * nu), slack=r(B, N, nx), Lxx=r(B, N, nx * nx), Lxu=r(B, N, nx * nu),
               Luu=r(B, N, nu * nu), Lx=r(B, N, nx), Lu=r(B, N, nu), Lxx_N=r(B, nx * nx), Lx_N=r(B, nx),
               G0=r(B, nx * nx), g0=r(B, nx), Hxx0=r(B, nx * nx))
    if a.exact:
        arr.update(Hxx=r(B, N, nx * nx), Hxu=r(B, N, nx * nu), Huu=r(B, N, nu * nu))
    if nc:
        arr.update(cJx=r(B, N, nc * nx), cJu=r(B, N, nc * nu), Lv=r(B, N, nc), shifted=r(B, N, nc),
                   lo=torch.full((nc,), -float("inf"), dtype=torch.float64, device=dev),
                   hi=torch.zeros(nc, dtype=torch.float64, device=dev))
    s = gar.CudaRiccatiBatch(nx, nu, nc, 0, nx, N, B)
    stream = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        s.assemble(arr, 1e-6, 1e3, stream=stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        s.assemble(arr, 1e-6, 1e3, stream=stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    rd = sum(v.numel() for k, v in arr.items() if k not in ("lo", "hi")) * 8
    wr = (B * N * gar.stage_record_doubles(nx, nu, nc) + B * gar.term_record_doubles(nx, 0) + B * nx * nx + B * nx) * 8
    peak = 6561.3
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])
    gbs = (rd + wr) / (ms * 1e-3) / 1e9
    print(json.dumps({"kernel": "lq_assemble_stage_kernel + lq_assemble_term_kernel", "ms_per_launch": ms,
                      "knots_per_sec": B * (N + 1) / (ms * 1e-3), "bytes_read": rd, "bytes_written": wr,
                      "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak},
                      "config": {"nx": nx, "nu": nu, "nc": nc, "horizon": N, "batch": B, "exact_hessians": a.exact,
                                 "l2": "%.2f GB per launch exceeds the 126 MB L2" % ((rd + wr) / 1e9)}}))


if __name__ == "__main__":
    main()
