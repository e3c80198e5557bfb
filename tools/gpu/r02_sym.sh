#!/bin/bash
# triangle-packed host sweep: GPU parity (bit-equal to the plain host sweep) and the bench line with both e2e figures
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "triangle_packed or pipelined_host" 2>&1 | tail -3
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_c2.json 2> gpurun_out/r02_bench_c2.err
tail -1 gpurun_out/r02_bench_c2.json | python -c 'import sys,json
d=json.loads(sys.stdin.read()); print("c2 value %.4g ms %.4f frac %.3f" % (d["value"], d["ms_per_step"], d["roofline"]["frac"])); print(d["e2e"])'
tail -3 gpurun_out/r02_bench_c2.err
