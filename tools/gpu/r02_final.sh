#!/bin/bash
# Round-2 final single-GPU evidence: the whole GPU test suite, one bench line per BASELINE config,
# the leg-mode table and the CTA kernel's phase clocks.  ncu evidence: tools/gpu/r02_profile.sh.
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_final.txt 2>&1
echo "pytest exit $?" >> gpurun_out/r02_pytest_gpu_final.txt
tail -4 gpurun_out/r02_pytest_gpu_final.txt
for c in c2 c1 c3 c4 c5; do
  python bench.py --config $c --steps 20 --warmup 5 > gpurun_out/r02_bench_$c.json 2> gpurun_out/r02_bench_$c.err
  tail -1 gpurun_out/r02_bench_$c.json | python -c 'import sys,json
d=json.loads(sys.stdin.read()); print(sys.argv[1], "value %.4g ms %.4f frac %.3f e2e %.4g parity %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], (d.get("e2e") or {}).get("value", 0), d.get("parity")))' $c
done
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null; tail -1 gpurun_out/r02_bench_reference_arm.json | cut -c1-400
python tools/gpu/legs_bench.py > gpurun_out/legs_bench.log 2>&1; tail -20 gpurun_out/legs_bench.log | cut -c1-200
python tools/gpu/blk_phases.py 57 28 150 512 > gpurun_out/r02_c5_phases.txt 2>&1; tail -14 gpurun_out/r02_c5_phases.txt
