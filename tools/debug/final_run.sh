#!/bin/bash
# Round-end evidence run (one GPU): tests, bench lines, reference arm, ncu launch list + full capture.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/pytest_gpu.txt; cat gpurun_out/pytest_gpu.txt
timeout 600 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -1 gpurun_out/bench_c2.json | cut -c1-300
timeout 600 python bench.py --impl reference > gpurun_out/bench_reference_arm.json 2>/dev/null; tail -1 gpurun_out/bench_reference_arm.json | cut -c1-200
for c in c1 c3 c4 c5; do timeout 300 python bench.py --config $c --no-cpu --no-e2e --steps 20 --warmup 5 > gpurun_out/bench_$c.json 2>/dev/null; tail -1 gpurun_out/bench_$c.json | cut -c1-160; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:riccati -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --no-cpu --no-e2e --steps 3 --warmup 1 > /dev/null 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:riccati_sweep -c 1 -o gpurun_out/c2_default python bench.py --no-cpu --no-e2e --steps 1 --warmup 1 > gpurun_out/ncu_c2.log 2>&1; tail -1 gpurun_out/ncu_c2.log
