#!/bin/bash
# last evidence pass of round 2 (final build): launch list, full ncu capture of the C2 launch, C2 / C4 bench lines
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu --strong none > gpurun_out/r02_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:riccati_sweep_kernel -s 3 -c 1 -f -o gpurun_out/r02_c2_final \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-parity --strong none > /dev/null 2>&1
ncu -i gpurun_out/r02_c2_final.ncu-rep --page details 2>/dev/null | head -300 > gpurun_out/r02_c2_details.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_c2.json 2> gpurun_out/r02_bench_c2.err
python bench.py --config c4 --steps 20 --warmup 5 --no-cpu > gpurun_out/r02_bench_c4.json 2> gpurun_out/r02_bench_c4.err
for c in c2 c4; do tail -1 gpurun_out/r02_bench_$c.json | python -c 'import sys,json
d=json.loads(sys.stdin.read()); print(sys.argv[1], "value %.4g ms %.4f frac %.3f traffic %s e2e %.4g" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["traffic"], (d.get("e2e") or {}).get("value", 0)))' $c; done
grep -c riccati_sweep_kernel gpurun_out/r02_launches.csv
