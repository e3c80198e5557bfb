#!/bin/bash
# Round-2 ncu evidence: launch list of the default bench command, full captures of the C2 warp kernel and
# the C5 CTA kernel (source-level), traffic per config.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu --strong none > gpurun_out/r02_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:riccati_sweep_kernel -s 3 -c 1 -f -o gpurun_out/r02_c2 \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-parity --strong none > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:riccati_block_kernel -s 3 -c 1 -f -o gpurun_out/r02_c5 \
    python bench.py --config c5 --steps 2 --warmup 3 --no-cpu --no-e2e --no-parity --strong none > /dev/null 2>&1
for c in c3 c4; do
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:riccati_sweep_kernel -s 3 -c 1 --csv --log-file gpurun_out/r02_traffic_$c.csv \
    python bench.py --config $c --steps 2 --warmup 3 --no-cpu --no-e2e --no-parity --strong none > /dev/null 2>&1
done
ls -la gpurun_out/r02_*
