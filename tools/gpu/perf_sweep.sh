#!/bin/bash
# kernel-only timings (device-timed, inputs resident): usage perf_sweep.sh <tag> "<cfg:variant> ..."
tag=$1; shift
mkdir -p gpurun_out
out=gpurun_out/perf_$tag.txt
: > $out
for cv in "$@"; do
  cfg=${cv%%:*}; var=${cv##*:}
  line=$(python bench.py --no-cpu --no-e2e --steps 30 --warmup 5 --config $cfg --variant $var 2>&1 | tail -1)
  echo "$cfg v$var $(echo "$line" | python -c 'import sys,json
try:
    d=json.loads(sys.stdin.read()); print("ms %.4f frac %.4f knots/s %.4g regs/ctas %s" % (d["ms_per_step"], d["roofline"]["frac"], d["value"], d["config"]["kernel"]))
except Exception as e: print("ERR", e)')" | tee -a $out
done
