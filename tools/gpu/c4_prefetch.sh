#!/bin/bash
# C4 (and C2): does the L2 prefetch of the records 4 knots ahead survive until it is used?  DRAM bytes and time
# per launch with the prefetch on (0), off (4) and at distance 2 (8).
mkdir -p gpurun_out
for cf in c4:0 c4:4 c4:8 c2:4; do
  c=${cf%%:*}; f=${cf##*:}
  AB2_DEBUG_FLAGS=$f ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:riccati_sweep_kernel -s 3 -c 1 --csv --log-file gpurun_out/pf_${c}_$f.csv \
    python bench.py --config $c --steps 2 --warmup 3 --no-cpu --no-e2e --no-parity --strong none > /dev/null 2>&1
  echo "$c flags $f: $(tail -3 gpurun_out/pf_${c}_$f.csv | awk -F'","' '{print $(NF-2), $NF}' | tr -d '"' | tr '\n' ' ')"
  AB2_DEBUG_FLAGS=$f python bench.py --config $c --steps 30 --warmup 5 --no-cpu --no-e2e --no-parity --strong none 2>/dev/null | tail -1 | python -c 'import sys,json
d=json.loads(sys.stdin.read()); print("   ms %.4f frac %.3f" % (d["ms_per_step"], d["roofline"]["frac"]))'
done
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
