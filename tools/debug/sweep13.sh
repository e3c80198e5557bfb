run() { timeout 120 python bench.py --no-cpu --no-e2e --steps 20 --warmup 5 "$@" 2>/dev/null | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); k=d['config']['kernel']; print('%-40s ms=%.4f frac=%.3f regs=%s ctas=%s thr=%s' % (' '.join(sys.argv[1:]), d['ms_per_step'], d['roofline']['frac'], k['regs_per_thread'], k['ctas_per_sm'], k['threads_per_cta']))
except Exception as e: print(' '.join(sys.argv[1:]), 'FAILED', e)" "$@"; }
for c in c1 c3; do
  for v in 0 1 4 5 6; do run --config $c --variant $v; done
  run --config $c --variant 1 --ctas-per-sm 7
  run --config $c --variant 6 --ctas-per-sm 8
done
