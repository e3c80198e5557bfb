run() { timeout 120 python bench.py --no-cpu --no-e2e --steps 20 --warmup 5 "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%-34s ms=%.4f frac=%.3f %s' % (' '.join(sys.argv[1:]), d['ms_per_step'], d['roofline']['frac'], d['config']['kernel']))" "$@"; }
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tensor or variants or matches" 2>&1 | tail -2
run
run --config c4
run --config c4 --variant 10
run --config c4 --variant 10 --ctas-per-sm 8
run --config c4 --variant 8
