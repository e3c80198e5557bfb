run() { timeout 120 python bench.py --no-cpu --no-e2e --steps 20 --warmup 5 "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%-30s ms=%.4f frac=%.3f' % (' '.join(sys.argv[1:]), d['ms_per_step'], d['roofline']['frac']))" "$@"; }
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
run
run --config c1
run --config c3
run --config c4
run --variant 8
