run() { timeout 120 python bench.py --no-cpu --no-e2e --steps 20 --warmup 5 "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%-45s ms=%.4f frac=%.3f k=%s' % (' '.join(sys.argv[1:]), d['ms_per_step'], d['roofline']['frac'], d['config']['kernel']))" "$@"; }
run
run --ctas-per-sm 7
run --ctas-per-sm 6
run --stagger-ns 200
run --stagger-ns 400
run --stagger-ns 800
run --ctas-per-sm 7 --stagger-ns 400
run --ctas-per-sm 7 --stagger-ns 800
run --variant 8
run --variant 8 --stagger-ns 500
run --config c4
run --config c4 --stagger-ns 800
run --config c1
run --config c3
