run() { timeout 120 python bench.py --no-cpu --no-e2e --steps 20 --warmup 5 "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%-30s ms=%.4f frac=%.3f' % (' '.join(sys.argv[1:]), d['ms_per_step'], d['roofline']['frac']))" "$@"; }
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tensor or variants or matches" 2>&1 | tail -2
run
run --config c4
timeout 600 ncu --set full --import-source on --clock-control none -k regex:riccati_sweep -c 1 -o gpurun_out/c2_now python bench.py --no-cpu --no-e2e --steps 1 --warmup 1 > gpurun_out/ncu_c2.log 2>&1; tail -1 gpurun_out/ncu_c2.log
