#!/bin/bash
# Round-2 GPU call 2: the new full-size / state-constraint / native-shape / 2x2-pivot tests,
# and the old suite with the tolerances restored to 1e-10 (collect every failure, do not stop).
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -s > gpurun_out/c2_fullsize.txt 2>&1
echo "fullsize exit $?" >> gpurun_out/c2_fullsize.txt
tail -40 gpurun_out/c2_fullsize.txt
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_lq_assemble.py -m gpu -q > gpurun_out/c2_parity.txt 2>&1
echo "parity exit $?" >> gpurun_out/c2_parity.txt
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/c2_parity.txt | head -40
