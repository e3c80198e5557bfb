#!/bin/bash
# Round-2 GPU call 1: full GPU test suite (incl. the merged parametric kernels), sanitizer
# passes over every kernel family, ncu of the FP64 micro-benchmark (what is the DMMA peak?).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c1_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c1_pytest.txt 2>&1
echo "pytest exit $?" >> gpurun_out/c1_pytest.txt
tail -5 gpurun_out/c1_pytest.txt
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/gpu/sanitize_cases.py \
      > gpurun_out/c1_sanitize_$tool.txt 2>&1
  echo "$tool exit $?" >> gpurun_out/c1_sanitize_$tool.txt
  tail -4 gpurun_out/c1_sanitize_$tool.txt
done
cd tools/micro && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o fp64_rates fp64_rates.cu && cd ../..
./tools/micro/fp64_rates > gpurun_out/c1_fp64_rates.txt 2>&1
timeout 300 ncu --clock-control none --metrics sm__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_tensor_op_dmma.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_op_dmma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,sm__cycles_elapsed.max,gpu__time_duration.sum,smsp__inst_executed.sum \
    --csv --log-file gpurun_out/c1_fp64_rates_ncu.csv ./tools/micro/fp64_rates > /dev/null 2>&1
echo "ncu exit $?"
timeout 300 ncu --clock-control none --query-metrics 2>/dev/null | grep -i -E "dmma|fp64|pipe_shared|pipe_tensor" > gpurun_out/c1_metric_names.txt
cat gpurun_out/c1_fp64_rates.txt
