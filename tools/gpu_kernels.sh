#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
for f in ${FILES:-gemm attention engine}; do
  timeout ${TEST_TIMEOUT:-1200} python -m pytest tests/test_gpu_$f.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/test_$f.log 2>&1
  echo "$f exit $?" >> gpurun_out/summary.txt
  tail -n 4 gpurun_out/test_$f.log >> gpurun_out/summary.txt
done
timeout 600 python tools/kernel_bench.py > gpurun_out/kernel_bench.log 2>&1; echo "kernel_bench exit $?" >> gpurun_out/summary.txt
timeout 300 python tools/prof_waits.py > gpurun_out/prof_waits.json 2>&1; echo "prof_waits exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; cat gpurun_out/kernel_bench.log
