#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/summary.txt
for f in elementwise engine flux_engine; do
  timeout 900 python -m pytest tests/test_gpu_$f.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/test_$f.log 2>&1
  echo "$f exit $?" >> gpurun_out/summary.txt; tail -n 3 gpurun_out/test_$f.log >> gpurun_out/summary.txt
done
timeout 600 python tools/flux_bench.py --steps 2 --warmup 1 > gpurun_out/flux_bench.log 2>&1; echo "flux_bench exit $?" >> gpurun_out/summary.txt; tail -n 1 gpurun_out/flux_bench.log >> gpurun_out/summary.txt
timeout 900 python bench.py --skip-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/summary.txt; tail -n 1 gpurun_out/bench.log >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
