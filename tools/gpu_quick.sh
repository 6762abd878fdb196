#!/bin/bash
# quick regression + bench: elementwise/engine tests, then the default bench
mkdir -p gpurun_out
: > gpurun_out/summary.txt
for f in ${FILES:-elementwise engine}; do
  timeout 1200 python -m pytest tests/test_gpu_$f.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/test_$f.log 2>&1
  echo "$f exit $?" >> gpurun_out/summary.txt; tail -n 2 gpurun_out/test_$f.log >> gpurun_out/summary.txt
done
timeout 1200 python bench.py ${BENCH_ARGS:---skip-cpu-baseline} > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/summary.txt; tail -n 1 gpurun_out/bench.log >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
