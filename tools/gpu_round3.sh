#!/bin/bash
# tests (one process per file) + kernel micro-bench + smoke + bench.py
mkdir -p gpurun_out
: > gpurun_out/summary.txt
for f in ${FILES:-elementwise gemm attention engine}; do
  timeout ${TEST_TIMEOUT:-1200} python -m pytest tests/test_gpu_$f.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/test_$f.log 2>&1
  echo "$f exit $?" >> gpurun_out/summary.txt
  tail -n 4 gpurun_out/test_$f.log >> gpurun_out/summary.txt
done
timeout 600 python tools/kernel_bench.py > gpurun_out/kernel_bench.log 2>&1; echo "kernel_bench exit $?" >> gpurun_out/summary.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt; tail -n 2 gpurun_out/smoke.log >> gpurun_out/summary.txt
if [ -z "$SKIP_BENCH" ]; then
  timeout 1200 python bench.py --steps ${BSTEPS:-2} --warmup ${BWARM:-3} --batch ${BBATCH:-4} ${BEXTRA} > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/summary.txt; tail -n 3 gpurun_out/bench.log >> gpurun_out/summary.txt
fi
cat gpurun_out/summary.txt; cat gpurun_out/kernel_bench.log
