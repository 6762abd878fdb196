#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
for f in ${FILES:-attention}; do
  FFB200_LIB=${LIBOVERRIDE:-} timeout 600 python -m pytest tests/test_gpu_$f.py -m gpu -q --tb=short -p no:cacheprovider ${PYTEST_ARGS:-} > gpurun_out/test_$f.log 2>&1
  echo "$f exit $?" >> gpurun_out/summary.txt; tail -n 6 gpurun_out/test_$f.log >> gpurun_out/summary.txt
done
timeout 300 python tools/kernel_bench.py flux > gpurun_out/kernel_bench_flux.log 2>&1; echo "kernel_bench flux exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; cat gpurun_out/kernel_bench_flux.log | tail -5
