#!/bin/bash
# One gpurun call: every -m gpu test file in its own process (a trapped kernel kills only that file), then micro-benchmarks.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
(ls /root/reference | head -3) > gpurun_out/ref_exists.txt 2>&1
: > gpurun_out/summary.txt
for f in ${FILES:-elementwise gemm attention engine}; do
  timeout ${TEST_TIMEOUT:-900} python -m pytest tests/test_gpu_$f.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/test_$f.log 2>&1
  echo "$f exit $?" >> gpurun_out/summary.txt
  tail -n 3 gpurun_out/test_$f.log >> gpurun_out/summary.txt
done
if [ -z "$SKIP_BENCH" ]; then
  timeout 600 python tools/kernel_bench.py > gpurun_out/kernel_bench.log 2>&1
  echo "kernel_bench exit $?" >> gpurun_out/summary.txt
fi
cat gpurun_out/summary.txt
