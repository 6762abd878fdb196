#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
for f in ${FILES:-gemm engine}; do
  timeout ${TEST_TIMEOUT:-1200} python -m pytest tests/test_gpu_$f.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/test_$f.log 2>&1
  echo "$f exit $?" >> gpurun_out/summary.txt
  tail -n 4 gpurun_out/test_$f.log >> gpurun_out/summary.txt
done
timeout 600 python tools/kernel_bench.py > gpurun_out/kernel_bench.log 2>&1; echo "kernel_bench exit $?" >> gpurun_out/summary.txt
for k in ${PROF:-attn gemm}; do
  pat=gemm_bf16; [ "$k" = "attn" ] && pat=attention_kernel
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$pat -s 2 -c 1 -f -o gpurun_out/prof_$k python tools/prof_kernels.py $k > gpurun_out/ncu_$k.log 2>&1
  echo "ncu $k exit $?" >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt; cat gpurun_out/kernel_bench.log
