#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
timeout 1200 python -m pytest tests/test_gpu_engine.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/test_engine.log 2>&1; echo "engine exit $?" >> gpurun_out/summary.txt; tail -n 6 gpurun_out/test_engine.log >> gpurun_out/summary.txt
for k in attn gemm gemm_up; do
  pat=gemm_bf16; [ "$k" = "attn" ] && pat=attention_kernel
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$pat -s 2 -c 1 -f -o gpurun_out/prof_$k python tools/prof_kernels.py $k > gpurun_out/ncu_$k.log 2>&1
  echo "ncu $k exit $?" >> gpurun_out/summary.txt
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:final_step -s 2 -c 1 -f -o gpurun_out/prof_final_step python bench.py --steps 1 --warmup 1 --batch 2 --num-inference-steps 2 --no-graph --skip-cpu-baseline > gpurun_out/ncu_final.log 2>&1; echo "ncu final exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
