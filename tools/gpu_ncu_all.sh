#!/bin/bash
# one `ncu --set full` capture per kernel of the final build (never a bench number); reports come back in gpurun_out/
mkdir -p gpurun_out
: > gpurun_out/summary.txt
NCU="ncu --set full --clock-control none --import-source on -f"
for k in attn gemm_b16 gemm_up_b16 gemm_out_b16; do
  pat=gemm_bf16; [ "$k" = "attn" ] && pat=attention_kernel
  timeout 600 $NCU -k regex:$pat -s 2 -c 1 -o gpurun_out/prof_$k python tools/prof_kernels.py $k > gpurun_out/ncu_$k.log 2>&1
  echo "ncu $k exit $?" >> gpurun_out/summary.txt
done
B="python bench.py --steps 1 --warmup 1 --batch 8 --num-inference-steps 2 --no-graph --skip-cpu-baseline"
timeout 900 $NCU -k regex:final_step -s 2 -c 1 -o gpurun_out/prof_final_step $B > gpurun_out/ncu_final.log 2>&1; echo "ncu final exit $?" >> gpurun_out/summary.txt
timeout 900 $NCU -k regex:ln_modulate -s 4 -c 1 -o gpurun_out/prof_ln_modulate $B > gpurun_out/ncu_ln.log 2>&1; echo "ncu ln exit $?" >> gpurun_out/summary.txt
timeout 900 $NCU -k regex:small_linear -s 5 -c 1 -o gpurun_out/prof_small_linear $B > gpurun_out/ncu_sl.log 2>&1; echo "ncu small_linear exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; ls -la gpurun_out/*.ncu-rep
