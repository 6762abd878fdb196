#!/bin/bash
# Everything the first GPU hour of the next round needs, in ONE gpurun call (getting a box is charged every time):
#   bash tools/build_variants.sh          # CPU, before the call
#   /usr/local/graft/bin/gpurun --timeout 3300 -- bash tools/gpu_round2.sh
# Stages are separate processes with their own timeouts (a trapped kernel only poisons its own CUDA context; every wait in the kernels is
# bounded, so a protocol bug ends in a tagged trap after ~4 s, not in a hang).  Results: gpurun_out/round2_summary.txt + per-stage logs.
mkdir -p gpurun_out
SUM=gpurun_out/round2_summary.txt; : > $SUM
stage() { echo "=== $1" | tee -a $SUM; }
stage "1 regression: validated suite + smoke + bench"
bash tools/gpu_full.sh > gpurun_out/stage1.log 2>&1; cat gpurun_out/summary.txt >> $SUM
stage "2 attention experiments (tools/gpu_maxfree.sh)"
timeout 1500 bash tools/gpu_maxfree.sh > gpurun_out/stage2.log 2>&1
cat gpurun_out/variants_tests.log gpurun_out/variants.log >> $SUM 2>/dev/null
stage "3 VAE decode, first run (tools/gpu_vae.sh)"
timeout 900 bash tools/gpu_vae.sh > gpurun_out/stage3.log 2>&1; tail -n 12 gpurun_out/stage3.log >> $SUM
stage "4 Wan2.1 engine, first run (tools/gpu_wan.sh)"
timeout 2000 bash tools/gpu_wan.sh > gpurun_out/stage4.log 2>&1; tail -n 12 gpurun_out/stage4.log >> $SUM
cat $SUM
