#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 2 -c 1 -f -o gpurun_out/prof_attn python tools/prof_kernels.py attn > gpurun_out/ncu_attn.log 2>&1
echo "ncu exit $?"; ls -la gpurun_out/prof_attn.ncu-rep
