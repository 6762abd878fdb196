#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/dbg_attention.py > gpurun_out/dbg_attention.log 2>&1; echo "dbg exit $?"; cat gpurun_out/dbg_attention.log
FILES="${FILES:-attention}" bash tools/gpu_round5.sh
