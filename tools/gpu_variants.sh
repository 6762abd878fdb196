#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/variants.log
for l in flow_factory_b200/libffb200.so flow_factory_b200/libffb200_exp_*.so; do
  FFB200_LIB=$PWD/$l timeout 120 python tools/attn_variants.py >> gpurun_out/variants.log 2>&1
done
cat gpurun_out/variants.log
