#!/bin/bash
# Round 2, GPU call 10+: variant libraries (flow_factory_b200/libffb200_exp_*.so) against the product library, pre-scaled keys.
mkdir -p gpurun_out
: > gpurun_out/r12_variants.log
for l in flow_factory_b200/libffb200.so flow_factory_b200/libffb200_exp_*.so; do
  ATT_PRE=1 FFB200_LIB=$PWD/$l timeout 120 python tools/attn_variants.py >> gpurun_out/r12_variants.log 2>&1
done
cat gpurun_out/r12_variants.log | cut -c1-260
