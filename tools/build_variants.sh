#!/bin/bash
# Builds the A/B attention libraries that tools/gpu_r2_attn.sh times against the product library (CPU only, ~40 s each, 4 in parallel).
# They are git-ignored (*_exp_*.so) but travel to the GPU box with gpurun like the product library.
# Current set: the polynomial share of the exp2 work (softmax.cuh: FFB_ATT_POLY_NUM of every 8 element pairs).
set -e
cd "$(dirname "$0")/.."
rm -f flow_factory_b200/libffb200_exp_*.so
B="nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC"
S=flow_factory_b200/csrc/ffb200.cu
O=flow_factory_b200/libffb200_exp
(
for n in ${POLY_SET:-0 1 3 4}; do echo "-DFFB_ATT_POLY_NUM=$n|p$n"; done
for extra in "$@"; do echo "$extra"; done
) | xargs -P 4 -I{} bash -c 'IFS="|" read -r flags name <<< "{}"; '"$B"' $flags -o '"$O"'_$name.so '"$S"' && echo built $name'
ls -la flow_factory_b200/libffb200_exp_*.so
