#!/bin/bash
# Builds the experimental attention libraries that tools/gpu_maxfree.sh A/Bs (CPU only, ~40 s each, 4 in parallel).
# They are git-ignored (*_exp_*.so) but travel to the GPU box with gpurun like the product library.
set -e
cd "$(dirname "$0")/.."
B="nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC"
S=flow_factory_b200/csrc/ffb200.cu
O=flow_factory_b200/libffb200_exp
run() { echo "build $1"; shift; "$@"; }
(
for n in 2 3 4; do echo "-DFFB_ATT_MAXFREE -DFFB_ATT_POLY_NUM=$n|maxfree_p$n"; done
echo "-DFFB_ATT_SUMMMA|summma"
echo "-DFFB_ATT_SUMMMA -DFFB_ATT_SUMMMA_NOWAIT|summma_nowait"
echo "-DFFB_ATT_SUMMMA -DFFB_ATT_MAXFREE|summma_maxfree"
echo "-DFFB_ATT_SUMMMA -DFFB_ATT_MAXFREE -DFFB_ATT_SUMMMA_NOWAIT|summma_maxfree_nowait"
for n in 1 2; do echo "-DFFB_ATT_SUMMMA -DFFB_ATT_MAXFREE -DFFB_ATT_POLY_NUM=$n|summma_maxfree_p$n"; done   # less FMA-pipe work left: the best polynomial share moves down
echo "-DFFB_ATT_STAGGER=700|stagger700"
echo "-DFFB_ATT_SPLIT|split128"              # head_dim 128 only: column-split softmax (max-free + tensor-core row sum built in)
echo "-DFFB_ATT_SPLIT -DFFB_ATT_POLY_NUM=2|split128_p2"
) | xargs -P 4 -I{} bash -c 'IFS="|" read -r flags name <<< "{}"; '"$B"' $flags -o '"$O"'_$name.so '"$S"' && echo built $name'
ls -la flow_factory_b200/libffb200_exp_*.so
