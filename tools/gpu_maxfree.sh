#!/bin/bash
# A/B of the max-free online softmax experiment (softmax.cuh, -DFFB_ATT_MAXFREE) against the product kernels.
# Build first (CPU):  for n in 2 3 4; do nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC \
#                          -DFFB_ATT_MAXFREE -DFFB_ATT_POLY_NUM=$n -o flow_factory_b200/libffb200_exp_maxfree_p$n.so flow_factory_b200/csrc/ffb200.cu; done
#                     cp flow_factory_b200/libffb200_exp_maxfree_p3.so flow_factory_b200/libffb200_exp_maxfree.so
#                     nvcc ... -DFFB_ATT_STAGGER=700 -o flow_factory_b200/libffb200_exp_stagger700.so ...   (lockstep test, product softmax)
#                     (with the row-max pass gone the ALU pipe has room: the best polynomial share of exp2 may move from 3 of 8)
# Run:                /usr/local/graft/bin/gpurun --timeout 900 -- bash tools/gpu_maxfree.sh
# Parity (attention, SD3.5 / FLUX / Qwen engines) with the experimental library, then isolated attention timings of both, then a bench.
mkdir -p gpurun_out; : > gpurun_out/variants.log
export V=$PWD/flow_factory_b200/libffb200_exp_maxfree.so
FFB200_LIB=$V timeout 600 python -m pytest tests/test_gpu_attention.py tests/test_gpu_engine.py tests/test_gpu_flux_engine.py tests/test_gpu_qwen_engine.py \
  -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/test_maxfree.log 2>&1; echo "maxfree tests exit $?"; tail -n 6 gpurun_out/test_maxfree.log
for l in flow_factory_b200/libffb200.so flow_factory_b200/libffb200_exp_maxfree*.so flow_factory_b200/libffb200_exp_stagger*.so; do FFB200_LIB=$PWD/$l timeout 120 python tools/attn_variants.py >> gpurun_out/variants.log 2>&1; done
cat gpurun_out/variants.log
FFB200_LIB=$V timeout 600 python bench.py --steps 2 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_maxfree.json 2> gpurun_out/bench_maxfree.err; tail -c 600 gpurun_out/bench_maxfree.json
