#!/bin/bash
# Round 2, GPU call 3: new tests (hooks, attention guard path), polynomial pattern / share A/B on pre-scaled keys, ncu --set full with
# source for the head_dim 64 and 128 attention kernels, host CPU topology + the reference arm on this box's cores.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_hooks.py tests/test_gpu_attention.py tests/test_gpu_engine.py tests/test_gpu_flux_engine.py tests/test_gpu_qwen_engine.py tests/test_gpu_elementwise.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/r3_tests.log 2>&1; echo "tests exit $?: $(tail -n 1 gpurun_out/r3_tests.log)"
: > gpurun_out/r3_variants.log
for l in flow_factory_b200/libffb200.so flow_factory_b200/libffb200_exp_*.so; do
  ATT_PRE=1 FFB200_LIB=$PWD/$l timeout 120 python tools/attn_variants.py >> gpurun_out/r3_variants.log 2>&1
done
ATT_PRE=0 timeout 120 python tools/attn_variants.py >> gpurun_out/r3_variants.log 2>&1
cat gpurun_out/r3_variants.log
for d in 64 128; do
  ATT_D=$d ATT_PRE=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:attention --launch-skip 2 --launch-count 1 -f -o gpurun_out/r3_att${d}_pre python tools/attn_one.py > gpurun_out/r3_ncu_att${d}.log 2>&1; echo "ncu d$d exit $?"
done
lscpu > gpurun_out/r3_lscpu.txt 2>&1; nproc >> gpurun_out/r3_lscpu.txt
timeout 900 python bench.py --impl reference --steps 3 --warmup 2 > gpurun_out/r3_bench_ref.log 2>&1; echo "ref arm exit $?"; tail -n 1 gpurun_out/r3_bench_ref.log
ls -la gpurun_out/*.ncu-rep
