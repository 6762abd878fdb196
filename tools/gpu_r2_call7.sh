#!/bin/bash
# Round 2, GPU call 7: head_dim-64 kernel layouts A/B: row3 (product), row2 (2 sub-tiles, double-buffered S), split2.
mkdir -p gpurun_out
for v in row2; do
  FFB200_ATT_VARIANT=$v timeout 600 python -m pytest tests/test_gpu_attention.py tests/test_gpu_engine.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/r7_tests_$v.log 2>&1; echo "tests[$v] exit $?: $(tail -n 1 gpurun_out/r7_tests_$v.log)"
done
: > gpurun_out/r7_variants.log
for l in flow_factory_b200/libffb200.so flow_factory_b200/libffb200_exp_p1.so; do
  for v in row3 row2 split2; do
    FFB200_ATT_VARIANT=$v ATT_PRE=1 FFB200_LIB=$PWD/$l timeout 120 python tools/attn_variants.py >> gpurun_out/r7_variants.log 2>&1
  done
done
cat gpurun_out/r7_variants.log
