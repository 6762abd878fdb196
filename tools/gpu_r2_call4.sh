#!/bin/bash
# Round 2, GPU call 4: software-pipelined softmax loops (d64 + d128) - full regression incl. the new C2-size parity tests, variant A/B,
# end-to-end bench.
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/r4_test_all.log 2>&1; echo "pytest -m gpu exit $?: $(tail -n 1 gpurun_out/r4_test_all.log)"
grep -E "FAILED|Error|passed|failed" gpurun_out/r4_test_all.log | tail -15
: > gpurun_out/r4_variants.log
for l in flow_factory_b200/libffb200.so flow_factory_b200/libffb200_exp_*.so; do
  ATT_PRE=1 FFB200_LIB=$PWD/$l timeout 120 python tools/attn_variants.py >> gpurun_out/r4_variants.log 2>&1
done
ATT_PRE=0 timeout 120 python tools/attn_variants.py >> gpurun_out/r4_variants.log 2>&1
cat gpurun_out/r4_variants.log
timeout 600 python bench.py --steps 2 --warmup 1 --skip-cpu-baseline > gpurun_out/r4_bench.log 2>&1; tail -n 1 gpurun_out/r4_bench.log | cut -c1-900
timeout 600 python tools/flux_bench.py --steps 1 --warmup 1 > gpurun_out/r4_flux_bench.log 2>&1; tail -n 1 gpurun_out/r4_flux_bench.log | cut -c1-600
cat gpurun_out/parity_c2_rollout.json 2>/dev/null | cut -c1-3000
cat gpurun_out/cross_path_ratio_tiny.json gpurun_out/cross_path_ratio_mid.json 2>/dev/null | cut -c1-600
