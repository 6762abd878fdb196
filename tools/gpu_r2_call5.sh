#!/bin/bash
# Round 2, GPU call 5: S-prefetch pipeline with prompt p_full release (the deferred release of call 4 cost 10-28 %), A/B against no prefetch;
# storage-dtype and clamp tests; bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_parity_c2.py tests/test_gpu_engine.py tests/test_gpu_elementwise.py tests/test_gpu_flux_engine.py tests/test_gpu_wan.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/r5_tests.log 2>&1; echo "tests exit $?: $(tail -n 1 gpurun_out/r5_tests.log)"
grep -E "^FAILED|^E  " gpurun_out/r5_tests.log | head -20
: > gpurun_out/r5_variants.log
for l in flow_factory_b200/libffb200.so flow_factory_b200/libffb200_exp_*.so; do
  ATT_PRE=1 FFB200_LIB=$PWD/$l timeout 120 python tools/attn_variants.py >> gpurun_out/r5_variants.log 2>&1
done
cat gpurun_out/r5_variants.log
timeout 600 python bench.py --steps 2 --warmup 1 --skip-cpu-baseline > gpurun_out/r5_bench.log 2>&1; tail -n 1 gpurun_out/r5_bench.log | cut -c1-400
