#!/bin/bash
# Round 2, GPU call 18: the head_dim-128 attention with the peeled / single-common-body loop against the previous one
# (libffb200_exp_oldatt128.so), polynomial share variants of the new head_dim-64 loop (p1, p3, c3), FLUX.1 parity + rollout.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_flux_ops.py tests/test_gpu_flux_engine.py tests/test_gpu_qwen_engine.py tests/test_gpu_wan.py -q -m gpu --tb=short -p no:cacheprovider -x > gpurun_out/r18_tests.log 2>&1; echo "pytest exit $?: $(tail -n 1 gpurun_out/r18_tests.log)"
grep -E "^FAILED|^E  " gpurun_out/r18_tests.log | head -20
for rep in 1 2; do
for lib in libffb200.so libffb200_exp_oldatt128.so libffb200_exp_p1.so libffb200_exp_p3.so libffb200_exp_c3.so; do
    ATT_PRE=1 FFB200_LIB=$PWD/flow_factory_b200/$lib timeout 300 python tools/attn_variants.py 2>/dev/null | tee -a gpurun_out/r18_attn_variants.jsonl | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('  %-28s pre=%d  d64 %.0f TFLOP/s (%.3f ms)   d128 %.0f (%.3f ms)  err %.2e %.2e' % (d['lib'], d['prescaled_keys'], d['tflops'], d['ms'], d['d128_tflops'], d['d128_ms'], d['trend_rel_err_d64'], d['trend_rel_err_d128']))"
done
done
for lib in libffb200.so libffb200_exp_oldatt128.so; do
  FFB200_LIB=$PWD/flow_factory_b200/$lib timeout 900 python bench.py --config flux1 --steps 1 --warmup 1 > gpurun_out/r18_bench_flux1_$lib.log 2>&1; echo "flux bench $lib exit $?"
  tail -n 1 gpurun_out/r18_bench_flux1_$lib.log | cut -c1-160
done
