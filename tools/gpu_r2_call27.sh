#!/bin/bash
# Round 2, GPU call 27: the range proof of the polynomial exp2 slots from the RMSNorm weights (ffb200_attention_normed / the SD3.5 engine's
# attention launches): parity (bit-identical to the checked kernel), isolated timing on the same input, whole rollout with and without
# (FFB200_NO_SCORE_BOUND=1).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_engine.py tests/test_gpu_parity_c2.py tests/test_gpu_hooks.py tests/test_gpu_flux_ops.py tests/test_gpu_stepwise.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/r27_tests.log 2>&1; echo "pytest exit $?: $(tail -n 1 gpurun_out/r27_tests.log)"
grep -E "^FAILED|^E  " gpurun_out/r27_tests.log | head -20
for rep in 1 2 3; do
  ATT_PRE=1 timeout 300 python tools/attn_variants.py 2>/dev/null | tee -a gpurun_out/r27_attn_variants.jsonl | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('  d64 %.0f   normed (proof) %.0f   same input, checked %.0f   d128 %.0f' % (d['tflops'], d['d64_normed_tflops'], d['d64_same_input_checked_tflops'], d['d128_tflops']))"
done
for m in "" 1; do
  env ${m:+FFB200_NO_SCORE_BOUND=1} timeout 900 python bench.py --skip-cpu-baseline --steps 2 --warmup 2 > gpurun_out/r27_bench_nobound$m.log 2>&1; echo "bench NO_SCORE_BOUND=${m:-0} exit $?"
  tail -n 1 gpurun_out/r27_bench_nobound$m.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('  value %.4f e2e %.4f  att %.0f (hot %.0f)  gemm %.0f (hot %.0f)  clk %s' % (d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['achieved_after_rollouts'], d['roofline_gemm']['achieved'], d['roofline_gemm']['achieved_after_rollouts'], d['clocks']['sm_mhz']))"
done
