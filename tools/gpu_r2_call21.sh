#!/bin/bash
# Round 2, GPU call 21: product = call 20's + warp-uniform warp index in the GEMM + 2 of 8 clustered at head_dim 128; A/B of dropping the warp
# barrier in front of the per-tile arrives (nosw) and of the two-at-a-time steady-state loop with compile-time barrier parities (u2).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_gemm.py tests/test_gpu_engine.py tests/test_gpu_flux_ops.py -q -m gpu --tb=short -p no:cacheprovider -x > gpurun_out/r21_tests.log 2>&1; echo "pytest exit $?: $(tail -n 1 gpurun_out/r21_tests.log)"
grep -E "^FAILED|^E  " gpurun_out/r21_tests.log | head -20
for v in nosw u2 u2nosw; do
  FFB200_LIB=$PWD/flow_factory_b200/libffb200_exp_$v.so timeout 600 python -m pytest tests/test_gpu_attention.py -q -m gpu --tb=line -p no:cacheprovider > gpurun_out/r21_tests_$v.log 2>&1; echo "attention tests with $v: exit $? $(tail -n 1 gpurun_out/r21_tests_$v.log)"
done
for rep in 1 2 3; do
for lib in libffb200.so libffb200_exp_nosw.so libffb200_exp_u2.so libffb200_exp_u2nosw.so; do
    ATT_PRE=1 FFB200_LIB=$PWD/flow_factory_b200/$lib timeout 300 python tools/attn_variants.py 2>/dev/null | tee -a gpurun_out/r21_attn_variants.jsonl | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('  %-28s pre=%d  d64 %.0f TFLOP/s (%.3f ms)   d128 %.0f (%.3f ms)  err %.2e %.2e' % (d['lib'], d['prescaled_keys'], d['tflops'], d['ms'], d['d128_tflops'], d['d128_ms'], d['trend_rel_err_d64'], d['trend_rel_err_d128']))"
done
done
for lib in libffb200.so libffb200_exp_u2nosw.so; do
  FFB200_LIB=$PWD/flow_factory_b200/$lib timeout 900 python bench.py --skip-cpu-baseline --steps 2 --warmup 2 > gpurun_out/r21_bench_$lib.log 2>&1; echo "bench $lib exit $?"
  tail -n 1 gpurun_out/r21_bench_$lib.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('  value %.4f e2e %.4f  att %.0f (hot %.0f)  gemm %.0f (hot %.0f)  clk %s' % (d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['achieved_after_rollouts'], d['roofline_gemm']['achieved'], d['roofline_gemm']['achieved_after_rollouts'], d['clocks']['sm_mhz']))"
done
