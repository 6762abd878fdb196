#!/bin/bash
# Round 2, GPU call 20: product = lean loops + 3 of 8 clustered + elected-lane arrives + warp-uniform warp index; polynomial share re-checked on
# that base (c2, c4, p4); the same warp-uniform trick in the GEMM kernel (libffb200_exp_gemmuwarp.so).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_engine.py tests/test_gpu_flux_ops.py tests/test_gpu_flux_engine.py tests/test_gpu_wan.py -q -m gpu --tb=short -p no:cacheprovider -x > gpurun_out/r20_tests.log 2>&1; echo "pytest exit $?: $(tail -n 1 gpurun_out/r20_tests.log)"
grep -E "^FAILED|^E  " gpurun_out/r20_tests.log | head -20
for rep in 1 2; do
for lib in libffb200.so libffb200_exp_c2.so libffb200_exp_c4.so libffb200_exp_p4.so; do
    ATT_PRE=1 FFB200_LIB=$PWD/flow_factory_b200/$lib timeout 300 python tools/attn_variants.py 2>/dev/null | tee -a gpurun_out/r20_attn_variants.jsonl | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('  %-28s pre=%d  d64 %.0f TFLOP/s (%.3f ms)   d128 %.0f (%.3f ms)  err %.2e %.2e' % (d['lib'], d['prescaled_keys'], d['tflops'], d['ms'], d['d128_tflops'], d['d128_ms'], d['trend_rel_err_d64'], d['trend_rel_err_d128']))"
done
done
show='
import sys, json
print("   " + "  ".join("%s %.0f/%.0f" % (d["name"], d["tflops"], d["cublas_tflops"]) for d in map(json.loads, sys.stdin)))'
for lib in libffb200.so libffb200_exp_gemmuwarp.so libffb200.so libffb200_exp_gemmuwarp.so; do
  FFB200_LIB=$PWD/flow_factory_b200/$lib timeout 600 python tools/kernel_bench.py 2>/dev/null | tee -a gpurun_out/r20_kernel_bench_$lib.jsonl | grep '"gemm"' | python -c "$show"
done
for lib in libffb200.so libffb200_exp_gemmuwarp.so; do
  FFB200_LIB=$PWD/flow_factory_b200/$lib timeout 900 python bench.py --skip-cpu-baseline --steps 2 --warmup 2 > gpurun_out/r20_bench_$lib.log 2>&1; echo "bench $lib exit $?"
  tail -n 1 gpurun_out/r20_bench_$lib.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('  value %.4f e2e %.4f  att %.0f (hot %.0f)  gemm %.0f (hot %.0f)  clk %s' % (d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['achieved_after_rollouts'], d['roofline_gemm']['achieved'], d['roofline_gemm']['achieved_after_rollouts'], d['clocks']['sm_mhz']))"
done
