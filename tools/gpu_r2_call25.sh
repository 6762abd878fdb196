#!/bin/bash
# Round 2, GPU call 25: head_dim-128 attention with the row sum on the tensor core (-DFFB_ATT128_TCSUM; polynomial share 2 / 3 / 4 of 8)
# against the product: parity of every head_dim-128 user (attention, FLUX.1, Qwen-Image, Wan2.1 tests) with the TCSUM library, timing.
mkdir -p gpurun_out
for v in tcsum; do
  FFB200_LIB=$PWD/flow_factory_b200/libffb200_exp_$v.so timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_flux_ops.py tests/test_gpu_flux_engine.py tests/test_gpu_qwen_engine.py tests/test_gpu_wan.py tests/test_gpu_hooks.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/r25_tests_$v.log 2>&1; echo "tests with $v: exit $? $(tail -n 1 gpurun_out/r25_tests_$v.log)"
  grep -E "^FAILED|^E  " gpurun_out/r25_tests_$v.log | head -12
done
for rep in 1 2; do
for lib in libffb200.so libffb200_exp_tcsum.so libffb200_exp_tcsum_c3.so libffb200_exp_tcsum_c4.so; do
    ATT_PRE=1 FFB200_LIB=$PWD/flow_factory_b200/$lib timeout 300 python tools/attn_variants.py 2>/dev/null | tee -a gpurun_out/r25_attn_variants.jsonl | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('  %-28s pre=%d  d64 %.0f TFLOP/s (%.3f ms)   d128 %.0f (%.3f ms)  err %.2e %.2e' % (d['lib'], d['prescaled_keys'], d['tflops'], d['ms'], d['d128_tflops'], d['d128_ms'], d['trend_rel_err_d64'], d['trend_rel_err_d128']))"
done
done
for lib in libffb200.so libffb200_exp_tcsum.so; do
  FFB200_LIB=$PWD/flow_factory_b200/$lib timeout 900 python bench.py --config wan21 --steps 1 --warmup 1 > gpurun_out/r25_bench_wan21_$lib.log 2>&1; echo "wan bench $lib exit $?"
  tail -n 1 gpurun_out/r25_bench_wan21_$lib.log | cut -c1-160
done
