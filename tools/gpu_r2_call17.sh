#!/bin/bash
# Round 2, GPU call 17: attention with peeled first / last tiles, ONE common-path body in the steady-state tiles and 64-register helper warps
# against the previous kernel (libffb200_exp_oldatt.so, built from the last commit's attention.cu): parity, isolated timing (engine layout
# and general path), whole rollout.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_engine.py tests/test_gpu_parity_c2.py tests/test_gpu_hooks.py -q -m gpu --tb=short -p no:cacheprovider -x > gpurun_out/r17_tests.log 2>&1; echo "pytest exit $?: $(tail -n 1 gpurun_out/r17_tests.log)"
grep -E "^FAILED|^E  " gpurun_out/r17_tests.log | head -20
for rep in 1 2; do
for lib in libffb200.so libffb200_exp_oldatt.so; do
  for pre in 1 0; do
    ATT_PRE=$pre FFB200_LIB=$PWD/flow_factory_b200/$lib timeout 300 python tools/attn_variants.py 2>/dev/null | tee -a gpurun_out/r17_attn_variants.jsonl | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('  %-26s pre=%d  d64 %.0f TFLOP/s (%.3f ms)   d128 %.0f   err %.2e' % (d['lib'], d['prescaled_keys'], d['tflops'], d['ms'], d['d128_tflops'], d['trend_rel_err_d64']))"
  done
done
done
for lib in libffb200.so libffb200_exp_oldatt.so; do
  FFB200_LIB=$PWD/flow_factory_b200/$lib timeout 900 python bench.py --skip-cpu-baseline --steps 2 --warmup 2 > gpurun_out/r17_bench_$lib.log 2>&1; echo "bench $lib exit $?"
  tail -n 1 gpurun_out/r17_bench_$lib.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('  value %.4f e2e %.4f  att %.0f (hot %.0f)  gemm %.0f (hot %.0f)  clk %s' % (d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['achieved_after_rollouts'], d['roofline_gemm']['achieved'], d['roofline_gemm']['achieved_after_rollouts'], d['clocks']['sm_mhz']))"
done
ATT_D=64 ATT_PRE=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:attention --launch-skip 2 --launch-count 1 -f -o gpurun_out/r17_att64 python tools/attn_one.py > gpurun_out/r17_ncu_att64.log 2>&1; echo "ncu d64 exit $?"
