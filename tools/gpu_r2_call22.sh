#!/bin/bash
# Round 2, GPU call 22: persistent, prefetching ln_modulate (D <= 1536) against the one-row-per-warp kernel (FFB200_LN_PERSISTENT=0):
# parity, isolated bandwidth, whole rollout.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_elementwise.py tests/test_gpu_engine.py tests/test_gpu_flux_engine.py -q -m gpu --tb=short -p no:cacheprovider -x > gpurun_out/r22_tests.log 2>&1; echo "pytest exit $?: $(tail -n 1 gpurun_out/r22_tests.log)"
grep -E "^FAILED|^E  " gpurun_out/r22_tests.log | head -20
for rep in 1 2; do for m in 1 0; do FFB200_LN_PERSISTENT=$m timeout 300 python tools/ln_bench.py 2>/dev/null | tee -a gpurun_out/r22_ln_bench.jsonl | cut -c1-330; done; done
for m in 1 0; do
  FFB200_LN_PERSISTENT=$m timeout 900 python bench.py --skip-cpu-baseline --steps 2 --warmup 2 > gpurun_out/r22_bench_ln$m.log 2>&1; echo "bench persistent=$m exit $?"
  tail -n 1 gpurun_out/r22_bench_ln$m.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('  value %.4f e2e %.4f  att %.0f (hot %.0f)  gemm %.0f (hot %.0f)  clk %s' % (d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['achieved_after_rollouts'], d['roofline_gemm']['achieved'], d['roofline_gemm']['achieved_after_rollouts'], d['clocks']['sm_mhz']))"
done
