#!/bin/bash
# Round 2, GPU call 24: compute-sanitizer (memcheck + racecheck) on small cases of every kernel of the final tree; the general-path lean tile
# body; kernel_bench with cuBLAS / cuDNN SDPA beside every kernel (engine layout and general path).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_hooks.py tests/test_gpu_flux_ops.py tests/test_gpu_engine.py -q -m gpu --tb=short -p no:cacheprovider -x > gpurun_out/r24_tests.log 2>&1; echo "pytest exit $?: $(tail -n 1 gpurun_out/r24_tests.log)"
grep -E "^FAILED|^E  " gpurun_out/r24_tests.log | head -20
bash tools/gpu_sanitize.sh
timeout 600 python tools/kernel_bench.py > gpurun_out/r24_kernel_bench.jsonl 2> gpurun_out/r24_kernel_bench.err; echo "kernel_bench exit $?"
timeout 300 python tools/kernel_bench.py flux >> gpurun_out/r24_kernel_bench.jsonl 2>> gpurun_out/r24_kernel_bench.err; echo "kernel_bench flux exit $?"
grep attention gpurun_out/r24_kernel_bench.jsonl | cut -c1-330
for pre in 1 0; do ATT_PRE=$pre timeout 300 python tools/attn_variants.py 2>/dev/null | tee -a gpurun_out/r24_attn_variants.jsonl | cut -c1-260; done
timeout 900 python bench.py --skip-cpu-baseline --steps 2 --warmup 2 > gpurun_out/r24_bench.log 2>&1; echo "bench exit $?"
tail -n 1 gpurun_out/r24_bench.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('  value %.4f e2e %.4f  att %.0f (hot %.0f)  gemm %.0f (hot %.0f)  clk %s' % (d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['achieved_after_rollouts'], d['roofline_gemm']['achieved'], d['roofline_gemm']['achieved_after_rollouts'], d['clocks']['sm_mhz']))"
