#!/bin/bash
# Round 2, GPU call 15: the remote tmem_empty arrive without MEMBAR.ALL.GPU; epilogue split policy A/B (FFB200_GEMM_EPI_SPLIT = 1 | 2 | auto)
# against the previous GEMM (libffb200_exp_oldgemm.so); full -m gpu suite with the tightened tolerances (measured errors -> parity_measured.jsonl).
mkdir -p gpurun_out; rm -f gpurun_out/parity_measured.jsonl
timeout 1500 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/r15_tests.log 2>&1; echo "pytest exit $?: $(tail -n 1 gpurun_out/r15_tests.log)"
grep -E "^FAILED|^E  " gpurun_out/r15_tests.log | head -20
show='
import sys, json
print("   " + "  ".join("%s %.0f/%.0f" % (d["name"], d["tflops"], d["cublas_tflops"]) for d in map(json.loads, sys.stdin)))'
for mode in 1 2 auto old; do
  lib=libffb200.so; [ $mode = old ] && lib=libffb200_exp_oldgemm.so
  env=""; [ $mode = 1 -o $mode = 2 ] && env="FFB200_GEMM_EPI_SPLIT=$mode"
  env $env FFB200_LIB=$PWD/flow_factory_b200/$lib timeout 600 python tools/kernel_bench.py > gpurun_out/r15_kernel_bench_$mode.jsonl 2> gpurun_out/r15_kernel_bench_$mode.err; echo "kernel_bench mode $mode exit $?"
  grep '"gemm"' gpurun_out/r15_kernel_bench_$mode.jsonl | python -c "$show"
done
for mode in auto 2 old; do
  lib=libffb200.so; [ $mode = old ] && lib=libffb200_exp_oldgemm.so
  env=""; [ $mode = 2 ] && env="FFB200_GEMM_EPI_SPLIT=2"
  env $env FFB200_LIB=$PWD/flow_factory_b200/$lib timeout 900 python bench.py --skip-cpu-baseline --steps 2 --warmup 2 > gpurun_out/r15_bench_$mode.log 2>&1; echo "bench mode $mode exit $?"
  tail -n 1 gpurun_out/r15_bench_$mode.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('  value %.4f e2e %.4f  att %.0f (hot %.0f)  gemm %.0f (hot %.0f)  clk %s' % (d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['achieved_after_rollouts'], d['roofline_gemm']['achieved'], d['roofline_gemm']['achieved_after_rollouts'], d['clocks']['sm_mhz']))"
done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_bf16 --launch-skip 2 --launch-count 1 -f -o gpurun_out/r15_gemm_mlp_up python tools/prof_kernels.py gemm_up_b16 > gpurun_out/r15_ncu_gemm_up.log 2>&1; echo "ncu gemm_up exit $?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_bf16 --launch-skip 2 --launch-count 1 -f -o gpurun_out/r15_gemm_attn_out python tools/prof_kernels.py gemm_out_b16 > gpurun_out/r15_ncu_gemm_out.log 2>&1; echo "ncu gemm_out exit $?"
