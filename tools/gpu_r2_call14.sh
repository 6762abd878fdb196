#!/bin/bash
# Round 2, GPU call 14: the GEMM with eight epilogue warps (two per TMEM lane quadrant) - parity, A/B against the previous kernel
# (libffb200_exp_oldgemm.so, built from the last commit's gemm.cu) in isolation and on the whole rollout, ncu --set full of the MLP-up and
# attention launches at the bench shapes (DRAM traffic for profiles/ncu_traffic.json), SIMT rate micro-benchmark incl. 16-bit MUFU forms.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_flux_ops.py tests/test_gpu_engine.py tests/test_gpu_flux_engine.py tests/test_gpu_wan.py -q -m gpu --tb=short -p no:cacheprovider -x > gpurun_out/r14_tests.log 2>&1; echo "pytest exit $?: $(tail -n 1 gpurun_out/r14_tests.log)"
grep -E "^FAILED|^E  " gpurun_out/r14_tests.log | head -20
for lib in libffb200.so libffb200_exp_oldgemm.so; do
  FFB200_LIB=$PWD/flow_factory_b200/$lib timeout 600 python tools/kernel_bench.py > gpurun_out/r14_kernel_bench_$lib.jsonl 2> gpurun_out/r14_kernel_bench_$lib.err; echo "kernel_bench $lib exit $?"
  grep '"gemm"' gpurun_out/r14_kernel_bench_$lib.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('  %-14s %7.1f TFLOP/s  (cuBLAS %7.1f)' % (d['name'], d['tflops'], d['cublas_tflops']))"
done
for lib in libffb200.so libffb200_exp_oldgemm.so; do
  FFB200_LIB=$PWD/flow_factory_b200/$lib timeout 900 python bench.py --skip-cpu-baseline --steps 2 --warmup 2 > gpurun_out/r14_bench_$lib.log 2>&1; echo "bench $lib exit $?"
  tail -n 1 gpurun_out/r14_bench_$lib.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('  value %.4f e2e %.4f  att %.0f (hot %.0f)  gemm %.0f (hot %.0f)  clk %s' % (d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['achieved_after_rollouts'], d['roofline_gemm']['achieved'], d['roofline_gemm']['achieved_after_rollouts'], d['clocks']['sm_mhz']))"
done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_bf16 --launch-skip 2 --launch-count 1 -f -o gpurun_out/r14_gemm_mlp_up python tools/prof_kernels.py gemm_up_b16 > gpurun_out/r14_ncu_gemm_up.log 2>&1; echo "ncu gemm_up exit $?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_bf16 --launch-skip 2 --launch-count 1 -f -o gpurun_out/r14_gemm_attn_out python tools/prof_kernels.py gemm_out_b16 > gpurun_out/r14_ncu_gemm_out.log 2>&1; echo "ncu gemm_out exit $?"
ATT_D=64 ATT_PRE=1 ATT_B=16 timeout 600 ncu --set full --clock-control none -k regex:attention --launch-skip 2 --launch-count 1 -f -o gpurun_out/r14_att64_b16 python tools/attn_one.py > gpurun_out/r14_ncu_att64_b16.log 2>&1; echo "ncu att b16 exit $?"
timeout 120 tools/experiments/pipe_rates.bin > gpurun_out/r14_pipe_rates.jsonl 2>&1; echo "pipe rates exit $?"; grep -E "f16|bf16x2|tanh|rcp|mufu_ex2\"" gpurun_out/r14_pipe_rates.jsonl | grep '"warps_per_smsp": 4'
