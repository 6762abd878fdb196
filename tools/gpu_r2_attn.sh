#!/bin/bash
# Round 2, GPU call 2: regression of the whole -m gpu suite on the new softmax (no per-tile max, zero reference, pre-scaled keys), A/B of
# the polynomial share with and without pre-scaled keys, end-to-end bench with / without key pre-scaling, first VAE / Wan timings.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/r2_test_all.log 2>&1; echo "pytest -m gpu exit $?: $(tail -n 1 gpurun_out/r2_test_all.log)"
: > gpurun_out/r2_variants.log
for l in flow_factory_b200/libffb200.so flow_factory_b200/libffb200_exp_*.so; do
  for pre in 0 1; do
    ATT_PRE=$pre FFB200_LIB=$PWD/$l timeout 120 python tools/attn_variants.py >> gpurun_out/r2_variants.log 2>&1
  done
done
cat gpurun_out/r2_variants.log
timeout 600 python bench.py --steps 2 --warmup 1 --skip-cpu-baseline > gpurun_out/r2_bench_pre.log 2>&1; tail -n 1 gpurun_out/r2_bench_pre.log
FFB200_NO_PRESCALE=1 timeout 600 python bench.py --steps 2 --warmup 1 --skip-cpu-baseline > gpurun_out/r2_bench_nopre.log 2>&1; tail -n 1 gpurun_out/r2_bench_nopre.log
timeout 300 python tools/vae_bench.py --res 1024 --batch 2 --steps 3 > gpurun_out/r2_vae_bench_1024.log 2>&1; tail -n 2 gpurun_out/r2_vae_bench_1024.log
timeout 300 python tools/vae_bench.py --res 1024 --batch 8 --steps 3 > gpurun_out/r2_vae_bench_1024_b8.log 2>&1; tail -n 2 gpurun_out/r2_vae_bench_1024_b8.log
timeout 900 python tools/wan_bench.py --steps 1 --warmup 1 --num-inference-steps 10 > gpurun_out/r2_wan_bench.log 2>&1; tail -n 2 gpurun_out/r2_wan_bench.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_smoke.log 2>&1; echo "smoke exit $?: $(tail -n 1 gpurun_out/r2_smoke.log)"
