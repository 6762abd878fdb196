#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
timeout 900 python bench.py --steps ${BSTEPS:-2} --warmup ${BWARM:-3} --batch ${BBATCH:-4} > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 2 -c 1 -f -o gpurun_out/prof_attn python tools/prof_kernels.py attn > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn exit $?" >> gpurun_out/summary.txt
# launch list of ONE denoise step (no graph so every kernel is visible): warm-up rollout = 2 steps, timed = 2 steps
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --batch 2 --num-inference-steps 2 --no-graph --skip-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu launches exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -n 2 gpurun_out/bench.log
