#!/bin/bash
# two-GPU check of the prompt-sharded path (one all-gather per rollout) for both models; run with `gpurun --gpus 2`
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_n2.log 2>&1; echo "bench n2 exit $?"; tail -n 1 gpurun_out/bench_n2.log
timeout 900 $TR --master-port 29512 tools/flux_bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/flux_bench_n2.log 2>&1; echo "flux n2 exit $?"; tail -n 1 gpurun_out/flux_bench_n2.log
