#!/bin/bash
# Round 2, 2-GPU call (gpurun --gpus 2): the multi-GPU path over real NCCL - SD3.5 bench at N=2, the FSDP2-sharded weight intake of the
# 20 B Qwen-Image model (ONE all-gather, timed) + its rollout, FLUX.1 at N=2.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/n2_bench_sd35.log 2>&1; echo "sd35 n2 exit $?"; grep '^{' gpurun_out/n2_bench_sd35.log | tail -n 1 | cut -c1-400
timeout 1200 $TR --master-port 29512 bench.py --config qwen_image --gpus 2 --steps 1 --warmup 1 > gpurun_out/n2_bench_qwen.log 2>&1; echo "qwen n2 exit $?"; grep '^{' gpurun_out/n2_bench_qwen.log | tail -n 1 | cut -c1-900
timeout 900 $TR --master-port 29513 bench.py --config flux1 --gpus 2 --steps 1 --warmup 1 > gpurun_out/n2_bench_flux1.log 2>&1; echo "flux n2 exit $?"; grep '^{' gpurun_out/n2_bench_flux1.log | tail -n 1 | cut -c1-400
timeout 900 $TR --master-port 29514 bench.py --config wan21 --gpus 2 --steps 1 --warmup 1 > gpurun_out/n2_bench_wan21.log 2>&1; echo "wan n2 exit $?"; grep '^{' gpurun_out/n2_bench_wan21.log | tail -n 1 | cut -c1-400
tail -n 5 gpurun_out/n2_bench_qwen.log | cut -c1-300
