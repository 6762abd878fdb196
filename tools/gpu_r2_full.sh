#!/bin/bash
TAG=${TAG:-16}
# Round 2, GPU call ${TAG} (TAG env): the driver's round-end sequence on the finalised attention kernels (all -m gpu tests, smoke, default bench incl. the
# CPU baseline), the other configs' rollout lines, and the ncu artefacts (launch list of one step, --set full of both attention kernels).
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/r${TAG}_test_all.log 2>&1; echo "pytest -m gpu exit $?: $(tail -n 1 gpurun_out/r${TAG}_test_all.log)"
grep -E "^FAILED|^E  " gpurun_out/r${TAG}_test_all.log | head -20
timeout 600 python __graft_entry__.py smoke > gpurun_out/r${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -n 3 gpurun_out/r${TAG}_smoke.log
timeout 1200 python bench.py > gpurun_out/r${TAG}_bench.log 2>&1; echo "bench exit $?"; tail -n 1 gpurun_out/r${TAG}_bench.log | cut -c1-600
timeout 600 python bench.py --config flux1 --steps 1 --warmup 1 > gpurun_out/r${TAG}_bench_flux1.log 2>&1; tail -n 1 gpurun_out/r${TAG}_bench_flux1.log | cut -c1-300
timeout 900 python bench.py --config wan21 --steps 1 --warmup 1 > gpurun_out/r${TAG}_bench_wan21.log 2>&1; tail -n 1 gpurun_out/r${TAG}_bench_wan21.log | cut -c1-300
timeout 900 python bench.py --config qwen_image --steps 1 --warmup 1 > gpurun_out/r${TAG}_bench_qwen.log 2>&1; tail -n 1 gpurun_out/r${TAG}_bench_qwen.log | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/r${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --batch 8 --num-inference-steps 2 --no-graph --skip-cpu-baseline > gpurun_out/r${TAG}_launchlist_bench.log 2>&1; echo "launch list exit $?"
for d in 64 128; do
  ATT_D=$d ATT_PRE=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:attention --launch-skip 2 --launch-count 1 -f -o gpurun_out/r${TAG}_att${d} python tools/attn_one.py > gpurun_out/r${TAG}_ncu_att${d}.log 2>&1; echo "ncu d$d exit $?"
done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_bf16 --launch-skip 2 --launch-count 1 -f -o gpurun_out/r${TAG}_gemm_mlp_up python tools/prof_kernels.py gemm_up_b16 > gpurun_out/r${TAG}_ncu_gemm_up.log 2>&1; echo "ncu gemm_up exit $?"
ATT_D=64 ATT_PRE=1 ATT_B=16 timeout 600 ncu --set full --clock-control none -k regex:attention --launch-skip 2 --launch-count 1 -f -o gpurun_out/r${TAG}_att64_b16 python tools/attn_one.py > gpurun_out/r${TAG}_ncu_att64_b16.log 2>&1; echo "ncu att b16 exit $?"
