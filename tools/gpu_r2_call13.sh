#!/bin/bash
# Round 2, GPU call 13: the driver's round-end sequence on the finalised attention kernels (all -m gpu tests, smoke, default bench incl. the
# CPU baseline), the other configs' rollout lines, and the ncu artefacts (launch list of one step, --set full of both attention kernels).
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/r13_test_all.log 2>&1; echo "pytest -m gpu exit $?: $(tail -n 1 gpurun_out/r13_test_all.log)"
grep -E "^FAILED|^E  " gpurun_out/r13_test_all.log | head -20
timeout 600 python __graft_entry__.py smoke > gpurun_out/r13_smoke.log 2>&1; echo "smoke exit $?"; tail -n 3 gpurun_out/r13_smoke.log
timeout 1200 python bench.py > gpurun_out/r13_bench.log 2>&1; echo "bench exit $?"; tail -n 1 gpurun_out/r13_bench.log | cut -c1-600
timeout 600 python bench.py --config flux1 --steps 1 --warmup 1 > gpurun_out/r13_bench_flux1.log 2>&1; tail -n 1 gpurun_out/r13_bench_flux1.log | cut -c1-300
timeout 900 python bench.py --config wan21 --steps 1 --warmup 1 > gpurun_out/r13_bench_wan21.log 2>&1; tail -n 1 gpurun_out/r13_bench_wan21.log | cut -c1-300
timeout 900 python bench.py --config qwen_image --steps 1 --warmup 1 > gpurun_out/r13_bench_qwen.log 2>&1; tail -n 1 gpurun_out/r13_bench_qwen.log | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/r13_launches.csv python bench.py --steps 1 --warmup 1 --batch 8 --num-inference-steps 2 --no-graph --skip-cpu-baseline > gpurun_out/r13_launchlist_bench.log 2>&1; echo "launch list exit $?"
for d in 64 128; do
  ATT_D=$d ATT_PRE=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:attention --launch-skip 2 --launch-count 1 -f -o gpurun_out/r13_att${d} python tools/attn_one.py > gpurun_out/r13_ncu_att${d}.log 2>&1; echo "ncu d$d exit $?"
done
