#!/bin/bash
# Round 2, first GPU call: everything that was written after round 1's GPU budget ran out, each stage in its own process.
mkdir -p gpurun_out
export FFB200_PENDING=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_first_smi.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_stepwise.py -q --tb=short -p no:cacheprovider > gpurun_out/r2_stepwise.log 2>&1; echo "stepwise exit $?: $(tail -n 1 gpurun_out/r2_stepwise.log)"
timeout 600 python -m pytest tests/test_gpu_vae.py -q --tb=short -p no:cacheprovider > gpurun_out/r2_vae.log 2>&1; echo "vae exit $?: $(tail -n 1 gpurun_out/r2_vae.log)"
timeout 600 python -m pytest tests/test_gpu_wan.py -q --tb=short -p no:cacheprovider > gpurun_out/r2_wan.log 2>&1; echo "wan exit $?: $(tail -n 1 gpurun_out/r2_wan.log)"
timeout 1200 bash tools/gpu_maxfree.sh > gpurun_out/r2_maxfree.log 2>&1
cat gpurun_out/variants_tests.log gpurun_out/variants.log
