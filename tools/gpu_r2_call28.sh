#!/bin/bash
# Round 2, GPU call 28 (last of the round's budget): the suites call 27 did not run, on the final tree, + smoke().
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_elementwise.py tests/test_gpu_flux_engine.py tests/test_gpu_qwen_engine.py tests/test_gpu_vae.py tests/test_gpu_wan.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/r28_tests.log 2>&1; echo "pytest exit $?: $(tail -n 1 gpurun_out/r28_tests.log)"
grep -E "^FAILED|^E  " gpurun_out/r28_tests.log | head -12
timeout 200 python __graft_entry__.py smoke > gpurun_out/r28_smoke.log 2>&1; echo "smoke exit $?"; tail -n 3 gpurun_out/r28_smoke.log
