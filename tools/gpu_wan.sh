#!/bin/bash
# First GPU run of the Wan2.1 T2V engine (written after round 1's GPU budget was spent):
#   /usr/local/graft/bin/gpurun --timeout 900 -- bash tools/gpu_wan.sh
mkdir -p gpurun_out
export FFB200_PENDING=1
timeout 700 python -m pytest tests/test_gpu_wan.py -x -q 2>&1 | tail -40 | tee gpurun_out/wan_tests.log
python - <<'PY' 2>&1 | tee gpurun_out/wan_deverr.log
from flow_factory_b200 import _lib
import ctypes as C
buf = (C.c_uint * 4)()
print("device error word:", _lib.lib().ffb200_device_error(C.byref(buf)), [hex(x) for x in buf])
PY
timeout 1500 python tools/wan_bench.py --steps 1 --warmup 1 --num-inference-steps 10 2>&1 | tail -3 | tee gpurun_out/wan_bench_10step.json
