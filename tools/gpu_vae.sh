#!/bin/bash
# First GPU run of the VAE decode kernels (written after round 1's GPU budget was spent):
#   /usr/local/graft/bin/gpurun --timeout 900 -- bash tools/gpu_vae.sh
# The parity module is gated on FFB200_PENDING=1 until it has passed once.
mkdir -p gpurun_out
export FFB200_PENDING=1
timeout 300 python -m pytest tests/test_gpu_stepwise.py -q 2>&1 | tail -5 | tee gpurun_out/stepwise_tests.log
timeout 600 python -m pytest tests/test_gpu_vae.py -x -q 2>&1 | tail -40 | tee gpurun_out/vae_tests.log
python - <<'PY' 2>&1 | tee gpurun_out/vae_deverr.log
from flow_factory_b200 import _lib
import ctypes as C
buf = (C.c_uint * 4)()
print("device error word:", _lib.lib().ffb200_device_error(C.byref(buf)), [hex(x) for x in buf])
PY
timeout 300 python tools/vae_bench.py --res 512 --batch 2 --steps 3 2>&1 | tail -3 | tee gpurun_out/vae_bench_512.json
timeout 300 python tools/vae_bench.py --res 1024 --batch 2 --steps 3 2>&1 | tail -3 | tee gpurun_out/vae_bench_1024.json
