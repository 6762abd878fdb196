#!/bin/bash
# Round 2, GPU call 6: column-split head_dim-64 attention (two warps per row, double-buffered S) - parity, A/B against the row-per-thread
# kernel (FFB200_ATT_ROW=1) and across the polynomial share, bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_engine.py tests/test_gpu_parity_c2.py tests/test_gpu_hooks.py tests/test_gpu_elementwise.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/r6_tests.log 2>&1; echo "tests exit $?: $(tail -n 1 gpurun_out/r6_tests.log)"
grep -E "^FAILED|^E  " gpurun_out/r6_tests.log | head -20
python - <<'PY' 2>&1 | tail -2
from flow_factory_b200 import _lib
import ctypes as C
buf = (C.c_uint * 4)()
print("device error word:", _lib.lib().ffb200_device_error(C.byref(buf)), [hex(x) for x in buf])
PY
: > gpurun_out/r6_variants.log
for l in flow_factory_b200/libffb200.so flow_factory_b200/libffb200_exp_*.so; do
  ATT_PRE=1 FFB200_LIB=$PWD/$l timeout 120 python tools/attn_variants.py >> gpurun_out/r6_variants.log 2>&1
done
FFB200_ATT_ROW=1 ATT_PRE=1 timeout 120 python tools/attn_variants.py >> gpurun_out/r6_variants.log 2>&1
ATT_PRE=0 timeout 120 python tools/attn_variants.py >> gpurun_out/r6_variants.log 2>&1
cat gpurun_out/r6_variants.log
timeout 600 python bench.py --steps 2 --warmup 1 --skip-cpu-baseline > gpurun_out/r6_bench.log 2>&1; tail -n 1 gpurun_out/r6_bench.log | cut -c1-400
FFB200_ATT_ROW=1 timeout 600 python bench.py --steps 2 --warmup 1 --skip-cpu-baseline > gpurun_out/r6_bench_row.log 2>&1; tail -n 1 gpurun_out/r6_bench_row.log | cut -c1-400
