#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:attention_kernel -s 2 -c 1 -o gpurun_out/prof_attn_b16 python tools/prof_kernels.py attn_b16 > gpurun_out/ncu_attn_b16.log 2>&1; echo "ncu exit $?"
timeout 900 python bench.py --skip-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench exit $?"; tail -n 1 gpurun_out/bench.log | cut -c1-200
python - <<'PY'
import json
l=open('gpurun_out/bench.log').read().strip().splitlines()[-1]
d=json.loads(l); print(json.dumps({k:d[k] for k in ('value','e2e','roofline','roofline_gemm','clocks')})[:1800])
PY
