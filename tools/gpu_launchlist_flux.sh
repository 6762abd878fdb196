#!/bin/bash
# ncu launch list of ONE FLUX.1-dev denoise step (compare shares, never a bench value)
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/launches_flux.csv \
  python tools/flux_bench.py --steps 1 --warmup 1 --batch ${B:-2} --num-inference-steps 2 --no-graph > gpurun_out/launchlist_flux.log 2>&1
echo "ncu exit $?"; tail -n 2 gpurun_out/launchlist_flux.log | cut -c1-300; wc -l gpurun_out/launches_flux.csv
