#!/bin/bash
# ncu launch list of ONE denoise step at the bench batch (cold numbers: compare shares, never a bench value)
mkdir -p gpurun_out
B=${B:-8}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/launches_b$B.csv \
  python bench.py --steps 1 --warmup 1 --batch $B --num-inference-steps 2 --no-graph --skip-cpu-baseline > gpurun_out/launchlist_b$B.log 2>&1
echo "ncu exit $?"; tail -n 3 gpurun_out/launchlist_b$B.log; wc -l gpurun_out/launches_b$B.csv
