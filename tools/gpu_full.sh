#!/bin/bash
# the driver's round-end sequence: all -m gpu tests in one pytest process, smoke, bench (+ reference arm)
mkdir -p gpurun_out
: > gpurun_out/summary.txt
timeout 1800 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/test_all.log 2>&1; echo "pytest -m gpu exit $?" >> gpurun_out/summary.txt; tail -n 3 gpurun_out/test_all.log >> gpurun_out/summary.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt; tail -n 1 gpurun_out/smoke.log >> gpurun_out/summary.txt
timeout 1200 python bench.py > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/summary.txt; tail -n 1 gpurun_out/bench.log >> gpurun_out/summary.txt
if [ -n "$WITH_REF" ]; then
  timeout 900 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "bench ref exit $?" >> gpurun_out/summary.txt; tail -n 1 gpurun_out/bench_ref.log >> gpurun_out/summary.txt
fi
cat gpurun_out/summary.txt
