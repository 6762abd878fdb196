#!/bin/bash
# round-end check: compute-sanitizer on small cases of every kernel, then the driver's sequence (all gpu tests, smoke, bench)
bash tools/gpu_sanitize.sh > gpurun_out/sanitize_summary.txt 2>&1; cat gpurun_out/sanitize_summary.txt
bash tools/gpu_full.sh
