#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
[ -n "$LIBOVERRIDE" ] && export FFB200_LIB="$LIBOVERRIDE"
for f in ${FILES:-attention}; do
  timeout 900 python -m pytest tests/test_gpu_$f.py -m gpu -q --tb=short -p no:cacheprovider ${PYTEST_ARGS:-} > gpurun_out/test_$f.log 2>&1
  echo "$f exit $?" >> gpurun_out/summary.txt; tail -n 30 gpurun_out/test_$f.log >> gpurun_out/summary.txt
done
if [ -n "$WITH_BENCH" ]; then timeout 300 python tools/kernel_bench.py flux > gpurun_out/kernel_bench_flux.log 2>&1; echo "kernel_bench flux exit $?" >> gpurun_out/summary.txt; tail -3 gpurun_out/kernel_bench_flux.log >> gpurun_out/summary.txt; fi
cat gpurun_out/summary.txt; for j in gpurun_out/flux_fwd_*.json; do [ -f "$j" ] && cat "$j" && echo; done
if [ -n "$WITH_FLUX_BENCH" ]; then timeout 600 python tools/flux_bench.py --batch ${FLUX_B:-2} > gpurun_out/flux_bench.log 2>&1; echo "flux_bench exit $?"; tail -3 gpurun_out/flux_bench.log; fi
