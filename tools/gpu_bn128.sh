#!/bin/bash
# A/B of the KV-128 attention experiment (-DFFB_ATT_BN128 build) against the product kernel: parity tests + isolated timing
mkdir -p gpurun_out; : > gpurun_out/variants.log
export V=$PWD/flow_factory_b200/libffb200_exp_bn128.so
FFB200_LIB=$V timeout 300 python -m pytest tests/test_gpu_attention.py tests/test_gpu_engine.py -m gpu -q --tb=short -p no:cacheprovider -k "not d128" > gpurun_out/test_bn128.log 2>&1; echo "bn128 tests exit $?"; tail -n 4 gpurun_out/test_bn128.log
for l in flow_factory_b200/libffb200.so $V; do FFB200_LIB=$l timeout 120 python tools/attn_variants.py >> gpurun_out/variants.log 2>&1; done
cat gpurun_out/variants.log
