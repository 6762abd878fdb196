#!/bin/bash
# compute-sanitizer memcheck + racecheck on small shapes of every kernel (slow: tiny cases only)
mkdir -p gpurun_out
SEL='test_gemm_identity_layout or test_gemm_batched_ragged or test_gemm_gate_residual or test_gemm_qkv_rmsnorm or matches_sdpa[2-77-2] or matches_sdpa[1-333-3] or test_ln_modulate or test_small_linear or test_sde_step_vs_reference_golden or forward_cfg_batching or test_fused_final_step_equals_unfused_composition[True-Flow-SDE] or attention_d128_matches_sdpa[2-77-2] or attention_d128_matches_sdpa[1-333-3] or attention_d128_strided or qkv_rmsnorm_rope_epilogue[2-333-2-256-5] or test_ln_modulate_d3072 or test_flux_forward_matches_oracle[tiny3] or test_flux_adapter_inference or test_gemm_gelu or test_gemm_rowtable or test_conv3x3_matches_torch[False-2-8-8-16-32] or test_conv3x3_matches_torch[True-3-5-6-64-8] or test_conv1x1_matches_linear[256-64-128] or test_group_norm_matches_torch[2-77-32-8-True] or test_decode_tiny_golden or test_op_rms_rope or test_op_layer_norm_and_gate_residual or test_op_attention_cross or test_step_and_rollout_consistency'
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 99 python -m pytest tests -m gpu -q -p no:cacheprovider -k "$SEL" > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool exit $?"; grep -E "ERROR SUMMARY|passed|failed|Error" gpurun_out/sanitize_$tool.log | tail -n 4
done
