"""ctypes binding of the C ABI in include/ffb200.h.  The CUDA library is mandatory: there is no CPU or
PyTorch fallback - loading fails loudly when libffb200.so is missing (build with `python -m flow_factory_b200.build`)."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FFB200_LIB", os.path.join(_HERE, "libffb200.so"))

vp, ci, cf, cll, cull = C.c_void_p, C.c_int, C.c_float, C.c_longlong, C.c_ulonglong


class ModelConfig(C.Structure):
    _fields_ = [(n, ci) for n in ("num_layers", "num_heads", "patch_size", "in_channels", "joint_attention_dim",
                                  "pooled_projection_dim", "pos_embed_max_size", "num_dual_layers")]


LAYER_FIELDS = ("qkv_w", "qkv_b", "norm_q", "norm_k", "add_qkv_w", "add_qkv_b", "norm_added_q", "norm_added_k",
                "out_w", "out_b", "add_out_w", "add_out_b", "qkv2_w", "qkv2_b", "norm_q2", "norm_k2", "out2_w", "out2_b",
                "ff1_w", "ff1_b", "ff2_w", "ff2_b", "cff1_w", "cff1_b", "cff2_w", "cff2_b")


class LayerWeights(C.Structure):
    _fields_ = [(n, vp) for n in LAYER_FIELDS]


GLOBAL_FIELDS = ("pe_w", "pe_b", "pos_embed", "t1_w", "t1_b", "t2_w", "t2_b", "p1_w", "p1_b", "p2_w", "p2_b",
                 "ctx_w", "ctx_b", "mod_w", "mod_b", "proj_w", "proj_b")


class Weights(C.Structure):
    _fields_ = [(n, vp) for n in GLOBAL_FIELDS] + [("layers", C.POINTER(LayerWeights))]


class StepCoef(C.Structure):
    _fields_ = [(n, cf) for n in ("t_model", "sigma", "sigma_prev", "dt", "noise_level", "std_dev_t", "c_x", "c_v",
                                  "noise_scale", "two_var", "log_norm", "cps_a", "cps_b")] + \
               [(n, ci) for n in ("dynamics", "compute_log_prob", "store_slot", "logp_slot")]


class StepArgs(C.Structure):
    _fields_ = [("latents", vp), ("coef", StepCoef), ("guidance_scale", cf), ("noise", vp), ("seed", cull),
                ("step_index", ci), ("next_latents", vp), ("out_next_latents", vp), ("out_mean", vp),
                ("out_log_prob", vp), ("out_noise_pred", vp), ("overflow_flag", vp)]


class RolloutArgs(C.Structure):
    _fields_ = [("num_steps", ci), ("coefs", C.POINTER(StepCoef)), ("guidance_scale", cf), ("x0", vp), ("noise", vp),
                ("seed", cull), ("all_latents", vp), ("n_latent_slots", ci), ("store_initial_slot", ci),
                ("log_probs", vp), ("n_logp_slots", ci), ("final_latents", vp), ("overflow_flag", vp), ("use_graph", ci)]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: the B200 rollout engine has no fallback path. "
                           "Build it with `python -m flow_factory_b200.build` (needs nvcc, sm_100a).")
    L = C.CDLL(LIB_PATH)
    L.ffb200_last_error.restype = C.c_char_p
    L.ffb200_last_launch_count.restype = cll
    L.ffb200_plan_workspace_bytes.restype = cll
    L.ffb200_plan_workspace_bytes.argtypes = [vp]
    L.ffb200_device_error.argtypes = [C.POINTER(C.c_uint * 4)]
    L.ffb200_engine_create.argtypes = [C.POINTER(ModelConfig), C.POINTER(Weights), C.POINTER(vp)]
    L.ffb200_engine_set_weights.argtypes = [vp, C.POINTER(Weights)]
    L.ffb200_engine_destroy.argtypes = [vp]
    L.ffb200_engine_destroy.restype = None
    L.ffb200_engine_mod_rows.argtypes = [vp]
    L.ffb200_plan_create.argtypes = [vp, ci, ci, ci, ci, ci, C.POINTER(vp)]
    L.ffb200_plan_destroy.argtypes = [vp]
    L.ffb200_plan_destroy.restype = None
    L.ffb200_plan_set_prompts.argtypes = [vp, vp, vp, vp]
    L.ffb200_transformer_forward.argtypes = [vp, vp, cf, vp, vp]
    L.ffb200_step.argtypes = [vp, C.POINTER(StepArgs), vp]
    L.ffb200_rollout.argtypes = [vp, C.POINTER(RolloutArgs), vp]
    L.ffb200_rollout_host.argtypes = [vp, C.POINTER(RolloutArgs), vp, vp, vp]
    L.ffb200_linear.argtypes = [vp, ci, ci, cll, ci, ci, vp, ci, vp, vp, cll, ci, ci, ci, vp, cll, vp, vp, ci, cf, vp, vp]
    L.ffb200_linear_qkv_rope.argtypes = [vp, ci, ci, cll, ci, ci, vp, ci, vp, vp, cll, ci, ci, vp, vp, ci, cf, vp, vp, ci, vp]
    L.ffb200_attention.argtypes = [vp, ci, ci, ci, vp, vp]
    L.ffb200_attention_ex.argtypes = [vp, ci, ci, ci, ci, vp, ci, vp]
    L.ffb200_attention_scaled.argtypes = [vp, ci, ci, ci, ci, vp, ci, cf, ci, vp]
    L.ffb200_attention_normed.argtypes = [vp, ci, ci, ci, vp, vp, vp, vp, vp, vp]
    L.ffb200_ln_modulate.argtypes = [vp, ci, ci, ci, cf, vp, vp, vp, vp, vp, vp, cll, vp]
    L.ffb200_small_linear.argtypes = [vp, ci, ci, cll, vp, vp, ci, vp, cll, vp, cll, ci, vp]
    L.ffb200_sde_step.argtypes = [vp, vp, ci, ci, ci, ci, C.POINTER(StepCoef), vp, cull, ci, vp, vp, vp, vp, vp, vp]
    L.ffb200_sde_step_ex.argtypes = [vp, vp, ci, ci, ci, ci, C.POINTER(StepCoef), vp, cull, ci, vp, vp, vp, vp, vp, ci, vp]
    L.ffb200_plan_set_latent_dtype.argtypes = [vp, ci]
    _lib = L
    return L


EXPORTED_SYMBOLS = (
    "ffb200_last_error", "ffb200_device_error", "ffb200_abi_version", "ffb200_debug_read_prof", "ffb200_engine_create",
    "ffb200_engine_set_weights", "ffb200_engine_destroy", "ffb200_engine_mod_rows", "ffb200_plan_create",
    "ffb200_plan_destroy", "ffb200_plan_workspace_bytes", "ffb200_plan_set_prompts", "ffb200_transformer_forward",
    "ffb200_step", "ffb200_rollout", "ffb200_rollout_host", "ffb200_last_launch_count", "ffb200_linear",
    "ffb200_linear_qkv_rope", "ffb200_attention", "ffb200_attention_ex", "ffb200_attention_scaled", "ffb200_attention_normed", "ffb200_ln_modulate", "ffb200_small_linear", "ffb200_sde_step", "ffb200_sde_step_ex", "ffb200_plan_set_latent_dtype",
    # FLUX.1: bound in flow_factory_b200/flux.py
    "ffb200_flux_engine_create", "ffb200_flux_engine_set_weights", "ffb200_flux_engine_destroy", "ffb200_flux_engine_mod_rows",
    "ffb200_flux_plan_create", "ffb200_flux_plan_create_ex", "ffb200_flux_set_text_lengths", "ffb200_flux_plan_destroy", "ffb200_flux_plan_workspace_bytes", "ffb200_flux_set_prompts",
    "ffb200_flux_forward", "ffb200_flux_step", "ffb200_flux_rollout",
    # Wan2.1 T2V: bound in flow_factory_b200/wan.py
    "ffb200_wan_engine_create", "ffb200_wan_engine_set_weights", "ffb200_wan_engine_destroy", "ffb200_wan_plan_create",
    "ffb200_wan_plan_destroy", "ffb200_wan_plan_workspace_bytes", "ffb200_wan_set_prompts", "ffb200_wan_forward", "ffb200_wan_step",
    "ffb200_wan_rollout", "ffb200_wan_rms_rope", "ffb200_wan_layer_norm", "ffb200_wan_gate_residual", "ffb200_wan_patchify",
    "ffb200_attention_cross",
    # VAE decode: bound in flow_factory_b200/vae.py
    "ffb200_vae_weight_count", "ffb200_vae_decoder_create", "ffb200_vae_decoder_destroy", "ffb200_vae_decoder_workspace_bytes",
    "ffb200_vae_decode", "ffb200_conv2d_nhwc", "ffb200_group_norm_nhwc")


def check(code: int, what: str = "ffb200") -> None:
    if code != 0:
        msg = lib().ffb200_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed: {msg}")
