"""Qwen-Image rollout engine (SURVEY.md 8f row 4 / BASELINE config 5): the dual-stream FLUX engine in its `variant = 1` mode.

Mirrors, for the no-grad rollout path only:
  QwenImageTransformer2DModel.forward     DF/models/transformers/transformer_qwenimage.py:878-993
  QwenImageAdapter.forward                FF/models/qwen_image/qwen_image.py:476-600 (true CFG + per-token norm rescale)
What is NOT covered yet (loud errors, no silent approximation): prompts padded to different lengths (key masks), zero_cond_t /
additional_t_cond / layer-3D RoPE checkpoints, FSDP2-sharded weights (the 20 B model, 40 GB in bf16, is replicated per GPU)."""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import torch

from . import _lib
from .flux import FluxEngineConfig, FluxPlan, FluxRolloutEngine, _L, flux_make_schedule  # noqa: F401
from .scheduler import make_step_coef


def qwen_engine_config(cfg) -> FluxEngineConfig:
    """`cfg`: anything with the QwenImageTransformer2DModel config attributes (transformer_qwenimage.py:797-811)."""
    g = (lambda k, d=None: cfg.get(k, d)) if isinstance(cfg, dict) else (lambda k, d=None: getattr(cfg, k, d))
    if g("attention_head_dim") != 128 or g("in_channels") != 64 or g("patch_size") != 2 or (g("out_channels") or 64) != 16:
        raise ValueError("Qwen-Image: head_dim 128, packed in_channels 64, patch 2 x 2 x 16 output channels")
    for flag in ("zero_cond_t", "use_additional_t_cond", "use_layer3d_rope", "guidance_embeds"):
        if g(flag, False):
            raise NotImplementedError(f"Qwen-Image option {flag} is not on the accelerated path")
    return FluxEngineConfig(num_layers=g("num_layers"), num_single_layers=0, num_heads=g("num_attention_heads"), in_channels=64,
                            joint_attention_dim=g("joint_attention_dim"), pooled_projection_dim=8, guidance_embeds=False,
                            axes_dims_rope=tuple(g("axes_dims_rope")), variant=1)


def qwen_rope_tables(h2: int, w2: int, n_text: int, axes_dim: Sequence[int], theta: float = 10000.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """QwenEmbedRope with scale_rope=True for one frame (transformer_qwenimage.py:195-345): complex frequencies for the image grid
    (centred: negative indices for the first half of each axis) and for the text tokens (positions max(h2, w2)//2 ...), returned as
    the engine's table format: cos / sin fp32 [n_text + h2*w2, 128], every value repeated twice, TEXT rows first."""
    pos_index = torch.arange(4096)
    neg_index = torch.arange(4096).flip(0) * -1 - 1

    def params(index, dim):
        freqs = torch.outer(index, 1.0 / torch.pow(theta, torch.arange(0, dim, 2).to(torch.float32).div(dim)))
        return torch.polar(torch.ones_like(freqs), freqs)

    pos = torch.cat([params(pos_index, a) for a in axes_dim], dim=1)
    neg = torch.cat([params(neg_index, a) for a in axes_dim], dim=1)
    fp = pos.split([x // 2 for x in axes_dim], dim=1)
    fn = neg.split([x // 2 for x in axes_dim], dim=1)
    f_frame = fp[0][0:1].view(1, 1, 1, -1).expand(1, h2, w2, -1)
    f_h = torch.cat([fn[1][-(h2 - h2 // 2):], fp[1][: h2 // 2]], dim=0).view(1, h2, 1, -1).expand(1, h2, w2, -1)
    f_w = torch.cat([fn[2][-(w2 - w2 // 2):], fp[2][: w2 // 2]], dim=0).view(1, 1, w2, -1).expand(1, h2, w2, -1)
    vid = torch.cat([f_frame, f_h, f_w], dim=-1).reshape(h2 * w2, -1)
    start = max(h2 // 2, w2 // 2)
    freqs = torch.cat([pos[start: start + n_text], vid], dim=0)          # [S, 64] complex64
    cos = freqs.real.float().repeat_interleave(2, dim=1).contiguous()
    sin = freqs.imag.float().repeat_interleave(2, dim=1).contiguous()
    return cos, sin


class QwenRolloutEngine(FluxRolloutEngine):
    """QwenImageTransformer2DModel.state_dict() + config -> native engine; plans may carry a true-CFG batch."""

    def __init__(self, model_config, state_dict: Dict[str, torch.Tensor], device: Optional[torch.device] = None):
        cfg = model_config if isinstance(model_config, FluxEngineConfig) else qwen_engine_config(model_config)
        assert cfg.variant == 1
        super().__init__(cfg, state_dict, device)

    def rope_tables(self, h2: int, w2: int, n_text: int):
        return qwen_rope_tables(h2, w2, n_text, self.cfg.axes_dims_rope)

    def set_prompts(self, plan: FluxPlan, prompt_embeds: torch.Tensor, negative_prompt_embeds: Optional[torch.Tensor] = None,
                    guidance_scale: float = 1.0, prompt_lengths: Optional[Sequence[int]] = None,
                    negative_lengths: Optional[Sequence[int]] = None) -> None:
        """prompt_embeds [B, Nt, joint_dim] (right-padded to the plan's Nt); with plan.cfg the negative embeddings (same Nt) form the
        first half of the batch.  `*_lengths`: valid leading tokens per sample (encoder_hidden_states_mask of the reference); the padded
        tail is masked as attention keys."""
        if plan.cfg:
            if negative_prompt_embeds is None or negative_prompt_embeds.shape != prompt_embeds.shape:
                raise ValueError("true CFG needs negative_prompt_embeds of the same (unpadded) shape as prompt_embeds")
            pe = torch.cat([negative_prompt_embeds, prompt_embeds], dim=0)
        else:
            pe = prompt_embeds
        pe = pe.to(device=self.device, dtype=torch.bfloat16).contiguous()
        assert pe.shape == (plan.batch * (2 if plan.cfg else 1), plan.n_text, self.cfg.joint_attention_dim), pe.shape
        plan._keep = [pe]
        st = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(_L().ffb200_flux_set_prompts(plan.handle, pe.data_ptr(), None, float(guidance_scale), st), "ffb200_flux_set_prompts")
        full = [plan.n_text] * plan.batch
        lens = (list(negative_lengths or full) if plan.cfg else []) + list(prompt_lengths or full)
        import ctypes as C
        arr = (C.c_int * len(lens))(*[int(v) for v in lens])
        _lib.check(_L().ffb200_flux_set_text_lengths(plan.handle, arr, st), "ffb200_flux_set_text_lengths")

    @staticmethod
    def t_model(timestep: float, latents_dtype: torch.dtype = torch.float16) -> float:
        """What Timesteps(scale=1000) sees: `t.to(latents.dtype) / 1000` (qwen_image.py:520, 556) cast to the hidden dtype bf16
        (transformer_qwenimage.py:924)."""
        return float((torch.tensor(float(timestep)).to(latents_dtype) / 1000).to(torch.bfloat16).float())

    def transformer_forward(self, plan: FluxPlan, packed_latents: torch.Tensor, timestep: float) -> torch.Tensor:
        """-> noise prediction bf16 [B, Ni, 64] (after the CFG combine when plan.cfg)."""
        x = packed_latents.to(device=self.device, dtype=torch.float16).contiguous()
        assert tuple(x.shape) == (plan.batch, plan.n_img, 64)
        out = torch.empty_like(x, dtype=torch.bfloat16)
        st = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(_L().ffb200_flux_forward(plan.handle, x.data_ptr(), self.t_model(timestep), out.data_ptr(), st), "ffb200_flux_forward")
        return out

    def make_coefs(self, plan: FluxPlan, num_steps: int, noise_level: float, sde_steps: Sequence[int], dynamics: str = "ODE",
                   compute_log_prob: bool = False, store_slots: Optional[Sequence[int]] = None,
                   logp_slots: Optional[Sequence[int]] = None):
        """DGPO rollouts are ODE without log-probs (FF/trainers/dgpo.py:865-886); SDE dynamics work the same way as for FLUX.1."""
        timesteps, sigmas = flux_make_schedule(num_steps, plan.n_img)
        sde = set(int(i) for i in sde_steps)
        coefs = []
        for i in range(num_steps):
            nl = noise_level if (i in sde and dynamics != "ODE") else 0.0
            coefs.append(make_step_coef(float(sigmas[i]), float(sigmas[i + 1]), nl, float(sigmas[1]), dynamics,
                                        t_model=self.t_model(float(timesteps[i])), compute_log_prob=compute_log_prob and nl > 0,
                                        store_slot=-1 if store_slots is None else int(store_slots[i]),
                                        logp_slot=-1 if logp_slots is None else int(logp_slots[i])))
        return timesteps, sigmas, coefs
