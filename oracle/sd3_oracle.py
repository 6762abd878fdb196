"""CPU/torch restatement of the reference rollout path (TEST INFRASTRUCTURE - see oracle/__init__.py).

Every function cites the reference file:line it follows.  All paths relative to /root/reference:
  FF = src/flow_factory ; DF = diffusers/src/diffusers

The functions use the same torch functional ops, in the same order, as the reference modules, so
that running them under `torch.autocast(device, torch.bfloat16)` reproduces the reference's
autocast numerics on either device (CPU autocast keeps layer_norm in bf16, CUDA autocast promotes it
to fp32 - SURVEY.md section 8(a)); run without autocast on fp32 weights they give the fp32 ground truth.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------
# Model config (DF/models/transformers/transformer_sd3.py:117-141)
# ----------------------------------------------------------------------------------------------
@dataclass
class SD3Config:
    sample_size: int = 128
    patch_size: int = 2
    in_channels: int = 16
    num_layers: int = 24
    attention_head_dim: int = 64
    num_attention_heads: int = 24
    joint_attention_dim: int = 4096
    caption_projection_dim: int = 1536
    pooled_projection_dim: int = 2048
    out_channels: int = 16
    pos_embed_max_size: int = 384
    dual_attention_layers: Tuple[int, ...] = tuple(range(13))
    qk_norm: str = "rms_norm"

    @property
    def inner_dim(self) -> int:
        return self.num_attention_heads * self.attention_head_dim

    def ref_kwargs(self) -> dict:
        return dict(sample_size=self.sample_size, patch_size=self.patch_size, in_channels=self.in_channels,
                    num_layers=self.num_layers, attention_head_dim=self.attention_head_dim,
                    num_attention_heads=self.num_attention_heads, joint_attention_dim=self.joint_attention_dim,
                    caption_projection_dim=self.caption_projection_dim,
                    pooled_projection_dim=self.pooled_projection_dim, out_channels=self.out_channels,
                    pos_embed_max_size=self.pos_embed_max_size,
                    dual_attention_layers=tuple(self.dual_attention_layers), qk_norm=self.qk_norm)


def sd35_medium() -> SD3Config:
    """SD3.5-medium (HF model-card values)."""
    return SD3Config()


def tiny_config(num_layers: int = 2, heads: int = 2, dual: Sequence[int] = (0,), joint_dim: int = 64,
                pooled_dim: int = 32, pos_max: int = 16, sample_size: int = 16) -> SD3Config:
    return SD3Config(sample_size=sample_size, num_layers=num_layers, num_attention_heads=heads,
                     joint_attention_dim=joint_dim, caption_projection_dim=heads * 64,
                     pooled_projection_dim=pooled_dim, pos_embed_max_size=pos_max,
                     dual_attention_layers=tuple(dual))


# ----------------------------------------------------------------------------------------------
# 2-D sin-cos position table (DF/models/embeddings.py:264-384 get_2d_sincos_pos_embed, output_type="pt")
# ----------------------------------------------------------------------------------------------
def _sincos_1d(embed_dim: int, pos: torch.Tensor) -> torch.Tensor:
    # DF/models/embeddings.py:358-384 (flip_sin_to_cos=False): omega in float64
    omega = torch.arange(embed_dim // 2, dtype=torch.float64)
    omega /= embed_dim / 2.0
    omega = 1.0 / 10000 ** omega
    out = torch.outer(pos.reshape(-1), omega)
    return torch.concat([torch.sin(out), torch.cos(out)], dim=1)


def sincos_pos_embed_2d(embed_dim: int, grid_size: int, base_size: int, interpolation_scale: float = 1.0) -> torch.Tensor:
    gs = (grid_size, grid_size)
    grid_h = torch.arange(gs[0], dtype=torch.float32) / (gs[0] / base_size) / interpolation_scale
    grid_w = torch.arange(gs[1], dtype=torch.float32) / (gs[1] / base_size) / interpolation_scale
    grid = torch.meshgrid(grid_w, grid_h, indexing="xy")  # w first (DF embeddings.py:311)
    grid = torch.stack(grid, dim=0).reshape([2, 1, gs[1], gs[0]])
    emb_h = _sincos_1d(embed_dim // 2, grid[0])
    emb_w = _sincos_1d(embed_dim // 2, grid[1])
    return torch.concat([emb_h, emb_w], dim=1).float()  # (grid*grid, D)


# ----------------------------------------------------------------------------------------------
# Deterministic random-init weights with diffusers state-dict key names
# ----------------------------------------------------------------------------------------------
def make_weights(cfg: SD3Config, seed: int = 0, dtype: torch.dtype = torch.float32, device: str = "cpu") -> Dict[str, torch.Tensor]:
    """Random weights (seeded, checkpoint-independent) keyed like SD3Transformer2DModel.state_dict().
    Scale ~ 1/sqrt(fan_in) keeps activations O(1) through 24 layers; adaLN rows get a smaller scale.
    `device='cuda'` draws from a CUDA generator (different stream than CPU; used only for full-size GPU tests)."""
    g = torch.Generator(device=device).manual_seed(seed)
    _randn = torch.randn
    def randn(*shape, generator=None):
        return _randn(*shape, generator=generator, device=device)
    D = cfg.inner_dim
    w: Dict[str, torch.Tensor] = {}

    def lin(name: str, out_f: int, in_f: int, scale: float = 1.0, bias_std: float = 0.02):
        w[name + ".weight"] = randn(out_f, in_f, generator=g) * (scale / math.sqrt(in_f))
        w[name + ".bias"] = randn(out_f, generator=g) * bias_std

    p = cfg.patch_size
    w["pos_embed.proj.weight"] = randn(D, cfg.in_channels, p, p, generator=g) / math.sqrt(cfg.in_channels * p * p)
    w["pos_embed.proj.bias"] = randn(D, generator=g) * 0.02
    w["pos_embed.pos_embed"] = sincos_pos_embed_2d(D, cfg.pos_embed_max_size, cfg.sample_size // p).unsqueeze(0).to(device)
    lin("time_text_embed.timestep_embedder.linear_1", D, 256)
    lin("time_text_embed.timestep_embedder.linear_2", D, D)
    lin("time_text_embed.text_embedder.linear_1", D, cfg.pooled_projection_dim)
    lin("time_text_embed.text_embedder.linear_2", D, D)
    lin("context_embedder", cfg.caption_projection_dim, cfg.joint_attention_dim)
    for i in range(cfg.num_layers):
        pre = f"transformer_blocks.{i}."
        last = i == cfg.num_layers - 1
        dual = i in cfg.dual_attention_layers
        lin(pre + "norm1.linear", (9 if dual else 6) * D, D, scale=0.5)
        lin(pre + "norm1_context.linear", (2 if last else 6) * D, D, scale=0.5)
        for a in (["attn", "attn2"] if dual else ["attn"]):
            for nm in ("to_q", "to_k", "to_v"):
                lin(pre + f"{a}.{nm}", D, D)
            w[pre + f"{a}.norm_q.weight"] = 1.0 + 0.1 * randn(cfg.attention_head_dim, generator=g)
            w[pre + f"{a}.norm_k.weight"] = 1.0 + 0.1 * randn(cfg.attention_head_dim, generator=g)
            lin(pre + f"{a}.to_out.0", D, D)
            if a == "attn":
                for nm in ("add_q_proj", "add_k_proj", "add_v_proj"):
                    lin(pre + f"attn.{nm}", D, D)
                w[pre + "attn.norm_added_q.weight"] = 1.0 + 0.1 * randn(cfg.attention_head_dim, generator=g)
                w[pre + "attn.norm_added_k.weight"] = 1.0 + 0.1 * randn(cfg.attention_head_dim, generator=g)
                if not last:
                    lin(pre + "attn.to_add_out", D, D)
        lin(pre + "ff.net.0.proj", 4 * D, D)
        lin(pre + "ff.net.2", D, 4 * D)
        if not last:
            lin(pre + "ff_context.net.0.proj", 4 * D, D)
            lin(pre + "ff_context.net.2", D, 4 * D)
    lin("norm_out.linear", 2 * D, D, scale=0.5)
    lin("proj_out", p * p * cfg.out_channels, D)
    return {k: v.to(dtype) for k, v in w.items()}


# ----------------------------------------------------------------------------------------------
# Transformer forward (DF/models/transformers/transformer_sd3.py:249-345)
# ----------------------------------------------------------------------------------------------
def timestep_embedding(timesteps: torch.Tensor, dim: int = 256) -> torch.Tensor:
    """DF/models/embeddings.py:26-77 with flip_sin_to_cos=True, downscale_freq_shift=0 (Timesteps, 1309-1325)."""
    half = dim // 2
    exponent = -math.log(10000) * torch.arange(0, half, dtype=torch.float32, device=timesteps.device)
    exponent = exponent / (half - 0.0)
    emb = torch.exp(exponent)
    emb = timesteps[:, None].float() * emb[None, :]
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    return torch.cat([emb[:, half:], emb[:, :half]], dim=-1)


def _linear(w, name, x):
    return F.linear(x, w[name + ".weight"], w[name + ".bias"])


def _rms_norm(x: torch.Tensor, weight: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """DF/models/normalization.py:553-567."""
    input_dtype = x.dtype
    variance = x.to(torch.float32).pow(2).mean(-1, keepdim=True)
    x = x * torch.rsqrt(variance + eps)
    if weight.dtype in (torch.float16, torch.bfloat16):
        x = x.to(weight.dtype)
    return x * weight


def _attention(w, pre: str, cfg: SD3Config, hs: torch.Tensor, ehs: Optional[torch.Tensor], ctx_pre_only: bool):
    """JointAttnProcessor2_0.__call__ (DF/models/attention_processor.py:1429-1505)."""
    B = hs.shape[0]
    H, d = cfg.num_attention_heads, cfg.attention_head_dim
    q = _linear(w, pre + "to_q", hs).view(B, -1, H, d).transpose(1, 2)
    k = _linear(w, pre + "to_k", hs).view(B, -1, H, d).transpose(1, 2)
    v = _linear(w, pre + "to_v", hs).view(B, -1, H, d).transpose(1, 2)
    q = _rms_norm(q, w[pre + "norm_q.weight"])
    k = _rms_norm(k, w[pre + "norm_k.weight"])
    if ehs is not None:
        eq = _linear(w, pre + "add_q_proj", ehs).view(B, -1, H, d).transpose(1, 2)
        ek = _linear(w, pre + "add_k_proj", ehs).view(B, -1, H, d).transpose(1, 2)
        ev = _linear(w, pre + "add_v_proj", ehs).view(B, -1, H, d).transpose(1, 2)
        eq = _rms_norm(eq, w[pre + "norm_added_q.weight"])
        ek = _rms_norm(ek, w[pre + "norm_added_k.weight"])
        q = torch.cat([q, eq], dim=2)  # image tokens first, then text (1480-1482)
        k = torch.cat([k, ek], dim=2)
        v = torch.cat([v, ev], dim=2)
    o = F.scaled_dot_product_attention(q, k, v, dropout_p=0.0, is_causal=False)
    o = o.transpose(1, 2).reshape(B, -1, H * d).to(q.dtype)
    eo = None
    if ehs is not None:
        o, eo = o[:, : hs.shape[1]], o[:, hs.shape[1]:]
        if not ctx_pre_only:
            eo = _linear(w, pre + "to_add_out", eo)
    o = _linear(w, pre + "to_out.0", o)
    return o, eo


def _layer_norm(x: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    return F.layer_norm(x, (x.shape[-1],), None, None, eps)


def _ff(w, pre: str, x: torch.Tensor) -> torch.Tensor:
    """FeedForward gelu-approximate (DF/models/attention.py:1682-1742; activations.py:87-90)."""
    x = _linear(w, pre + "net.0.proj", x)
    x = F.gelu(x, approximate="tanh")
    return _linear(w, pre + "net.2", x)


def _block(w, i: int, cfg: SD3Config, hs: torch.Tensor, ehs: torch.Tensor, temb: torch.Tensor):
    """JointTransformerBlock.forward (DF/models/attention.py:681-748)."""
    pre = f"transformer_blocks.{i}."
    last = i == cfg.num_layers - 1
    dual = i in cfg.dual_attention_layers
    emb = _linear(w, pre + "norm1.linear", F.silu(temb))
    if dual:  # SD35AdaLayerNormZeroX (normalization.py:115-127)
        (shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp,
         shift_msa2, scale_msa2, gate_msa2) = emb.chunk(9, dim=1)
        nh = _layer_norm(hs)
        norm_hs = nh * (1 + scale_msa[:, None]) + shift_msa[:, None]
        norm_hs2 = nh * (1 + scale_msa2[:, None]) + shift_msa2[:, None]
    else:     # AdaLayerNormZero (normalization.py:157-170)
        shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp = emb.chunk(6, dim=1)
        norm_hs = _layer_norm(hs) * (1 + scale_msa[:, None]) + shift_msa[:, None]
    if last:  # AdaLayerNormContinuous (normalization.py:346-351)
        cemb = _linear(w, pre + "norm1_context.linear", F.silu(temb).to(ehs.dtype))
        c_scale, c_shift = torch.chunk(cemb, 2, dim=1)
        norm_ehs = _layer_norm(ehs) * (1 + c_scale)[:, None, :] + c_shift[:, None, :]
    else:
        cemb = _linear(w, pre + "norm1_context.linear", F.silu(temb))
        c_shift_msa, c_scale_msa, c_gate_msa, c_shift_mlp, c_scale_mlp, c_gate_mlp = cemb.chunk(6, dim=1)
        norm_ehs = _layer_norm(ehs) * (1 + c_scale_msa[:, None]) + c_shift_msa[:, None]

    attn_out, ctx_attn_out = _attention(w, pre + "attn.", cfg, norm_hs, norm_ehs, last)
    hs = hs + gate_msa.unsqueeze(1) * attn_out
    if dual:
        attn_out2, _ = _attention(w, pre + "attn2.", cfg, norm_hs2, None, False)
        hs = hs + gate_msa2.unsqueeze(1) * attn_out2
    norm_hs = _layer_norm(hs)
    norm_hs = norm_hs * (1 + scale_mlp[:, None]) + shift_mlp[:, None]
    hs = hs + gate_mlp.unsqueeze(1) * _ff(w, pre + "ff.", norm_hs)
    if last:
        ehs = None
    else:
        ehs = ehs + c_gate_msa.unsqueeze(1) * ctx_attn_out
        norm_ehs = _layer_norm(ehs)
        norm_ehs = norm_ehs * (1 + c_scale_mlp[:, None]) + c_shift_mlp[:, None]
        ehs = ehs + c_gate_mlp.unsqueeze(1) * _ff(w, pre + "ff_context.", norm_ehs)
    return ehs, hs


def transformer_forward(w: Dict[str, torch.Tensor], cfg: SD3Config, hidden_states: torch.Tensor,
                        encoder_hidden_states: torch.Tensor, pooled_projections: torch.Tensor,
                        timestep: torch.Tensor, return_intermediates: bool = False, max_layers: Optional[int] = None):
    """SD3Transformer2DModel.forward (DF/models/transformers/transformer_sd3.py:288-345).
    `max_layers` (timing only): run just the first blocks - used by bench.py's bounded CPU sample."""
    inter = {}
    p = cfg.patch_size
    height, width = hidden_states.shape[-2:]
    # PatchEmbed.forward (DF/models/embeddings.py:554-583) with cropped_pos_embed (531-552)
    lat = F.conv2d(hidden_states, w["pos_embed.proj.weight"], w["pos_embed.proj.bias"], stride=p)
    lat = lat.flatten(2).transpose(1, 2)
    hp, wp = height // p, width // p
    top, left = (cfg.pos_embed_max_size - hp) // 2, (cfg.pos_embed_max_size - wp) // 2
    pe = w["pos_embed.pos_embed"].reshape(1, cfg.pos_embed_max_size, cfg.pos_embed_max_size, -1)
    pe = pe[:, top: top + hp, left: left + wp, :].reshape(1, -1, pe.shape[-1])
    hs = (lat + pe).to(lat.dtype)
    # CombinedTimestepTextProjEmbeddings (embeddings.py:1592-1600)
    tproj = timestep_embedding(timestep, 256)
    temb_t = _linear(w, "time_text_embed.timestep_embedder.linear_2",
                     F.silu(_linear(w, "time_text_embed.timestep_embedder.linear_1",
                                    tproj.to(dtype=pooled_projections.dtype))))
    temb_p = _linear(w, "time_text_embed.text_embedder.linear_2",
                     F.silu(_linear(w, "time_text_embed.text_embedder.linear_1", pooled_projections)))
    temb = temb_t + temb_p
    ehs = _linear(w, "context_embedder", encoder_hidden_states)
    if return_intermediates:
        inter["temb"], inter["hs0"], inter["ehs0"] = temb, hs, ehs
    for i in range(cfg.num_layers if max_layers is None else min(max_layers, cfg.num_layers)):
        ehs, hs = _block(w, i, cfg, hs, ehs, temb)
        if return_intermediates:
            inter[f"hs_{i}"] = hs
            if ehs is not None:
                inter[f"ehs_{i}"] = ehs
    # norm_out: AdaLayerNormContinuous (normalization.py:346-351)
    emb = _linear(w, "norm_out.linear", F.silu(temb).to(hs.dtype))
    scale, shift = torch.chunk(emb, 2, dim=1)
    hs = _layer_norm(hs) * (1 + scale)[:, None, :] + shift[:, None, :]
    hs = _linear(w, "proj_out", hs)
    hs = hs.reshape(hs.shape[0], hp, wp, p, p, cfg.out_channels)
    hs = torch.einsum("nhwpqc->nchpwq", hs)
    out = hs.reshape(hs.shape[0], cfg.out_channels, hp * p, wp * p)
    return (out, inter) if return_intermediates else out


# ----------------------------------------------------------------------------------------------
# Scheduler (FF/scheduler/flow_match_euler_discrete.py, DF/schedulers/scheduling_flow_match_euler_discrete.py)
# ----------------------------------------------------------------------------------------------
def make_schedule(num_inference_steps: int, shift: float = 3.0, num_train_timesteps: int = 1000):
    """set_scheduler_timesteps (FF flow_match...py:49-77) + diffusers set_timesteps (DF ...:282-384) for the
    static-shift case (use_dynamic_shifting=False): sigmas passed in = linspace(1, 1/T, T) -> shifted
    s*sigma/(1+(s-1)*sigma) (DF :350) -> timesteps = sigmas*1000 ; sigmas gets a trailing 0 (DF :377)."""
    sig = np.linspace(1.0, 1 / num_inference_steps, num_inference_steps).astype(np.float32)
    sig = shift * sig / (1 + (shift - 1) * sig)
    sigmas = torch.from_numpy(sig).to(torch.float32)
    timesteps = sigmas * num_train_timesteps
    sigmas = torch.cat([sigmas, torch.zeros(1)])
    return timesteps, sigmas


def current_sde_steps(num_steps: int, sde_steps: Optional[Sequence[int]], num_sde_steps: Optional[int], seed: int) -> List[int]:
    """FF flow_match...py:126-160: default window = all steps but the last; random subset by seed."""
    steps = torch.tensor(list(sde_steps), dtype=torch.int64) if sde_steps is not None else torch.arange(0, num_steps - 1)
    n = num_sde_steps if num_sde_steps is not None else len(steps)
    if n >= len(steps):
        return steps.tolist()
    g = torch.Generator().manual_seed(seed)
    sel = torch.randperm(len(steps), generator=g)[:n]
    return steps[sel].tolist()


def sde_step(noise_pred: torch.Tensor, latents: torch.Tensor, sigma: float, sigma_prev: float, noise_level: float,
             sigma_max: float, dynamics_type: str = "Flow-SDE", noise: Optional[torch.Tensor] = None,
             next_latents: Optional[torch.Tensor] = None, compute_log_prob: bool = True):
    """FlowMatchEulerDiscreteSDEScheduler.step, `timestep_next` given (FF flow_match...py:299-438).
    Scalars are fp32 tensors exactly as `to_broadcast_tensor` builds them.  Returns dict of fp32 tensors."""
    in_dtype = latents.dtype
    v = noise_pred.float()
    x = latents.float()
    nl = next_latents.float() if next_latents is not None else None
    f32 = lambda s: torch.tensor(s, dtype=torch.float32, device=x.device).view(*([1] * x.ndim))   # to_broadcast_tensor
    sigma_t, sigma_p, eta = f32(sigma), f32(sigma_prev), f32(noise_level)
    dt = sigma_p - sigma_t
    log_prob = None
    if dynamics_type == "ODE":
        mean = x + v * dt
        std_dev_t = torch.zeros_like(sigma_t)
        if nl is None:
            nl = mean
        if compute_log_prob:
            log_prob = torch.zeros(x.shape[0], dtype=torch.float32, device=x.device)
    elif dynamics_type == "Flow-SDE":
        smax = f32(sigma_max)
        std_dev_t = torch.sqrt(sigma_t / (1 - torch.where(sigma_t == 1.0, smax, sigma_t))) * eta
        mean = x * (1 + std_dev_t ** 2 / (2 * sigma_t) * dt) + v * (1 + std_dev_t ** 2 * (1 - sigma_t) / (2 * sigma_t)) * dt
        if nl is None:
            nl = mean + std_dev_t * torch.sqrt(-1 * dt) * noise.float()
            nl = nl.to(in_dtype).float()
        if compute_log_prob:
            sv = std_dev_t * torch.sqrt(-1 * dt)
            lp = (-((nl - mean) ** 2) / (2 * sv ** 2) - torch.log(sv)
                  - torch.log(torch.sqrt(2 * torch.as_tensor(math.pi))))
            log_prob = lp.mean(dim=tuple(range(1, lp.ndim)))
    elif dynamics_type == "Dance-SDE":
        x0 = x - sigma_t * v
        std_dev_t = eta
        log_term = 0.5 * eta ** 2 * (x - x0 * (1 - sigma_t)) / sigma_t ** 2
        mean = x + (v + log_term) * dt
        if nl is None:
            nl = mean + std_dev_t * torch.sqrt(-1 * dt) * noise.float()
            nl = nl.to(in_dtype).float()
        if compute_log_prob:
            sv = std_dev_t * torch.sqrt(-1 * dt)
            lp = (-((nl - mean) ** 2) / (2 * sv ** 2) - torch.log(sv)
                  - torch.log(torch.sqrt(2 * torch.as_tensor(math.pi))))
            log_prob = lp.mean(dim=tuple(range(1, lp.ndim)))
    elif dynamics_type == "CPS":
        std_dev_t = sigma_p * torch.sin(eta * torch.pi / 2)
        x0 = x - sigma_t * v
        x1 = x + v * (1 - sigma_t)
        mean = x0 * (1 - sigma_p) + x1 * torch.sqrt(sigma_p ** 2 - std_dev_t ** 2)
        if nl is None:
            nl = mean + std_dev_t * noise.float()
            nl = nl.to(in_dtype).float()
        if compute_log_prob:
            lp = -((nl - mean) ** 2)
            log_prob = lp.mean(dim=tuple(range(1, lp.ndim)))
    else:
        raise ValueError(dynamics_type)
    return dict(next_latents=nl, next_latents_mean=mean, std_dev_t=std_dev_t, dt=dt, log_prob=log_prob, noise_pred=v)


def cast_latents(latents: torch.Tensor, target: torch.dtype = torch.float16) -> torch.Tensor:
    """BaseAdapter.cast_latents (FF/models/abc.py:172-182)."""
    if latents.dtype == target:
        return latents
    if target == torch.float16 and latents.abs().max().item() > 65504.0:
        latents = latents.clamp(-65504.0, 65504.0)
    return latents.to(target)


# ----------------------------------------------------------------------------------------------
# Trajectory sampler: loop body of SD3_5Adapter.inference / forward (FF/models/stable_diffusion/sd3_5.py:266-304, 392-446)
# ----------------------------------------------------------------------------------------------
def rollout(w, cfg: SD3Config, x0: torch.Tensor, prompt_embeds: torch.Tensor, pooled: torch.Tensor,
            neg_prompt_embeds: Optional[torch.Tensor], neg_pooled: Optional[torch.Tensor],
            num_inference_steps: int, guidance_scale: float, noise_level: float = 0.7, shift: float = 3.0,
            sde_step_indices: Optional[Sequence[int]] = None, dynamics_type: str = "Flow-SDE",
            noises: Optional[Sequence[torch.Tensor]] = None, compute_log_prob: bool = True,
            storage_dtype: torch.dtype = torch.float16, autocast: Optional[str] = None,
            max_steps: Optional[int] = None):
    """Returns dict(latents=[T+1 tensors], log_probs={step: (B,)}, noise_preds=[...]).
    `noises[i]` is the fp32 N(0,1) tensor the reference would draw at step i (randn_tensor, FF :350-357);
    `autocast`: None (plain), 'cpu' or 'cuda' -> torch.autocast(autocast, bfloat16) around the transformer."""
    timesteps, sigmas = make_schedule(num_inference_steps, shift)
    if sde_step_indices is None:
        sde_step_indices = list(range(num_inference_steps - 1))
    sigma_max = float(sigmas[1])
    latents = cast_latents(x0, storage_dtype)
    do_cfg = guidance_scale > 1.0 and neg_prompt_embeds is not None and neg_pooled is not None
    out = dict(latents=[latents], log_probs={}, noise_preds=[], timesteps=timesteps, sigmas=sigmas)
    B = latents.shape[0]
    nsteps = num_inference_steps if max_steps is None else min(max_steps, num_inference_steps)
    for i in range(nsteps):
        t = timesteps[i]
        t_next = timesteps[i + 1] if i + 1 < num_inference_steps else torch.tensor(0.0)
        nl = noise_level if i in sde_step_indices else 0.0
        timestep = t.expand(B).to(device=latents.device, dtype=latents.dtype)   # sd3_5.py:394 (fp16-rounded t)
        if do_cfg:
            pe = torch.cat([neg_prompt_embeds, prompt_embeds], dim=0)  # sd3_5.py:409-413
            pp = torch.cat([neg_pooled, pooled], dim=0)
            li = torch.cat([latents, latents], dim=0)
            ti = timestep.repeat(2)
        else:
            pe, pp, li, ti = prompt_embeds, pooled, latents, timestep
        if autocast is not None:
            with torch.autocast(autocast, dtype=torch.bfloat16):
                v = transformer_forward(w, cfg, li, pe, pp, ti)
        else:
            v = transformer_forward(w, cfg, li.to(w["proj_out.weight"].dtype), pe, pp, ti)
        if do_cfg:
            vu, vc = v.chunk(2)
            v = vu + guidance_scale * (vc - vu)                        # sd3_5.py:431-433 (in v's dtype)
        clp = compute_log_prob and nl > 0
        r = sde_step(v, latents, float(t) / 1000 if False else (t / 1000).item(), (t_next / 1000).item(), nl,
                     sigma_max, dynamics_type, noise=None if noises is None else noises[i], compute_log_prob=clp)
        latents = cast_latents(r["next_latents"], storage_dtype)
        out["latents"].append(latents)
        out["noise_preds"].append(v)
        if clp:
            out["log_probs"][i] = r["log_prob"]
    return out


# FLOP model
def flops_per_forward(cfg: SD3Config, ni: int, nt: int) -> Tuple[float, float]:
    D, L, L2 = cfg.inner_dim, cfg.num_layers, len(cfg.dual_attention_layers)
    S = ni + nt
    linear = L * 24 * ni * D * D + (L * 6 + (L - 1) * 18) * nt * D * D + L2 * 8 * ni * D * D
    attn = L * 4 * S * S * D + L2 * 4 * ni * ni * D
    return float(linear), float(attn)


# ----------------------------------------------------------------------------------------------
# Seeded synthetic inputs shared by the golden generator, the tests and bench.py
# ----------------------------------------------------------------------------------------------
def make_inputs(cfg: SD3Config, batch: int, lat_h: int, lat_w: int, n_text: int, seed: int = 1):
    g = torch.Generator().manual_seed(seed)
    return dict(
        prompt_embeds=torch.randn(batch, n_text, cfg.joint_attention_dim, generator=g),
        pooled=torch.randn(batch, cfg.pooled_projection_dim, generator=g),
        neg_prompt_embeds=torch.randn(batch, n_text, cfg.joint_attention_dim, generator=g),
        neg_pooled=torch.randn(batch, cfg.pooled_projection_dim, generator=g),
        x0=torch.randn(batch, cfg.in_channels, lat_h, lat_w, generator=g),
    )


def make_noises(num_steps: int, shape, seed: int = 123) -> List[torch.Tensor]:
    """The fp32 N(0,1) draws the reference makes from the global CPU RNG after torch.manual_seed(seed):
    one randn_tensor(noise_pred.shape, fp32) per step (FF flow_match...py:350-357)."""
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(shape, generator=g, dtype=torch.float32) for _ in range(num_steps)]
