"""CPU/torch restatement of the Wan2.1 text-to-video transformer (TEST INFRASTRUCTURE - see oracle/__init__.py).

SURVEY.md section 8(f) "next" row 4 / BASELINE config 4.  Groundwork only: there is no native engine for this model yet; the oracle
is pinned against the imported reference (tests/golden/make_golden.py -> tests/golden/wan_tiny.pt) so that the engine of a later
round starts from a checked specification.  Paths relative to /root/reference: DF = diffusers/src/diffusers.
Covers WanTransformer3DModel.forward for T2V (no image conditioning, scalar timestep per sample):
DF/models/transformers/transformer_wan.py:629-740 ; WanTransformerBlock 462-505 ; WanAttnProcessor 78-162 ; WanRotaryPosEmbed 354-417 ;
WanTimeTextImageEmbedding 330-351.  The rollout's scheduler math is the same Euler / SDE step as the other models
(FF/scheduler/unipc_multistep.py:290-421) and is covered by sd3_oracle.sde_step.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Tuple

import torch
import torch.nn.functional as F

from .sd3_oracle import _linear, timestep_embedding


@dataclass
class WanConfig:
    patch_size: Tuple[int, int, int] = (1, 2, 2)
    num_attention_heads: int = 12          # Wan2.1-T2V-1.3B (HF model card; the in-tree defaults are the 14 B model)
    attention_head_dim: int = 128
    in_channels: int = 16
    out_channels: int = 16
    text_dim: int = 4096
    freq_dim: int = 256
    ffn_dim: int = 8960
    num_layers: int = 30
    cross_attn_norm: bool = True
    eps: float = 1e-6
    rope_max_seq_len: int = 1024

    @property
    def inner_dim(self) -> int:
        return self.num_attention_heads * self.attention_head_dim

    def ref_kwargs(self) -> dict:
        return dict(patch_size=tuple(self.patch_size), num_attention_heads=self.num_attention_heads,
                    attention_head_dim=self.attention_head_dim, in_channels=self.in_channels, out_channels=self.out_channels,
                    text_dim=self.text_dim, freq_dim=self.freq_dim, ffn_dim=self.ffn_dim, num_layers=self.num_layers,
                    cross_attn_norm=self.cross_attn_norm, qk_norm="rms_norm_across_heads", eps=self.eps,
                    rope_max_seq_len=self.rope_max_seq_len)


def wan21_t2v_1_3b() -> WanConfig:
    return WanConfig()


def tiny_wan_config(num_layers: int = 2, heads: int = 2, text_dim: int = 64, ffn_dim: int = 320) -> WanConfig:
    return WanConfig(num_attention_heads=heads, text_dim=text_dim, ffn_dim=ffn_dim, num_layers=num_layers, rope_max_seq_len=64)


def make_wan_weights(cfg: WanConfig, seed: int = 0, dtype: torch.dtype = torch.float32) -> Dict[str, torch.Tensor]:
    """Random weights keyed like WanTransformer3DModel.state_dict() (T2V: no image embedder / added kv projections)."""
    g = torch.Generator().manual_seed(seed)
    D = cfg.inner_dim
    w: Dict[str, torch.Tensor] = {}

    def lin(name: str, out_f: int, in_f: int, scale: float = 1.0):
        w[name + ".weight"] = torch.randn(out_f, in_f, generator=g) * (scale / math.sqrt(in_f))
        w[name + ".bias"] = torch.randn(out_f, generator=g) * 0.02

    pt, ph, pw = cfg.patch_size
    w["patch_embedding.weight"] = torch.randn(D, cfg.in_channels, pt, ph, pw, generator=g) / math.sqrt(cfg.in_channels * pt * ph * pw)
    w["patch_embedding.bias"] = torch.randn(D, generator=g) * 0.02
    lin("condition_embedder.time_embedder.linear_1", D, cfg.freq_dim)
    lin("condition_embedder.time_embedder.linear_2", D, D)
    lin("condition_embedder.time_proj", 6 * D, D, 0.5)
    lin("condition_embedder.text_embedder.linear_1", D, cfg.text_dim)
    lin("condition_embedder.text_embedder.linear_2", D, D)
    for i in range(cfg.num_layers):
        pre = f"blocks.{i}."
        w[pre + "scale_shift_table"] = torch.randn(1, 6, D, generator=g) / D ** 0.5
        for a in ("attn1", "attn2"):
            for nm in ("to_q", "to_k", "to_v", "to_out.0"):
                lin(pre + f"{a}.{nm}", D, D)
            w[pre + f"{a}.norm_q.weight"] = 1.0 + 0.1 * torch.randn(D, generator=g)
            w[pre + f"{a}.norm_k.weight"] = 1.0 + 0.1 * torch.randn(D, generator=g)
        if cfg.cross_attn_norm:
            w[pre + "norm2.weight"] = 1.0 + 0.1 * torch.randn(D, generator=g)
            w[pre + "norm2.bias"] = 0.02 * torch.randn(D, generator=g)
        lin(pre + "ffn.net.0.proj", cfg.ffn_dim, D)
        lin(pre + "ffn.net.2", D, cfg.ffn_dim)
    w["scale_shift_table"] = torch.randn(1, 2, D, generator=g) / D ** 0.5
    lin("proj_out", cfg.out_channels * pt * ph * pw, D)
    return {k: v.to(dtype) for k, v in w.items()}


def wan_rope(cfg: WanConfig, ppf: int, pph: int, ppw: int, theta: float = 10000.0):
    """WanRotaryPosEmbed (354-417): (cos, sin) fp32 [1, ppf*pph*ppw, 1, head_dim], frequencies in float64, values repeated pairwise."""
    d = cfg.attention_head_dim
    h_dim = w_dim = 2 * (d // 6)
    t_dim = d - h_dim - w_dim
    cos_l, sin_l = [], []
    for dim in (t_dim, h_dim, w_dim):
        freqs = 1.0 / (theta ** (torch.arange(0, dim, 2, dtype=torch.float64) / dim))
        freqs = torch.outer(torch.arange(cfg.rope_max_seq_len), freqs)
        cos_l.append(freqs.cos().repeat_interleave(2, dim=1).float())
        sin_l.append(freqs.sin().repeat_interleave(2, dim=1).float())
    expand = lambda t, n, shape: t[:n].view(*shape, -1).expand(ppf, pph, ppw, -1)
    cos = torch.cat([expand(cos_l[0], ppf, (ppf, 1, 1)), expand(cos_l[1], pph, (1, pph, 1)), expand(cos_l[2], ppw, (1, 1, ppw))], dim=-1)
    sin = torch.cat([expand(sin_l[0], ppf, (ppf, 1, 1)), expand(sin_l[1], pph, (1, pph, 1)), expand(sin_l[2], ppw, (1, 1, ppw))], dim=-1)
    return cos.reshape(1, ppf * pph * ppw, 1, -1), sin.reshape(1, ppf * pph * ppw, 1, -1)


def _apply_rope(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """WanAttnProcessor's inner apply_rotary_emb (103-114): computed in the tensor's own dtype."""
    x1, x2 = x.unflatten(-1, (-1, 2)).unbind(-1)
    c, s = cos[..., 0::2], sin[..., 1::2]
    out = torch.empty_like(x)
    out[..., 0::2] = x1 * c - x2 * s
    out[..., 1::2] = x1 * s + x2 * c
    return out.type_as(x)


def _wan_attention(w, pre: str, cfg: WanConfig, hs, ehs, rope):
    """WanAttnProcessor.__call__ (78-162), T2V: rms_norm ACROSS heads (torch.nn.RMSNorm over the inner dim) before the head split."""
    H = cfg.num_attention_heads
    kv = hs if ehs is None else ehs
    q = _linear(w, pre + "to_q", hs)
    k = _linear(w, pre + "to_k", kv)
    v = _linear(w, pre + "to_v", kv)
    q = F.rms_norm(q, (q.shape[-1],), w[pre + "norm_q.weight"], cfg.eps)
    k = F.rms_norm(k, (k.shape[-1],), w[pre + "norm_k.weight"], cfg.eps)
    q, k, v = q.unflatten(2, (H, -1)), k.unflatten(2, (H, -1)), v.unflatten(2, (H, -1))
    if rope is not None:
        q, k = _apply_rope(q, *rope), _apply_rope(k, *rope)
    o = F.scaled_dot_product_attention(q.permute(0, 2, 1, 3), k.permute(0, 2, 1, 3), v.permute(0, 2, 1, 3), dropout_p=0.0,
                                       is_causal=False).permute(0, 2, 1, 3)
    o = o.flatten(2, 3).type_as(q)
    return _linear(w, pre + "to_out.0", o)


def _fp32_ln(x, weight=None, bias=None, eps=1e-6):
    """FP32LayerNorm (DF/models/normalization.py:84-93)."""
    return F.layer_norm(x.float(), (x.shape[-1],), None if weight is None else weight.float(), None if bias is None else bias.float(),
                        eps).to(x.dtype)


def _wan_block(w, i: int, cfg: WanConfig, hs, ehs, temb6, rope):
    pre = f"blocks.{i}."
    shift_msa, scale_msa, gate_msa, c_shift, c_scale, c_gate = (w[pre + "scale_shift_table"] + temb6.float()).chunk(6, dim=1)
    n = (_fp32_ln(hs.float(), eps=cfg.eps) * (1 + scale_msa) + shift_msa).type_as(hs)
    a = _wan_attention(w, pre + "attn1.", cfg, n, None, rope)
    hs = (hs.float() + a * gate_msa).type_as(hs)
    if cfg.cross_attn_norm:
        n = _fp32_ln(hs.float(), w[pre + "norm2.weight"], w[pre + "norm2.bias"], cfg.eps).type_as(hs)
    else:
        n = hs
    hs = hs + _wan_attention(w, pre + "attn2.", cfg, n, ehs, None)
    n = (_fp32_ln(hs.float(), eps=cfg.eps) * (1 + c_scale) + c_shift).type_as(hs)
    ff = _linear(w, pre + "ffn.net.2", F.gelu(_linear(w, pre + "ffn.net.0.proj", n), approximate="tanh"))
    return (hs.float() + ff.float() * c_gate).type_as(hs)


def wan_forward(w: Dict[str, torch.Tensor], cfg: WanConfig, hidden_states: torch.Tensor, timestep: torch.Tensor,
                encoder_hidden_states: torch.Tensor) -> torch.Tensor:
    """WanTransformer3DModel.forward (629-740), T2V.  hidden_states [B, C, F, H, W]; timestep [B] on the 0..1000 scale."""
    B, C, Fr, Hh, Ww = hidden_states.shape
    pt, ph, pw = cfg.patch_size
    ppf, pph, ppw = Fr // pt, Hh // ph, Ww // pw
    # freqs_cos / freqs_sin are registered (non-persistent) BUFFERS of WanRotaryPosEmbed (389-390): `model.to(bf16)` /
    # `from_pretrained(torch_dtype=bf16)` casts them with the weights, so a bf16 model rotates with bf16-rounded tables
    pw_ = w["patch_embedding.weight"]
    rope = tuple(t.to(device=pw_.device, dtype=pw_.dtype) for t in wan_rope(cfg, ppf, pph, ppw))    # buffers move with the module
    hs = F.conv3d(hidden_states, w["patch_embedding.weight"], w["patch_embedding.bias"], stride=cfg.patch_size)
    hs = hs.flatten(2).transpose(1, 2)
    # WanTimeTextImageEmbedding (330-351)
    tproj = timestep_embedding(timestep, cfg.freq_dim)
    tdt = w["condition_embedder.time_embedder.linear_1.weight"].dtype
    if tproj.dtype != tdt:
        tproj = tproj.to(tdt)
    temb = _linear(w, "condition_embedder.time_embedder.linear_2",
                   F.silu(_linear(w, "condition_embedder.time_embedder.linear_1", tproj))).type_as(encoder_hidden_states)
    temb6 = _linear(w, "condition_embedder.time_proj", F.silu(temb)).unflatten(1, (6, -1))
    ehs = _linear(w, "condition_embedder.text_embedder.linear_2",
                  F.gelu(_linear(w, "condition_embedder.text_embedder.linear_1", encoder_hidden_states), approximate="tanh"))
    for i in range(cfg.num_layers):
        hs = _wan_block(w, i, cfg, hs, ehs, temb6, rope)
    shift, scale = (w["scale_shift_table"] + temb.unsqueeze(1)).chunk(2, dim=1)
    hs = (_fp32_ln(hs.float(), eps=cfg.eps) * (1 + scale) + shift).type_as(hs)
    hs = _linear(w, "proj_out", hs)
    hs = hs.reshape(B, ppf, pph, ppw, pt, ph, pw, -1).permute(0, 7, 1, 4, 2, 5, 3, 6)
    return hs.flatten(6, 7).flatten(4, 5).flatten(2, 3)


def make_wan_inputs(cfg: WanConfig, batch: int, frames: int, lat_h: int, lat_w: int, n_text: int, seed: int = 1):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(batch, cfg.in_channels, frames, lat_h, lat_w, generator=g),
            torch.randn(batch, n_text, cfg.text_dim, generator=g))
