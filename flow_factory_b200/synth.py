"""Synthetic model + inputs for benchmarking without checkpoints: SD3.5-medium architecture constants, random-init
weights with SD3Transformer2DModel state-dict key names (generated directly on the device in the target dtype),
synthetic prompt embeddings, and the FLOP model of BASELINE.md section 2."""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Tuple

import torch


@dataclass
class SD3ModelConfig:
    """DF/models/transformers/transformer_sd3.py:117-141 (HF model-card values for SD3.5-medium as defaults)."""
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


def sd35_medium() -> SD3ModelConfig:
    return SD3ModelConfig()


def pos_embed_table(embed_dim: int, grid: int, base_size: int) -> torch.Tensor:
    """2-D sin-cos table the PatchEmbed buffer holds (DF/models/embeddings.py:264-384): [grid*grid, D] fp32,
    first half from the row coordinate... (w-major meshgrid as diffusers builds it)."""
    coords = torch.arange(grid, dtype=torch.float32) / (grid / base_size)
    gw, gh = torch.meshgrid(coords, coords, indexing="xy")

    def one_d(dim, pos):
        omega = 1.0 / 10000 ** (torch.arange(dim // 2, dtype=torch.float64) / (dim / 2.0))
        out = torch.outer(pos.reshape(-1).to(torch.float64), omega)
        return torch.cat([torch.sin(out), torch.cos(out)], dim=1)
    return torch.cat([one_d(embed_dim // 2, gw), one_d(embed_dim // 2, gh)], dim=1).float()


def random_weights(cfg: SD3ModelConfig, seed: int = 0, dtype: torch.dtype = torch.bfloat16, device: str = "cuda") -> Dict[str, torch.Tensor]:
    g = torch.Generator(device=device).manual_seed(seed)
    D, p = cfg.inner_dim, cfg.patch_size
    w: Dict[str, torch.Tensor] = {}

    def rn(*shape, std=1.0):
        return (torch.randn(*shape, generator=g, device=device) * std).to(dtype)

    def lin(name, out_f, in_f, scale=1.0):
        w[name + ".weight"] = rn(out_f, in_f, std=scale / math.sqrt(in_f))
        w[name + ".bias"] = rn(out_f, std=0.02)

    w["pos_embed.proj.weight"] = rn(D, cfg.in_channels, p, p, std=1 / math.sqrt(cfg.in_channels * p * p))
    w["pos_embed.proj.bias"] = rn(D, std=0.02)
    w["pos_embed.pos_embed"] = pos_embed_table(D, cfg.pos_embed_max_size, cfg.sample_size // p).unsqueeze(0).to(device)
    lin("time_text_embed.timestep_embedder.linear_1", D, 256)
    lin("time_text_embed.timestep_embedder.linear_2", D, D)
    lin("time_text_embed.text_embedder.linear_1", D, cfg.pooled_projection_dim)
    lin("time_text_embed.text_embedder.linear_2", D, D)
    lin("context_embedder", cfg.caption_projection_dim, cfg.joint_attention_dim)
    for i in range(cfg.num_layers):
        pre, last, dual = f"transformer_blocks.{i}.", i == cfg.num_layers - 1, i in cfg.dual_attention_layers
        lin(pre + "norm1.linear", (9 if dual else 6) * D, D, 0.5)
        lin(pre + "norm1_context.linear", (2 if last else 6) * D, D, 0.5)
        for a in (["attn", "attn2"] if dual else ["attn"]):
            for nm in ("to_q", "to_k", "to_v", "to_out.0"):
                lin(pre + f"{a}.{nm}", D, D)
            for nm in ("norm_q", "norm_k"):
                w[pre + f"{a}.{nm}.weight"] = (1.0 + 0.1 * torch.randn(64, generator=g, device=device)).to(dtype)
        for nm in ("add_q_proj", "add_k_proj", "add_v_proj"):
            lin(pre + f"attn.{nm}", D, D)
        for nm in ("norm_added_q", "norm_added_k"):
            w[pre + f"attn.{nm}.weight"] = (1.0 + 0.1 * torch.randn(64, generator=g, device=device)).to(dtype)
        if not last:
            lin(pre + "attn.to_add_out", D, D)
        lin(pre + "ff.net.0.proj", 4 * D, D)
        lin(pre + "ff.net.2", D, 4 * D)
        if not last:
            lin(pre + "ff_context.net.0.proj", 4 * D, D)
            lin(pre + "ff_context.net.2", D, 4 * D)
    lin("norm_out.linear", 2 * D, D, 0.5)
    lin("proj_out", p * p * cfg.out_channels, D)
    return w


def random_inputs(cfg: SD3ModelConfig, batch: int, lat_h: int, lat_w: int, n_text: int, seed: int, device) -> Dict[str, torch.Tensor]:
    g = torch.Generator(device=device).manual_seed(seed)
    rn = lambda *s: torch.randn(*s, generator=g, device=device)
    return dict(prompt_embeds=rn(batch, n_text, cfg.joint_attention_dim).bfloat16(), pooled=rn(batch, cfg.pooled_projection_dim).bfloat16(),
                neg_prompt_embeds=rn(batch, n_text, cfg.joint_attention_dim).bfloat16(), neg_pooled=rn(batch, cfg.pooled_projection_dim).bfloat16(),
                x0=rn(batch, cfg.in_channels, lat_h, lat_w).bfloat16())


def flops_per_forward(cfg: SD3ModelConfig, ni: int, nt: int) -> Tuple[float, float]:
    """(linear, attention) FLOPs of one transformer forward of one sample, MAC = 2 FLOP."""
    D, L, L2 = cfg.inner_dim, cfg.num_layers, len(cfg.dual_attention_layers)
    S = ni + nt
    linear = L * 24 * ni * D * D + (L * 6 + (L - 1) * 18) * nt * D * D + L2 * 8 * ni * D * D
    attn = L * 4 * S * S * D + L2 * 4 * ni * ni * D
    return float(linear), float(attn)
