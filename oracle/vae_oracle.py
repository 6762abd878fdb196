"""CPU/torch restatement of the VAE decode that follows the rollout (TEST INFRASTRUCTURE - see oracle/__init__.py).

SURVEY.md section 8(f) "next" row 3: SD3_5Adapter.decode_latents (FF/models/stable_diffusion/sd3_5.py:161-172) ->
AutoencoderKL.decode -> Decoder.forward (DF/models/autoencoders/vae.py:279-316): conv_in, UNetMidBlock2D (ResnetBlock2D, single-head
attention with GroupNorm, ResnetBlock2D), UpDecoderBlock2D x4 (3 ResnetBlock2D + nearest 2x upsample + conv), GroupNorm, SiLU, conv_out.
Groundwork only: there is no native decoder yet; pinned against the imported reference (tests/golden/vae_tiny.pt).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Tuple

import torch
import torch.nn.functional as F


@dataclass
class VaeConfig:
    latent_channels: int = 16
    out_channels: int = 3
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_num_groups: int = 32
    scaling_factor: float = 1.5305
    shift_factor: float = 0.0609

    def ref_kwargs(self) -> dict:
        n = len(self.block_out_channels)
        return dict(in_channels=3, out_channels=self.out_channels, down_block_types=("DownEncoderBlock2D",) * n,
                    up_block_types=("UpDecoderBlock2D",) * n, block_out_channels=tuple(self.block_out_channels),
                    layers_per_block=self.layers_per_block, latent_channels=self.latent_channels, norm_num_groups=self.norm_num_groups,
                    use_quant_conv=False, use_post_quant_conv=False, scaling_factor=self.scaling_factor, shift_factor=self.shift_factor)


def sd35_vae() -> VaeConfig:
    """SD3 / SD3.5 VAE (HF model card values: 16 latent channels, no quant convs)."""
    return VaeConfig()


def tiny_vae_config() -> VaeConfig:
    return VaeConfig(latent_channels=4, block_out_channels=(32, 64), layers_per_block=1, norm_num_groups=8)


def make_vae_decoder_weights(cfg: VaeConfig, seed: int = 0, dtype: torch.dtype = torch.float32) -> Dict[str, torch.Tensor]:
    """Random decoder weights keyed like AutoencoderKL.state_dict() (`decoder.*` only)."""
    g = torch.Generator().manual_seed(seed)
    w: Dict[str, torch.Tensor] = {}

    def conv(name, o, i, k):
        w[name + ".weight"] = torch.randn(o, i, k, k, generator=g) / math.sqrt(i * k * k)
        w[name + ".bias"] = torch.randn(o, generator=g) * 0.02

    def lin(name, o, i):
        w[name + ".weight"] = torch.randn(o, i, generator=g) / math.sqrt(i)
        w[name + ".bias"] = torch.randn(o, generator=g) * 0.02

    def norm(name, c):
        w[name + ".weight"] = 1.0 + 0.1 * torch.randn(c, generator=g)
        w[name + ".bias"] = 0.05 * torch.randn(c, generator=g)

    def resnet(pre, cin, cout):
        norm(pre + "norm1", cin); conv(pre + "conv1", cout, cin, 3)
        norm(pre + "norm2", cout); conv(pre + "conv2", cout, cout, 3)
        if cin != cout:
            conv(pre + "conv_shortcut", cout, cin, 1)

    rev = list(reversed(cfg.block_out_channels))
    top = rev[0]
    conv("decoder.conv_in", top, cfg.latent_channels, 3)
    resnet("decoder.mid_block.resnets.0.", top, top)
    norm("decoder.mid_block.attentions.0.group_norm", top)
    for nm in ("to_q", "to_k", "to_v", "to_out.0"):
        lin("decoder.mid_block.attentions.0." + nm, top, top)
    resnet("decoder.mid_block.resnets.1.", top, top)
    prev = top
    for i, ch in enumerate(rev):
        for j in range(cfg.layers_per_block + 1):
            resnet(f"decoder.up_blocks.{i}.resnets.{j}.", prev if j == 0 else ch, ch)
        if i != len(rev) - 1:
            conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", ch, ch, 3)
        prev = ch
    norm("decoder.conv_norm_out", rev[-1])
    conv("decoder.conv_out", cfg.out_channels, rev[-1], 3)
    return {k: v.to(dtype) for k, v in w.items()}


def _resnet(w, pre: str, x: torch.Tensor, groups: int) -> torch.Tensor:
    """ResnetBlock2D.forward without time embedding (DF/models/resnet.py:319-377), eps 1e-6, output_scale_factor 1."""
    h = F.group_norm(x, groups, w[pre + "norm1.weight"], w[pre + "norm1.bias"], 1e-6)
    h = F.conv2d(F.silu(h), w[pre + "conv1.weight"], w[pre + "conv1.bias"], padding=1)
    h = F.group_norm(h, groups, w[pre + "norm2.weight"], w[pre + "norm2.bias"], 1e-6)
    h = F.conv2d(F.silu(h), w[pre + "conv2.weight"], w[pre + "conv2.bias"], padding=1)
    if pre + "conv_shortcut.weight" in w:
        x = F.conv2d(x, w[pre + "conv_shortcut.weight"], w[pre + "conv_shortcut.bias"])
    return (x + h) / 1.0


def _mid_attention(w, pre: str, x: torch.Tensor, groups: int) -> torch.Tensor:
    """Attention(heads=1, residual_connection=True, group norm) through AttnProcessor2_0 (DF/models/attention_processor.py)."""
    B, C, H, W = x.shape
    residual = x
    h = x.view(B, C, H * W).transpose(1, 2)
    h = F.group_norm(h.transpose(1, 2), groups, w[pre + "group_norm.weight"], w[pre + "group_norm.bias"], 1e-6).transpose(1, 2)
    q = F.linear(h, w[pre + "to_q.weight"], w[pre + "to_q.bias"]).view(B, -1, 1, C).transpose(1, 2)
    k = F.linear(h, w[pre + "to_k.weight"], w[pre + "to_k.bias"]).view(B, -1, 1, C).transpose(1, 2)
    v = F.linear(h, w[pre + "to_v.weight"], w[pre + "to_v.bias"]).view(B, -1, 1, C).transpose(1, 2)
    o = F.scaled_dot_product_attention(q, k, v, dropout_p=0.0, is_causal=False)
    o = o.transpose(1, 2).reshape(B, -1, C).to(q.dtype)
    o = F.linear(o, w[pre + "to_out.0.weight"], w[pre + "to_out.0.bias"])
    o = o.transpose(-1, -2).reshape(B, C, H, W)
    return (o + residual) / 1.0


def vae_decode(w: Dict[str, torch.Tensor], cfg: VaeConfig, latents: torch.Tensor) -> torch.Tensor:
    """decode_latents (sd3_5.py:166-169) + AutoencoderKL.decode: latents [B, C, h, w] -> image [B, 3, 8h, 8w] (before postprocess)."""
    g = cfg.norm_num_groups
    z = (latents / cfg.scaling_factor) + cfg.shift_factor
    x = F.conv2d(z, w["decoder.conv_in.weight"], w["decoder.conv_in.bias"], padding=1)
    x = _resnet(w, "decoder.mid_block.resnets.0.", x, g)
    x = _mid_attention(w, "decoder.mid_block.attentions.0.", x, g)
    x = _resnet(w, "decoder.mid_block.resnets.1.", x, g)
    n = len(cfg.block_out_channels)
    for i in range(n):
        for j in range(cfg.layers_per_block + 1):
            x = _resnet(w, f"decoder.up_blocks.{i}.resnets.{j}.", x, g)
        if i != n - 1:   # Upsample2D: nearest 2x then conv (DF/models/upsampling.py)
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
            x = F.conv2d(x, w[f"decoder.up_blocks.{i}.upsamplers.0.conv.weight"], w[f"decoder.up_blocks.{i}.upsamplers.0.conv.bias"], padding=1)
    x = F.group_norm(x, g, w["decoder.conv_norm_out.weight"], w["decoder.conv_norm_out.bias"], 1e-6)
    return F.conv2d(F.silu(x), w["decoder.conv_out.weight"], w["decoder.conv_out.bias"], padding=1)


def vae_decode_flops(cfg: VaeConfig, lat_h: int, lat_w: int) -> float:
    """Convolution + attention FLOPs of one decode (MAC = 2 FLOP)."""
    rev = list(reversed(cfg.block_out_channels))
    h, w_, fl = lat_h, lat_w, 0.0
    conv = lambda cin, cout, k, hh, ww: 2.0 * hh * ww * cin * cout * k * k
    top = rev[0]
    fl += conv(cfg.latent_channels, top, 3, h, w_) + 4 * conv(top, top, 3, h, w_) + 4 * 2.0 * h * w_ * top * top + 4.0 * (h * w_) ** 2 * top
    prev = top
    for i, ch in enumerate(rev):
        for j in range(cfg.layers_per_block + 1):
            cin = prev if j == 0 else ch
            fl += conv(cin, ch, 3, h, w_) + conv(ch, ch, 3, h, w_) + (conv(cin, ch, 1, h, w_) if cin != ch else 0.0)
        if i != len(rev) - 1:
            h, w_ = 2 * h, 2 * w_
            fl += conv(ch, ch, 3, h, w_)
        prev = ch
    return fl + conv(rev[-1], cfg.out_channels, 3, h, w_)
