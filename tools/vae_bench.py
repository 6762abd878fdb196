#!/usr/bin/env python
"""VAE decode timing: SD3.5 VAE architecture, random weights, `--batch` latents of `--res`^2 images per call.

  python tools/vae_bench.py --res 1024 --batch 4 --steps 5

Device-timed with CUDA events after warm-up; prints one JSON line (ms per image, TFLOP/s from the conv + attention FLOP model)."""
from __future__ import annotations

import argparse
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch


def rand_decoder_state_dict(cfg, device, seed=0):
    """Random `decoder.*` entries keyed like AutoencoderKL.state_dict()."""
    g = torch.Generator(device=device).manual_seed(seed)
    sd = {}

    def conv(n, o, i, k):
        sd[n + ".weight"] = (torch.randn(o, i, k, k, generator=g, device=device) / math.sqrt(i * k * k)).bfloat16()
        sd[n + ".bias"] = (torch.randn(o, generator=g, device=device) * 0.02).bfloat16()

    def lin(n, o, i):
        sd[n + ".weight"] = (torch.randn(o, i, generator=g, device=device) / math.sqrt(i)).bfloat16()
        sd[n + ".bias"] = (torch.randn(o, generator=g, device=device) * 0.02).bfloat16()

    def norm(n, c):
        sd[n + ".weight"] = (1 + 0.1 * torch.randn(c, generator=g, device=device)).bfloat16()
        sd[n + ".bias"] = (0.05 * torch.randn(c, generator=g, device=device)).bfloat16()

    def resnet(p, ci, co):
        norm(p + "norm1", ci); conv(p + "conv1", co, ci, 3); norm(p + "norm2", co); conv(p + "conv2", co, co, 3)
        if ci != co:
            conv(p + "conv_shortcut", co, ci, 1)

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
    return sd


def decode_flops(cfg, h, w):
    rev = list(reversed(cfg.block_out_channels))
    conv = lambda ci, co, k, hh, ww: 2.0 * hh * ww * ci * co * k * k
    top = rev[0]
    fl = conv(cfg.latent_channels, top, 3, h, w) + 4 * conv(top, top, 3, h, w) + 4 * 2.0 * h * w * top * top + 4.0 * (h * w) ** 2 * top
    prev = top
    for i, ch in enumerate(rev):
        for j in range(cfg.layers_per_block + 1):
            ci = prev if j == 0 else ch
            fl += conv(ci, ch, 3, h, w) + conv(ch, ch, 3, h, w) + (conv(ci, ch, 1, h, w) if ci != ch else 0.0)
        if i != len(rev) - 1:
            h, w = 2 * h, 2 * w
            fl += conv(ch, ch, 3, h, w)
        prev = ch
    return fl + conv(rev[-1], cfg.out_channels, 3, h, w)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--res", type=int, default=1024)
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    a = ap.parse_args()
    from flow_factory_b200.vae import B200VaeDecoder, VaeDecoderConfig
    dev = torch.device("cuda", 0)
    cfg = VaeDecoderConfig()
    h = a.res // 8
    dec = B200VaeDecoder(cfg, rand_decoder_state_dict(cfg, dev), h, h, batch=a.batch, device=dev)
    lat = torch.randn(a.batch, cfg.latent_channels, h, h, device=dev).half()
    for _ in range(a.warmup):
        img = dec.decode(lat)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        img = dec.decode(lat)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    fl = decode_flops(cfg, h, h)
    print(json.dumps({"metric": f"VAE decode {a.res}^2 (SD3.5 VAE architecture, random weights)", "ms_per_image": ms / a.batch, "batch": a.batch,
                      "images_per_s": a.batch / (ms / 1e3), "tflops": fl * a.batch / (ms / 1e3) / 1e12, "flops_per_image": fl,
                      "gpu_launches_per_call": B200VaeDecoder.last_launch_count(), "workspace_GB": dec.workspace_bytes() / 1e9,
                      "finite": bool(torch.isfinite(img.float()).all()), "image_abs_mean": float(img.float().abs().mean())}), flush=True)


if __name__ == "__main__":
    main()
