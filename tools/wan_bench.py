#!/usr/bin/env python
"""Wan2.1-T2V-1.3B rollout bench (BASELINE config 4: 480 x 832, 81 frames, 50 steps, CFG 5, prompt-sharded over the GPUs of a node).

  python tools/wan_bench.py --gpus 1 --steps 1 --warmup 1
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port P tools/wan_bench.py --gpus 8

Same contract as bench.py / flux_bench.py (one rank per GPU, CUDA events, max over ranks, rank 0 prints one JSON line): Wan2.1 1.3 B
architecture (30 blocks, D = 1536, head_dim 128, ffn 8960), random-init weights created on the device, S = 32 760 video tokens + 512
text tokens, true CFG as a batch of 2, Flow-SDE on the UniPC flow-sigma schedule.  A "step" is ONE ROLLOUT of `batch` prompts per rank
through `B200Wan21Adapter.inference`.  First GPU run pending (see flow_factory_b200/wan.py)."""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch


def rand_state_dict(cfg, device, seed=0):
    """Random WanTransformer3DModel.state_dict() (diffusers key names), bf16, drawn on the device."""
    g = torch.Generator(device=device).manual_seed(seed)
    D = cfg.inner_dim
    sd = {}

    def lin(name, o, i, scale=1.0):
        sd[name + ".weight"] = torch.randn(o, i, generator=g, device=device, dtype=torch.bfloat16) * (scale / math.sqrt(i))
        sd[name + ".bias"] = torch.randn(o, generator=g, device=device, dtype=torch.bfloat16) * 0.02

    pt, ph, pw = cfg.patch_size
    sd["patch_embedding.weight"] = torch.randn(D, cfg.in_channels, pt, ph, pw, generator=g, device=device, dtype=torch.bfloat16) / math.sqrt(cfg.in_channels * pt * ph * pw)
    sd["patch_embedding.bias"] = torch.randn(D, generator=g, device=device, dtype=torch.bfloat16) * 0.02
    lin("condition_embedder.time_embedder.linear_1", D, cfg.freq_dim); lin("condition_embedder.time_embedder.linear_2", D, D)
    lin("condition_embedder.time_proj", 6 * D, D, 0.5)
    lin("condition_embedder.text_embedder.linear_1", D, cfg.text_dim); lin("condition_embedder.text_embedder.linear_2", D, D)
    for i in range(cfg.num_layers):
        p = f"blocks.{i}."
        sd[p + "scale_shift_table"] = (torch.randn(1, 6, D, generator=g, device=device) / D ** 0.5).bfloat16()
        for a in ("attn1", "attn2"):
            for nm in ("to_q", "to_k", "to_v", "to_out.0"):
                lin(p + f"{a}.{nm}", D, D)
            for nm in ("norm_q", "norm_k"):
                sd[p + f"{a}.{nm}.weight"] = (1.0 + 0.1 * torch.randn(D, generator=g, device=device)).bfloat16()
        sd[p + "norm2.weight"] = (1.0 + 0.1 * torch.randn(D, generator=g, device=device)).bfloat16()
        sd[p + "norm2.bias"] = (0.02 * torch.randn(D, generator=g, device=device)).bfloat16()
        lin(p + "ffn.net.0.proj", cfg.ffn_dim, D); lin(p + "ffn.net.2", D, cfg.ffn_dim)
    sd["scale_shift_table"] = (torch.randn(1, 2, D, generator=g, device=device) / D ** 0.5).bfloat16()
    lin("proj_out", cfg.out_channels * pt * ph * pw, D)
    return sd


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--batch", type=int, default=1, help="prompts per rank per rollout")
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--width", type=int, default=832)
    ap.add_argument("--frames", type=int, default=81)
    ap.add_argument("--n-text", type=int, default=512)
    ap.add_argument("--num-inference-steps", type=int, default=50)
    ap.add_argument("--guidance", type=float, default=5.0)
    ap.add_argument("--num-sde-steps", type=int, default=1)
    ap.add_argument("--no-graph", action="store_true")
    a = ap.parse_args()
    import torch.distributed as dist
    from flow_factory_b200.dist import all_gather_rollout
    from flow_factory_b200.scheduler import UniPCMultistepSDEScheduler
    from flow_factory_b200.trajectory import compute_trajectory_indices
    from flow_factory_b200.wan import WanEngineConfig, WanRolloutEngine
    from flow_factory_b200.wan_adapter import B200Wan21Adapter

    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = WanEngineConfig()
    T, B = a.num_inference_steps, a.batch
    sched = UniPCMultistepSDEScheduler(noise_level=0.7, flow_shift=3.0, num_sde_steps=a.num_sde_steps, seed=42)
    sd = rand_state_dict(cfg, dev)
    adapter = B200Wan21Adapter(cfg, sd, device=dev, scheduler=sched, rng="philox", use_graph=not a.no_graph)
    adapter.rollout()
    del sd
    sched.set_timesteps(T)
    traj_idx = compute_trajectory_indices(sched.train_timesteps, T)
    g = torch.Generator(device=dev).manual_seed(1 + rank)
    pe = torch.randn(B, a.n_text, cfg.text_dim, generator=g, device=dev).bfloat16()
    ne = torch.randn(B, a.n_text, cfg.text_dim, generator=g, device=dev).bfloat16()
    shape = adapter.latent_shape(B, a.height, a.width, a.frames)
    x0 = torch.randn(shape, generator=g, device=dev).half()
    kw = dict(height=a.height, width=a.width, num_frames=a.frames, num_inference_steps=T, guidance_scale=a.guidance, compute_log_prob=True,
              trajectory_indices=traj_idx)

    def rollout_device():
        s = adapter.inference(prompt_embeds=pe, negative_prompt_embeds=ne, latents=x0, **kw)
        if world > 1:
            all_gather_rollout(torch.stack([x.all_latents for x in s]), torch.stack([x.log_probs for x in s]))
        return s

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        s = rollout_device()
    barrier()
    launches = WanRolloutEngine.last_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        s = rollout_device()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    value = world * B * a.steps / (float(ms) / 1e3)
    host_pe, host_ne = pe.cpu().pin_memory(), ne.cpu().pin_memory()

    def rollout_e2e():
        ss = adapter.inference(prompt_embeds=host_pe.to(dev, non_blocking=True), negative_prompt_embeds=host_ne.to(dev, non_blocking=True), **kw)
        lat = torch.stack([x.all_latents for x in ss]); lp = torch.stack([x.log_probs for x in ss])
        fin = torch.stack([x.extra_kwargs["final_latents"] for x in ss])
        if world > 1:
            lat, lp = all_gather_rollout(lat, lp)
        out = (lat.cpu(), lp.cpu(), fin.cpu())
        return sum(t.numel() * t.element_size() for t in out)

    d2h = rollout_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        rollout_e2e()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    if rank == 0:
        D, S, L, ffn = cfg.inner_dim, shape[2] * (shape[3] // 2) * (shape[4] // 2), cfg.num_layers, cfg.ffn_dim
        lin = L * (2 * S * D * (6 * D + 2 * ffn))                      # q|k|v, out, q2, out2, ffn (the cached text k|v are not per step)
        att = L * (4.0 * S * S * D + 4.0 * S * a.n_text * D)
        fl_latent = (lin + att) * T * (2 if a.guidance > 1 else 1)
        lp0 = s[0].log_probs
        print(json.dumps({
            "metric": f"rollout latents/sec Wan2.1-T2V-1.3B {a.height}x{a.width}x{a.frames} {T}-step", "value": value, "unit": "latents/s",
            "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": float(ms) / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "impl": "b200",
            "config": {"workload": f"Wan2.1-T2V-1.3B architecture {a.height}x{a.width}, {a.frames} frames, {T}-step GRPO rollout (Flow-SDE, noise 0.7, "
                                   f"num_sde_steps {a.num_sde_steps}, UniPC flow-sigma schedule, flow_shift 3), true CFG {a.guidance}, {a.n_text} text tokens, "
                                   f"{S} video tokens, random-init weights",
                       "per_rank_batch": B, "global_batch": B * world, "parallelism": f"dp{world} (prompt-sharded, 1 all-gather/rollout)",
                       "cuda_graph": not a.no_graph, "rng": "in-kernel Philox4x32-10"},
            "e2e": {"value": world * B * a.steps / float(e2e_s), "unit": "latents/s",
                    "h2d_bytes_per_step": (host_pe.numel() + host_ne.numel()) * 2, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches * a.steps),
            "whole_step_achieved_tflops_per_gpu": value * fl_latent / 1e12 / world, "flops_per_latent": fl_latent,
            "finite": bool(torch.isfinite(s[0].all_latents.float()).all() and torch.isfinite(lp0).all()),
            "log_probs_sample0": lp0.flatten().tolist(), "weights_GB": adapter.engine.weights.nbytes() / 1e9}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
