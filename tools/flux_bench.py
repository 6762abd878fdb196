#!/usr/bin/env python
"""FLUX.1-dev rollout bench (BASELINE config 3: FLUX.1-dev 1024^2 28-step GRPO rollout, prompt-sharded over the GPUs of a node).

  python tools/flux_bench.py --gpus 1 --steps 2 --warmup 1
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port P tools/flux_bench.py --gpus 8

Same contract as bench.py (one rank per GPU, device-timed with CUDA events, max over ranks, rank 0 prints one JSON line), for the
"next" row of SURVEY 8f: FLUX.1-dev architecture (19 dual + 38 single blocks, D = 3072, head_dim 128), random-init weights created
on the device (no checkpoints offline), 4096 image + 512 text tokens, embedded guidance 3.5 (no CFG batch), Flow-SDE with the
resolution-dependent shift.  A "step" is ONE ROLLOUT of `batch` prompts per rank through `B200Flux1Adapter.inference`.
`value`: inputs resident in HBM; `e2e`: pinned host prompt embeddings in, host results out, copies inside the timed region.
bench.py stays the driver's bench (config C2); this is the measurement tool of the FLUX.1 row."""
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
    """Random FluxTransformer2DModel.state_dict() (diffusers key names), bf16, drawn on the device."""
    g = torch.Generator(device=device).manual_seed(seed)
    D, d = cfg.inner_dim, 128
    sd = {}

    def lin(name, o, i, scale=1.0):
        sd[name + ".weight"] = torch.randn(o, i, generator=g, device=device, dtype=torch.bfloat16) * (scale / math.sqrt(i))
        sd[name + ".bias"] = torch.randn(o, generator=g, device=device, dtype=torch.bfloat16) * 0.02

    def rms(name):
        sd[name + ".weight"] = (1.0 + 0.1 * torch.randn(d, generator=g, device=device)).bfloat16()

    lin("x_embedder", D, 64); lin("context_embedder", D, cfg.joint_attention_dim)
    for n, i in (("timestep_embedder", 256), ("guidance_embedder", 256), ("text_embedder", cfg.pooled_projection_dim)):
        lin(f"time_text_embed.{n}.linear_1", D, i); lin(f"time_text_embed.{n}.linear_2", D, D)
    for i in range(cfg.num_layers):
        p = f"transformer_blocks.{i}."
        lin(p + "norm1.linear", 6 * D, D, 0.5); lin(p + "norm1_context.linear", 6 * D, D, 0.5)
        for nm in ("to_q", "to_k", "to_v", "add_q_proj", "add_k_proj", "add_v_proj", "to_out.0", "to_add_out"):
            lin(p + "attn." + nm, D, D)
        for nm in ("norm_q", "norm_k", "norm_added_q", "norm_added_k"):
            rms(p + "attn." + nm)
        for ff in ("ff", "ff_context"):
            lin(p + ff + ".net.0.proj", 4 * D, D); lin(p + ff + ".net.2", D, 4 * D)
    for i in range(cfg.num_single_layers):
        p = f"single_transformer_blocks.{i}."
        lin(p + "norm.linear", 3 * D, D, 0.5); lin(p + "proj_mlp", 4 * D, D); lin(p + "proj_out", D, 5 * D)
        for nm in ("to_q", "to_k", "to_v"):
            lin(p + "attn." + nm, D, D)
        rms(p + "attn.norm_q"); rms(p + "attn.norm_k")
    lin("norm_out.linear", 2 * D, D, 0.5); lin("proj_out", 64, D)
    return sd


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--batch", type=int, default=2, help="prompts per rank per rollout")
    ap.add_argument("--res", type=int, default=1024)
    ap.add_argument("--n-text", type=int, default=512)
    ap.add_argument("--num-inference-steps", type=int, default=28)
    ap.add_argument("--guidance", type=float, default=3.5)
    ap.add_argument("--num-sde-steps", type=int, default=1)
    ap.add_argument("--no-graph", action="store_true")
    a = ap.parse_args()
    import torch.distributed as dist
    from flow_factory_b200.dist import all_gather_rollout
    from flow_factory_b200.flux import FluxEngineConfig, FluxRolloutEngine
    from flow_factory_b200.flux_adapter import B200Flux1Adapter
    from flow_factory_b200.scheduler import FlowMatchEulerDiscreteSDEScheduler
    from flow_factory_b200.trajectory import compute_trajectory_indices

    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = FluxEngineConfig()
    T, B = a.num_inference_steps, a.batch
    sched = FlowMatchEulerDiscreteSDEScheduler(noise_level=0.7, use_dynamic_shifting=True, num_sde_steps=a.num_sde_steps, seed=42)
    sd = rand_state_dict(cfg, dev)
    adapter = B200Flux1Adapter(cfg, sd, device=dev, scheduler=sched, rng="philox", use_graph=not a.no_graph)
    adapter.rollout()
    del sd
    ni = (a.res // 16) ** 2
    sched.set_timesteps(T, seq_len=ni)
    traj_idx = compute_trajectory_indices(sched.train_timesteps, T)
    g = torch.Generator(device=dev).manual_seed(1 + rank)
    pe = torch.randn(B, a.n_text, cfg.joint_attention_dim, generator=g, device=dev).bfloat16()
    pooled = torch.randn(B, cfg.pooled_projection_dim, generator=g, device=dev).bfloat16()
    x0 = torch.randn(B, ni, 64, generator=g, device=dev).half()
    kw = dict(height=a.res, width=a.res, num_inference_steps=T, guidance_scale=a.guidance, compute_log_prob=True, trajectory_indices=traj_idx)

    def rollout_device():
        s = adapter.inference(prompt_embeds=pe, pooled_prompt_embeds=pooled, latents=x0, **kw)
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
    launches = FluxRolloutEngine.last_launch_count()
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

    host_pe, host_pooled = pe.cpu().pin_memory(), pooled.cpu().pin_memory()

    def rollout_e2e():
        ss = adapter.inference(prompt_embeds=host_pe.to(dev, non_blocking=True), pooled_prompt_embeds=host_pooled.to(dev, non_blocking=True), **kw)
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
        S, D = ni + a.n_text, cfg.inner_dim
        lin = cfg.num_layers * (2 * S * D * 12 * D) + cfg.num_single_layers * (2 * S * D * 7 * D + 2 * S * 5 * D * D)
        att = (cfg.num_layers + cfg.num_single_layers) * 4.0 * S * S * D
        fl_latent = (lin + att) * T
        lp0 = s[0].log_probs
        print(json.dumps({
            "metric": f"rollout latents/sec FLUX.1-dev {a.res}^2 {T}-step", "value": value, "unit": "latents/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": float(ms) / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "impl": "b200",
            "config": {"workload": f"FLUX.1-dev architecture {a.res}x{a.res} {T}-step GRPO rollout (Flow-SDE, noise 0.7, num_sde_steps {a.num_sde_steps}, dynamic shift), "
                                   f"embedded guidance {a.guidance}, {a.n_text} text tokens, random-init weights",
                       "per_rank_batch": B, "global_batch": B * world, "parallelism": f"dp{world} (prompt-sharded, 1 all-gather/rollout)",
                       "cuda_graph": not a.no_graph, "rng": "in-kernel Philox4x32-10"},
            "e2e": {"value": world * B * a.steps / float(e2e_s), "unit": "latents/s",
                    "h2d_bytes_per_step": host_pe.numel() * 2 + host_pooled.numel() * 2, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches * a.steps),
            "whole_step_achieved_tflops_per_gpu": value * fl_latent / 1e12 / world, "flops_per_latent": fl_latent,
            "finite": bool(torch.isfinite(s[0].all_latents.float()).all() and torch.isfinite(lp0).all()),
            "log_probs_sample0": lp0.flatten().tolist(), "weights_GB": adapter.engine.weights.nbytes() / 1e9}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
