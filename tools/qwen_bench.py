#!/usr/bin/env python
"""Qwen-Image (BASELINE config 5 model: 60 dual-stream blocks, D = 3072, head_dim 128, 20 B parameters) rollout on ONE B200:
random-init weights created on the device (40 GB bf16, replicated - no FSDP2 shard is needed at 180 GB), 1024^2 (4096 image tokens),
DGPO-style rollout = ODE, no log-prob, true CFG (2 forwards per step as one batch of 2B + per-token norm rescale).
Developer measurement for the 'next' row 4, not bench.py."""
import argparse, json, math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flow_factory_b200.flux import FluxEngineConfig
from flow_factory_b200.qwen import QwenRolloutEngine


def rand_state_dict(cfg, device, seed=0):
    g = torch.Generator(device=device).manual_seed(seed)
    D, J = cfg.inner_dim, cfg.joint_attention_dim
    sd = {}
    def lin(name, o, i, scale=1.0):
        sd[name + ".weight"] = torch.randn(o, i, generator=g, device=device, dtype=torch.bfloat16) * (scale / math.sqrt(i))
        sd[name + ".bias"] = torch.randn(o, generator=g, device=device, dtype=torch.bfloat16) * 0.02
    lin("img_in", D, 64); lin("txt_in", D, J)
    sd["txt_norm.weight"] = (1.0 + 0.1 * torch.randn(J, generator=g, device=device)).bfloat16()
    lin("time_text_embed.timestep_embedder.linear_1", D, 256); lin("time_text_embed.timestep_embedder.linear_2", D, D)
    for i in range(cfg.num_layers):
        p = f"transformer_blocks.{i}."
        lin(p + "img_mod.1", 6 * D, D, 0.5); lin(p + "txt_mod.1", 6 * D, D, 0.5)
        for nm in ("to_q", "to_k", "to_v", "add_q_proj", "add_k_proj", "add_v_proj", "to_out.0", "to_add_out"):
            lin(p + "attn." + nm, D, D)
        for nm in ("norm_q", "norm_k", "norm_added_q", "norm_added_k"):
            sd[p + f"attn.{nm}.weight"] = (1.0 + 0.1 * torch.randn(128, generator=g, device=device)).bfloat16()
        for mlp in ("img_mlp", "txt_mlp"):
            lin(p + mlp + ".net.0.proj", 4 * D, D); lin(p + mlp + ".net.2", D, 4 * D)
    lin("norm_out.linear", 2 * D, D, 0.5); lin("proj_out", 64, D)
    return sd


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1); ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--res", type=int, default=1024); ap.add_argument("--n-text", type=int, default=256)
    ap.add_argument("--guidance", type=float, default=4.0); ap.add_argument("--reps", type=int, default=1)
    a = ap.parse_args()
    dev = torch.device("cuda")
    cfg = FluxEngineConfig(num_layers=60, num_single_layers=0, num_heads=24, joint_attention_dim=3584, pooled_projection_dim=8,
                           guidance_embeds=False, variant=1)
    t0 = time.time()
    sd = rand_state_dict(cfg, dev)
    eng = QwenRolloutEngine(cfg, sd, dev)
    del sd
    h2 = w2 = a.res // 16
    cfg_on = a.guidance > 1.0
    plan = eng.plan(a.batch, h2, w2, a.n_text, cfg=cfg_on)
    g = torch.Generator(device=dev).manual_seed(1)
    pe = torch.randn(a.batch, a.n_text, 3584, generator=g, device=dev).bfloat16()
    npe = torch.randn(a.batch, a.n_text, 3584, generator=g, device=dev).bfloat16()
    x0 = torch.randn(a.batch, h2 * w2, 64, generator=g, device=dev).half()
    eng.set_prompts(plan, pe, npe if cfg_on else None, a.guidance)
    T = a.steps
    ts, sig, coefs = eng.make_coefs(plan, T, 0.0, [], dynamics="ODE", store_slots=[(0 if i == T - 1 else -1) for i in range(T)])
    setup_s = time.time() - t0
    r = eng.rollout(plan, x0, coefs, 1, -1, 0)       # warm-up (captures the graph)
    torch.cuda.synchronize()
    times = []
    for _ in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = eng.rollout(plan, x0, coefs, 1, -1, 0); e1.record(); torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = sorted(times)[len(times) // 2]
    S, D = h2 * w2 + a.n_text, cfg.inner_dim
    fwd = cfg.num_layers * (2 * S * D * 12 * D + 4.0 * S * S * D)
    fl = fwd * T * (2 if cfg_on else 1) * a.batch
    print(json.dumps({"model": "Qwen-Image 20B architecture (random init)", "res": a.res, "batch": a.batch, "steps": T, "true_cfg": cfg_on,
                      "n_text": a.n_text, "ms_per_rollout": ms, "latents_per_s": a.batch / (ms / 1e3), "tflops": fl / ms / 1e9,
                      "pflop_per_latent": fl / a.batch / 1e15, "finite": bool(torch.isfinite(r["final_latents"].float()).all()),
                      "launches": eng.last_launch_count(), "weights_GB": eng.weights.nbytes() / 1e9,
                      "workspace_GB": plan.workspace_bytes / 1e9, "setup_s": setup_s}))


if __name__ == "__main__":
    main()
