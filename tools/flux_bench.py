"""FLUX.1-dev (BASELINE config 3 model, 19 dual + 38 single blocks, D = 3072) rollout on ONE B200: random-init weights created
on the device, 1024^2 (4096 image + 512 text tokens), T denoise steps.  Developer measurement for the 'next' row, not bench.py."""
import argparse, json, math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flow_factory_b200.flux import FluxEngineConfig, FluxRolloutEngine


def rand_state_dict(cfg, device, seed=0):
    g = torch.Generator(device=device).manual_seed(seed)
    D, d = cfg.inner_dim, 128
    sd = {}
    def lin(name, o, i, scale=1.0):
        sd[name + ".weight"] = (torch.randn(o, i, generator=g, device=device, dtype=torch.bfloat16) * (scale / math.sqrt(i)))
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
    ap.add_argument("--batch", type=int, default=2); ap.add_argument("--steps", type=int, default=28)
    ap.add_argument("--res", type=int, default=1024); ap.add_argument("--n-text", type=int, default=512)
    ap.add_argument("--reps", type=int, default=2)
    a = ap.parse_args()
    dev = torch.device("cuda")
    cfg = FluxEngineConfig()
    t0 = time.time()
    sd = rand_state_dict(cfg, dev)
    eng = FluxRolloutEngine(cfg, sd, dev)
    del sd
    h2 = w2 = a.res // 16
    plan = eng.plan(a.batch, h2, w2, a.n_text)
    g = torch.Generator(device=dev).manual_seed(1)
    pe = torch.randn(a.batch, a.n_text, cfg.joint_attention_dim, generator=g, device=dev).bfloat16()
    pooled = torch.randn(a.batch, cfg.pooled_projection_dim, generator=g, device=dev).bfloat16()
    x0 = torch.randn(a.batch, h2 * w2, 64, generator=g, device=dev).half()
    eng.set_prompts(plan, pe, pooled, 3.5)
    T = a.steps
    ts, sig, coefs = eng.make_coefs(plan, T, 0.7, [1], store_slots=[(0 if i == T - 1 else -1) for i in range(T)], logp_slots=[(0 if i == 1 else -1) for i in range(T)])
    setup_s = time.time() - t0
    r = eng.rollout(plan, x0, coefs, 1, -1, 1, seed=3)          # warm-up (captures the graph)
    torch.cuda.synchronize()
    times = []
    for _ in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = eng.rollout(plan, x0, coefs, 1, -1, 1, seed=3); e1.record(); torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = sorted(times)[len(times) // 2]
    S, D, Ni = h2 * w2 + a.n_text, cfg.inner_dim, h2 * w2
    lin = cfg.num_layers * (2 * S * D * 12 * D) + cfg.num_single_layers * (2 * S * D * 7 * D + 2 * S * 5 * D * D)
    att = (cfg.num_layers + cfg.num_single_layers) * 4.0 * S * S * D
    fl = (lin + att) * T * a.batch
    print(json.dumps({"model": "FLUX.1-dev (random init)", "res": a.res, "batch": a.batch, "steps": T, "ms_per_rollout": ms,
                      "latents_per_s": a.batch / (ms / 1e3), "tflops": fl / ms / 1e9, "flops_per_latent_T": (lin + att) * T / 1e12,
                      "finite": bool(torch.isfinite(r["final_latents"].float()).all()), "log_prob": r["log_probs"].flatten().tolist(),
                      "launches": eng.last_launch_count(), "weights_GB": eng.weights.nbytes() / 1e9,
                      "workspace_GB": plan.workspace_bytes / 1e9, "setup_s": setup_s}))


if __name__ == "__main__":
    main()
