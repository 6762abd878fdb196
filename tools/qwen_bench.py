#!/usr/bin/env python
"""Qwen-Image rollout bench (BASELINE config 5: Qwen-Image 20B 1024^2 50-step DGPO rollout, 8xB200 + FSDP2 parameter shard).

  python tools/qwen_bench.py --gpus 1 --steps 1 --warmup 1
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port P tools/qwen_bench.py --gpus 8

Same contract as bench.py (one rank per GPU, device-timed with CUDA events, max over ranks, rank 0 prints one JSON line).  Qwen-Image
architecture (60 dual-stream blocks, D = 3072, head_dim 128, 20 B parameters), random-init weights created on the device, 4096 image tokens,
DGPO-style rollout = ODE without log-probs (FF/trainers/dgpo.py:865-886), true CFG (one batch of 2B + per-token norm rescale).
With N > 1 ranks the weights START as an FSDP2-style shard (every parameter a DTensor Shard(0) over the N ranks - what `fully_shard` leaves the
trainer with, config/accelerate_configs/fsdp*.yaml) and the engine takes them in through ONE all-gather per rollout
(flow_factory_b200.dist.gather_sharded_state_dict); its device time is reported as `fsdp2_gather_ms` next to the rollout."""
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
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--batch", type=int, default=1, help="prompts per rank per rollout")
    ap.add_argument("--res", type=int, default=1024)
    ap.add_argument("--n-text", type=int, default=256)
    ap.add_argument("--num-inference-steps", type=int, default=50)
    ap.add_argument("--guidance", type=float, default=4.0)
    ap.add_argument("--layers", type=int, default=60, help="transformer blocks (60 = the 20 B model; fewer only for quick checks)")
    a = ap.parse_args()
    import torch.distributed as dist
    from flow_factory_b200.dist import all_gather_rollout, gather_sharded_state_dict
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = FluxEngineConfig(num_layers=a.layers, num_single_layers=0, num_heads=24, joint_attention_dim=3584, pooled_projection_dim=8,
                           guidance_embeds=False, variant=1)
    t0 = time.time()
    sd = rand_state_dict(cfg, dev)                       # same seed on every rank: identical replicas
    gather_ms = None
    if world > 1:
        # FSDP2 layout: keep only this rank's dim-0 shard of every parameter, as DTensors on a 1-D mesh
        from torch.distributed.device_mesh import init_device_mesh
        from torch.distributed.tensor import DTensor, Shard
        mesh = init_device_mesh("cuda", (world,))
        sharded = {}
        for k in list(sd.keys()):
            full = sd.pop(k)
            chunk = -(-full.shape[0] // world)
            lo, hi = min(rank * chunk, full.shape[0]), min((rank + 1) * chunk, full.shape[0])
            sharded[k] = DTensor.from_local(full[lo:hi].clone(), mesh, [Shard(0)], run_check=False, shape=full.shape, stride=full.stride())
            del full
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sd = gather_sharded_state_dict(sharded, dtype=torch.bfloat16)          # the ONE collective
        e1.record(); torch.cuda.synchronize()
        g = torch.tensor([e0.elapsed_time(e1)], device=dev)
        dist.all_reduce(g, op=dist.ReduceOp.MAX)
        gather_ms = float(g)
        del sharded
    eng = QwenRolloutEngine(cfg, sd, dev)
    del sd
    torch.cuda.empty_cache()
    h2 = w2 = a.res // 16
    cfg_on = a.guidance > 1.0
    B, T = a.batch, a.num_inference_steps
    plan = eng.plan(B, h2, w2, a.n_text, cfg=cfg_on)
    g = torch.Generator(device=dev).manual_seed(1 + rank)
    pe = torch.randn(B, a.n_text, 3584, generator=g, device=dev).bfloat16()
    npe = torch.randn(B, a.n_text, 3584, generator=g, device=dev).bfloat16()
    x0 = torch.randn(B, h2 * w2, 64, generator=g, device=dev).half()
    ts, sig, coefs = eng.make_coefs(plan, T, 0.0, [], dynamics="ODE", store_slots=[(0 if i == T - 1 else -1) for i in range(T)])
    setup_s = time.time() - t0

    def rollout_device(p_, n_):
        eng.set_prompts(plan, p_, n_ if cfg_on else None, a.guidance)
        r = eng.rollout(plan, x0, coefs, 1, -1, 0)
        if world > 1:
            all_gather_rollout(r["all_latents"], torch.zeros(B, 1, device=dev))
        return r

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(1, a.warmup)):
        r = rollout_device(pe, npe)
    barrier()
    launches = eng.last_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        r = rollout_device(pe, npe)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    value = world * B * a.steps / (float(ms) / 1e3)
    # end to end: pinned host prompt embeddings in, final latents to the host, copies inside the timed region
    hpe, hnpe = pe.cpu().pin_memory(), npe.cpu().pin_memory()
    barrier()
    t1 = time.perf_counter()
    d2h = 0
    for _ in range(a.steps):
        rr = rollout_device(hpe.to(dev, non_blocking=True), hnpe.to(dev, non_blocking=True))
        out = rr["final_latents"].cpu()
        d2h = out.numel() * out.element_size()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t1], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    if rank == 0:
        S, D = h2 * w2 + a.n_text, cfg.inner_dim
        fwd = cfg.num_layers * (2 * S * D * 12 * D + 4.0 * S * S * D)
        fl_latent = fwd * T * (2 if cfg_on else 1)
        print(json.dumps({
            "metric": f"rollout latents/sec Qwen-Image 20B {a.res}^2 {T}-step DGPO", "value": value, "unit": "latents/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": float(ms) / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "impl": "b200",
            "config": {"workload": f"Qwen-Image architecture ({cfg.num_layers} dual blocks, D 3072, head_dim 128) {a.res}x{a.res} {T}-step DGPO rollout "
                                   f"(ODE, no log-prob), true CFG {a.guidance}, {a.n_text} text tokens, random-init weights",
                       "per_rank_batch": B, "global_batch": B * world,
                       "parallelism": f"dp{world} (prompt-sharded, 1 all-gather/rollout)" + ("; weights from an FSDP2 Shard(0) layout through ONE all-gather" if world > 1 else ""),
                       "cuda_graph": True},
            "e2e": {"value": world * B * a.steps / float(e2e_s), "unit": "latents/s", "h2d_bytes_per_step": 2 * hpe.numel() * 2, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches * a.steps), "fsdp2_gather_ms": gather_ms,
            "whole_step_achieved_tflops_per_gpu": value * fl_latent / 1e12 / world, "flops_per_latent": fl_latent,
            "finite": bool(torch.isfinite(r["final_latents"].float()).all()), "weights_GB": eng.weights.nbytes() / 1e9,
            "workspace_GB": plan.workspace_bytes / 1e9, "setup_s": setup_s}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
