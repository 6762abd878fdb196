"""-m gpu parity of the FLUX.1 rollout path: engine forward / step / rollout through the C ABI vs the oracle
(oracle/flux_oracle.py, pinned bit-exact against the reference's FluxTransformer2DModel on CPU) and the committed golden fixture."""
import os
import pytest
import torch

pytestmark = pytest.mark.gpu

from tests.gpu_util import err_report, dump
from flow_factory_b200.flux import FluxRolloutEngine, model_scalar, pack_latents, flux_make_schedule
from oracle import flux_oracle as FO
from oracle import sd3_oracle as O


def _truth(cfg, w32, lat, pe, pooled, t, img_ids, guidance, dtype_sem=True):
    """fp32 oracle forward fed with the scalars the engine's bf16 semantics see (transformer_flux.py:679-682)."""
    B = lat.shape[0]
    t_eff = torch.full((B,), model_scalar(float(t) / 1000) / 1000)
    g_eff = torch.full((B,), model_scalar(guidance, torch.float16) / 1000)
    with torch.no_grad():
        return FO.flux_forward(w32, cfg, lat.float(), pe.float(), pooled.float(), t_eff, img_ids, torch.zeros(pe.shape[1], 3), guidance=g_eff)


def _setup(cfg, B, lh, lw, nt, seed):
    w32 = FO.make_flux_weights(cfg, seed=seed)
    wb = {k: v.bfloat16() for k, v in w32.items()}
    lat, pe, pooled, img_ids, _ = FO.make_flux_inputs(cfg, B, lh, lw, nt, seed=seed + 1)
    lat = lat.half()
    eng = FluxRolloutEngine(cfg, wb)
    plan = eng.plan(B, lh // 2, lw // 2, nt)
    return w32, wb, lat, pe.bfloat16(), pooled.bfloat16(), img_ids, eng, plan


@pytest.mark.parametrize("name", ["tiny", "tiny3", "mid"])
def test_flux_forward_matches_oracle(name):
    cfg, (B, lh, lw, nt), seed = {
        "tiny": (FO.tiny_flux_config(), (2, 8, 8, 7), 0),
        "tiny3": (FO.tiny_flux_config(num_layers=2, num_single_layers=1, heads=3, joint_dim=96, pooled_dim=48), (1, 12, 8, 21), 5),
        "mid": (FO.tiny_flux_config(num_layers=2, num_single_layers=3, heads=4, joint_dim=256, pooled_dim=64), (2, 32, 48, 77), 9),
    }[name]
    w32, wb, lat, pe, pooled, img_ids, eng, plan = _setup(cfg, B, lh, lw, nt, seed)
    t, gs = 612.0, 3.5
    eng.set_prompts(plan, pe, pooled, gs)
    got = eng.transformer_forward(plan, lat, t).float().cpu()
    torch.cuda.synchronize()
    # fp32 truth on the bf16-rounded weights / inputs, and the reference's own bf16 (CPU autocast) error against it
    wt = {k: v.float() for k, v in wb.items()}
    truth = _truth(cfg, wt, lat, pe, pooled, t, img_ids, gs)
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        B_ = lat.shape[0]
        ref_bf16 = FO.flux_forward(wb, cfg, lat.bfloat16(), pe, pooled, torch.full((B_,), t / 1000), img_ids.bfloat16(),
                                   torch.zeros(nt, 3).bfloat16(), guidance=torch.full((B_,), gs).half()).float()
    e_eng = float((got - truth).abs().max()); e_ref = float((ref_bf16 - truth).abs().max())
    scale = float(truth.abs().max())
    rep = dict(tag=f"flux_fwd_{name}", e_engine=e_eng, e_ref_bf16=e_ref, truth_absmax=scale,
               mean_engine=float((got - truth).abs().mean()), mean_ref=float((ref_bf16 - truth).abs().mean()))
    dump(f"flux_fwd_{name}.json", rep)
    assert torch.isfinite(got).all()
    assert e_eng <= max(1.3 * e_ref, 0.005 * scale), rep     # measured on B200: 0.60 - 0.98 x the bf16 reference's own error
    assert rep["mean_engine"] <= max(3.0 * rep["mean_ref"], 0.004 * scale), rep


def test_flux_step_and_rollout_match_golden(golden_dir):
    """T=4 Flow-SDE rollout of the tiny model: engine (bf16) vs the fixture minted from the REAL reference in fp32
    (tests/golden/flux_tiny.pt) - schedule bit-exact, latents within bf16 model error, log-probs <= 1e-3 relative given the
    trajectory (teacher-forced on the golden latents), graph replay == eager."""
    g = torch.load(os.path.join(golden_dir, "flux_tiny.pt"), weights_only=False)["rollout_fp32"]
    cfg = FO.tiny_flux_config()
    w32, wb, lat, pe, pooled, img_ids, eng, plan = _setup(cfg, 2, 8, 8, 7, 0)
    T, gs, nl = 4, 3.5, 0.7
    eng.set_prompts(plan, pe, pooled, gs)
    sde = O.current_sde_steps(T, None, None, 42)
    ts, sig, coefs = eng.make_coefs(plan, T, nl, sde, store_slots=[1, 2, 3, 4], logp_slots=[0, 1, 2, -1])
    assert torch.equal(ts, g["timesteps"]) and torch.equal(sig, g["sigmas"])
    noises = torch.stack(O.make_noises(T, tuple(lat.shape), seed=123))
    outs = {}
    for graph in (False, True):
        r = eng.rollout(plan, lat, coefs, n_latent_slots=T + 1, store_initial_slot=0, n_logp_slots=T - 1, noise=noises, use_graph=graph)
        torch.cuda.synchronize()
        outs[graph] = {k: v.cpu() for k, v in r.items()}
    assert torch.equal(outs[False]["all_latents"], outs[True]["all_latents"])
    assert torch.equal(outs[False]["log_probs"], outs[True]["log_probs"])
    r = outs[True]
    assert int(r["overflow"]) == 0
    assert torch.equal(r["all_latents"][:, 0], lat)
    assert torch.equal(r["all_latents"][:, T], r["final_latents"])
    # latents: the engine's distance to the fp32 reference trajectory must be of the size of the reference's OWN bf16 distance
    # (oracle under CPU autocast, same noise) - 4 SDE steps amplify the per-step bf16 model error
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        rb = FO.flux_rollout(wb, cfg, lat.bfloat16(), pe, pooled, img_ids, T, gs, noise_level=nl, noises=list(noises))
    for i in range(T + 1):
        d = float((r["all_latents"][:, i].float() - g["latents"][i].float()).abs().max())
        d_ref = float((rb["latents"][i].float() - g["latents"][i].float()).abs().max())
        assert d <= max(3.0 * d_ref, 0.05), (i, d, d_ref)
    # log-prob parity proper: teacher-force each SDE step on the golden (x_t, x_{t+1}) pair -> same Gaussian, only v differs
    for i in sorted(g["log_probs"]):
        out = eng.step(plan, g["latents"][i], coefs[i], next_latents=g["latents"][i + 1])
        torch.cuda.synchronize()
        lp, ref = out["log_prob"].cpu(), g["log_probs"][i]
        rel = float(((lp - ref).abs() / ref.abs().clamp_min(1e-6)).max())
        assert rel <= 5e-2, (i, lp.tolist(), ref.tolist())
        # and exactly (<= 1e-3 rel) against the oracle's scheduler.step given the engine's own noise prediction
        o = O.sde_step(out["noise_pred"].cpu(), g["latents"][i], float(sig[i]), float(sig[i + 1]), nl, float(sig[1]), "Flow-SDE",
                       next_latents=g["latents"][i + 1], compute_log_prob=True)
        rel2 = float(((lp - o["log_prob"]).abs() / o["log_prob"].abs().clamp_min(1e-6)).max())
        assert rel2 <= 1e-3, (i, rel2)
    assert eng.last_launch_count() > 0


def test_flux_adapter_inference_and_forward_api():
    """B200Flux1Adapter mirrors Flux1Adapter.inference / forward (flux1.py:152-349): sample fields, index maps, the packed-latent
    trajectory, and the GRPO invariant - replaying a stored (x_t, x_{t+1}) pair through forward() reproduces the rollout log-prob."""
    from flow_factory_b200.flux_adapter import B200Flux1Adapter
    from flow_factory_b200.scheduler import FlowMatchEulerDiscreteSDEScheduler
    cfg = FO.tiny_flux_config()
    wb = {k: v.bfloat16() for k, v in FO.make_flux_weights(cfg, seed=0).items()}
    sch = FlowMatchEulerDiscreteSDEScheduler(noise_level=0.7, use_dynamic_shifting=True, dynamics_type="Flow-SDE", num_sde_steps=2, seed=7)
    ad = B200Flux1Adapter(cfg, wb, scheduler=sch)
    g = torch.Generator().manual_seed(5)
    B, nt, T = 3, 9, 6
    pe = torch.randn(B, nt, cfg.joint_attention_dim, generator=g).bfloat16().cuda()
    pooled = torch.randn(B, cfg.pooled_projection_dim, generator=g).bfloat16().cuda()
    torch.manual_seed(11)
    samples = ad.inference(prompt=["a"] * B, height=128, width=192, num_inference_steps=T, guidance_scale=3.5, prompt_embeds=pe,
                           pooled_prompt_embeds=pooled, compute_log_prob=True, trajectory_indices="all")
    torch.cuda.synchronize()
    assert len(samples) == B
    s0 = samples[0]
    Ni = (128 // 16) * (192 // 16)
    assert tuple(s0.all_latents.shape) == (T + 1, Ni, 64) and s0.all_latents.dtype == torch.float16
    assert s0.latent_index_map.tolist() == list(range(T + 1))
    assert tuple(s0.img_ids.shape) == (Ni, 3) and s0.height == 128 and s0.width == 192
    sde_steps = sorted(sch.current_sde_steps.tolist())
    assert len(sde_steps) == 2 and s0.log_probs.numel() == 2
    assert torch.isfinite(s0.log_probs).all() and torch.isfinite(s0.all_latents.float()).all()
    # GRPO ratio == 1: teacher-forced replay of each SDE step
    lat = torch.stack([s.all_latents for s in samples])          # [B, T+1, Ni, 64]
    for k, i in enumerate(sde_steps):
        out = ad.forward(t=s0.timesteps[i], t_next=(s0.timesteps[i + 1] if i + 1 < T else torch.tensor(0.0)), latents=lat[:, i],
                         next_latents=lat[:, i + 1], prompt_embeds=pe, pooled_prompt_embeds=pooled, img_ids=s0.img_ids,
                         guidance_scale=3.5, noise_level=sch.noise_level, compute_log_prob=True)
        lp_roll = torch.stack([s.log_probs[k] for s in samples])
        torch.testing.assert_close(out.log_prob, lp_roll, rtol=1e-5, atol=1e-6)
