"""-m gpu parity at the BENCHMARKED size and across paths (VERDICT r1 "close the parity gaps at the benchmarked size"):

* the 30-step, CFG, graph-replayed rollout that bench.py times (config C2: SD3.5-medium architecture, 1024^2, fp16 latent storage) against
  the oracle under the reference's numerics on this device (bf16 weights + CUDA autocast) and against an fp32 ground truth, position by
  position of the kept trajectory;
* the CROSS-PATH importance ratio: the engine's rollout `log_probs` against a teacher-forced replay of the same stored transitions
  through the REFERENCE numerics (oracle, bf16 CUDA autocast) - what integration/ff_b200_glue.py creates (rollout on the engine, autograd
  replay on diffusers).  SURVEY 7.2 #1 / grpo.py:263-276 need |ratio - 1| <= 1e-4; the measured values (written to gpurun_out, copied to
  profiles/) are 1e-4 .. 6e-4, so the glue implements option (b) of that section: old_log_prob is re-evaluated through the reference forward.
* the fp16 storage clamp of cast_latents (FF/models/abc.py:172-182) driven past +-65504.
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import sd3_oracle as O                     # noqa: E402  (the checker)
from tests.gpu_util import dump                         # noqa: E402

DEV = "cuda"


def _rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20))


def _adapter(cfg, w32, **sched_kw):
    from flow_factory_b200 import FlowMatchEulerDiscreteSDEScheduler
    from flow_factory_b200.adapter import B200SD3_5Adapter
    sch = FlowMatchEulerDiscreteSDEScheduler(noise_level=0.7, shift=3.0, **sched_kw)
    ad = B200SD3_5Adapter(cfg, w32, device=DEV, scheduler=sch, rng="torch")
    ad.rollout()
    return ad


def _replay_ratio(cfg, wb, inp, samples, ts, sig, g, sde_steps):
    """Teacher-forced log-prob of every stored SDE transition through the oracle (reference numerics, bf16 CUDA autocast), compared with
    the log-prob the engine recorded while sampling.  Returns max |exp(new - old) - 1| and the per-step values."""
    B = len(samples)
    worst, rows = 0.0, []
    lmap, pmap = samples[0].latent_index_map.tolist(), samples[0].log_prob_index_map.tolist()
    for i in sde_steps:
        if lmap[i] < 0 or lmap[i + 1] < 0 or pmap[i] < 0:
            continue
        x_t = torch.stack([s.all_latents[lmap[i]] for s in samples]).to(DEV)
        x_n = torch.stack([s.all_latents[lmap[i + 1]] for s in samples]).to(DEV)
        old = torch.stack([s.log_probs[pmap[i]] for s in samples]).to(DEV)
        timestep = ts[i].expand(B).to(device=DEV, dtype=x_t.dtype)
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            v = O.transformer_forward(wb, cfg, torch.cat([x_t, x_t]), torch.cat([inp["neg_prompt_embeds"], inp["prompt_embeds"]]),
                                      torch.cat([inp["neg_pooled"], inp["pooled"]]), timestep.repeat(2))
        vu, vc = v.chunk(2)
        v = vu + g * (vc - vu)
        r = O.sde_step(v, x_t, (ts[i] / 1000).item(), (ts[i + 1] / 1000).item() if i + 1 < len(ts) else 0.0, 0.7, float(sig[1]),
                       next_latents=x_n, compute_log_prob=True)
        ratio = torch.exp(r["log_prob"] - old)
        d = float((ratio - 1).abs().max())
        rows.append({"step": i, "max_abs_ratio_minus_1": d, "old": old.tolist(), "new": r["log_prob"].tolist()})
        worst = max(worst, d)
    return worst, rows


@pytest.mark.parametrize("name,kw,B,lh,lw,nt,T", [
    ("tiny", dict(), 2, 16, 16, 13, 6),
    ("mid", dict(num_layers=4, heads=4, dual=(0, 1), joint_dim=256, pooled_dim=128, pos_max=48, sample_size=64), 2, 32, 32, 77, 8),
])
def test_cross_path_ratio_small(name, kw, B, lh, lw, nt, T):
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = O.tiny_config(**kw)
    w32 = O.make_weights(cfg, seed=5)
    inp = {k: v.to(DEV).bfloat16() for k, v in O.make_inputs(cfg, B, lh, lw, nt, seed=6).items()}
    ad = _adapter(cfg, w32)
    samples = ad.inference(height=lh * 8, width=lw * 8, num_inference_steps=T, guidance_scale=4.5, prompt_embeds=inp["prompt_embeds"],
                           pooled_prompt_embeds=inp["pooled"], negative_prompt_embeds=inp["neg_prompt_embeds"],
                           negative_pooled_prompt_embeds=inp["neg_pooled"], compute_log_prob=True, trajectory_indices="all", latents=inp["x0"])
    ts, sig = O.make_schedule(T, 3.0)
    wb = {k: v.to(DEV).bfloat16() for k, v in w32.items()}
    worst, rows = _replay_ratio(cfg, wb, inp, samples, ts, sig, 4.5, range(T - 1))
    dump(f"cross_path_ratio_{name}.json", {"max_abs_ratio_minus_1": worst, "meets_1e-4": worst <= 1e-4, "steps": rows})
    assert len(rows) == T - 1
    # measured on a B200 (round 2): 2.0e-4 (tiny, T=6) / 5.9e-4 (mid, T=8): two independent bf16 forwards, amplified by CFG, at the large dt
    # of a short schedule.  That is ABOVE GRPO's 1e-4 clip range - which is why the reference-side glue re-evaluates old_log_prob through
    # the reference forward (integration/ff_b200_glue.py, SURVEY 7.2 #1 option b).  The bound here guards against regressions.
    assert worst <= 1.5e-3, worst


def test_c2_rollout_vs_reference_numerics_and_cross_path_ratio():
    """Config C2 as bench.py runs it (30 steps, CFG 4.5, CUDA graph replay, caller-provided noise = the reference's noise stream), B=2."""
    torch.backends.cuda.matmul.allow_tf32 = False
    from flow_factory_b200.trajectory import compute_trajectory_indices
    cfg = O.sd35_medium()
    T, B, g, lat, nt = 30, 2, 4.5, 128, 333
    w32 = O.make_weights(cfg, seed=0, device=DEV)
    inp = {k: v.to(DEV).bfloat16() for k, v in O.make_inputs(cfg, B, lat, lat, nt, seed=1).items()}
    noises = torch.stack(O.make_noises(T, (B, 16, lat, lat), seed=123)).to(DEV)
    ad = _adapter(cfg, w32, num_sde_steps=3, seed=7)
    ts_host = ad.scheduler.set_timesteps(T, seq_len=4096)
    sde_now = sorted(ad.scheduler.current_sde_steps.tolist())
    traj = compute_trajectory_indices(ad.scheduler.train_timesteps, T)
    samples = ad.inference(height=1024, width=1024, num_inference_steps=T, guidance_scale=g, prompt_embeds=inp["prompt_embeds"],
                           pooled_prompt_embeds=inp["pooled"], negative_prompt_embeds=inp["neg_prompt_embeds"],
                           negative_pooled_prompt_embeds=inp["neg_pooled"], compute_log_prob=True, trajectory_indices=traj,
                           latents=inp["x0"], noise=noises)
    torch.cuda.synchronize()
    assert ad._last_overflow is None or int(ad._last_overflow) == 0
    ts, sig = O.make_schedule(T, 3.0)
    assert torch.equal(ts, ts_host)
    # reference numerics and fp32 truth with the same noise and the same SDE window
    wb = {k: v.bfloat16() for k, v in w32.items()}
    with torch.no_grad():
        rb = O.rollout(wb, cfg, inp["x0"], inp["prompt_embeds"], inp["pooled"], inp["neg_prompt_embeds"], inp["neg_pooled"], T, g,
                       sde_step_indices=sde_now, noises=list(noises), autocast="cuda")
    del wb
    f32 = {k: v.float() for k, v in inp.items()}
    with torch.no_grad():
        r32 = O.rollout(w32, cfg, f32["x0"], f32["prompt_embeds"], f32["pooled"], f32["neg_prompt_embeds"], f32["neg_pooled"], T, g,
                        sde_step_indices=sde_now, noises=list(noises))
    lmap, pmap = samples[0].latent_index_map.tolist(), samples[0].log_prob_index_map.tolist()
    rep = {"sde_steps": sde_now, "positions": []}
    for pos, slot in enumerate(lmap):
        if slot < 0:
            continue
        got = torch.stack([s.all_latents[slot] for s in samples]).float()
        e_eng, e_ref = _rel(got, r32["latents"][pos]), _rel(rb["latents"][pos], r32["latents"][pos])
        rep["positions"].append({"position": pos, "err_engine_vs_fp32": e_eng, "err_ref_bf16_vs_fp32": e_ref, "err_engine_vs_ref": _rel(got, rb["latents"][pos])})
        assert e_eng <= 3.0 * e_ref + 2e-3, (pos, e_eng, e_ref)
    final = torch.stack([s.final_latents for s in samples]).float()
    rep["final"] = {"err_engine_vs_fp32": _rel(final, r32["latents"][T]), "err_ref_bf16_vs_fp32": _rel(rb["latents"][T], r32["latents"][T])}
    assert rep["final"]["err_engine_vs_fp32"] <= 3.0 * rep["final"]["err_ref_bf16_vs_fp32"] + 2e-3
    lp_rows = []
    for i in sde_now:
        if pmap[i] < 0:
            continue
        lp = torch.stack([s.log_probs[pmap[i]] for s in samples])
        rel = float(((lp - rb["log_probs"][i]).abs() / rb["log_probs"][i].abs()).max())
        lp_rows.append({"step": i, "engine": lp.tolist(), "reference_bf16": rb["log_probs"][i].tolist(), "max_rel": rel})
        assert rel <= 1e-3, (i, rel)                      # north star: log-probs <= 1e-3 rel
    rep["log_probs"] = lp_rows
    assert lp_rows, "no SDE step inside the kept trajectory"
    # cross-path ratio on the stored transitions of this very rollout
    wb = {k: v.bfloat16() for k, v in w32.items()}
    worst, rows = _replay_ratio(cfg, wb, inp, samples, ts, sig, g, sde_now)
    rep["cross_path_ratio"] = {"max_abs_ratio_minus_1": worst, "meets_1e-4": worst <= 1e-4, "steps": rows}
    dump("parity_c2_rollout.json", rep)
    # measured (round 2): 2.7e-6 at step 0, 1.1e-4 at step 18, 2.1e-4 at step 21 - at / above the 1e-4 clip range for late steps, hence the
    # glue's reference-path recomputation of old_log_prob (see test_cross_path_ratio_small); regression bound:
    assert rows and worst <= 6e-4, worst


def test_fp16_storage_clamp_and_overflow_flag():
    """cast_latents (FF/models/abc.py:172-182): next_latents beyond the fp16 range are clamped to +-65504 (never inf) and the sticky
    overflow flag is raised; in-range values are untouched and the flag stays 0.  Driven through the engine step with latents near the
    fp16 maximum and a velocity that pushes part of them over."""
    from flow_factory_b200 import FlowMatchEulerDiscreteSDEScheduler, RolloutEngine
    cfg = O.tiny_config()
    w32 = O.make_weights(cfg, seed=0)
    inp = {k: v.to(DEV) for k, v in O.make_inputs(cfg, 2, 16, 16, 13, seed=2).items()}
    eng = RolloutEngine(cfg, w32, DEV)
    sch = FlowMatchEulerDiscreteSDEScheduler(noise_level=0.7, shift=3.0, dynamics_type="Flow-SDE")
    ts = sch.set_timesteps(8, seq_len=64)
    plan = eng.plan(2, False, 16, 16, 13)
    eng.set_prompts(plan, inp["prompt_embeds"], inp["pooled"])
    # the fused final-step kernel: latents at the edge of the fp16 range, caller noise large enough to push rows 8.. of sample 0 above
    # +65504 and of sample 1 below -65504; rows 0-7 stay far inside
    x = torch.zeros(2, 16, 16, 16, device=DEV)
    x[0, :, 8:], x[1, :, 8:] = 65000.0, -65000.0
    x[:, :, :8] = torch.randn(2, 16, 8, 16, device=DEV)
    noise = torch.zeros(2, 16, 16, 16, device=DEV)
    noise[0, :, 8:], noise[1, :, 8:] = 1.0e5, -1.0e5
    coef = sch.step_coef(ts[2], ts[3], 0.7, t_model=float(ts[2].half()))
    r = eng.step(plan, x.half(), coef, 1.0, noise=noise)
    nxt, mean = r["next_latents"].float(), r["next_latents_mean"]
    want = mean + coef.noise_scale * noise
    assert torch.isfinite(nxt).all()
    assert bool((want[0, :, 8:] > 65504.0).all()) and bool((want[1, :, 8:] < -65504.0).all())      # the test does drive the clamp
    assert bool((nxt[0, :, 8:] == 65504.0).all()) and bool((nxt[1, :, 8:] == -65504.0).all())
    assert torch.equal(nxt[:, :, :8], want[:, :, :8].half().float())
    assert int(r["overflow"]) != 0
    r_ok = eng.step(plan, x.half() * 1e-3, coef, 1.0, noise=noise * 1e-4)
    assert int(r_ok["overflow"]) == 0 and float(r_ok["next_latents"].float().abs().max()) < 65504.0
    # a direct drive of the standalone step kernel with a velocity that certainly overflows: exact clamp semantics
    v = torch.zeros(2, 16, 16, 16, device=DEV).bfloat16()
    v[0, 0] = -3.0e6                                      # x + v dt with dt < 0 -> far above +65504
    v[1, 1] = 3.0e6
    xs = torch.zeros(2, 16, 16, 16, device=DEV).half()
    sch2 = FlowMatchEulerDiscreteSDEScheduler(noise_level=0.7, shift=3.0, dynamics_type="Flow-SDE")
    sch2.set_timesteps(8, seq_len=64)
    o2 = sch2.step(noise_pred=v, timestep=ts[2], latents=xs, timestep_next=ts[3], noise_level=0.7, compute_log_prob=True,
                   noise=torch.zeros(2, 16, 16, 16, device=DEV))
    n2 = o2.next_latents
    assert float(o2.next_latents_mean[0, 0].min()) > 65504.0 and float(o2.next_latents_mean[1, 1].max()) < -65504.0
    assert torch.isfinite(n2).all()
    assert bool((n2[0, 0] == 65504.0).all()) and bool((n2[1, 1] == -65504.0).all())      # the reference: clamp(-65504, 65504) then .to(fp16)
    untouched = torch.ones_like(n2, dtype=torch.bool); untouched[0, 0] = False; untouched[1, 1] = False
    assert torch.equal(n2[untouched], o2.next_latents_mean[untouched].half().float())
