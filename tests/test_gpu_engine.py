"""-m gpu parity: the whole MMDiT forward, the fused step and the T-step rollout through the C ABI, against
(a) golden fixtures minted from the reference, (b) the oracle run on the same device, (c) invariants at full size."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import sd3_oracle as O
from tests.gpu_util import dump


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), weights_only=False)


def _engine(cfg, w):
    from flow_factory_b200 import RolloutEngine
    return RolloutEngine(cfg, w, "cuda")


def _oracle_fwd(cfg, w32, x, pe, pp, t, mode):
    """mode 'fp32' = ground truth, 'bf16' = reference numerics on this device (bf16 weights + CUDA autocast)."""
    dev = "cuda"
    tt = torch.full((x.shape[0],), float(t), device=dev)
    if mode == "fp32":
        w = {k: v.to(dev) for k, v in w32.items()}
        with torch.no_grad():
            return O.transformer_forward(w, cfg, x.float().to(dev), pe.float().to(dev), pp.float().to(dev), tt)
    w = {k: v.to(dev).bfloat16() for k, v in w32.items()}
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        return O.transformer_forward(w, cfg, x.half().to(dev), pe.bfloat16().to(dev), pp.bfloat16().to(dev), tt.half())


def _rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm())


@pytest.mark.parametrize("name,kw,B,lh,lw,nt", [
    ("tiny", dict(), 2, 16, 16, 13),
    ("tiny3", dict(num_layers=3, heads=3, dual=(0, 1), joint_dim=96, pooled_dim=48, pos_max=24, sample_size=32), 1, 24, 16, 21),
    ("mid", dict(num_layers=4, heads=4, dual=(0, 1), joint_dim=256, pooled_dim=128, pos_max=48, sample_size=64), 2, 32, 32, 77),
])
def test_forward_vs_oracle(name, kw, B, lh, lw, nt):
    """Engine error against the fp32 truth must be comparable to the bf16 reference's own error."""
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = O.tiny_config(**kw)
    w32 = O.make_weights(cfg, seed=3)
    inp = O.make_inputs(cfg, B, lh, lw, nt, seed=4)
    t = float(torch.tensor(612.3).half())
    eng = _engine(cfg, w32)
    plan = eng.plan(B, False, lh, lw, nt)
    eng.set_prompts(plan, inp["prompt_embeds"], inp["pooled"])
    v = eng.transformer_forward(plan, inp["x0"].half(), t)
    torch.cuda.synchronize()
    truth = _oracle_fwd(cfg, w32, inp["x0"].half(), inp["prompt_embeds"].bfloat16(), inp["pooled"].bfloat16(), t, "fp32")
    ref = _oracle_fwd(cfg, w32, inp["x0"], inp["prompt_embeds"], inp["pooled"], t, "bf16")
    e_eng, e_ref, e_cross = _rel(v, truth), _rel(ref, truth), _rel(v, ref)
    dump(f"fwd_{name}.json", dict(err_engine_vs_fp32=e_eng, err_ref_bf16_vs_fp32=e_ref, err_engine_vs_ref=e_cross,
                                  launches=eng.last_launch_count()))
    assert not torch.isnan(v).any()
    assert e_eng <= 1.3 * e_ref + 5e-4, (e_eng, e_ref, e_cross)


def test_forward_cfg_batching_order():
    """CFG: uncond half first, both halves share the latents (sd3_5.py:409-413)."""
    cfg = O.tiny_config()
    w32 = O.make_weights(cfg, seed=0)
    inp = O.make_inputs(cfg, 2, 16, 16, 13, seed=1)
    eng = _engine(cfg, w32)
    t = 988.5
    p2 = eng.plan(2, True, 16, 16, 13)
    eng.set_prompts(p2, inp["prompt_embeds"], inp["pooled"], inp["neg_prompt_embeds"], inp["neg_pooled"])
    v2 = eng.transformer_forward(p2, inp["x0"].half(), t)
    p1 = eng.plan(2, False, 16, 16, 13)
    eng.set_prompts(p1, inp["neg_prompt_embeds"], inp["neg_pooled"])
    vu = eng.transformer_forward(p1, inp["x0"].half(), t)
    eng.set_prompts(p1, inp["prompt_embeds"], inp["pooled"])
    vc = eng.transformer_forward(p1, inp["x0"].half(), t)
    assert torch.equal(v2[:2], vu) and torch.equal(v2[2:], vc)


def test_rollout_tiny_vs_reference_golden(golden_dir):
    """T=4, CFG 4.5, Flow-SDE, fp16 storage, same noise stream: latents and log-probs against the fixture minted from
    the reference loop (bf16 autocast) and its fp32 twin.  log-prob tolerance: 1e-3 relative (north-star)."""
    g = _load(golden_dir, "rollout_tiny.pt")
    from flow_factory_b200.adapter import B200SD3_5Adapter
    from flow_factory_b200 import FlowMatchEulerDiscreteSDEScheduler
    cfg = O.tiny_config()
    w32 = O.make_weights(cfg, seed=0)
    inp = O.make_inputs(cfg, 2, 16, 16, 13, seed=1)
    noises = torch.stack(O.make_noises(4, (2, 16, 16, 16), seed=123))
    ad = B200SD3_5Adapter(cfg, w32, scheduler=FlowMatchEulerDiscreteSDEScheduler(noise_level=0.7, shift=3.0))
    ad.rollout()
    for use_graph in (False, True):
        ad.use_graph = use_graph
        samples = ad.inference(height=128, width=128, num_inference_steps=4, guidance_scale=4.5,
                               prompt_embeds=inp["prompt_embeds"].cuda(), pooled_prompt_embeds=inp["pooled"].cuda(),
                               negative_prompt_embeds=inp["neg_prompt_embeds"].cuda(), negative_pooled_prompt_embeds=inp["neg_pooled"].cuda(),
                               compute_log_prob=True, trajectory_indices="all", latents=inp["x0"].bfloat16().cuda(), noise=noises.cuda())
        assert len(samples) == 2
        assert torch.equal(samples[0].timesteps, g["bf16"]["timesteps"])
        rep = {}
        for b in range(2):
            s = samples[b]
            assert s.all_latents.shape == (5, 16, 16, 16) and s.log_probs.shape == (3,)
            assert s.latent_index_map.tolist() == [0, 1, 2, 3, 4]
            for pos in range(5):
                got = s.all_latents[pos].float().cpu()
                t32, tb = g["fp32"]["latents"][pos][b].float(), g["bf16"]["latents"][pos][b].float()
                e_eng, e_ref = float((got - t32).abs().max()), float((tb - t32).abs().max())
                rep[f"b{b}_pos{pos}"] = (e_eng, e_ref)
                assert e_eng <= 3 * e_ref + 4e-3, (b, pos, e_eng, e_ref)
            for i in range(3):
                lp, lp_ref = float(s.log_probs[i]), float(g["bf16"]["log_probs"][i][b])
                assert abs(lp - lp_ref) <= 1e-3 * abs(lp_ref), (i, lp, lp_ref)
        dump(f"rollout_tiny_graph{int(use_graph)}.json", rep)


def test_teacher_forced_replay_reproduces_rollout_log_probs():
    """The GRPO invariant ratio == exp(new_logp - old_logp) == 1 before the first update (grpo.py:271): replaying a stored
    transition through forward(next_latents=stored) returns the rollout's log-prob."""
    from flow_factory_b200.adapter import B200SD3_5Adapter
    from flow_factory_b200 import FlowMatchEulerDiscreteSDEScheduler
    from flow_factory_b200.trajectory import compute_trajectory_indices
    cfg = O.tiny_config()
    w32 = O.make_weights(cfg, seed=0)
    inp = {k: v.cuda() for k, v in O.make_inputs(cfg, 2, 16, 16, 13, seed=1).items()}
    sch = FlowMatchEulerDiscreteSDEScheduler(noise_level=0.7, shift=3.0, num_sde_steps=2, seed=5)
    ad = B200SD3_5Adapter(cfg, w32, scheduler=sch, rng="philox")
    ad.rollout()
    sch.set_timesteps(6, seq_len=64)
    train_steps = sch.train_timesteps.tolist()
    idx = compute_trajectory_indices(train_steps, 6)
    kw = dict(prompt_embeds=inp["prompt_embeds"], pooled_prompt_embeds=inp["pooled"],
              negative_prompt_embeds=inp["neg_prompt_embeds"], negative_pooled_prompt_embeds=inp["neg_pooled"], guidance_scale=4.5)
    samples = ad.inference(height=128, width=128, num_inference_steps=6, compute_log_prob=True, trajectory_indices=idx,
                           latents=inp["x0"].bfloat16(), **kw)
    st = type(samples[0]).stack(samples)
    for i in sorted(train_steps):
        li, lj = int(samples[0].latent_index_map[i]), int(samples[0].latent_index_map[i + 1])
        lp_slot = int(samples[0].log_prob_index_map[i])
        assert li >= 0 and lj >= 0 and lp_slot >= 0
        t = sch.timesteps[i]
        tn = sch.timesteps[i + 1] if i + 1 < 6 else torch.tensor(0.0)
        out = ad.forward(t=t, t_next=tn, latents=st["all_latents"][:, li], next_latents=st["all_latents"][:, lj],
                         noise_level=sch.noise_level, compute_log_prob=True, return_kwargs=["log_prob", "next_latents_mean"], **kw)
        old = st["log_probs"][:, lp_slot]
        ratio = torch.exp(out.log_prob - old)
        assert float((ratio - 1).abs().max()) <= 1e-5, (i, ratio)


def test_adapter_trajectory_subset_and_no_logprob():
    from flow_factory_b200.adapter import B200SD3_5Adapter
    cfg = O.tiny_config()
    w32 = O.make_weights(cfg, seed=0)
    inp = {k: v.cuda() for k, v in O.make_inputs(cfg, 1, 16, 16, 13, seed=2).items()}
    ad = B200SD3_5Adapter(cfg, w32, rng="philox")
    s = ad.inference(height=128, width=128, num_inference_steps=5, guidance_scale=1.0, prompt_embeds=inp["prompt_embeds"],
                     pooled_prompt_embeds=inp["pooled"], compute_log_prob=False, trajectory_indices=[-1])[0]
    assert s.all_latents.shape[0] == 1 and s.latent_index_map.tolist() == [-1, -1, -1, -1, -1, 0]
    assert s.log_probs is None and s.log_prob_index_map is None
    assert torch.equal(s.all_latents[0], s.extra_kwargs["final_latents"])
    ad.eval()   # ODE sampling, noise_level 0 everywhere (scheduler.eval)
    s2 = ad.inference(height=128, width=128, num_inference_steps=5, guidance_scale=1.0, prompt_embeds=inp["prompt_embeds"],
                      pooled_prompt_embeds=inp["pooled"], compute_log_prob=True, trajectory_indices=None,
                      latents=torch.zeros(1, 16, 16, 16, device="cuda"))[0]
    assert s2.all_latents is None and s2.latent_index_map is None


def test_forward_sd35_medium_1024_vs_oracle():
    """Full-size config C2 (24 layers, 13 dual, D=1536, 1024^2 -> 4096+333 tokens), one CFG forward, random-init weights."""
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = O.sd35_medium()
    w32 = O.make_weights(cfg, seed=0, device="cuda")
    inp = O.make_inputs(cfg, 1, 128, 128, 333, seed=1)
    t = 612.5
    eng = _engine(cfg, w32)
    plan = eng.plan(1, False, 128, 128, 333)
    eng.set_prompts(plan, inp["prompt_embeds"], inp["pooled"])
    v = eng.transformer_forward(plan, inp["x0"].half(), t)
    torch.cuda.synchronize()
    ref = _oracle_fwd(cfg, w32, inp["x0"], inp["prompt_embeds"], inp["pooled"], t, "bf16")
    wb = None
    truth = _oracle_fwd(cfg, w32, inp["x0"].half(), inp["prompt_embeds"].bfloat16(), inp["pooled"].bfloat16(), t, "fp32")
    e_eng, e_ref, e_cross = _rel(v, truth), _rel(ref, truth), _rel(v, ref)
    dump("fwd_sd35_medium_1024.json", dict(err_engine_vs_fp32=e_eng, err_ref_bf16_vs_fp32=e_ref, err_engine_vs_ref=e_cross,
                                           launches=eng.last_launch_count(), workspace_gb=plan.workspace_bytes / 2 ** 30))
    assert not torch.isnan(v).any()
    assert e_eng <= 1.3 * e_ref + 5e-4, (e_eng, e_ref, e_cross)


@pytest.mark.parametrize("dyn", ["Flow-SDE", "CPS", "ODE"])
@pytest.mark.parametrize("cfg_on", [True, False])
def test_fused_final_step_equals_unfused_composition(dyn, cfg_on):
    """The fused proj_out+CFG+unpatchify+scheduler.step epilogue (final_step.cu) against the same result assembled from
    pieces that are each pinned elsewhere: plain proj_out GEMM -> unpatchify -> bf16 CFG (torch) -> sde_step kernel
    (which is bit-checked against the reference-minted fixtures)."""
    from flow_factory_b200 import FlowMatchEulerDiscreteSDEScheduler
    cfg = O.tiny_config(num_layers=2, heads=2, dual=(0,), pos_max=24, sample_size=32)
    w32 = O.make_weights(cfg, seed=9)
    B, lh, lw, nt = 3, 24, 16, 17                    # 96 tokens: ragged 128-token tile
    inp = {k: v.cuda() for k, v in O.make_inputs(cfg, B, lh, lw, nt, seed=10).items()}
    eng = _engine(cfg, w32)
    sch = FlowMatchEulerDiscreteSDEScheduler(noise_level=0.7, shift=3.0, dynamics_type=dyn)
    ts = sch.set_timesteps(8, seq_len=96)
    plan = eng.plan(B, cfg_on, lh, lw, nt)
    eng.set_prompts(plan, inp["prompt_embeds"], inp["pooled"], inp["neg_prompt_embeds"] if cfg_on else None,
                    inp["neg_pooled"] if cfg_on else None)
    x = inp["x0"].half()
    g = 4.5 if cfg_on else 1.0
    noise = torch.randn(B, 16, lh, lw, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))
    t_model = float(ts[2].half())
    for use_noise in (True, False):
        coef = sch.step_coef(ts[2], ts[3], 0.7, t_model=t_model)
        r = eng.step(plan, x, coef, g, noise=noise if use_noise else None, seed=77, step_index=0)
        v = eng.transformer_forward(plan, x, t_model)
        if cfg_on:
            vu, vc = v.chunk(2)
            v = vu + g * (vc - vu)                    # bf16 tensor ops, as sd3_5.py:431-433
        assert torch.equal(r["noise_pred"], v)
        ref = sch.step(noise_pred=v, timestep=ts[2], latents=x, timestep_next=ts[3], noise_level=0.7, compute_log_prob=True,
                       noise=noise if use_noise else None, seed=77, step_index=0)
        torch.testing.assert_close(r["next_latents_mean"], ref.next_latents_mean, rtol=0, atol=0)
        if dyn == "ODE":
            assert torch.equal(r["next_latents"].float(), ref.next_latents.half().float())
        else:
            assert torch.equal(r["next_latents"].float(), ref.next_latents)      # same Philox stream in both kernels
        torch.testing.assert_close(r["log_prob"], ref.log_prob, rtol=1e-6, atol=1e-7)


def test_refresh_weights_tracks_the_trainer():
    """Weights move under the engine (optimizer / EMA / LoRA merge): refresh_weights re-packs in place."""
    cfg = O.tiny_config()
    w32 = O.make_weights(cfg, seed=0)
    inp = O.make_inputs(cfg, 1, 16, 16, 13, seed=1)
    eng = _engine(cfg, w32)
    plan = eng.plan(1, False, 16, 16, 13)
    eng.set_prompts(plan, inp["prompt_embeds"], inp["pooled"])
    v0 = eng.transformer_forward(plan, inp["x0"].half(), 500.0).clone()
    w2 = {k: v.clone() for k, v in w32.items()}
    w2["transformer_blocks.1.ff.net.2.weight"] *= 1.5
    w2["transformer_blocks.0.attn.to_q.weight"] += 0.05
    eng.refresh_weights(w2)
    v1 = eng.transformer_forward(plan, inp["x0"].half(), 500.0)
    assert not torch.equal(v0, v1)
    truth = _oracle_fwd(cfg, w2, inp["x0"].half(), inp["prompt_embeds"].bfloat16(), inp["pooled"].bfloat16(), 500.0, "fp32")
    ref = _oracle_fwd(cfg, w2, inp["x0"], inp["prompt_embeds"], inp["pooled"], 500.0, "bf16")
    assert _rel(v1, truth) <= 1.3 * _rel(ref, truth) + 5e-4
    eng.refresh_weights(w32)
    assert torch.equal(eng.transformer_forward(plan, inp["x0"].half(), 500.0), v0)


def test_rollout_host_entry_matches_device_entry():
    """ffb200_rollout_host (host buffers in/out, copies inside the call) == ffb200_rollout on device buffers (same Philox seed)."""
    from flow_factory_b200 import FlowMatchEulerDiscreteSDEScheduler
    from flow_factory_b200.trajectory import plan_slots
    cfg = O.tiny_config()
    w32 = O.make_weights(cfg, seed=0)
    inp = O.make_inputs(cfg, 2, 16, 16, 13, seed=1)
    eng = _engine(cfg, w32)
    sch = FlowMatchEulerDiscreteSDEScheduler(noise_level=0.7, shift=3.0)
    T = 5
    ts = sch.set_timesteps(T, seq_len=64)
    has = [i < T - 1 for i in range(T)]
    lat_slot, lp_slot, _, _ = plan_slots("all", T, has)
    coefs = [sch.step_coef(ts[i], ts[i + 1] if i + 1 < T else torch.tensor(0.0), 0.7 if has[i] else 0.0, compute_log_prob=has[i],
                           t_model=float(ts[i].half()), store_slot=lat_slot[i + 1], logp_slot=lp_slot[i]) for i in range(T)]
    plan = eng.plan(2, True, 16, 16, 13)
    pe = torch.cat([inp["neg_prompt_embeds"], inp["prompt_embeds"]]).bfloat16().contiguous()
    pp = torch.cat([inp["neg_pooled"], inp["pooled"]]).bfloat16().contiguous()
    x0 = inp["x0"].half().contiguous()
    rh = eng.rollout_host(plan, x0, pe, pp, coefs, 4.5, T + 1, 0, T - 1, seed=99, use_graph=True)
    eng.set_prompts(plan, inp["prompt_embeds"], inp["pooled"], inp["neg_prompt_embeds"], inp["neg_pooled"])
    rd = eng.rollout(plan, x0, coefs, 4.5, T + 1, 0, T - 1, seed=99, use_graph=False)
    torch.cuda.synchronize()
    assert torch.equal(rh["all_latents"], rd["all_latents"].cpu())
    assert torch.equal(rh["log_probs"], rd["log_probs"].cpu())
    assert torch.equal(rh["final_latents"], rd["final_latents"].cpu())
    assert torch.equal(rd["all_latents"][:, 0].cpu(), x0) and torch.equal(rd["all_latents"][:, T], rd["final_latents"])
    assert int(rh["overflow"]) == 0


@pytest.mark.parametrize("B,nt,lh,lw", [(3, 77, 32, 32), (1, 333, 16, 48), (5, 1, 16, 16)])
def test_odd_batch_text_length_and_aspect(B, nt, lh, lw):
    cfg = O.tiny_config(num_layers=2, heads=2, dual=(0,), pos_max=32, sample_size=32)
    w32 = O.make_weights(cfg, seed=11)
    inp = O.make_inputs(cfg, B, lh, lw, nt, seed=12)
    eng = _engine(cfg, w32)
    plan = eng.plan(B, False, lh, lw, nt)
    eng.set_prompts(plan, inp["prompt_embeds"], inp["pooled"])
    v = eng.transformer_forward(plan, inp["x0"].half(), 250.0)
    truth = _oracle_fwd(cfg, w32, inp["x0"].half(), inp["prompt_embeds"].bfloat16(), inp["pooled"].bfloat16(), 250.0, "fp32")
    ref = _oracle_fwd(cfg, w32, inp["x0"], inp["prompt_embeds"], inp["pooled"], 250.0, "bf16")
    assert not torch.isnan(v).any()
    assert _rel(v, truth) <= 1.3 * _rel(ref, truth) + 5e-4, (_rel(v, truth), _rel(ref, truth))


def test_dance_sde_step_through_the_engine():
    from flow_factory_b200 import FlowMatchEulerDiscreteSDEScheduler
    cfg = O.tiny_config()
    w32 = O.make_weights(cfg, seed=0)
    inp = {k: v.cuda() for k, v in O.make_inputs(cfg, 2, 16, 16, 13, seed=1).items()}
    eng = _engine(cfg, w32)
    sch = FlowMatchEulerDiscreteSDEScheduler(noise_level=0.5, shift=3.0, dynamics_type="Dance-SDE")
    ts = sch.set_timesteps(10, seq_len=64)
    plan = eng.plan(2, False, 16, 16, 13)
    eng.set_prompts(plan, inp["prompt_embeds"], inp["pooled"])
    x = inp["x0"].half()
    noise = torch.randn(2, 16, 16, 16, device="cuda")
    coef = sch.step_coef(ts[4], ts[5], 0.5, t_model=float(ts[4].half()))
    r = eng.step(plan, x, coef, 1.0, noise=noise)
    ro = O.sde_step(r["noise_pred"], x, (ts[4] / 1000).item(), (ts[5] / 1000).item(), 0.5, float(sch.sigmas[1]), "Dance-SDE", noise=noise)
    torch.testing.assert_close(r["next_latents_mean"], ro["next_latents_mean"], rtol=1e-6, atol=1e-6)
    assert torch.equal(r["next_latents"].float(), ro["next_latents"])
    torch.testing.assert_close(r["log_prob"], ro["log_prob"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("storage,dtype", [("bf16", torch.bfloat16), ("fp32", torch.float32), (None, torch.bfloat16)])
def test_rollout_latent_storage_dtypes(storage, dtype):
    """latent_storage_dtype other than fp16 (cast_latents, FF/models/abc.py:172-182; None = the transformer dtype): the kept trajectory
    comes back in that dtype, the timestep fed to the model is rounded through it (sd3_5.py:394), and the rollout tracks the oracle run with
    the same storage dtype as closely as the bf16 reference tracks fp32; the in-rollout log-probs equal a teacher-forced replay."""
    torch.backends.cuda.matmul.allow_tf32 = False
    from flow_factory_b200 import FlowMatchEulerDiscreteSDEScheduler
    from flow_factory_b200.adapter import B200SD3_5Adapter
    cfg = O.tiny_config()
    w32 = O.make_weights(cfg, seed=0)
    inp = {k: v.cuda() for k, v in O.make_inputs(cfg, 2, 16, 16, 13, seed=1).items()}
    T = 4
    noises = torch.stack(O.make_noises(T, (2, 16, 16, 16), seed=123)).cuda()
    ad = B200SD3_5Adapter(cfg, w32, scheduler=FlowMatchEulerDiscreteSDEScheduler(noise_level=0.7, shift=3.0), latent_storage_dtype=storage)
    ad.rollout()
    kw = dict(height=128, width=128, num_inference_steps=T, guidance_scale=4.5, prompt_embeds=inp["prompt_embeds"], pooled_prompt_embeds=inp["pooled"],
              negative_prompt_embeds=inp["neg_prompt_embeds"], negative_pooled_prompt_embeds=inp["neg_pooled"], compute_log_prob=True,
              trajectory_indices="all", latents=inp["x0"].bfloat16(), noise=noises)
    samples = ad.inference(**kw)
    assert samples[0].all_latents.dtype == dtype and samples[0].final_latents.dtype == dtype
    wb = {k: v.cuda().bfloat16() for k, v in w32.items()}
    bf = {k: v.bfloat16() for k, v in inp.items()}
    with torch.no_grad():
        rb = O.rollout(wb, cfg, bf["x0"], bf["prompt_embeds"], bf["pooled"], bf["neg_prompt_embeds"], bf["neg_pooled"], T, 4.5,
                       noises=list(noises), autocast="cuda", storage_dtype=dtype)
        r32 = O.rollout({k: v.cuda() for k, v in w32.items()}, cfg, bf["x0"].float(), bf["prompt_embeds"].float(), bf["pooled"].float(),
                        bf["neg_prompt_embeds"].float(), bf["neg_pooled"].float(), T, 4.5, noises=list(noises), storage_dtype=dtype)
    for b in range(2):
        got = samples[b].all_latents[-1].float()
        e_eng = float((got - r32["latents"][-1][b].float()).abs().max())
        e_ref = float((rb["latents"][-1][b].float() - r32["latents"][-1][b].float()).abs().max())
        assert e_eng <= 3 * e_ref + 4e-3, (e_eng, e_ref)
        for jj, i in enumerate(sorted(rb["log_probs"])):
            lp, ref = float(samples[b].log_probs[jj]), float(rb["log_probs"][i][b])
            assert abs(lp - ref) <= 1e-3 * abs(ref), (lp, ref)
    # teacher-forced replay of stored transition 1 -> 2 through forward(): the same log-prob
    sch = ad.scheduler
    ts = sch.set_timesteps(T, seq_len=64)
    xt = torch.stack([s_.all_latents[1] for s_ in samples]); xn = torch.stack([s_.all_latents[2] for s_ in samples])
    out = ad.forward(t=ts[1], t_next=ts[2], latents=xt, next_latents=xn, prompt_embeds=inp["prompt_embeds"], pooled_prompt_embeds=inp["pooled"],
                     negative_prompt_embeds=inp["neg_prompt_embeds"], negative_pooled_prompt_embeds=inp["neg_pooled"], guidance_scale=4.5,
                     noise_level=0.7, compute_log_prob=True)
    old = torch.stack([s_.log_probs[1] for s_ in samples])
    assert float((torch.exp(out.log_prob - old) - 1).abs().max()) <= 1e-5
