"""Parity of the Wan2.1 T2V engine (csrc/wan_*.cu, the cross-attention instantiation of the head-dim-128 attention kernel; SURVEY 8f
row 4) against the pinned oracle, through the C ABI.

First B200 run: round 2 (every kernel-level case and the rollout consistency green at the first attempt; the two forward-vs-oracle cases
failed on the ORACLE's rope tables living on the CPU - fixed in oracle/wan_oracle.py)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from flow_factory_b200 import wan as W                              # noqa: E402
from flow_factory_b200.scheduler import UniPCMultistepSDEScheduler  # noqa: E402
from flow_factory_b200.wan_adapter import B200Wan21Adapter          # noqa: E402
from oracle import sd3_oracle as O                                   # noqa: E402  (the checker)
from oracle import wan_oracle as WO                                  # noqa: E402

DEV = "cuda"


def _rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-12))


def _setup(layers=2, heads=2, B=2, Fr=3, H=8, Wd=12, nt=70, seed=3):
    ocfg = WO.tiny_wan_config(num_layers=layers, heads=heads)
    cfg = W.WanEngineConfig(num_layers=layers, num_attention_heads=heads, text_dim=ocfg.text_dim, ffn_dim=ocfg.ffn_dim,
                            rope_max_seq_len=ocfg.rope_max_seq_len)
    w = {k: v.bfloat16().float() for k, v in WO.make_wan_weights(ocfg, seed=seed).items()}
    lat, pe = WO.make_wan_inputs(ocfg, B, Fr, H, Wd, nt, seed=seed + 1)
    neg = torch.randn(pe.shape, generator=torch.Generator().manual_seed(seed + 2))
    return ocfg, cfg, w, lat.half(), pe.bfloat16(), neg.bfloat16()


def _oracle_pred(w, ocfg, lat16, t, pe, dtype):
    wd = {k: v.to(DEV, dtype) for k, v in w.items()}
    x = lat16.float().to(DEV)
    tt = torch.full((lat16.shape[0],), float(t), device=DEV)
    with torch.no_grad():
        if dtype == torch.float32:
            return WO.wan_forward(wd, ocfg, x, tt, pe.float().to(DEV))
        with torch.autocast("cuda", dtype=torch.bfloat16):
            return WO.wan_forward(wd, ocfg, x.bfloat16(), tt, pe.to(DEV))


# ------------------------------------------------------------------------------------------------ kernel by kernel (isolates a failure)
def test_op_patchify():
    x = torch.randn(2, 16, 3, 8, 12, generator=torch.Generator().manual_seed(0)).half().to(DEV)
    out = W.patchify(x, (1, 2, 2), reps=2)
    B, C, Fr, H, Wd = x.shape
    ref = x.reshape(B, C, Fr, 1, H // 2, 2, Wd // 2, 2).permute(0, 2, 4, 6, 1, 3, 5, 7).reshape(B * Fr * (H // 2) * (Wd // 2), C * 4).bfloat16()
    assert torch.equal(out, torch.cat([ref, ref], 0))


@pytest.mark.parametrize("D,rope", [(256, True), (1536, True), (384, False)])
def test_op_rms_rope(D, rope):
    g = torch.Generator(device=DEV).manual_seed(D)
    B, S = 2, 77
    qkv = torch.randn(B, S, 3 * D, generator=g, device=DEV).bfloat16()
    w = (1 + 0.1 * torch.randn(D, generator=g, device=DEV)).bfloat16()
    cos, sin = W.wan_rope_tables(W.WanEngineConfig(rope_max_seq_len=64), 7, 1, 11)     # 77 tokens, bf16-rounded values
    cos, sin = cos.to(DEV), sin.to(DEV)
    x = qkv.clone()
    W.rms_rope_(x, D, w, 1e-6, cos if rope else None, sin if rope else None, col=D)     # the k block of a fused q|k|v row
    k = qkv[..., D:2 * D]
    ref = torch.nn.functional.rms_norm(k.float(), (D,), w.float(), 1e-6).bfloat16()
    if rope:
        rh = ref.reshape(B, S, D // 128, 128)
        c, s_ = cos.bfloat16()[None, :, None, 0::2], sin.bfloat16()[None, :, None, 1::2]
        x1, x2 = rh[..., 0::2], rh[..., 1::2]
        o = torch.empty_like(rh)
        o[..., 0::2] = x1 * c - x2 * s_                                                # bf16 tensor arithmetic, as the reference
        o[..., 1::2] = x1 * s_ + x2 * c
        ref = o.reshape(B, S, D)
    assert torch.equal(x[..., :D], qkv[..., :D]) and torch.equal(x[..., 2 * D:], qkv[..., 2 * D:])   # neighbours untouched
    torch.testing.assert_close(x[..., D:2 * D].float(), ref.float(), rtol=1.6e-2, atol=1e-2)
    assert _rel(x[..., D:2 * D], ref) < 3e-3


def test_op_layer_norm_and_gate_residual():
    g = torch.Generator(device=DEV).manual_seed(3)
    B, S, D = 2, 100, 1536
    x = (torch.randn(B, S, D, generator=g, device=DEV) * 2 + 0.3).bfloat16()
    scale, shift = torch.randn(B, D, generator=g, device=DEV) * 0.2, torch.randn(B, D, generator=g, device=DEV) * 0.2
    ln = torch.nn.functional.layer_norm(x.float(), (D,), None, None, 1e-6)
    out = W.layer_norm(x, 1e-6, scale=scale, shift=shift)
    torch.testing.assert_close(out.float(), (ln * (1 + scale[:, None]) + shift[:, None]).bfloat16().float(), rtol=1.6e-2, atol=1e-2)
    wt, bs = (1 + 0.1 * torch.randn(D, generator=g, device=DEV)).bfloat16(), (0.1 * torch.randn(D, generator=g, device=DEV)).bfloat16()
    out = W.layer_norm(x, 1e-6, weight=wt, bias=bs)
    torch.testing.assert_close(out.float(), (ln * wt.float() + bs.float()).bfloat16().float(), rtol=1.6e-2, atol=1e-2)
    y = torch.randn(B, S, D, generator=g, device=DEV).bfloat16()
    gate = torch.randn(B, D, generator=g, device=DEV)
    h = x.clone()
    W.gate_residual_(h, y, gate)
    assert torch.equal(h, (x.float() + y.float() * gate[:, None]).bfloat16())           # one rounding, exactly the reference expression


@pytest.mark.parametrize("Sq,Skv,H", [(300, 70, 2), (256, 512, 3), (1000, 64, 1)])
def test_op_attention_cross(Sq, Skv, H):
    g = torch.Generator(device=DEV).manual_seed(Sq + Skv)
    B, D = 2, 128 * H
    q = torch.randn(B, Sq, D, generator=g, device=DEV).bfloat16()
    kv = torch.randn(B, Skv, 2 * D, generator=g, device=DEV).bfloat16()
    out = W.attention_cross(q, kv, H)
    sp = lambda t: t.reshape(B, t.shape[1], H, 128).transpose(1, 2).float()
    ref = torch.nn.functional.scaled_dot_product_attention(sp(q), sp(kv[..., :D]), sp(kv[..., D:])).transpose(1, 2).reshape(B, Sq, D)
    assert torch.isfinite(out.float()).all()
    assert _rel(out, ref) < 8e-3


# ------------------------------------------------------------------------------------------------ engine
@pytest.mark.parametrize("geom", [dict(), dict(layers=3, heads=3, B=1, Fr=2, H=16, Wd=20, nt=512)])
def test_forward_matches_oracle(geom):
    ocfg, cfg, w, lat, pe, neg = _setup(**geom)
    eng = W.WanRolloutEngine(cfg, w, device=DEV)
    B, _, Fr, H, Wd = lat.shape
    plan = eng.plan(B, Fr, H, Wd, pe.shape[1], cfg=False)
    eng.set_prompts(plan, pe)
    v = eng.transformer_forward(plan, lat, 875.0)
    truth, ref16 = _oracle_pred(w, ocfg, lat, 875.0, pe, torch.float32), _oracle_pred(w, ocfg, lat, 875.0, pe, torch.bfloat16)
    assert torch.isfinite(v.float()).all()
    e_eng, e_ref = _rel(v, truth), _rel(ref16, truth)
    assert e_eng <= 1.3 * e_ref + 5e-4, (e_eng, e_ref)
    # true CFG as one batch of 2B: u + g (c - u) in bf16 (wan2_t2v.py:526)
    planc = eng.plan(B, Fr, H, Wd, pe.shape[1], cfg=True)
    eng.set_prompts(planc, pe, neg)
    vc = eng.transformer_forward(planc, lat, 875.0, guidance_scale=5.0)
    tu = _oracle_pred(w, ocfg, lat, 875.0, neg, torch.float32)
    truth_c = tu + 5.0 * (truth - tu)
    ru = _oracle_pred(w, ocfg, lat, 875.0, neg, torch.bfloat16)
    ref_c = (ru + 5.0 * (ref16 - ru))
    assert _rel(vc, truth_c) <= 1.3 * _rel(ref_c, truth_c) + 1e-3


def test_step_and_rollout_consistency():
    ocfg, cfg, w, lat, pe, neg = _setup()
    sch = UniPCMultistepSDEScheduler(noise_level=0.7, flow_shift=3.0, num_sde_steps=2, seed=1)
    ad = B200Wan21Adapter(cfg, w, device=DEV, scheduler=sch, rng="torch", use_graph=True)
    ad.rollout()
    B, _, Fr, H, Wd = lat.shape
    T = 4
    noise = torch.randn(T, *lat.shape, generator=torch.Generator().manual_seed(9))
    kw = dict(height=H * 8, width=Wd * 8, num_frames=(Fr - 1) * 4 + 1, num_inference_steps=T, guidance_scale=5.0, prompt_embeds=pe.to(DEV),
              negative_prompt_embeds=neg.to(DEV), compute_log_prob=True, latents=lat.to(DEV), noise=noise.to(DEV))
    s_graph = ad.inference(**kw)
    ad.use_graph = False
    s_eager = ad.inference(**kw)
    for a, b in zip(s_graph, s_eager):
        assert torch.equal(a.all_latents, b.all_latents) and torch.equal(a.log_probs, b.log_probs)
    assert s_graph[0].timesteps.dtype == torch.int64 and s_graph[0].all_latents.shape[1:] == lat.shape[1:]
    # forward() reproduces rollout step 0 given the same noise: next latents bit for bit, log-prob to 1e-5
    ts = sch.set_timesteps(T)
    nl = sch.noise_level if 0 in set(sch.current_sde_steps.tolist()) else 0.0
    out = ad.forward(t=ts[0], t_next=ts[1], latents=lat.to(DEV), prompt_embeds=pe.to(DEV), negative_prompt_embeds=neg.to(DEV), guidance_scale=5.0,
                     noise_level=nl, compute_log_prob=nl > 0, noise=noise[0].to(DEV))
    i1 = int(s_graph[0].latent_index_map[1])
    if i1 >= 0:
        assert torch.equal(out.next_latents.half()[0], s_graph[0].all_latents[i1])
    # the scheduler arithmetic on the engine's own noise prediction == the oracle step (reference-pinned in test_host_logic_wan.py)
    c = sch.step_coef(ts[0], ts[1], nl)
    o = O.sde_step(out.noise_pred.cpu(), lat, c.sigma, c.sigma_prev, nl, float(sch.sigmas[1]), "Flow-SDE", noise=noise[0])
    torch.testing.assert_close(out.next_latents_mean.cpu(), o["next_latents_mean"], rtol=1e-6, atol=1e-6)
