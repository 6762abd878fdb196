"""-m gpu parity of the Qwen-Image path: the dual-stream engine in variant-1 mode vs the oracle
(oracle/qwen_oracle.py, pinned bit-exact against the reference's QwenImageTransformer2DModel on CPU)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from tests.gpu_util import dump
from flow_factory_b200.qwen import QwenRolloutEngine, qwen_rope_tables
from oracle import qwen_oracle as QO
from oracle import sd3_oracle as O


def test_qwen_rope_tables_match_the_oracle():
    cos, sin = qwen_rope_tables(6, 4, 9, (16, 56, 56))
    vid, txt = QO.qwen_rope(1, 6, 4, 9, (16, 56, 56))
    f = torch.cat([txt, vid])
    assert torch.equal(cos[:, 0::2], f.real) and torch.equal(cos[:, 1::2], f.real)
    assert torch.equal(sin[:, 0::2], f.imag) and torch.equal(sin[:, 1::2], f.imag)


def _truth(cfg, wt, lat, pe, t_eff, h2, w2):
    with torch.no_grad():
        return QO.qwen_forward(wt, cfg, lat.float(), pe.float(), torch.full((lat.shape[0],), t_eff), (1, h2, w2))


@pytest.mark.parametrize("name", ["tiny", "mid"])
def test_qwen_forward_and_true_cfg_match_oracle(name):
    cfg, (B, h2, w2, nt), seed = {"tiny": (QO.tiny_qwen_config(), (2, 6, 4, 9), 0),
                                  "mid": (QO.tiny_qwen_config(num_layers=3, heads=4, joint_dim=256), (2, 24, 16, 37), 4)}[name]
    w32 = QO.make_qwen_weights(cfg, seed=seed)
    wb = {k: v.bfloat16() for k, v in w32.items()}
    wt = {k: v.float() for k, v in wb.items()}
    lat, pe = QO.make_qwen_inputs(cfg, B, h2, w2, nt, seed=seed + 1)
    _, npe = QO.make_qwen_inputs(cfg, B, h2, w2, nt, seed=seed + 7)
    lat, pe, npe = lat.half(), pe.bfloat16(), npe.bfloat16()
    eng = QwenRolloutEngine(cfg, wb)
    t = 612.0
    t_eff = eng.t_model(t)
    # --- plain forward
    plan = eng.plan(B, h2, w2, nt)
    eng.set_prompts(plan, pe)
    got = eng.transformer_forward(plan, lat, t).float().cpu()
    torch.cuda.synchronize()
    truth = _truth(cfg, wt, lat, pe, t_eff, h2, w2)
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        ref_bf16 = QO.qwen_forward(wb, cfg, lat.bfloat16(), pe, torch.full((B,), t / 1000), (1, h2, w2)).float()
    e_eng, e_ref, scale = float((got - truth).abs().max()), float((ref_bf16 - truth).abs().max()), float(truth.abs().max())
    rep = dict(tag=f"qwen_fwd_{name}", e_engine=e_eng, e_ref_bf16=e_ref, truth_absmax=scale,
               mean_engine=float((got - truth).abs().mean()), mean_ref=float((ref_bf16 - truth).abs().mean()))
    dump(f"qwen_fwd_{name}.json", rep)
    assert torch.isfinite(got).all()
    assert e_eng <= max(1.3 * e_ref, 0.005 * scale), rep     # measured on B200: 0.60 - 0.98 x the bf16 reference's own error
    assert rep["mean_engine"] <= max(3.0 * rep["mean_ref"], 0.004 * scale), rep
    # --- true CFG with per-token norm rescale (qwen_image.py:580-587)
    gs = 4.0
    planc = eng.plan(B, h2, w2, nt, cfg=True)
    eng.set_prompts(planc, pe, npe, gs)
    gotc = eng.transformer_forward(planc, lat, t).float().cpu()
    torch.cuda.synchronize()
    truth_neg = _truth(cfg, wt, lat, npe, t_eff, h2, w2)
    truth_c = QO.true_cfg_combine(truth, truth_neg, gs)
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        ref_neg = QO.qwen_forward(wb, cfg, lat.bfloat16(), npe, torch.full((B,), t / 1000), (1, h2, w2))
        ref_c = QO.true_cfg_combine(ref_bf16.bfloat16(), ref_neg, gs).float()
    e_eng_c, e_ref_c = float((gotc - truth_c).abs().max()), float((ref_c - truth_c).abs().max())
    repc = dict(tag=f"qwen_cfg_{name}", e_engine=e_eng_c, e_ref_bf16=e_ref_c, truth_absmax=float(truth_c.abs().max()))
    dump(f"qwen_cfg_{name}.json", repc)
    assert e_eng_c <= max(1.3 * e_ref_c, 0.005 * repc["truth_absmax"]), repc


def test_qwen_ode_rollout_matches_oracle_loop():
    """DGPO-style rollout (ODE, no log-prob, true CFG): engine vs the oracle loop (fp32 truth and bf16 CPU autocast), graph == eager."""
    cfg = QO.tiny_qwen_config()
    w32 = QO.make_qwen_weights(cfg, seed=0)
    wb = {k: v.bfloat16() for k, v in w32.items()}
    B, h2, w2, nt, T, gs = 2, 6, 4, 9, 4, 4.0
    lat, pe = QO.make_qwen_inputs(cfg, B, h2, w2, nt, seed=1)
    _, npe = QO.make_qwen_inputs(cfg, B, h2, w2, nt, seed=8)
    lat, pe, npe = lat.half(), pe.bfloat16(), npe.bfloat16()
    eng = QwenRolloutEngine(cfg, wb)
    plan = eng.plan(B, h2, w2, nt, cfg=True)
    eng.set_prompts(plan, pe, npe, gs)
    ts, sig, coefs = eng.make_coefs(plan, T, 0.0, [], dynamics="ODE", store_slots=[1, 2, 3, 4])
    outs = {}
    for graph in (False, True):
        r = eng.rollout(plan, lat, coefs, n_latent_slots=T + 1, store_initial_slot=0, n_logp_slots=0, use_graph=graph)
        torch.cuda.synchronize()
        outs[graph] = r["all_latents"].cpu()
    assert torch.equal(outs[False], outs[True])

    def oracle_loop(w, autocast):
        x = lat
        xs = [x]
        for i in range(T):
            tm = torch.full((B,), float(ts[i]) / 1000)
            if autocast:
                with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
                    vp = QO.qwen_forward(w, cfg, x.bfloat16(), pe, tm, (1, h2, w2))
                    vn = QO.qwen_forward(w, cfg, x.bfloat16(), npe, tm, (1, h2, w2))
                    v = QO.true_cfg_combine(vp, vn, gs)
            else:
                with torch.no_grad():
                    tme = torch.full((B,), eng.t_model(float(ts[i])))
                    vp = QO.qwen_forward(w, cfg, x.float(), pe.float(), tme, (1, h2, w2))
                    vn = QO.qwen_forward(w, cfg, x.float(), npe.float(), tme, (1, h2, w2))
                    v = QO.true_cfg_combine(vp, vn, gs)
            r = O.sde_step(v, x, float(sig[i]), float(sig[i + 1]), 0.0, float(sig[1]), "ODE", compute_log_prob=False)
            x = O.cast_latents(r["next_latents"], torch.float16)
            xs.append(x)
        return xs
    truth = oracle_loop({k: v.float() for k, v in wb.items()}, False)
    refb = oracle_loop(wb, True)
    for i in range(T + 1):
        d = float((outs[True][:, i].float() - truth[i].float()).abs().max())
        d_ref = float((refb[i].float() - truth[i].float()).abs().max())
        assert d <= max(3.0 * d_ref, 0.05), (i, d, d_ref)


def test_qwen_adapter_inference_api():
    """B200QwenImageAdapter mirrors QwenImageAdapter.inference / forward (qwen_image.py:290-600): DGPO-style ODE rollout with true CFG,
    sample fields, refusal of padded prompts, and forward() reproducing a rollout step."""
    from flow_factory_b200.qwen_adapter import B200QwenImageAdapter
    cfg = QO.tiny_qwen_config()
    wb = {k: v.bfloat16() for k, v in QO.make_qwen_weights(cfg, seed=0).items()}
    ad = B200QwenImageAdapter(cfg, wb)
    g = torch.Generator().manual_seed(5)
    B, nt, T = 2, 9, 5
    pe = torch.randn(B, nt, cfg.joint_attention_dim, generator=g).bfloat16().cuda()
    npe = torch.randn(B, nt, cfg.joint_attention_dim, generator=g).bfloat16().cuda()
    mask = torch.ones(B, nt, dtype=torch.long).cuda()
    torch.manual_seed(3)
    samples = ad.inference(prompt=["a"] * B, negative_prompt=[""] * B, height=96, width=64, num_inference_steps=T, guidance_scale=4.0,
                           prompt_embeds=pe, prompt_embeds_mask=mask, negative_prompt_embeds=npe, negative_prompt_embeds_mask=mask,
                           compute_log_prob=False, trajectory_indices="all")
    torch.cuda.synchronize()
    s0 = samples[0]
    Ni = (96 // 16) * (64 // 16)
    assert tuple(s0.all_latents.shape) == (T + 1, Ni, 64) and s0.log_probs is None and s0.img_shapes == [(1, 6, 4)]
    assert torch.isfinite(s0.all_latents.float()).all()
    lat = torch.stack([s.all_latents for s in samples])
    out = ad.forward(t=s0.timesteps[2], t_next=s0.timesteps[3], latents=lat[:, 2], prompt_embeds=pe, prompt_embeds_mask=mask,
                     img_shapes=[[(1, 6, 4)]] * B, negative_prompt_embeds=npe, negative_prompt_embeds_mask=mask, guidance_scale=4.0,
                     compute_log_prob=False)
    torch.cuda.synchronize()
    # ODE: next_latents is the un-rounded mean; its fp16 storage cast is what the rollout kept
    assert torch.equal(out.next_latents.half(), lat[:, 3])
    # ragged prompts (list form) are right-padded and masked; a non-prefix mask is refused
    rag = ad.inference(height=96, width=64, num_inference_steps=2, guidance_scale=1.0, prompt_embeds=[pe[0], pe[1, :6]],
                       prompt_embeds_mask=[mask[0], mask[1, :6]])
    torch.cuda.synchronize()
    assert rag[1].prompt_embeds_mask.sum() == 6 and torch.isfinite(rag[1].all_latents.float()).all()
    bad = mask.clone(); bad[1, 2] = 0
    with pytest.raises(NotImplementedError):
        ad.inference(height=96, width=64, num_inference_steps=2, prompt_embeds=pe, prompt_embeds_mask=bad)


def test_qwen_key_padding_mask_matches_oracle():
    """encoder_hidden_states_mask -> joint attention mask (transformer_qwenimage.py:941-958): right-padded prompts, padded keys masked.
    Engine vs the oracle's masked forward (itself pinned against the reference with a mask, tests/golden/qwen_tiny.pt)."""
    cfg = QO.tiny_qwen_config(num_layers=2, heads=2, joint_dim=64)
    w32 = QO.make_qwen_weights(cfg, seed=2)
    wb = {k: v.bfloat16() for k, v in w32.items()}
    wt = {k: v.float() for k, v in wb.items()}
    B, h2, w2, nt = 3, 10, 12, 150          # text rows span three 64-wide KV tiles: partially and fully masked tiles
    lat, pe = QO.make_qwen_inputs(cfg, B, h2, w2, nt, seed=3)
    lat, pe = lat.half(), pe.bfloat16()
    lens = [150, 70, 5]
    mask = (torch.arange(nt)[None, :] < torch.tensor(lens)[:, None]).float()
    eng = QwenRolloutEngine(cfg, wb)
    plan = eng.plan(B, h2, w2, nt)
    eng.set_prompts(plan, pe, prompt_lengths=lens)
    t = 431.0
    got = eng.transformer_forward(plan, lat, t).float().cpu()
    torch.cuda.synchronize()
    with torch.no_grad():
        te = torch.full((B,), eng.t_model(t))
        truth = QO.qwen_forward(wt, cfg, lat.float(), pe.float(), te, (1, h2, w2), encoder_hidden_states_mask=mask)
        unmasked = QO.qwen_forward(wt, cfg, lat.float(), pe.float(), te, (1, h2, w2))
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        ref_bf16 = QO.qwen_forward(wb, cfg, lat.bfloat16(), pe, torch.full((B,), t / 1000), (1, h2, w2), encoder_hidden_states_mask=mask).float()
    e_eng, e_ref = float((got - truth).abs().max()), float((ref_bf16 - truth).abs().max())
    effect = float((unmasked - truth).abs().max())
    rep = dict(tag="qwen_masked", e_engine=e_eng, e_ref_bf16=e_ref, mask_effect=effect, truth_absmax=float(truth.abs().max()))
    dump("qwen_masked.json", rep)
    d_unmasked = float((got - unmasked).abs().max())
    rep["engine_vs_unmasked"] = d_unmasked
    dump("qwen_masked.json", rep)
    assert effect > 2.5 * e_ref, rep                       # the mask matters on this input ...
    assert e_eng <= 1.6 * e_ref + 0.005, rep               # ... the engine is as close to the MASKED truth as the bf16 reference ...
    assert d_unmasked >= 0.6 * effect, rep                 # ... and not to the unmasked one
    assert float((got[0] - unmasked[0]).abs().max()) <= 1.6 * e_ref + 0.005   # the unpadded sample is untouched
    eng.set_prompts(plan, pe)                                            # resetting the lengths restores the unmasked result
    got2 = eng.transformer_forward(plan, lat, t).float().cpu()
    assert float((got2 - unmasked).abs().max()) <= max(3.0 * e_ref, 0.02 * rep["truth_absmax"])
