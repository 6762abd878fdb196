"""Host logic of the Wan2.1 T2V path (flow_factory_b200/wan.py, scheduler.UniPCMultistepSDEScheduler) on CPU:
* the UniPC flow-sigma schedule and the per-step scalars against a fixture minted from the REAL reference scheduler
  (tests/golden/make_golden.py::golden_wan_schedule), and the oracle's step arithmetic against the reference step on that schedule;
* the RoPE tables against the pinned oracle;
* the packed-weight layout and the engine's dataflow (csrc/wan_engine.cu: im2col order, fused q|k|v, cached cross-attention k|v,
  modulation-vector layout, "nhwpqc" unpatchify over the [C, F*H, W] view) by running a torch model of that dataflow against the
  pinned oracle forward."""
import math
import os

import pytest
import torch
import torch.nn.functional as F

from flow_factory_b200 import wan as W
from flow_factory_b200.scheduler import UniPCMultistepSDEScheduler
from oracle import sd3_oracle as O
from oracle import wan_oracle as WO


@pytest.fixture(scope="module")
def sched_golden(golden_dir):
    return torch.load(os.path.join(golden_dir, "wan_schedule.pt"), weights_only=False)


def test_unipc_schedule_bit_exact(sched_golden):
    for T, shift in ((10, 3.0), (50, 3.0), (20, 5.0)):
        s = UniPCMultistepSDEScheduler(noise_level=0.7, num_sde_steps=2, seed=5, flow_shift=shift)
        ts = s.set_timesteps(T)
        e = sched_golden[f"T{T}_s{shift}"]
        assert ts.dtype == torch.int64 and torch.equal(ts, e["timesteps"])
        assert torch.equal(s.sigmas, e["sigmas"])
        assert torch.equal(s.current_sde_steps, e["sde"])
        assert torch.equal(s.get_noise_levels(), e["noise_levels"])
        assert s.index_for_timestep(ts[3]) == 3


@pytest.mark.parametrize("dyn", ["Flow-SDE", "Dance-SDE", "CPS", "ODE"])
def test_unipc_step_scalars_and_oracle_step(sched_golden, dyn):
    """sigma = int timestep / 1000 (not sigmas[i]); sigma_max = sigmas[1].  Scalars bit exact; the oracle's step reproduces the reference."""
    x, v = sched_golden["x"], sched_golden["v"]
    s = UniPCMultistepSDEScheduler(noise_level=0.7, flow_shift=3.0, dynamics_type=dyn)
    s.set_timesteps(10)
    for i in (0, 4, 9):
        e = sched_golden[f"{dyn}_{i}"]
        c = s.step_coef(e["t"], e["tn"], 0.7, compute_log_prob=True)
        assert c.dt == float(e["dt"].flatten()[0])
        assert c.std_dev_t == pytest.approx(float(e["std_dev_t"].flatten()[0]), rel=0, abs=0)
        o = O.sde_step(v, x, c.sigma, c.sigma_prev, 0.7, e["sigma_max"], dyn, noise=e["noise"])
        torch.testing.assert_close(o["next_latents_mean"], e["mean"], rtol=0, atol=0)
        torch.testing.assert_close(o["next_latents"], e["next_latents"], rtol=0, atol=0)
        if e["log_prob"] is not None:
            torch.testing.assert_close(o["log_prob"], e["log_prob"], rtol=1e-6, atol=1e-6)


def test_rope_tables_match_oracle():
    ocfg = WO.tiny_wan_config()
    cfg = W.WanEngineConfig(num_layers=ocfg.num_layers, num_attention_heads=ocfg.num_attention_heads, text_dim=ocfg.text_dim,
                            ffn_dim=ocfg.ffn_dim, rope_max_seq_len=ocfg.rope_max_seq_len)
    cos_o, sin_o = WO.wan_rope(ocfg, 3, 4, 5)
    cos, sin = W.wan_rope_tables(cfg, 3, 4, 5, table_dtype=torch.float32)
    assert torch.equal(cos, cos_o.reshape(60, 128)) and torch.equal(sin, sin_o.reshape(60, 128))
    cb, sb = W.wan_rope_tables(cfg, 3, 4, 5)                                   # the bf16 module's buffers
    assert torch.equal(cb, cos_o.reshape(60, 128).bfloat16().float()) and torch.equal(sb, sin_o.reshape(60, 128).bfloat16().float())
    with pytest.raises(ValueError):
        W.wan_rope_tables(cfg, 3, 4, 1000)


# ------------------------------------------------------------------------------------------------ torch model of csrc/wan_engine.cu
def _patchify(x, pt, ph, pw):
    """wan_patchify_kernel: rows (b, f', h', w'), columns (c, dt, dh, dw)."""
    B, C, Fr, H, Wd = x.shape
    x = x.reshape(B, C, Fr // pt, pt, H // ph, ph, Wd // pw, pw).permute(0, 2, 4, 6, 1, 3, 5, 7)
    return x.reshape(B, (Fr // pt) * (H // ph) * (Wd // pw), C * pt * ph * pw)


def _rms_rope(x, weight, eps, cos=None, sin=None):
    """wan_rms_rope_kernel on fp32 data (no bf16 roundings): RMSNorm over the full width, then per 128-wide head the pair rotation."""
    y = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps) * weight
    if cos is None:
        return y
    S, D = y.shape[-2], y.shape[-1]
    yh = y.reshape(*y.shape[:-1], D // 128, 128)
    x1, x2 = yh[..., 0::2], yh[..., 1::2]
    c, s = cos[:, None, 0::2], sin[:, None, 1::2]
    out = torch.empty_like(yh)
    out[..., 0::2] = x1 * c - x2 * s
    out[..., 1::2] = x1 * s + x2 * c
    return out.reshape(y.shape)


def _sdpa(q, k, v, heads):
    B, Sq, D = q.shape
    sp = lambda t: t.reshape(B, t.shape[1], heads, D // heads).transpose(1, 2)
    return F.scaled_dot_product_attention(sp(q), sp(k), sp(v)).transpose(1, 2).reshape(B, Sq, D)


def _ln(x, eps):
    return F.layer_norm(x, (x.shape[-1],), None, None, eps)


def _engine_model_forward(cfg: W.WanEngineConfig, g, layers, latents, t, prompt_embeds, cos, sin):
    D, Hh = cfg.inner_dim, cfg.num_attention_heads
    B, C, Fr, H, Wd = latents.shape
    h = _patchify(latents, *cfg.patch_size) @ g["pe_w"].t() + g["pe_b"]
    tproj = O.timestep_embedding(t, cfg.freq_dim)
    temb = F.silu(tproj @ g["t1_w"].t() + g["t1_b"]) @ g["t2_w"].t() + g["t2_b"]
    temb6 = F.silu(temb) @ g["tp_w"].t() + g["tp_b"]                                  # [B, 6 D]
    ctx = F.gelu(prompt_embeds @ g["x1_w"].t() + g["x1_b"], approximate="tanh") @ g["x2_w"].t() + g["x2_b"]
    for lw in layers:
        mod = (lw["table"][None, :] + temb6).reshape(B, 6, D)                          # wan_mod_vectors layout [b][6][D]
        kv2 = ctx @ lw["kv2_w"].t() + lw["kv2_b"]                                       # cached per prompt set
        k2, v2 = _rms_rope(kv2[..., :D], lw["norm_k2"], cfg.eps), kv2[..., D:]
        a1 = _ln(h, cfg.eps) * (1 + mod[:, 1:2]) + mod[:, 0:1]
        qkv = a1 @ lw["qkv_w"].t() + lw["qkv_b"]
        q = _rms_rope(qkv[..., :D], lw["norm_q"], cfg.eps, cos, sin)
        k = _rms_rope(qkv[..., D:2 * D], lw["norm_k"], cfg.eps, cos, sin)
        y = _sdpa(q, k, qkv[..., 2 * D:], Hh) @ lw["out_w"].t() + lw["out_b"]
        h = h + y * mod[:, 2:3]
        a1 = _ln(h, cfg.eps) * lw["norm2_w"] + lw["norm2_b"]
        q2 = _rms_rope(a1 @ lw["q2_w"].t() + lw["q2_b"], lw["norm_q2"], cfg.eps)
        h = h + (_sdpa(q2, k2, v2, Hh) @ lw["out2_w"].t() + lw["out2_b"])
        a1 = _ln(h, cfg.eps) * (1 + mod[:, 4:5]) + mod[:, 3:4]
        y = F.gelu(a1 @ lw["ff1_w"].t() + lw["ff1_b"], approximate="tanh") @ lw["ff2_w"].t() + lw["ff2_b"]
        h = h + y * mod[:, 5:6]
    fin = (g["table"].reshape(1, 2, D) + temb[:, None, :])                              # wan_final_mod layout [b][2][D]
    a1 = _ln(h, cfg.eps) * (1 + fin[:, 1:2]) + fin[:, 0:1]
    vtok = a1 @ g["proj_w"].t() + g["proj_b"]                                            # [B, S, ph*pw*C] "pqc"
    # sde_step_kernel's unpatchify over the [C, F*H, W] view: tok = (y / p) * wp + x / p, n = ((y % p) * p + x % p) * C + c
    p = cfg.patch_size[1]
    Hv, wp = Fr * H, Wd // p
    out = torch.empty(B, C, Hv, Wd)
    yy, xx = torch.meshgrid(torch.arange(Hv), torch.arange(Wd), indexing="ij")
    tok = (yy // p) * wp + xx // p
    for c in range(C):
        n = ((yy % p) * p + xx % p) * C + c
        out[:, c] = vtok[:, tok, n]
    return out.reshape(B, C, Fr, H, Wd)


def test_packed_engine_model_matches_oracle():
    ocfg = WO.tiny_wan_config(num_layers=2, heads=2)
    cfg = W.WanEngineConfig(num_layers=2, num_attention_heads=2, text_dim=ocfg.text_dim, ffn_dim=ocfg.ffn_dim, rope_max_seq_len=ocfg.rope_max_seq_len)
    w = WO.make_wan_weights(ocfg, seed=2)
    g, layers = W.pack_wan_state_dict(cfg, w)
    assert set(g) == set(W.GLOBAL_FIELDS) and all(set(l) == set(W.LAYER_FIELDS) for l in layers)
    D = cfg.inner_dim
    assert g["pe_w"].shape == (D, 64) and layers[0]["qkv_w"].shape == (3 * D, D) and layers[0]["kv2_w"].shape == (2 * D, D)
    assert layers[0]["table"].shape == (6 * D,) and g["table"].shape == (2 * D,)
    lat, pe = WO.make_wan_inputs(ocfg, 2, 3, 4, 8, 5, seed=1)
    t = torch.tensor([875.0, 875.0])
    cos, sin = W.wan_rope_tables(cfg, 3, 2, 4, table_dtype=torch.float32)
    with torch.no_grad():
        ref = WO.wan_forward(w, ocfg, lat, t, pe)
        got = _engine_model_forward(cfg, g, layers, lat, t, pe, cos, sin)
    torch.testing.assert_close(got, ref, rtol=2e-4, atol=2e-4)


def test_config_guards_and_engine_needs_cuda():
    d = dict(num_layers=30, num_attention_heads=12, attention_head_dim=128, in_channels=16, out_channels=16, text_dim=4096, freq_dim=256,
             ffn_dim=8960, patch_size=(1, 2, 2), eps=1e-6, rope_max_seq_len=1024, qk_norm="rms_norm_across_heads", cross_attn_norm=True,
             image_dim=None, added_kv_proj_dim=None)
    assert W.WanEngineConfig.from_model_config(d) == W.WanEngineConfig()
    for bad in (dict(image_dim=1280), dict(qk_norm="rms_norm"), dict(cross_attn_norm=False), dict(attention_head_dim=64)):
        with pytest.raises(NotImplementedError):
            W.WanEngineConfig.from_model_config({**d, **bad})
    with pytest.raises(NotImplementedError):
        UniPCMultistepSDEScheduler(use_dynamic_shifting=True)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            W.WanRolloutEngine(W.WanEngineConfig(), {})


# ------------------------------------------------------------------------------------------------ adapter bookkeeping with a stub engine
class _StubPlan:
    def __init__(self, batch, frames, height, width, n_text, cfg):
        self.batch, self.frames, self.height, self.width, self.n_text, self.cfg = batch, frames, height, width, n_text, cfg
        self.latent_shape = (batch, 16, frames, height, width)


class _StubEngine:
    """Stands in for WanRolloutEngine (which needs a B200): records the calls, returns tensors of the contract's shapes."""

    def __init__(self, model_config, state_dict, device):
        self.device, self.cfg, self.calls = torch.device("cpu"), model_config, []

    def plan(self, batch, frames, height, width, n_text, cfg=True):
        return _StubPlan(batch, frames, height, width, n_text, cfg)

    def set_prompts(self, plan, pe, neg=None):
        self.calls.append(("set_prompts", tuple(pe.shape), None if neg is None else tuple(neg.shape)))

    def rollout(self, plan, x0, coefs, guidance, n_lat, store_initial, n_lp, noise=None, seed=0, use_graph=True):
        self.calls.append(("rollout", len(coefs), guidance, n_lat, store_initial, n_lp, None if noise is None else tuple(noise.shape)))
        self.coefs = coefs
        B = plan.batch
        return dict(all_latents=torch.zeros(B, max(n_lat, 1), *plan.latent_shape[1:], dtype=torch.float16) if n_lat else None,
                    log_probs=torch.zeros(B, max(n_lp, 1)) if n_lp else None, final_latents=torch.zeros(plan.latent_shape, dtype=torch.float16),
                    overflow=torch.zeros(1, dtype=torch.int32))

    def step(self, plan, latents, coef, guidance, noise=None, next_latents=None, seed=0):
        self.calls.append(("step", coef.sigma, coef.sigma_prev, guidance))
        z = torch.zeros(plan.latent_shape)
        return dict(next_latents=z.half(), next_latents_mean=z, log_prob=torch.zeros(plan.batch), noise_pred=z.bfloat16(), overflow=torch.zeros(1))


def test_adapter_bookkeeping_with_stub_engine(monkeypatch):
    from flow_factory_b200 import wan_adapter as WA
    from flow_factory_b200.trajectory import compute_trajectory_indices
    monkeypatch.setattr(WA, "WanRolloutEngine", _StubEngine)
    sch = UniPCMultistepSDEScheduler(noise_level=0.7, flow_shift=3.0, num_sde_steps=2, seed=5)
    ad = WA.B200Wan21Adapter(W.WanEngineConfig(), {}, device="cpu", scheduler=sch, rng="torch")
    ad.rollout()
    assert ad.latent_shape(2, 480, 832, 81) == (2, 16, 21, 60, 104)
    T = 10
    sch.set_timesteps(T)
    idx = compute_trajectory_indices(sch.train_timesteps, T)
    pe, neg = torch.zeros(2, 12, 4096), torch.zeros(2, 12, 4096)
    out = ad.inference(prompt=["a", "b"], height=64, width=96, num_frames=9, num_inference_steps=T, guidance_scale=5.0, prompt_embeds=pe,
                       negative_prompt_embeds=neg, compute_log_prob=True, trajectory_indices=idx)
    eng = ad.engine
    assert eng.calls[0] == ("set_prompts", (2, 12, 4096), (2, 12, 4096))
    kind, n_coef, g, n_lat, store0, n_lp, nshape = eng.calls[1]
    assert (kind, n_coef, g) == ("rollout", T, 5.0) and nshape == (T, 2, 16, 3, 8, 12)
    sde = sorted(sch.current_sde_steps.tolist())
    assert n_lp == len(sde) == 2
    # per-step scalars: integer timesteps feed the model, sigma = t / 1000, noise only on the selected SDE steps
    ts = sch.timesteps
    for i, c in enumerate(eng.coefs):
        assert c.t_model == float(ts[i]) and c.sigma == pytest.approx(float(ts[i]) / 1000, rel=1e-7)
        assert (c.noise_level > 0) == (i in sde) and bool(c.compute_log_prob) == (i in sde)
        assert (c.logp_slot >= 0) == (i in sde)
    assert len(out) == 2 and out[0].timesteps.dtype == torch.int64 and out[0].prompt == "a" and out[1].prompt == "b"
    assert out[0].log_probs.shape == (n_lp,) and out[0].all_latents.shape == (n_lat, 16, 3, 8, 12)
    lm, pm = out[0].latent_index_map, out[0].log_prob_index_map
    assert lm.shape == (T + 1,) and int((lm >= 0).sum()) == n_lat and int((pm >= 0).sum()) == n_lp
    # no CFG when the scale is <= 1 or there is no negative prompt (wan2_t2v.py:487-495)
    eng.calls.clear()
    ad.inference(height=64, width=96, num_frames=9, num_inference_steps=4, guidance_scale=1.0, prompt_embeds=pe, negative_prompt_embeds=neg)
    assert eng.calls[0][2] is None and eng.calls[1][2] == 1.0
    # forward(): one step on the integer timesteps
    eng.calls.clear()
    sch.set_timesteps(T)
    with pytest.raises(ValueError, match="timestep_next"):        # the reference reads an integer timestep without t_next as a step INDEX
        ad.forward(t=sch.timesteps[3], latents=torch.zeros(2, 16, 3, 8, 12), prompt_embeds=pe, negative_prompt_embeds=neg, guidance_scale=5.0,
                   noise_level=0.7)
    eng.calls.clear()
    o = ad.forward(t=sch.timesteps[3], t_next=sch.timesteps[4], latents=torch.zeros(2, 16, 3, 8, 12), prompt_embeds=pe, negative_prompt_embeds=neg,
                   guidance_scale=5.0, noise_level=0.7, compute_log_prob=True)
    st = [c for c in eng.calls if c[0] == "step"][0]
    assert st[1] == pytest.approx(float(sch.timesteps[3]) / 1000) and st[2] == pytest.approx(float(sch.timesteps[4]) / 1000) and st[3] == 5.0
    assert o.std_dev_t.shape == (2, 1, 1, 1, 1) and o.next_latents.dtype == torch.float32
    # GRPOTrainer.evaluate (grpo.py:110-119): per-prompt CPU generators, no trajectory, no log-probs
    eng.calls.clear()
    gens = [torch.Generator().manual_seed(s) for s in (1, 2)]
    ev = ad.inference(height=64, width=96, num_frames=9, num_inference_steps=4, guidance_scale=5.0, prompt_embeds=pe, negative_prompt_embeds=neg,
                      compute_log_prob=False, trajectory_indices=None, generator=gens)
    assert eng.calls[1][3] == 0 and eng.calls[1][5] == 0                       # no latent slots, no log-prob slots
    assert ev[0].all_latents is None and ev[0].log_probs is None and ev[0].latent_index_map is None
    ad.eval()                                                                  # eval mode = diffusers' UniPC multistep solver in the reference
    with pytest.raises(NotImplementedError, match="UniPC"):
        ad.inference(height=64, width=96, num_frames=9, num_inference_steps=4, prompt_embeds=pe)
    ad.rollout()
    # per-step callbacks (GRPO-Guard): the step loop over forward(), integer timesteps, final t_next = 0
    eng.calls.clear()
    sch.set_timesteps(T)
    cb = ad.inference(height=64, width=96, num_frames=9, num_inference_steps=T, guidance_scale=5.0, prompt_embeds=pe, negative_prompt_embeds=neg,
                      compute_log_prob=True, trajectory_indices=idx, extra_call_back_kwargs=["next_latents_mean"])
    steps = [c for c in eng.calls if c[0] == "step"]
    assert len(steps) == T and steps[-1][2] == 0.0 and steps[0][1] == pytest.approx(float(sch.timesteps[0]) / 1000)
    assert cb[0].callback_index_map.shape == (T,) and cb[0].next_latents_mean.shape[1:] == (16, 3, 8, 12)
    assert cb[0].all_latents.shape[0] == len(idx) and cb[0].log_probs.shape[0] == len([i for i in sde if i in set(idx)])
    for bad in (dict(guidance_scale_2=3.0), dict(attention_kwargs={"a": 1}), dict(extra_call_back_kwargs=["prompt_embeds"])):
        with pytest.raises(NotImplementedError):
            ad.inference(height=64, width=96, num_frames=9, num_inference_steps=4, prompt_embeds=pe, **bad)
