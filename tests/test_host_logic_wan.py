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
