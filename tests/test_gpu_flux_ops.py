"""-m gpu parity of the FLUX.1 building blocks against the oracle's torch restatement:
fused q|k|v + RMSNorm + RoPE GEMM epilogue (head_dim 128), LayerNorm-modulate at D = 3072, and the joint attention chain."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from tests.gpu_util import err_report, dump, device_error
from flow_factory_b200 import ops
from oracle import flux_oracle as FO


def _qkv_rope_ref(x, W, b, nq, nk, cos, sin, H, eps=1e-6):
    """to_q/to_k/to_v -> unflatten -> torch.nn.RMSNorm -> apply_rotary_emb (transformer_flux.py:87-117), fp32 math on the
    bf16-rounded tensors exactly where the reference holds bf16 tensors."""
    D = W.shape[0] // 3
    y = torch.nn.functional.linear(x.float(), W.float(), b.float()).bfloat16()            # nn.Linear output (bf16)
    q, k, v = y.split(D, dim=-1)
    out = []
    for t, w in ((q, nq), (k, nk)):
        t = t.unflatten(-1, (H, 128))
        t = torch.nn.functional.rms_norm(t.float(), (128,), w.float(), eps).bfloat16()   # bf16((x*rs)*w)
        out.append(FO.apply_rope(t, cos, sin).flatten(-2))
    out.append(v)
    return torch.cat(out, dim=-1)


@pytest.mark.parametrize("B,S,H,K,off", [(1, 128, 1, 128, 0), (2, 333, 2, 256, 5), (1, 1000, 3, 384, 512)])
def test_qkv_rmsnorm_rope_epilogue(B, S, H, K, off):
    g = torch.Generator(device="cuda").manual_seed(S + H)
    D = 128 * H
    x = torch.randn(B, S, K, device="cuda", generator=g).bfloat16()
    W = (torch.randn(3 * D, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    b = (0.1 * torch.randn(3 * D, device="cuda", generator=g)).bfloat16()
    nq = (1 + 0.1 * torch.randn(128, device="cuda", generator=g)).bfloat16()
    nk = (1 + 0.1 * torch.randn(128, device="cuda", generator=g)).bfloat16()
    ids = torch.zeros(off + S, 3)
    ids[:, 1] = torch.arange(off + S) // 7
    ids[:, 2] = torch.arange(off + S) % 13
    cos, sin = FO.rope_tables(ids, (16, 56, 56))
    cos, sin = cos.cuda().contiguous(), sin.cuda().contiguous()
    # the GEMM writes rows [off, off+S) of a joint [B, off+S, 3D] buffer and reads rope rows [off, off+S)
    out = torch.full((B, off + S, 3 * D), 3.0, device="cuda", dtype=torch.bfloat16)
    ops.linear_qkv_rope(x, W, b, out, nq, nk, cos, sin, num_batch=B, rows_per_batch=S, out_batch_stride=(off + S) * 3 * D,
                        out_row_offset=off, rope_row_offset=off)
    torch.cuda.synchronize()
    ref = _qkv_rope_ref(x, W, b, nq, nk, cos[off:], sin[off:], H)
    got = out[:, off:]
    rep = err_report(got.reshape(-1, 3 * D), ref.reshape(-1, 3 * D).float(), f"qkv_rope_{B}_{S}_{H}")
    ok = rep["n_nan"] == 0 and rep["max_abs"] <= 2.1e-2 and rep.get("mean_abs", 0) <= 2e-3
    if not ok:
        rep["device_error"] = device_error()
        dump(f"diag_qkv_rope_{B}_{S}_{H}.json", rep)
    assert ok, rep
    assert bool((out[:, :off] == 3.0).all())                       # rows before the offset untouched
    # v block: plain Linear, bit-exact against the bf16-rounded fp32 product is not guaranteed (accumulation order) -> tolerance only


def test_ln_modulate_d3072():
    g = torch.Generator(device="cuda").manual_seed(3)
    B, R, D = 2, 300, 3072
    x = (2 * torch.randn(B, R, D, device="cuda", generator=g) + 0.3).bfloat16()
    mod = (0.2 * torch.randn(B, 2 * D, device="cuda", generator=g)).bfloat16()
    out = torch.empty_like(x)
    ops.ln_modulate(x, mod[:, :D], mod[:, D:], out, mod_batch_stride=2 * D)
    torch.cuda.synchronize()
    xf = x.float()
    y = torch.nn.functional.layer_norm(xf, (D,), None, None, 1e-6)
    ref = y * (1 + mod[:, D:].float())[:, None].bfloat16().float() + mod[:, :D].float()[:, None]
    assert float((out.float() - ref).abs().max()) <= 4e-2


def test_flux_joint_attention_chain_matches_oracle():
    """One FluxAttnProcessor call (transformer_flux.py:83-139) for a dual-stream block: two fused qkv+RMSNorm+RoPE GEMMs writing
    [text ; image] rows of one buffer, head_dim-128 attention, vs the oracle under fp32 math on the same bf16 weights."""
    cfg = FO.tiny_flux_config(heads=2)
    w = {k: v.cuda() for k, v in FO.make_flux_weights(cfg, seed=3, dtype=torch.bfloat16).items()}
    B, nt, ni, D, H = 2, 37, 200, cfg.inner_dim, cfg.num_attention_heads
    g = torch.Generator(device="cuda").manual_seed(9)
    hs = torch.randn(B, ni, D, device="cuda", generator=g).bfloat16()
    ehs = torch.randn(B, nt, D, device="cuda", generator=g).bfloat16()
    ids = torch.zeros(nt + ni, 3); ids[nt:, 1] = torch.arange(ni) // 20; ids[nt:, 2] = torch.arange(ni) % 20
    cos, sin = FO.rope_tables(ids, cfg.axes_dims_rope)
    cos, sin = cos.cuda().contiguous(), sin.cuda().contiguous()
    pre = "transformer_blocks.0.attn."
    cat = lambda names, suf: torch.cat([w[pre + n + suf] for n in names]).contiguous()
    S = nt + ni
    qkv = torch.empty(B, S, 3 * D, device="cuda", dtype=torch.bfloat16)
    ops.linear_qkv_rope(ehs, cat(("add_q_proj", "add_k_proj", "add_v_proj"), ".weight"), cat(("add_q_proj", "add_k_proj", "add_v_proj"), ".bias"),
                        qkv, w[pre + "norm_added_q.weight"], w[pre + "norm_added_k.weight"], cos, sin, num_batch=B, rows_per_batch=nt,
                        out_batch_stride=S * 3 * D, out_row_offset=0, rope_row_offset=0)
    ops.linear_qkv_rope(hs, cat(("to_q", "to_k", "to_v"), ".weight"), cat(("to_q", "to_k", "to_v"), ".bias"),
                        qkv, w[pre + "norm_q.weight"], w[pre + "norm_k.weight"], cos, sin, num_batch=B, rows_per_batch=ni,
                        out_batch_stride=S * 3 * D, out_row_offset=nt, rope_row_offset=nt)
    att = ops.attention(qkv, H, head_dim=128)
    torch.cuda.synchronize()
    # oracle: same weights, fp32 math (no autocast), before the output projections
    w32 = {k: v.float() for k, v in w.items()}
    import oracle.flux_oracle as M
    captured = {}
    orig = M._linear
    def spy(ww, name, x):
        if name.endswith("to_out.0") or name.endswith("to_add_out"):
            captured[name] = x
        return orig(ww, name, x)
    M._linear = spy
    try:
        with torch.no_grad():
            M._flux_attention(w32, pre, cfg, hs.float(), ehs.float(), cos, sin, eps=1e-6)
    finally:
        M._linear = orig
    ref = torch.cat([captured[pre + "to_add_out"], captured[pre + "to_out.0"]], dim=1)
    rep = err_report(att.reshape(-1, D), ref.reshape(-1, D), "flux_attn_chain")
    assert rep["n_nan"] == 0 and rep["max_abs"] <= 5.2e-3, rep
