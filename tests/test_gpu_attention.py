"""-m gpu parity: tcgen05 flash attention over the fused token-major qkv buffer vs torch SDPA (fp32 math)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from tests.gpu_util import ptr, stream, err_report, dump, device_error
from flow_factory_b200 import _lib


def _ref(qkv, H):
    B, S, _ = qkv.shape
    D = 64 * H
    q, k, v = qkv.float().split(D, dim=2)
    sp = lambda t: t.reshape(B, S, H, 64).transpose(1, 2)
    o = torch.nn.functional.scaled_dot_product_attention(sp(q), sp(k), sp(v))
    return o.transpose(1, 2).reshape(B, S, D)


@pytest.mark.parametrize("B,S,H", [(1, 40, 1), (2, 64, 2), (1, 65, 1), (1, 128, 1), (1, 256, 2), (2, 77, 2), (1, 333, 3), (2, 589, 2), (1, 4429, 4), (2, 4096, 24)])
def test_attention_matches_sdpa(B, S, H):
    g = torch.Generator(device="cuda").manual_seed(S + H)
    qkv = torch.randn(B, S, 3 * 64 * H, device="cuda", generator=g).bfloat16()
    out = torch.full((B, S, 64 * H), float("nan"), device="cuda", dtype=torch.bfloat16)
    _lib.check(_lib.lib().ffb200_attention(ptr(qkv), B, S, H, ptr(out), stream()), "ffb200_attention")
    torch.cuda.synchronize()
    ref = _ref(qkv, H)
    rep = err_report(out.reshape(-1, 64 * H), ref.reshape(-1, 64 * H), f"attn_{B}_{S}_{H}")
    ok = rep["n_nan"] == 0 and rep["max_abs"] <= 5e-3 * max(1.0, rep["ref_absmax"])     # measured on B200: <= 4.2e-3 of that scale (bf16 P and bf16 output rounding; profiles/r02_parity_measured.jsonl)
    if not ok:
        rep["device_error"] = device_error()
        dump(f"diag_attn_{B}_{S}_{H}.json", rep)
    assert ok, rep


def test_attention_peaked_softmax_and_scale():
    """Large-magnitude q/k (after RMSNorm |q| ~ 8): exercises the running-max rescale across KV tiles."""
    B, S, H = 1, 700, 2
    g = torch.Generator(device="cuda").manual_seed(11)
    qkv = torch.randn(B, S, 3 * 64 * H, device="cuda", generator=g)
    qkv[..., : 2 * 64 * H] *= 3.0
    qkv = qkv.bfloat16()
    out = torch.empty((B, S, 64 * H), device="cuda", dtype=torch.bfloat16)
    _lib.check(_lib.lib().ffb200_attention(ptr(qkv), B, S, H, ptr(out), stream()), "ffb200_attention")
    torch.cuda.synchronize()
    ref = _ref(qkv, H)
    assert float((out.float() - ref).abs().max()) <= 3e-2


def _ref_d(qkv, H, d):
    B, S, _ = qkv.shape
    D = d * H
    q, k, v = qkv.float().split(D, dim=2)
    sp = lambda t: t.reshape(B, S, H, d).transpose(1, 2)
    o = torch.nn.functional.scaled_dot_product_attention(sp(q), sp(k), sp(v))
    return o.transpose(1, 2).reshape(B, S, D)


@pytest.mark.parametrize("B,S,H", [(1, 40, 1), (2, 64, 2), (1, 65, 1), (1, 128, 1), (2, 77, 2), (1, 333, 3), (2, 640, 2), (1, 4608, 3), (2, 2100, 24)])
def test_attention_d128_matches_sdpa(B, S, H):
    """head_dim 128 (FLUX.1, transformer_flux.py:118-125): two-panel Q/K/V tiles, N = 128 MN-major V operand."""
    from flow_factory_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(S + H)
    qkv = torch.randn(B, S, 3 * 128 * H, device="cuda", generator=g).bfloat16()
    out = torch.full((B, S, 128 * H), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.attention(qkv, H, out, head_dim=128)
    torch.cuda.synchronize()
    ref = _ref_d(qkv, H, 128)
    rep = err_report(out.reshape(-1, 128 * H), ref.reshape(-1, 128 * H), f"attn128_{B}_{S}_{H}")
    ok = rep["n_nan"] == 0 and rep["max_abs"] <= 5e-3 * max(1.0, rep["ref_absmax"])   # measured on B200: <= 3.1e-3 of that scale
    if not ok:
        rep["device_error"] = device_error()
        dump(f"diag_attn128_{B}_{S}_{H}.json", rep)
    assert ok, rep


def test_attention_d128_strided_output_and_peaked_softmax():
    """Output written as the first 128*H columns of a wider row (the single-stream block's [attn | mlp] buffer,
    transformer_flux.py:400); large-magnitude q/k exercise the running-max rescale of the 128-column accumulator."""
    from flow_factory_b200 import ops
    B, S, H = 1, 700, 2
    g = torch.Generator(device="cuda").manual_seed(11)
    qkv = torch.randn(B, S, 3 * 128 * H, device="cuda", generator=g)
    qkv[..., : 2 * 128 * H] *= 2.0
    qkv = qkv.bfloat16()
    wide = torch.full((B, S, 128 * H + 192), 7.0, device="cuda", dtype=torch.bfloat16)
    ops.attention(qkv, H, wide, head_dim=128, out_row_stride=128 * H + 192)
    torch.cuda.synchronize()
    ref = _ref_d(qkv, H, 128)
    assert float((wide[..., : 128 * H].float() - ref).abs().max()) <= 3e-2
    assert bool((wide[..., 128 * H:] == 7.0).all())        # columns beyond the head block untouched


LOG2E = 1.4426950408889634


@pytest.mark.parametrize("d,B,S,H", [(64, 1, 40, 2), (128, 1, 64, 1), (64, 2, 130, 1), (64, 2, 589, 2), (64, 1, 4429, 3), (128, 2, 640, 2), (128, 1, 2100, 3)])
def test_attention_prescaled_keys(d, B, S, H):
    """The engines' layout: k already carries softmax_scale * log2(e) (folded into the key RMSNorm multiply, ONE bf16 rounding), the
    kernel takes q.k as base-2 exponents without any per-score multiply.  Reference: fp32 softmax of q.k' in base 2."""
    from flow_factory_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(7 * S + H + d)
    q = torch.randn(B, S, H * d, device="cuda", generator=g)
    k = torch.randn(B, S, H * d, device="cuda", generator=g)
    v = torch.randn(B, S, H * d, device="cuda", generator=g)
    kp = (k * (d ** -0.5 * LOG2E)).bfloat16()                       # what the QKV epilogue stores
    qkv = torch.cat([q.bfloat16(), kp, v.bfloat16()], dim=-1).contiguous()
    out = ops.attention(qkv, H, head_dim=d, k_prescaled=True)
    torch.cuda.synchronize()
    sp = lambda t: t.float().reshape(B, S, H, d).transpose(1, 2)
    ref = torch.nn.functional.scaled_dot_product_attention(sp(q.bfloat16()), sp(kp), sp(v.bfloat16()), scale=1.0 / LOG2E)   # e^(x ln 2) = 2^x
    ref = ref.transpose(1, 2).reshape(B, S, H * d)
    assert torch.isfinite(out.float()).all()
    assert float((out.float() - ref).abs().max()) <= 2e-2
    # and the same numbers as the general path on the unscaled keys, up to the one differently-placed bf16 rounding of k
    gen = ops.attention(torch.cat([q.bfloat16(), k.bfloat16(), v.bfloat16()], dim=-1).contiguous(), H, head_dim=d)
    assert float((out.float() - gen.float()).norm() / gen.float().norm()) <= 6e-3


@pytest.mark.parametrize("d", [64, 128])
@pytest.mark.parametrize("pre", [False, True])
def test_attention_reference_shift_on_rising_scores(d, pre):
    """No per-tile row maximum (softmax.cuh): the reference starts at 0 and moves by an exact power of two when the running sum passes
    2^64.  Scores rising along the keys by ~90 nats force that path (and the accumulator rescale in TMEM) on every row."""
    from flow_factory_b200 import ops
    B, S, H = 1, 1536, 2
    g = torch.Generator(device="cuda").manual_seed(1)
    q = torch.randn(B, S, H, d, generator=g, device="cuda")
    k = torch.randn(B, S, H, d, generator=g, device="cuda")
    q[..., 0] = 8.0
    k[..., 0] = torch.linspace(0, 11.0 * d ** 0.5, S, device="cuda")[None, :, None]      # + 88 nats = 127 exponent units along the sequence
    v = torch.randn(B, S, H, d, generator=g, device="cuda")
    scale = d ** -0.5
    kk = (k * (scale * LOG2E)) if pre else k
    qkv = torch.cat([q.reshape(B, S, H * d), kk.reshape(B, S, H * d), v.reshape(B, S, H * d)], -1).bfloat16().contiguous()
    out = ops.attention(qkv, H, head_dim=d, k_prescaled=pre)
    torch.cuda.synchronize()
    qb, kb, vb = (t.reshape(B, S, H, d).transpose(1, 2).float() for t in qkv.split(H * d, dim=-1))
    ref = torch.nn.functional.scaled_dot_product_attention(qb, kb, vb, scale=(1.0 / LOG2E) if pre else scale).transpose(1, 2).reshape(B, S, H * d)
    assert torch.isfinite(out.float()).all()
    assert float((out.float() - ref).norm() / ref.norm()) <= 4e-3


def test_attention_large_first_tile_takes_a_nonzero_reference():
    """First-tile scores beyond 2^+-32: the reference is that tile's row maximum (general subtract path), not 0."""
    from flow_factory_b200 import ops
    B, S, H, d = 1, 400, 2, 64
    g = torch.Generator(device="cuda").manual_seed(3)
    q = torch.randn(B, S, H, d, generator=g, device="cuda")
    k = torch.randn(B, S, H, d, generator=g, device="cuda")
    q[..., 1] = 20.0
    k[..., 1] = 20.0                                                                        # + 50 nats on every score
    k[:, :, 1, 1] = -20.0                                                                   # head 1: - 50 nats (far below 2^-32)
    v = torch.randn(B, S, H, d, generator=g, device="cuda")
    qkv = torch.cat([q.reshape(B, S, H * d), k.reshape(B, S, H * d), v.reshape(B, S, H * d)], -1).bfloat16().contiguous()
    out = ops.attention(qkv, H, head_dim=d)
    torch.cuda.synchronize()
    qb, kb, vb = (t.reshape(B, S, H, d).transpose(1, 2).float() for t in qkv.split(H * d, dim=-1))
    ref = torch.nn.functional.scaled_dot_product_attention(qb, kb, vb).transpose(1, 2).reshape(B, S, H * d)
    assert torch.isfinite(out.float()).all()
    assert float((out.float() - ref).norm() / ref.norm()) <= 4e-3


@pytest.mark.parametrize("B,S,H", [(1, 333, 3), (2, 589, 2), (1, 4429, 2)])
@pytest.mark.parametrize("w_scale", [1.0, 2.5, 4.0])
def test_attention_normed_range_proof_is_bit_identical(B, S, H, w_scale):
    """ffb200_attention_normed: q / k heads produced by a per-head RMSNorm with weights wq / wk (attention_processor.py:1456-1473).  The
    kernel proves |q.k'| <= 64 max|wq| max|wk| scale*log2(e) and, when that is within the polynomial exp2's range, skips the per-tile
    range check: the output must equal the checked kernel's bit for bit - with weights small enough for the proof (w_scale 1, 2.5:
    bounds 12 .. 73) and with weights too large for it (w_scale 4: 188 > 120, the check stays)."""
    from flow_factory_b200 import ops
    d = 64
    g = torch.Generator(device="cuda").manual_seed(11 * S + H)
    wq = (w_scale * (1.0 + 0.1 * torch.randn(d, device="cuda", generator=g)).clamp(0.5, 1.0)).bfloat16()     # max|w| <= w_scale
    wk = (w_scale * (1.0 + 0.1 * torch.randn(d, device="cuda", generator=g)).clamp(0.5, 1.0)).bfloat16()
    wqt, wkt = (0.9 * wq.float()).bfloat16(), (0.9 * wk.float()).bfloat16()

    def rms(x, w):                                             # the QKV epilogue: bf16(bf16(x * rsqrt(mean x^2 + eps)) * w)
        xf = x.float().reshape(B, -1, H, d)
        n = (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6)).bfloat16().float()
        return (n * w.float()).reshape(B, -1, H * d)
    n_text = S // 5
    x = torch.randn(B, S, 3 * H * d, device="cuda", generator=g)
    q = torch.cat([rms(x[:, n_text:, : H * d], wq), rms(x[:, :n_text, : H * d], wqt)], dim=1).bfloat16()
    k = torch.cat([rms(x[:, n_text:, H * d: 2 * H * d], wk), rms(x[:, :n_text, H * d: 2 * H * d], wkt)], dim=1)
    kp = (k * (d ** -0.5 * LOG2E)).bfloat16()                  # keys pre-scaled, ONE rounding
    qkv = torch.cat([q, kp, x[..., 2 * H * d:].bfloat16()], dim=-1).contiguous()
    checked = ops.attention(qkv, H, head_dim=d, k_prescaled=True)
    proven = ops.attention_normed(qkv, H, wq, wk, wqt, wkt)
    torch.cuda.synchronize()
    assert torch.isfinite(proven.float()).all()
    assert torch.equal(checked, proven)
    sp = lambda t: t.float().reshape(B, S, H, d).transpose(1, 2)
    ref = torch.nn.functional.scaled_dot_product_attention(sp(q), sp(kp), sp(qkv[..., 2 * H * d:]), scale=1.0 / LOG2E).transpose(1, 2).reshape(B, S, H * d)
    assert float((proven.float() - ref).abs().max()) <= 2e-2 * max(1.0, float(ref.abs().max()))
    # the bound itself: no score of this input exceeds it
    bound = 64 * 1.016 * float(wq.float().abs().max()) * float(wk.float().abs().max()) * d ** -0.5 * LOG2E
    assert float((sp(q) @ sp(kp).transpose(-1, -2)).abs().max()) <= bound
