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


@pytest.mark.parametrize("B,S,H", [(1, 128, 1), (1, 256, 2), (2, 77, 2), (1, 333, 3), (2, 589, 2), (1, 4429, 4), (2, 4096, 24)])
def test_attention_matches_sdpa(B, S, H):
    g = torch.Generator(device="cuda").manual_seed(S + H)
    qkv = torch.randn(B, S, 3 * 64 * H, device="cuda", generator=g).bfloat16()
    out = torch.full((B, S, 64 * H), float("nan"), device="cuda", dtype=torch.bfloat16)
    _lib.check(_lib.lib().ffb200_attention(ptr(qkv), B, S, H, ptr(out), stream()), "ffb200_attention")
    torch.cuda.synchronize()
    ref = _ref(qkv, H)
    rep = err_report(out.reshape(-1, 64 * H), ref.reshape(-1, 64 * H), f"attn_{B}_{S}_{H}")
    ok = rep["n_nan"] == 0 and rep["max_abs"] <= 2e-2
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


@pytest.mark.parametrize("B,S,H", [(1, 128, 1), (2, 77, 2), (1, 333, 3), (2, 640, 2), (1, 4608, 3), (2, 2100, 24)])
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
    ok = rep["n_nan"] == 0 and rep["max_abs"] <= 2e-2
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
