"""-m gpu: the diffusers plug points of flow_factory_b200/diffusers_hooks.py on real kernels.  diffusers itself is not on the GPU box, so
the attention MODULE is a stand-in with the attributes `JointAttnProcessor2_0` reads (attention_processor.py:1429-1505); the real classes
are driven on CPU in tests/test_reference_hooks.py."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from flow_factory_b200 import diffusers_hooks as DH        # noqa: E402


def _sdpa(q, k, v, scale=None):
    return F.scaled_dot_product_attention(q.float().permute(0, 2, 1, 3), k.float().permute(0, 2, 1, 3), v.float().permute(0, 2, 1, 3),
                                          scale=scale).permute(0, 2, 1, 3)


@pytest.mark.parametrize("D,Sq,Skv,scale", [(64, 333, 333, None), (128, 640, 640, None), (128, 640, 640, 0.05), (128, 300, 77, None), (64, 589, 589, 0.2)])
def test_backend_function_matches_sdpa(D, Sq, Skv, scale):
    g = torch.Generator(device="cuda").manual_seed(D + Sq)
    B, H = 2, 3
    q = torch.randn(B, Sq, H, D, device="cuda", generator=g).bfloat16()
    k = torch.randn(B, Skv, H, D, device="cuda", generator=g).bfloat16()
    v = torch.randn(B, Skv, H, D, device="cuda", generator=g).bfloat16()
    out = DH.b200_attention_backend(q, k, v, scale=scale)
    assert out.shape == (B, Sq, H, D) and out.dtype == torch.bfloat16
    ref = _sdpa(q, k, v, scale)
    assert float((out.float() - ref).abs().max()) <= 2e-2
    assert float((out.float() - ref).norm() / ref.norm()) <= 4e-3


def test_backend_function_backward_is_sdpa_recomputation():
    B, S, H, D = 1, 200, 2, 64
    g = torch.Generator(device="cuda").manual_seed(5)
    mk = lambda: torch.randn(B, S, H, D, device="cuda", generator=g).bfloat16().requires_grad_(True)
    q, k, v = mk(), mk(), mk()
    o = DH.b200_attention_backend(q, k, v)
    o.float().square().sum().backward()
    q2, k2, v2 = (t.detach().clone().requires_grad_(True) for t in (q, k, v))
    o2 = F.scaled_dot_product_attention(q2.permute(0, 2, 1, 3), k2.permute(0, 2, 1, 3), v2.permute(0, 2, 1, 3)).permute(0, 2, 1, 3)
    o2.float().square().sum().backward()
    for a, b in ((q, q2), (k, k2), (v, v2)):
        assert float((a.grad.float() - b.grad.float()).norm() / b.grad.float().norm()) <= 2e-2      # upstream gradient differs by the bf16 forward


class _Attn(nn.Module):
    """The attributes of diffusers' `Attention` that the SD3 processors read."""

    def __init__(self, dim, heads, ctx_dim, context_pre_only, with_ctx):
        super().__init__()
        self.heads, self.context_pre_only = heads, context_pre_only
        self.to_q, self.to_k, self.to_v = nn.Linear(dim, dim), nn.Linear(dim, dim), nn.Linear(dim, dim)
        self.norm_q, self.norm_k = nn.RMSNorm(dim // heads, eps=1e-6), nn.RMSNorm(dim // heads, eps=1e-6)
        self.to_out = nn.ModuleList([nn.Linear(dim, dim), nn.Dropout(0.0)])
        if with_ctx:
            self.add_q_proj, self.add_k_proj, self.add_v_proj = nn.Linear(ctx_dim, dim), nn.Linear(ctx_dim, dim), nn.Linear(ctx_dim, dim)
            self.norm_added_q, self.norm_added_k = nn.RMSNorm(dim // heads, eps=1e-6), nn.RMSNorm(dim // heads, eps=1e-6)
            if not context_pre_only:
                self.to_add_out = nn.Linear(dim, dim)


def _reference_processor(attn, hs, ehs):
    """JointAttnProcessor2_0.__call__ restated (1442-1505), fp32 math on the bf16-rounded parameters."""
    B = hs.shape[0]
    f = lambda m, x: F.linear(x, m.weight.float(), m.bias.float())
    n = lambda m, x: F.rms_norm(x, (x.shape[-1],), m.weight.float(), m.eps)
    sp = lambda t: t.view(B, -1, attn.heads, t.shape[-1] // attn.heads).transpose(1, 2)
    q, k, v = n(attn.norm_q, sp(f(attn.to_q, hs))), n(attn.norm_k, sp(f(attn.to_k, hs))), sp(f(attn.to_v, hs))
    if ehs is not None:
        cq, ck, cv = n(attn.norm_added_q, sp(f(attn.add_q_proj, ehs))), n(attn.norm_added_k, sp(f(attn.add_k_proj, ehs))), sp(f(attn.add_v_proj, ehs))
        q, k, v = torch.cat([q, cq], 2), torch.cat([k, ck], 2), torch.cat([v, cv], 2)
    o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, -1, q.shape[1] * q.shape[3])
    if ehs is not None:
        o, c = o[:, : hs.shape[1]], o[:, hs.shape[1]:]
        if not attn.context_pre_only:
            c = f(attn.to_add_out, c)
        return f(attn.to_out[0], o), c
    return f(attn.to_out[0], o)


@pytest.mark.parametrize("with_ctx,pre_only", [(True, False), (True, True), (False, False)])
def test_sd3_processor_fused_path(with_ctx, pre_only):
    torch.manual_seed(3)
    dim, heads, B, Ni, Nt = 256, 4, 2, 300, 45
    attn = _Attn(dim, heads, dim, pre_only, with_ctx).cuda().bfloat16()
    with torch.no_grad():
        for m in attn.modules():
            if isinstance(m, nn.RMSNorm):
                m.weight.copy_(1 + 0.1 * torch.randn_like(m.weight))
    hs = torch.randn(B, Ni, dim, device="cuda").bfloat16()
    ehs = torch.randn(B, Nt, dim, device="cuda").bfloat16() if with_ctx else None
    proc = DH.B200JointAttnProcessor()
    with torch.no_grad():
        assert proc._fusable(attn, hs, ehs)
        got = proc(attn, hs, encoder_hidden_states=ehs)
        ref = _reference_processor(attn, hs.float(), None if ehs is None else ehs.float())
    got, ref = (got if isinstance(got, tuple) else (got,)), (ref if isinstance(ref, tuple) else (ref,))
    assert len(got) == len(ref)
    for a, b in zip(got, ref):
        assert a.shape == b.shape and torch.isfinite(a.float()).all()
        assert float((a.float() - b).norm() / b.norm()) <= 1.2e-2
    # weights move under the processor (optimizer step): the packed copy follows the version counters
    with torch.no_grad():
        attn.to_q.weight.mul_(0.5)
        got2 = proc(attn, hs, encoder_hidden_states=ehs)
        ref2 = _reference_processor(attn, hs.float(), None if ehs is None else ehs.float())
    a, b = (got2[0] if isinstance(got2, tuple) else got2), (ref2[0] if isinstance(ref2, tuple) else ref2)
    assert float((a.float() - b).norm() / b.norm()) <= 1.2e-2


def test_sd3_processor_autograd_path_uses_the_backend_function():
    torch.manual_seed(4)
    attn = _Attn(128, 2, 128, False, True).cuda().bfloat16()
    hs = torch.randn(1, 70, 128, device="cuda").bfloat16().requires_grad_(True)
    ehs = torch.randn(1, 9, 128, device="cuda").bfloat16()
    proc = DH.B200JointAttnProcessor()
    assert not proc._fusable(attn, hs, ehs)
    out, ctx = proc(attn, hs, encoder_hidden_states=ehs)
    (out.float().square().sum() + ctx.float().square().sum()).backward()
    assert hs.grad is not None and torch.isfinite(hs.grad.float()).all() and attn.to_q.weight.grad is not None
    ref, _ = _reference_processor(attn, hs.detach().float(), ehs.float())
    assert float((out.detach().float() - ref).norm() / ref.norm()) <= 1.2e-2
