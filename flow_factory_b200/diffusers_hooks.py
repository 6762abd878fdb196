"""The finer-grained plug points of SURVEY.md 8(b): the B200 attention kernels behind diffusers' own attention hooks, so that a
reference transformer that is NOT replaced as a whole (the autograd replay in `optimize()`, a model family without a native engine)
still runs its attention on the tcgen05 kernels.

(i)  `install_attention_backend(name)` - overwrites one slot of diffusers' closed backend registry
     (`_AttentionBackendRegistry._backends[AttentionBackendName(name)]` + `_constraints` + `_supported_arg_names`,
     DF/models/attention_dispatch.py:275-293) with `b200_attention_backend`, whose signature is the registry's
     `(query, key, value, attn_mask, dropout_p, is_causal, scale, enable_gqa, return_lse, _parallel_config)` over (B, S, H, D) tensors
     (`_native_attention`, attention_dispatch.py:2930-2945).  Flow-Factory reaches it unchanged through
     `model.attn_backend: "<name>"` -> `BaseAdapter._set_attention_backend` -> `transformer.set_attention_backend(name)`
     (FF/models/abc.py:782-798, DF/models/modeling_utils.py:588-645) - FLUX / Wan / Qwen-Image processors call
     `dispatch_attention_fn(..., backend=self._attention_backend)` (transformer_flux.py:118-125).
(ii) `B200JointAttnProcessor` - an SD3 attention processor with `JointAttnProcessor2_0.__call__`'s contract
     (DF/models/attention_processor.py:1429-1505; SD3's processor does not go through the backend registry), installed with
     `model.set_attn_processor(B200JointAttnProcessor())` / `Attention.set_processor` (attention_processor.py:535-553).  Under no-grad on
     CUDA bf16 the q|k|v projections + per-head RMSNorm are ONE fused GEMM (`ops.linear(EPI_QKV_RMSNORM)`) writing the token-major joint
     [image ; text] buffer the attention kernel reads with TMA - no head transposes, no concat copies.

Forward is always the native kernel (no PyTorch fallback; unsupported arguments raise).  When autograd needs a backward (the training
replay) the gradient is obtained by recomputing the same attention with torch's SDPA inside `backward` - the reference's own operator,
exactly what runs today - so the hooks can stay installed for the whole training loop.  diffusers is imported lazily: the package itself
does not depend on it.
"""
from __future__ import annotations

from typing import Any, Dict, Optional, Tuple

import torch

from . import ops

SUPPORTED_HEAD_DIMS = (64, 128)
DEFAULT_BACKEND_SLOT = "_native_flash"      # a member of the closed AttentionBackendName enum without extra package requirements


def _check_device_dtype(query, key, value) -> None:
    if not (query.is_cuda and key.is_cuda and value.is_cuda):
        raise RuntimeError("b200 attention backend: CUDA tensors only (there is no CPU path)")
    if query.dtype != torch.bfloat16 or key.dtype != torch.bfloat16 or value.dtype != torch.bfloat16:
        raise NotImplementedError(f"b200 attention backend: bf16 only (got {query.dtype}); run the model under bf16 autocast / weights")


def _check_supported(query, key, value, attn_mask, dropout_p, is_causal, enable_gqa, return_lse, _parallel_config) -> None:
    _check_device_dtype(query, key, value)
    if query.dim() != 4 or key.shape != value.shape or query.shape[0] != key.shape[0] or query.shape[3] != key.shape[3]:
        raise ValueError("b200 attention backend: expected (B, S, H, D) query / key / value")
    if query.shape[3] not in SUPPORTED_HEAD_DIMS:
        raise NotImplementedError(f"b200 attention backend: head_dim {query.shape[3]} (supported: {SUPPORTED_HEAD_DIMS})")
    if query.shape[2] != key.shape[2] or enable_gqa:
        raise NotImplementedError("b200 attention backend: grouped-query attention is not implemented")
    if attn_mask is not None:
        raise NotImplementedError("b200 attention backend: attn_mask is not implemented (the native Qwen-Image engine masks padded text keys "
                                  "inside the kernel; through this hook pass unpadded sequences)")
    if dropout_p != 0.0 or is_causal:
        raise NotImplementedError("b200 attention backend: dropout / causal masking are not implemented (the DiT paths use neither)")
    if return_lse:
        raise ValueError("b200 attention backend does not support return_lse=True")
    if _parallel_config is not None:
        raise NotImplementedError("b200 attention backend: context parallelism is not implemented")
    if query.shape[1] != key.shape[1] and query.shape[3] != 128:
        raise NotImplementedError("b200 attention backend: cross-attention (different q / kv lengths) needs head_dim 128")


def _forward_native(query: torch.Tensor, key: torch.Tensor, value: torch.Tensor, scale: Optional[float]) -> torch.Tensor:
    B, Sq, H, D = query.shape
    Skv = key.shape[1]
    if Sq == Skv:
        # the kernels read one token-major [B, S, 3*H*D] buffer with TMA (a copy the fused engines never make: their QKV GEMM writes it)
        qkv = torch.cat([query.reshape(B, Sq, H * D), key.reshape(B, Skv, H * D), value.reshape(B, Skv, H * D)], dim=-1)
        out = ops.attention(qkv, H, head_dim=D, scale=scale)
    else:
        if scale is not None and abs(scale - D ** -0.5) > 1e-7:
            raise NotImplementedError("b200 attention backend: a custom scale with cross-attention is not implemented")
        from .wan import attention_cross
        kv = torch.cat([key.reshape(B, Skv, H * D), value.reshape(B, Skv, H * D)], dim=-1)
        out = attention_cross(query.reshape(B, Sq, H * D).contiguous(), kv, H)
    return out.view(B, Sq, H, D)


class _B200Attention(torch.autograd.Function):
    """Native forward; backward by recomputation through torch SDPA (the operator the reference trains with)."""

    @staticmethod
    def forward(ctx, query, key, value, scale):
        ctx.save_for_backward(query, key, value)
        ctx.scale = scale
        return _forward_native(query.detach(), key.detach(), value.detach(), scale)

    @staticmethod
    def backward(ctx, grad_out):
        query, key, value = ctx.saved_tensors
        with torch.enable_grad():
            q, k, v = (t.detach().requires_grad_(True) for t in (query, key, value))
            o = torch.nn.functional.scaled_dot_product_attention(q.permute(0, 2, 1, 3), k.permute(0, 2, 1, 3), v.permute(0, 2, 1, 3),
                                                                 dropout_p=0.0, is_causal=False, scale=ctx.scale).permute(0, 2, 1, 3)
            gq, gk, gv = torch.autograd.grad(o, (q, k, v), grad_out)
        return gq, gk, gv, None


def b200_attention_backend(query: torch.Tensor, key: torch.Tensor, value: torch.Tensor, attn_mask: Optional[torch.Tensor] = None,
                           dropout_p: float = 0.0, is_causal: bool = False, scale: Optional[float] = None, enable_gqa: bool = False,
                           return_lse: bool = False, _parallel_config: Any = None) -> torch.Tensor:
    """diffusers attention-backend function: (B, S, H, D) bf16 in, (B, S, H, D) out."""
    _check_supported(query, key, value, attn_mask, dropout_p, is_causal, enable_gqa, return_lse, _parallel_config)
    if torch.is_grad_enabled() and (query.requires_grad or key.requires_grad or value.requires_grad):
        return _B200Attention.apply(query, key, value, scale)
    return _forward_native(query, key, value, scale)


_installed: Dict[str, Tuple[Any, Any, Any]] = {}


def install_attention_backend(name: str = DEFAULT_BACKEND_SLOT):
    """Route diffusers' attention backend `name` to the B200 kernels.  Returns the AttentionBackendName member; then
    `transformer.set_attention_backend(name)` (what Flow-Factory's `model.attn_backend: <name>` does) activates it."""
    import inspect

    from diffusers.models.attention_dispatch import AttentionBackendName, _AttentionBackendRegistry as R
    member = AttentionBackendName(name)
    if name not in _installed:
        _installed[name] = (R._backends.get(member), R._constraints.get(member), R._supported_arg_names.get(member))
    R._backends[member] = b200_attention_backend
    R._constraints[member] = []                                     # the function validates its own arguments (and raises)
    R._supported_arg_names[member] = set(inspect.signature(b200_attention_backend).parameters.keys())
    return member


def uninstall_attention_backend(name: str = DEFAULT_BACKEND_SLOT) -> None:
    """Put the original function of slot `name` back."""
    from diffusers.models.attention_dispatch import AttentionBackendName, _AttentionBackendRegistry as R
    if name not in _installed:
        return
    member = AttentionBackendName(name)
    fn, cons, names = _installed.pop(name)
    for table, old in ((R._backends, fn), (R._constraints, cons), (R._supported_arg_names, names)):
        if old is None:
            table.pop(member, None)
        else:
            table[member] = old


class B200JointAttnProcessor:
    """SD3 / SD3.5 joint attention (`JointAttnProcessor2_0`, attention_processor.py:1429-1505) on the B200 kernels.

    Fused path (no grad, CUDA, bf16 module weights, head_dim 64, rms qk-norm): per stream ONE GEMM computes q|k|v, applies the per-head
    RMSNorm in its epilogue and writes rows [0, Ni) / [Ni, Ni + Nt) of the joint token-major buffer - image tokens first, as the
    reference concatenates them (1480-1482) - then the attention kernel, then `to_out` / `to_add_out` as the module's own Linear.
    Anything else (autograd, other head dims, no qk-norm) takes the same projections as the reference and calls the backend function."""

    def __init__(self):
        self._packed: Dict[int, Tuple[Tuple, Dict[str, torch.Tensor]]] = {}

    # -- fused weights, rebuilt when the module's parameters were written to (optimizer / EMA swaps bump the version counters)
    def _pack(self, attn, ctx: bool) -> Dict[str, torch.Tensor]:
        mods = [attn.to_q, attn.to_k, attn.to_v] + ([attn.add_q_proj, attn.add_k_proj, attn.add_v_proj] if ctx else [])
        key = tuple((m.weight.data_ptr(), m.weight._version, None if m.bias is None else m.bias._version) for m in mods)
        hit = self._packed.get(id(attn))
        if hit is not None and hit[0] == key:
            return hit[1]
        cat = lambda ms: torch.cat([m.weight.detach() for m in ms], dim=0).to(torch.bfloat16).contiguous()
        catb = lambda ms: torch.cat([(m.bias.detach() if m.bias is not None else torch.zeros(m.out_features, device=m.weight.device))
                                     for m in ms], dim=0).to(torch.bfloat16).contiguous()
        p = {"w": cat(mods[:3]), "b": catb(mods[:3]),
             "nq": attn.norm_q.weight.detach().to(torch.bfloat16).contiguous(), "nk": attn.norm_k.weight.detach().to(torch.bfloat16).contiguous()}
        if ctx:
            p.update({"cw": cat(mods[3:]), "cb": catb(mods[3:]),
                      "cnq": attn.norm_added_q.weight.detach().to(torch.bfloat16).contiguous(),
                      "cnk": attn.norm_added_k.weight.detach().to(torch.bfloat16).contiguous()})
        self._packed[id(attn)] = (key, p)
        return p

    @staticmethod
    def _fusable(attn, hidden_states, encoder_hidden_states) -> bool:
        if torch.is_grad_enabled() and (hidden_states.requires_grad or any(p.requires_grad for p in attn.to_q.parameters())):
            return False
        if not hidden_states.is_cuda or attn.to_q.weight.dtype != torch.bfloat16:
            return False
        inner = attn.to_q.out_features
        if inner // attn.heads != 64 or inner % 64 or attn.to_q.in_features % 8:
            return False
        norms = [attn.norm_q, attn.norm_k] + ([attn.norm_added_q, attn.norm_added_k] if encoder_hidden_states is not None else [])
        return all(n is not None and getattr(n, "weight", None) is not None and type(n).__name__ == "RMSNorm" for n in norms)

    def __call__(self, attn, hidden_states: torch.Tensor, encoder_hidden_states: Optional[torch.Tensor] = None,
                 attention_mask: Optional[torch.Tensor] = None, *args, **kwargs) -> torch.Tensor:
        if attention_mask is not None:
            raise NotImplementedError("B200JointAttnProcessor: attention_mask is not implemented (SD3 passes none)")
        B, Ni = hidden_states.shape[0], hidden_states.shape[1]
        ctx = encoder_hidden_states is not None
        heads = attn.heads
        if self._fusable(attn, hidden_states, encoder_hidden_states):
            w = self._pack(attn, ctx)
            D = attn.to_q.out_features
            Nt = encoder_hidden_states.shape[1] if ctx else 0
            S = Ni + Nt
            eps = float(getattr(attn.norm_q, "eps", 1e-6))
            qkv = torch.empty((B, S, 3 * D), dtype=torch.bfloat16, device=hidden_states.device)
            x = hidden_states.to(torch.bfloat16).contiguous()
            ops.linear(x, w["w"], w["b"], qkv, num_batch=B, rows_per_batch=Ni, a_batch_stride=Ni * x.shape[-1],
                       out_batch_stride=S * 3 * D, out_row_offset=0, epi=ops.EPI_QKV_RMSNORM, norm_q=w["nq"], norm_k=w["nk"], qk_dim=D, eps=eps)
            if ctx:
                c = encoder_hidden_states.to(torch.bfloat16).contiguous()
                ops.linear(c, w["cw"], w["cb"], qkv, num_batch=B, rows_per_batch=Nt, a_batch_stride=Nt * c.shape[-1],
                           out_batch_stride=S * 3 * D, out_row_offset=Ni, epi=ops.EPI_QKV_RMSNORM, norm_q=w["cnq"], norm_k=w["cnk"],
                           qk_dim=D, eps=eps)
            out = ops.attention(qkv, heads, head_dim=64).to(hidden_states.dtype)
        else:
            head_dim = attn.to_q.out_features // heads
            split = lambda t: t.view(B, -1, heads, head_dim)
            q, k, v = split(attn.to_q(hidden_states)), split(attn.to_k(hidden_states)), split(attn.to_v(hidden_states))
            if attn.norm_q is not None:
                q = attn.norm_q(q)
            if attn.norm_k is not None:
                k = attn.norm_k(k)
            if ctx:
                cq, ck, cv = (split(attn.add_q_proj(encoder_hidden_states)), split(attn.add_k_proj(encoder_hidden_states)),
                              split(attn.add_v_proj(encoder_hidden_states)))
                if attn.norm_added_q is not None:
                    cq = attn.norm_added_q(cq)
                if attn.norm_added_k is not None:
                    ck = attn.norm_added_k(ck)
                q, k, v = torch.cat([q, cq], dim=1), torch.cat([k, ck], dim=1), torch.cat([v, cv], dim=1)
            dt = q.dtype
            out = b200_attention_backend(q.to(torch.bfloat16), k.to(torch.bfloat16), v.to(torch.bfloat16)).reshape(B, -1, heads * head_dim).to(dt)
        if ctx:
            out, ctx_out = out[:, :Ni], out[:, Ni:]
            if not attn.context_pre_only:
                ctx_out = attn.to_add_out(ctx_out)
        out = attn.to_out[1](attn.to_out[0](out))
        return (out, ctx_out) if ctx else out


def install_sd3_attn_processor(transformer) -> "B200JointAttnProcessor":
    """`transformer.set_attn_processor(B200JointAttnProcessor())` on an SD3Transformer2DModel (one shared processor object: its packed
    weight cache is keyed per Attention module)."""
    proc = B200JointAttnProcessor()
    transformer.set_attn_processor(proc)
    return proc
