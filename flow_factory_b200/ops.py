"""Op-level Python entry points over the C ABI (one reference call site each; see include/ffb200.h).
Tensors are torch CUDA tensors (device memory + stream plumbing); the kernels are ours; no fallback."""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib

EPI_BIAS, EPI_BIAS_GELU, EPI_GATE_RESIDUAL, EPI_QKV_RMSNORM, EPI_BIAS_ADD_ROWTABLE = range(5)


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def linear(A, W, bias, out, *, num_batch=1, rows_per_batch=None, a_batch_stride=0, out_batch_stride=0, out_row_offset=0,
           epi=EPI_BIAS, gate=None, gate_batch_stride=0, norm_q=None, norm_k=None, qk_dim=0, eps=1e-6, row_table=None) -> None:
    """out = epilogue(A @ W^T + bias) on the tcgen05 GEMM (nn.Linear call sites of the MMDiT block)."""
    K, N = W.shape[1], W.shape[0]
    rows_per_batch = rows_per_batch if rows_per_batch is not None else A.numel() // K // num_batch
    code = _lib.lib().ffb200_linear(_ptr(A), num_batch, rows_per_batch, a_batch_stride, A.stride(-2), K, _ptr(W), N, _ptr(bias),
                                    _ptr(out), out_batch_stride, out_row_offset, out.stride(-2), epi, _ptr(gate),
                                    gate_batch_stride, _ptr(norm_q), _ptr(norm_k), qk_dim, eps, _ptr(row_table), _stream(A))
    _lib.check(code, "ffb200_linear")


def linear_qkv_rope(A, W, bias, out, norm_q, norm_k, rope_cos, rope_sin, *, num_batch=1, rows_per_batch=None, a_batch_stride=0,
                    out_batch_stride=0, out_row_offset=0, rope_row_offset=0, eps=1e-6) -> None:
    """FLUX.1 fused q|k|v projection (head_dim 128): Linear + RMSNorm(q, k heads) + interleaved-pair RoPE, token-major into `out`
    [.., S, 3*D] at row `out_row_offset`; rope tables fp32 [tokens, 128], GEMM row r uses table row rope_row_offset + r."""
    K, N = W.shape[1], W.shape[0]
    rows_per_batch = rows_per_batch if rows_per_batch is not None else A.numel() // K // num_batch
    code = _lib.lib().ffb200_linear_qkv_rope(_ptr(A), num_batch, rows_per_batch, a_batch_stride, A.stride(-2), K, _ptr(W), N,
                                             _ptr(bias), _ptr(out), out_batch_stride, out_row_offset, out.stride(-2), _ptr(norm_q),
                                             _ptr(norm_k), N // 3, eps, _ptr(rope_cos), _ptr(rope_sin), rope_row_offset, _stream(A))
    _lib.check(code, "ffb200_linear_qkv_rope")


def attention(qkv: torch.Tensor, num_heads: int, out: Optional[torch.Tensor] = None, head_dim: int = 64,
              out_row_stride: int = 0, scale: Optional[float] = None, k_prescaled: bool = False) -> torch.Tensor:
    """Joint attention over a fused token-major qkv buffer bf16 [B, S, 3*head_dim*H] -> bf16 [B, S, head_dim*H].
    head_dim 128 (FLUX.1) may write into a wider row (`out` [B, S, out_row_stride], columns [0, head_dim*H)).
    `scale`: softmax scale (default 1/sqrt(head_dim)); `k_prescaled`: the keys already carry scale * log2(e) (the engines' layout)."""
    B, S, _ = qkv.shape
    if out is None:
        out = torch.empty((B, S, out_row_stride or head_dim * num_heads), dtype=torch.bfloat16, device=qkv.device)
    if scale is not None or k_prescaled:
        _lib.check(_lib.lib().ffb200_attention_scaled(_ptr(qkv), B, S, num_heads, head_dim, _ptr(out), out_row_stride,
                                                      float(scale) if scale is not None else 0.0, int(k_prescaled), _stream(qkv)),
                   "ffb200_attention_scaled")
    elif head_dim == 64 and out_row_stride == 0:
        _lib.check(_lib.lib().ffb200_attention(_ptr(qkv), B, S, num_heads, _ptr(out), _stream(qkv)), "ffb200_attention")
    else:
        _lib.check(_lib.lib().ffb200_attention_ex(_ptr(qkv), B, S, num_heads, head_dim, _ptr(out), out_row_stride, _stream(qkv)),
                   "ffb200_attention_ex")
    return out


def attention_normed(qkv: torch.Tensor, num_heads: int, wq: torch.Tensor, wk: torch.Tensor, wq_text: Optional[torch.Tensor] = None,
                     wk_text: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """head_dim 64 attention over pre-scaled keys whose q / k heads came out of a per-head RMSNorm with the bf16 weights `wq` / `wk` [64]
    (and `wq_text` / `wk_text` for the text rows of a joint attention): the call the SD3.5 engine makes (ffb200_attention_normed)."""
    B, S, _ = qkv.shape
    if out is None:
        out = torch.empty((B, S, 64 * num_heads), dtype=torch.bfloat16, device=qkv.device)
    _lib.check(_lib.lib().ffb200_attention_normed(_ptr(qkv), B, S, num_heads, _ptr(out), _ptr(wq), _ptr(wk), _ptr(wq_text), _ptr(wk_text),
                                                  _stream(qkv)), "ffb200_attention_normed")
    return out


def ln_modulate(x, shift1, scale1, out1, shift2=None, scale2=None, out2=None, *, mod_batch_stride, eps=1e-6) -> None:
    B, R, D = x.shape
    _lib.check(_lib.lib().ffb200_ln_modulate(_ptr(x), B, R, D, eps, _ptr(shift1), _ptr(scale1), _ptr(out1), _ptr(shift2),
                                             _ptr(scale2), _ptr(out2), mod_batch_stride, _stream(x)), "ffb200_ln_modulate")


def small_linear(x, W, bias, out, *, addend=None, silu_input=False) -> None:
    B, K = x.shape
    N = W.shape[0]
    _lib.check(_lib.lib().ffb200_small_linear(_ptr(x), B, K, x.stride(0), _ptr(W), _ptr(bias), N, _ptr(out), out.stride(0),
                                              _ptr(addend), addend.stride(0) if addend is not None else 0, int(silu_input),
                                              _stream(x)), "ffb200_small_linear")
