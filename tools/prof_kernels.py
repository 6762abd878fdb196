"""A few launches of one kernel at the SD3.5-medium 1024^2 shape, for `ncu --set full` captures (never a bench number)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flow_factory_b200 import ops

which = sys.argv[1] if len(sys.argv) > 1 else "gemm"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
torch.manual_seed(0)
if which in ("attn", "attn_b16"):
    B, S, H = (16 if which == "attn_b16" else 2), 4429, 24
    qkv = torch.randn(B, S, 3 * 64 * H, device="cuda").bfloat16()
    out = torch.empty(B, S, 64 * H, device="cuda", dtype=torch.bfloat16)
    for _ in range(n):
        ops.attention(qkv, H, out)
else:
    shapes = {"gemm": (8192, 4608, 1536, 3), "gemm_up": (8192, 6144, 1536, 1), "gemm_out": (8192, 1536, 1536, 2),
              "gemm_b16": (65536, 4608, 1536, 3), "gemm_up_b16": (65536, 6144, 1536, 1), "gemm_out_b16": (65536, 1536, 1536, 2)}
    M, N, K, epi = shapes[which]
    A = torch.randn(M, K, device="cuda").bfloat16()
    W = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
    b = torch.randn(N, device="cuda").bfloat16()
    o = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    gate = torch.randn(1, N, device="cuda").bfloat16()
    nq = torch.ones(64, device="cuda").bfloat16()
    kw = dict(epi=epi)
    if epi == 2: kw.update(gate=gate, gate_batch_stride=N)
    if epi == 3: kw.update(norm_q=nq, norm_k=nq, qk_dim=N // 3)
    for _ in range(n):
        ops.linear(A, W, b, o, **kw)
torch.cuda.synchronize()
