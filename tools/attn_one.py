"""One attention launch series for ncu (tools/gpu_r2_ncu.sh): head_dim ATT_D (64 | 128) at the bench shape, keys pre-scaled when ATT_PRE=1."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flow_factory_b200 import ops

d = int(os.environ.get("ATT_D", "64"))
pre = os.environ.get("ATT_PRE", "1") == "1"
B, S, H = (int(os.environ.get("ATT_B", "8")), 4429, 24) if d == 64 else (int(os.environ.get("ATT_B", "2")), 4608, 24)
torch.manual_seed(0)
qkv = torch.randn(B, S, 3 * d * H, device="cuda")
if pre:
    qkv[..., d * H: 2 * d * H] *= d ** -0.5 * 1.4426950408889634
qkv = qkv.bfloat16()
out = torch.empty(B, S, d * H, device="cuda", dtype=torch.bfloat16)
normed = os.environ.get("ATT_NORMED", "0") == "1" and d == 64      # the engine's launch: RMS-normed heads + the range proof (ffb200_attention_normed)
if normed:
    x = torch.randn(B, S, 3, H, d, device="cuda")
    x[:, :, :2] = x[:, :, :2] * torch.rsqrt(x[:, :, :2].pow(2).mean(-1, keepdim=True) + 1e-6)
    x[:, :, 1] *= d ** -0.5 * 1.4426950408889634
    qkv = x.reshape(B, S, 3 * d * H).bfloat16()
    w = torch.ones(d, device="cuda").bfloat16()
for _ in range(4):
    if normed:
        ops.attention_normed(qkv, H, w, w, out=out)
    else:
        ops.attention(qkv, H, out, head_dim=d, k_prescaled=pre)
torch.cuda.synchronize()
print("done", d, pre, float(out.float().abs().mean()))
