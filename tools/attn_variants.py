"""Times ffb200_attention (B=8 S=4429 H=24) for the library named by FFB200_LIB.  Developer experiment aid."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flow_factory_b200 import ops
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
B, S, H = 8, 4429, 24
torch.manual_seed(0)
qkv = torch.randn(B, S, 3 * 64 * H, device="cuda").bfloat16()
out = torch.empty(B, S, 64 * H, device="cuda", dtype=torch.bfloat16)
ts = []
for i in range(9):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); ops.attention(qkv, H, out); b.record(); torch.cuda.synchronize()
    if i >= 2: ts.append(a.elapsed_time(b))
ts.sort()
print(json.dumps({"lib": os.path.basename(os.environ.get("FFB200_LIB", "libffb200.so")), "ms": ts[len(ts) // 2]}))
