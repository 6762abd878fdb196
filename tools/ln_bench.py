"""Times ffb200_ln_modulate at the SD3.5 bench shape (16 forward samples x 4096 image tokens x 1536; single and dual output) and checks it
against a torch model of the same operation order."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flow_factory_b200 import ops

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
B, R, D = 16, 4096, 1536
g = torch.Generator(device="cuda").manual_seed(0)
x = (torch.randn(B, R, D, device="cuda", generator=g) * 2).bfloat16()
mod = (torch.randn(B, 4 * D, device="cuda", generator=g) * 0.3).bfloat16()
sh1, sc1, sh2, sc2 = (mod[:, i * D:(i + 1) * D] for i in range(4))
o1, o2 = torch.empty_like(x), torch.empty_like(x)


def timed(dual):
    fn = (lambda: ops.ln_modulate(x, sh1, sc1, o1, sh2, sc2, o2, mod_batch_stride=4 * D)) if dual else (lambda: ops.ln_modulate(x, sh1, sc1, o1, mod_batch_stride=4 * D))
    for _ in range(3):
        fn()
    ts = []
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = sorted(ts)[len(ts) // 2]
    nbytes = x.numel() * 2 * (3 if dual else 2)
    return ms, nbytes / ms / 1e9


def ref(sh, sc):
    xf = x.float()
    mean = xf.mean(-1, keepdim=True)
    var = ((xf - mean) ** 2).mean(-1, keepdim=True)
    y = (xf - mean) * torch.rsqrt(var + 1e-6)
    return (y * (1.0 + sc.float()).bfloat16().float()[:, None, :] + sh.float()[:, None, :]).bfloat16()


s_ms, s_tbs = timed(False)
d_ms, d_tbs = timed(True)
e1 = float((o1.float() - ref(sh1, sc1).float()).abs().max()); e2 = float((o2.float() - ref(sh2, sc2).float()).abs().max())
print(json.dumps({"kernel": "ln_modulate", "persistent": os.environ.get("FFB200_LN_PERSISTENT", "1") != "0", "shape": [B, R, D],
                  "single_ms": s_ms, "single_TBps": s_tbs, "dual_ms": d_ms, "dual_TBps": d_tbs, "max_abs_vs_torch_single": e1, "max_abs_vs_torch_dual": e2}))
