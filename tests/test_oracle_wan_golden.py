"""The Wan2.1-T2V oracle (oracle/wan_oracle.py; SURVEY 8f row 4, groundwork for a later engine) pinned against fixtures minted from the
REAL reference (tests/golden/make_golden.py: vendored diffusers WanTransformer3DModel, CPU)."""
import os
import torch
from oracle import wan_oracle as WO


def test_wan_forward_fp32_and_autocast(golden_dir):
    e = torch.load(os.path.join(golden_dir, "wan_tiny.pt"), weights_only=False)["tiny"]
    cfg = WO.tiny_wan_config()
    w = WO.make_wan_weights(cfg, seed=0)
    assert sorted(w.keys()) == e["keys"]                       # key parity with WanTransformer3DModel.state_dict()
    B, fr, lh, lw, nt = e["shape"]
    lat, pe = WO.make_wan_inputs(cfg, B, fr, lh, lw, nt, seed=1)
    with torch.no_grad():
        y = WO.wan_forward(w, cfg, lat, e["t"], pe)
    torch.testing.assert_close(y, e["y32"], rtol=1e-5, atol=1e-5)
    wb = {k: v.bfloat16() for k, v in w.items()}
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        yb = WO.wan_forward(wb, cfg, lat.bfloat16(), e["t"], pe.bfloat16())
    assert torch.equal(yb, e["y_bf16_cpu_autocast"])            # same ops, same order -> bit exact on CPU
