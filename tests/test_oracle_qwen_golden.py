"""The Qwen-Image oracle (oracle/qwen_oracle.py; SURVEY 8f row 4, groundwork for a later engine) pinned against fixtures minted from the
REAL reference (tests/golden/make_golden.py: vendored diffusers QwenImageTransformer2DModel, CPU)."""
import os
import torch
from oracle import qwen_oracle as QO


def test_qwen_forward_fp32_masked_and_autocast(golden_dir):
    g = torch.load(os.path.join(golden_dir, "qwen_tiny.pt"), weights_only=False)
    e = g["tiny"]
    cfg = QO.tiny_qwen_config()
    w = QO.make_qwen_weights(cfg, seed=0)
    assert sorted(w.keys()) == e["keys"]                       # key parity with QwenImageTransformer2DModel.state_dict()
    B, h2, w2, nt = e["shape"]
    lat, pe = QO.make_qwen_inputs(cfg, B, h2, w2, nt, seed=1)
    with torch.no_grad():
        y = QO.qwen_forward(w, cfg, lat, pe, e["t"], (1, h2, w2))
        ym = QO.qwen_forward(w, cfg, lat, pe, e["t"], (1, h2, w2), encoder_hidden_states_mask=e["mask"])
    torch.testing.assert_close(y, e["y32"], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(ym, e["y32_masked"], rtol=1e-5, atol=1e-5)
    wb = {k: v.bfloat16() for k, v in w.items()}
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        yb = QO.qwen_forward(wb, cfg, lat.bfloat16(), pe.bfloat16(), e["t"].bfloat16(), (1, h2, w2))
    assert torch.equal(yb, e["y_bf16_cpu_autocast"])            # same ops, same order -> bit exact on CPU
    c = g["cfg"]
    torch.testing.assert_close(QO.true_cfg_combine(e["y32"], e["y32_masked"], c["gs"]), c["pred"], rtol=1e-6, atol=1e-6)
