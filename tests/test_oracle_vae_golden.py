"""The VAE-decode oracle (oracle/vae_oracle.py; SURVEY 8f row 3, groundwork for a later native decoder) pinned against a fixture minted
from the REAL reference (tests/golden/make_golden.py: vendored diffusers AutoencoderKL.decode, CPU)."""
import os
import torch
from oracle import vae_oracle as VO


def test_vae_decode_matches_reference(golden_dir):
    e = torch.load(os.path.join(golden_dir, "vae_tiny.pt"), weights_only=False)["tiny"]
    cfg = VO.tiny_vae_config()
    w = VO.make_vae_decoder_weights(cfg, seed=0)
    assert sorted(w.keys()) == e["keys"]                      # key parity with AutoencoderKL.state_dict()'s decoder.* entries
    with torch.no_grad():
        img = VO.vae_decode(w, cfg, e["lat"])
    assert tuple(img.shape) == (2, 3, 12, 20)                 # two levels -> one 2x upsample of the (6, 10) latent grid
    torch.testing.assert_close(img, e["img"], rtol=1e-5, atol=1e-5)


def test_vae_decode_bf16_autocast_bit_exact(golden_dir):
    """bf16 module under CPU autocast (the frozen VAE's dtype in the trainer): the oracle reproduces the reference bit for bit."""
    e = torch.load(os.path.join(golden_dir, "vae_tiny.pt"), weights_only=False)["tiny"]
    cfg = VO.tiny_vae_config()
    w = {k: v.bfloat16() for k, v in VO.make_vae_decoder_weights(cfg, seed=0).items()}
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        img = VO.vae_decode(w, cfg, e["lat"].to(torch.bfloat16))
    assert img.dtype == torch.bfloat16 and torch.equal(img, e["img_bf16"])


def test_vae_flop_model_full_size():
    fl = VO.vae_decode_flops(VO.sd35_vae(), 128, 128)
    assert 8e12 < fl < 13e12     # a 1024^2 decode is ~10.5 TFLOP: ~1.5 % of the 675 TFLOP rollout of one latent
