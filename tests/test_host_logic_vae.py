"""Host logic of the native VAE decode (flow_factory_b200/vae.py) on CPU: the packed-weight layout and the weight
ORDER consumed by csrc/vae_engine.cu are checked by running a torch model of the engine's dataflow (shifted-box implicit GEMM,
fused q|k projection, V^T produced by a GEMM with W_v as the A operand, bias of V folded behind the softmax) against the pinned
oracle (oracle/vae_oracle.py) - and the C library's own weight count (a host-only entry point) against the Python one."""
import ctypes as C
import math
import os

import pytest
import torch
import torch.nn.functional as F

from flow_factory_b200 import _lib
from flow_factory_b200 import vae as V
from oracle import vae_oracle as VO


def _conv_from_packed(x, wp, bias, cout, taps):
    """x NHWC fp32 [B, H, W, Cin]; wp packed [Cout, taps * Cin_pad]: the K loop of csrc/vae_conv.cu (tap-major, zero-filled halo)."""
    B, H, W, cin = x.shape
    cp = wp.shape[1] // taps
    xp = F.pad(x, (0, cp - cin))
    cols = []
    for tap in range(taps):
        dy, dx = (tap // 3 - 1, tap % 3 - 1) if taps == 9 else (0, 0)
        sh = torch.zeros_like(xp)
        hs, he = max(0, -dy), min(H, H - dy)
        ws, we = max(0, -dx), min(W, W - dx)
        sh[:, hs:he, ws:we] = xp[:, hs + dy:he + dy, ws + dx:we + dx]
        cols.append(sh)
    a = torch.cat(cols, -1).reshape(B * H * W, taps * cp)
    return (a @ wp.float().t() + bias.float()[:cout]).reshape(B, H, W, cout)


class _EngineModel:
    """Consumes the packed list exactly like VaeBuilder::walk (csrc/vae_engine.cu)."""

    def __init__(self, cfg: V.VaeDecoderConfig, weights):
        self.cfg, self.w, self.i = cfg, weights, 0

    def nxt(self):
        t = self.w[self.i]; self.i += 1
        return t

    def gn(self, x, silu):
        g, b = self.nxt().float(), self.nxt().float()
        y = F.group_norm(x.permute(0, 3, 1, 2), self.cfg.norm_num_groups, g, b, 1e-6).permute(0, 2, 3, 1)
        return F.silu(y) if silu else y

    def conv(self, x, cout, taps):
        return _conv_from_packed(x, self.nxt(), self.nxt(), cout, taps)

    def resnet(self, x, cin, cout):
        h = self.conv(self.gn(x, True), cout, 9)
        t = self.gn(h, True)
        c2w, c2b = self.nxt(), self.nxt()
        res = self.conv(x, cout, 1) if cin != cout else x
        return _conv_from_packed(t, c2w, c2b, cout, 9) + res

    def attention(self, x, c):
        B, H, W, _ = x.shape
        t = self.gn(x, False).reshape(B, H * W, c)
        qkw, qkb, vw, vb, ow, ob = (self.nxt() for _ in range(6))
        qk = t @ qkw.float().t() + qkb.float()
        out = []
        for b in range(B):
            vt = vw.float() @ t[b].t()                                   # V^T without its bias
            p = torch.softmax(qk[b, :, :c] @ qk[b, :, c:].t() / math.sqrt(c), -1)
            out.append(p @ vt.t() + vb.float()[:c])                      # bias of V behind the row-stochastic P
        o = torch.stack(out) @ ow.float().t() + ob.float()[:c]
        return o.reshape(B, H, W, c) + x

    def decode(self, latents):
        cfg = self.cfg
        rev = list(reversed(cfg.block_out_channels))
        z = (latents / cfg.scaling_factor + cfg.shift_factor).permute(0, 2, 3, 1)
        z = F.pad(z, (0, 64 - z.shape[-1]))                              # the engine's latent buffer has 64 channels
        x = self.conv(z, rev[0], 9)
        x = self.resnet(x, rev[0], rev[0]); x = self.attention(x, rev[0]); x = self.resnet(x, rev[0], rev[0])
        prev = rev[0]
        for i, ch in enumerate(rev):
            for j in range(cfg.layers_per_block + 1):
                x = self.resnet(x, prev if j == 0 else ch, ch)
            if i != len(rev) - 1:
                x = x.repeat_interleave(2, 1).repeat_interleave(2, 2)
                x = self.conv(x, ch, 9)
            prev = ch
        x = self.gn(x, True)
        img = self.conv(x, cfg.out_channels, 9)
        assert self.i == len(self.w)
        return img.permute(0, 3, 1, 2)


def _cfgs():
    tiny = VO.tiny_vae_config()
    shortcut = VO.VaeConfig(latent_channels=4, block_out_channels=(16, 32, 64), layers_per_block=1, norm_num_groups=8)
    return [tiny, shortcut]


@pytest.mark.parametrize("ocfg", _cfgs(), ids=["tiny", "three_levels"])
def test_packed_engine_model_matches_oracle(ocfg):
    w = {k: v.to(torch.bfloat16).float() for k, v in VO.make_vae_decoder_weights(ocfg, seed=3).items()}   # bf16-exact weights
    cfg = V.VaeDecoderConfig(latent_channels=ocfg.latent_channels, out_channels=ocfg.out_channels, block_out_channels=tuple(ocfg.block_out_channels),
                             layers_per_block=ocfg.layers_per_block, norm_num_groups=ocfg.norm_num_groups,
                             scaling_factor=ocfg.scaling_factor, shift_factor=ocfg.shift_factor)
    packed = V.pack_vae_decoder_weights(w, cfg)
    assert len(packed) == V.expected_weight_count(cfg)
    assert all(t.dtype == torch.bfloat16 and t.is_contiguous() for t in packed)
    lat = torch.randn(2, ocfg.latent_channels, 5, 6, generator=torch.Generator().manual_seed(4))
    with torch.no_grad():
        ref = VO.vae_decode(w, ocfg, lat)
        got = _EngineModel(cfg, packed).decode(lat)
    assert got.shape == ref.shape
    torch.testing.assert_close(got, ref, rtol=2e-4, atol=2e-4)


def test_pack_conv3x3_layout():
    w = torch.arange(2 * 3 * 9, dtype=torch.float32).reshape(2, 3, 3, 3)
    p = V.pack_conv3x3(w)
    assert p.shape == (2, 9 * 64) and p.dtype == torch.bfloat16
    for co in range(2):
        for tap in range(9):
            for c in range(3):
                assert float(p[co, tap * 64 + c]) == float(w[co, c, tap // 3, tap % 3])
            assert float(p[co, tap * 64 + 3:(tap + 1) * 64].abs().max()) == 0.0
    assert V.pad_vec8(torch.ones(3)).tolist() == [1.0, 1.0, 1.0, 0, 0, 0, 0, 0]
    with pytest.raises(ValueError):
        V.pack_conv3x3(torch.zeros(2, 3, 1, 1))


def test_config_from_diffusers_dict_and_guards():
    d = dict(latent_channels=16, out_channels=3, block_out_channels=[128, 256, 512, 512], layers_per_block=2, norm_num_groups=32,
             scaling_factor=1.5305, shift_factor=0.0609, use_post_quant_conv=False)
    cfg = V.VaeDecoderConfig.from_config(d)
    assert cfg == V.VaeDecoderConfig()
    assert V.expected_weight_count(cfg) == 136
    with pytest.raises(NotImplementedError):
        V.VaeDecoderConfig.from_config({**d, "use_post_quant_conv": True})
    w = VO.make_vae_decoder_weights(VO.tiny_vae_config(), seed=0)
    with pytest.raises(NotImplementedError):
        V.pack_vae_decoder_weights({**w, "post_quant_conv.weight": torch.zeros(1)}, V.VaeDecoderConfig(4, 3, (32, 64), 1, 8))
    with pytest.raises(ValueError, match="missing"):       # config deeper than the state dict
        V.pack_vae_decoder_weights(w, V.VaeDecoderConfig(4, 3, (32, 64), 2, 8))
    w2 = VO.make_vae_decoder_weights(VO.VaeConfig(4, 3, (32, 64), 2, 8), seed=0)
    with pytest.raises(ValueError, match="does not match"):  # state dict deeper than the config
        V.pack_vae_decoder_weights(w2, V.VaeDecoderConfig(4, 3, (32, 64), 1, 8))


@pytest.mark.skipif(not os.path.exists(_lib.LIB_PATH), reason="libffb200.so not built")
def test_c_weight_count_matches_python():
    """ffb200_vae_weight_count walks the same builder as ffb200_vae_decoder_create without touching the GPU."""
    L = V._L()
    for cfg in (V.VaeDecoderConfig(), V.VaeDecoderConfig(4, 3, (32, 64), 1, 8), V.VaeDecoderConfig(4, 3, (16, 32, 64), 1, 8),
                V.VaeDecoderConfig(16, 3, (128, 128, 256), 3, 32)):
        assert L.ffb200_vae_weight_count(C.byref(V._c_config(cfg))) == V.expected_weight_count(cfg)
    bad = V.VaeDecoderConfig(4, 3, (32, 48), 1, 8)          # mid-block width not a multiple of 64
    assert L.ffb200_vae_weight_count(C.byref(V._c_config(bad))) < 0
    assert b"multiple of 64" in L.ffb200_last_error()


def test_decoder_needs_cuda():
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    with pytest.raises(RuntimeError, match="no CPU path"):
        V.B200VaeDecoder(V.VaeDecoderConfig(), {}, 8, 8)


# ------------------------------------------------------------------------------------------------ tile-level model of csrc/vae_conv.cu
def _pick_tw_log2(H, W):
    """vae_pick_tw_log2 (csrc/vae_engine.cu)."""
    best, best_area = 7, None
    for l in range(7, 2, -1):
        tw, th = 1 << l, 128 >> l
        area = -(-W // tw) * tw * (-(-H // th)) * th
        if best_area is None or area < best_area:
            best, best_area = l, area
    return best


def _tma_box(x, b, h0, w0, c0, th, tw):
    """cp.async.bulk.tensor.4d tile mode on NHWC x: box {64, tw, th, 1} at (c0, w0, h0, b), zero fill outside the tensor; rows in
    box order (w fastest, then h) = the rows of the 128 x 64 A tile."""
    B, H, W, Cc = x.shape
    box = torch.zeros(th, tw, 64)
    if 0 <= b < B:
        for i in range(th):
            for j in range(tw):
                h, w = h0 + i, w0 + j
                if 0 <= h < H and 0 <= w < W:
                    n = max(0, min(64, Cc - c0))
                    box[i, j, :n] = x[b, h, w, c0:c0 + n]
    return box.reshape(th * tw, 64)


def _conv_tiled(x, wp, bias, cout, taps):
    """Walks m tiles in CTA pairs exactly like conv_bf16_kernel: conv_tile_origin, the tap / channel-block K loop, pixel_of, n_store."""
    B, H, W, cin = x.shape
    l2 = _pick_tw_log2(H, W)
    tw, th = 1 << l2, 128 >> l2
    tiles_w, tiles_h = -(-W // tw), -(-H // th)
    tiles_m = B * tiles_h * tiles_w
    kc = -(-cin // 64)
    N = -(-cout // 64) * 64
    wfull = torch.zeros(N, taps * kc * 64)
    wfull[:cout, :wp.shape[1]] = wp.float()                       # rows beyond Cout / columns beyond k_total: TMA zero fill
    out = torch.full((B, H, W, cout), float("nan"))
    writes = torch.zeros(B, H, W, dtype=torch.int32)
    for mp in range((tiles_m + 1) // 2):
        for rank in range(2):
            tm = 2 * mp + rank
            if tm >= tiles_m:
                b, h0, w0 = B, 0, 0                                # ghost tile
            else:
                b, rem = divmod(tm, tiles_h * tiles_w)
                ty, tx = divmod(rem, tiles_w)
                h0, w0 = ty * th, tx * tw
            acc = torch.zeros(128, N)
            kb = 0
            for tap in range(taps):
                dy, dx = (tap // 3 - 1, tap % 3 - 1) if taps == 9 else (0, 0)
                for kcb in range(kc):
                    a = _tma_box(x, b, h0 + dy, w0 + dx, kcb * 64, th, tw)
                    acc += a @ wfull[:, kb * 64:(kb + 1) * 64].t()
                    kb += 1
            for r in range(128):
                h, w = h0 + (r >> l2), w0 + (r & (tw - 1))
                if tm < tiles_m and h < H and w < W:
                    out[b, h, w] = acc[r, :cout] + bias.float()[:cout]
                    writes[b, h, w] += 1
    assert int(writes.min()) == 1 and int(writes.max()) == 1       # every output pixel written exactly once
    return out


@pytest.mark.parametrize("B,H,W,cin,cout,taps", [(1, 5, 6, 8, 8, 9), (2, 3, 20, 16, 24, 9), (1, 1, 130, 72, 8, 1), (3, 9, 9, 8, 72, 9),
                                                   (1, 16, 16, 8, 8, 9)])
def test_conv_tile_walk_matches_conv2d(B, H, W, cin, cout, taps):
    g = torch.Generator().manual_seed(B + H + W + cin + cout)
    x = torch.randn(B, H, W, cin, generator=g).bfloat16().float()
    k = 3 if taps == 9 else 1
    w = (torch.randn(cout, cin, k, k, generator=g) / (k * cin ** 0.5)).bfloat16().float()
    bias = torch.randn(cout, generator=g).bfloat16().float()
    wp = V.pack_conv3x3(w) if taps == 9 else V.pack_conv1x1(w)
    got = _conv_tiled(x, wp, V.pad_vec8(bias), cout, taps)
    ref = F.conv2d(x.permute(0, 3, 1, 2), w, bias, padding=k // 2).permute(0, 2, 3, 1)
    torch.testing.assert_close(got, ref, rtol=1e-4, atol=1e-4)
