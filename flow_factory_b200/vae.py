"""Host side of the native VAE decode: weight packing + the ctypes binding of `ffb200_vae_*`.

Mirrors `SD3_5Adapter.decode_latents` (FF/models/stable_diffusion/sd3_5.py:161-172): `latents.to(vae.dtype) / scaling_factor +
shift_factor -> AutoencoderKL.decode -> image_processor.postprocess(output_type="pt")`.

STATUS: validated on B200 in round 2 (tests/test_gpu_vae.py: layer ops and the whole decoder against the pinned oracle; 1024^2 decode
12.6 ms per image at batch 8, profiles/r02_vae_decode_1024.json).  The packing below is also unit-tested on CPU (tests/test_host_logic_vae.py).

Weight order handed to `ffb200_vae_decoder_create` (all bf16, contiguous):

    conv_in.w conv_in.b
    mid resnet 0 | mid attention | mid resnet 1
    for every up block: (layers_per_block + 1) resnets, then upsampler conv.w conv.b (all but the last block)
    conv_norm_out.gamma conv_norm_out.beta conv_out.w conv_out.b

    resnet    := norm1.gamma norm1.beta conv1.w conv1.b norm2.gamma norm2.beta conv2.w conv2.b [conv_shortcut.w conv_shortcut.b]
    attention := group_norm.gamma group_norm.beta  [to_q;to_k].w [to_q;to_k].b  to_v.w to_v.b  to_out.0.w to_out.0.b

3x3 kernels are packed [Cout][tap = ky*3+kx][Cin padded to a multiple of 64] (the K order of the implicit GEMM in csrc/vae_conv.cu),
1x1 kernels / linears stay [Cout][Cin], biases are zero-padded to a multiple of 8 entries.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Any, List, Mapping, Optional, Tuple

import torch

from . import _lib

vp, ci, cf = C.c_void_p, C.c_int, C.c_float


@dataclass
class VaeDecoderConfig:
    """The AutoencoderKL config fields the decoder depends on (DF/models/autoencoders/autoencoder_kl.py register_to_config)."""
    latent_channels: int = 16
    out_channels: int = 3
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_num_groups: int = 32
    scaling_factor: float = 1.5305
    shift_factor: float = 0.0609

    @classmethod
    def from_config(cls, cfg: Any) -> "VaeDecoderConfig":
        get = (lambda k, d=None: cfg.get(k, d)) if isinstance(cfg, Mapping) else (lambda k, d=None: getattr(cfg, k, d))
        if get("use_post_quant_conv", False):
            # SD3 / FLUX VAEs have no post_quant_conv; the SD1.x-style 1x1 conv in front of the decoder is not implemented
            raise NotImplementedError("VAE with post_quant_conv is not supported by the native decoder")
        shift = get("shift_factor", 0.0)
        return cls(latent_channels=int(get("latent_channels", 16)), out_channels=int(get("out_channels", 3)),
                   block_out_channels=tuple(int(c) for c in get("block_out_channels")), layers_per_block=int(get("layers_per_block", 2)),
                   norm_num_groups=int(get("norm_num_groups", 32)), scaling_factor=float(get("scaling_factor", 1.0)),
                   shift_factor=float(shift if shift is not None else 0.0))


class VaeConfigC(C.Structure):
    _fields_ = [("latent_channels", ci), ("out_channels", ci), ("num_blocks", ci), ("block_out_channels", ci * 8),
                ("layers_per_block", ci), ("norm_num_groups", ci), ("scaling_factor", cf), ("shift_factor", cf)]


def _c_config(cfg: VaeDecoderConfig) -> VaeConfigC:
    if not 1 <= len(cfg.block_out_channels) <= 8:
        raise ValueError("block_out_channels must have 1..8 entries")
    arr = (ci * 8)(*cfg.block_out_channels, *([0] * (8 - len(cfg.block_out_channels))))
    return VaeConfigC(cfg.latent_channels, cfg.out_channels, len(cfg.block_out_channels), arr, cfg.layers_per_block,
                      cfg.norm_num_groups, cfg.scaling_factor, cfg.shift_factor)


# ------------------------------------------------------------------------------------------------ packing (pure torch, CPU-testable)
def pack_conv3x3(weight: torch.Tensor) -> torch.Tensor:
    """[Cout, Cin, 3, 3] -> bf16 [Cout, 9 * Cin_pad] with K index tap * Cin_pad + c, tap = ky * 3 + kx, Cin_pad = ceil64(Cin)."""
    co, cin, kh, kw = weight.shape
    if (kh, kw) != (3, 3):
        raise ValueError(f"expected a 3x3 kernel, got {kh}x{kw}")
    cp = (cin + 63) // 64 * 64
    out = torch.zeros(co, 9, cp, dtype=torch.bfloat16, device=weight.device)
    out[:, :, :cin] = weight.permute(0, 2, 3, 1).reshape(co, 9, cin).to(torch.bfloat16)
    return out.reshape(co, 9 * cp).contiguous()


def pack_conv1x1(weight: torch.Tensor) -> torch.Tensor:
    """[Cout, Cin, 1, 1] or [Cout, Cin] -> bf16 [Cout, Cin]."""
    return weight.reshape(weight.shape[0], weight.shape[1]).to(torch.bfloat16).contiguous()


def pad_vec8(v: torch.Tensor) -> torch.Tensor:
    n = (v.numel() + 7) // 8 * 8
    out = torch.zeros(n, dtype=torch.bfloat16, device=v.device)
    out[: v.numel()] = v.to(torch.bfloat16)
    return out


def expected_weight_count(cfg: VaeDecoderConfig) -> int:
    rev = list(reversed(cfg.block_out_channels))
    n, prev = 2 + 8 + 8 + 8, rev[0]            # conv_in, mid resnet 0, attention, mid resnet 1
    for i, ch in enumerate(rev):
        for j in range(cfg.layers_per_block + 1):
            n += 8 + (2 if (prev if j == 0 else ch) != ch else 0)
        if i != len(rev) - 1:
            n += 2
        prev = ch
    return n + 4


def pack_vae_decoder_weights(state_dict: Mapping[str, torch.Tensor], cfg: VaeDecoderConfig, device=None) -> List[torch.Tensor]:
    """AutoencoderKL.state_dict() (`decoder.*` keys, any float dtype) -> the ordered bf16 tensor list described in the module docstring."""
    class _Need(dict):
        def __missing__(self, key):
            raise ValueError(f"state dict does not match the config: `{key}` is missing")

    sd = _Need({k: v for k, v in state_dict.items() if k.startswith("decoder.")})
    if any(k.startswith("post_quant_conv.") for k in state_dict):
        raise NotImplementedError("VAE with post_quant_conv is not supported by the native decoder")
    out: List[torch.Tensor] = []

    def dev(t: torch.Tensor) -> torch.Tensor:
        return t.to(device) if device is not None else t

    def conv3(name: str):
        out.append(dev(pack_conv3x3(sd[name + ".weight"]))); out.append(dev(pad_vec8(sd[name + ".bias"])))

    def conv1(name: str):
        out.append(dev(pack_conv1x1(sd[name + ".weight"]))); out.append(dev(pad_vec8(sd[name + ".bias"])))

    def norm(name: str):
        out.append(dev(sd[name + ".weight"].to(torch.bfloat16).contiguous())); out.append(dev(sd[name + ".bias"].to(torch.bfloat16).contiguous()))

    def resnet(pre: str):
        norm(pre + "norm1"); conv3(pre + "conv1"); norm(pre + "norm2"); conv3(pre + "conv2")
        if pre + "conv_shortcut.weight" in sd:
            conv1(pre + "conv_shortcut")

    rev = list(reversed(cfg.block_out_channels))
    if cfg.latent_channels > 64:
        raise ValueError("latent_channels > 64 is not supported (the latent tensor is stored with 64 channels)")
    conv3("decoder.conv_in")      # Cin padded to 64 == the channel count of the engine's NHWC latent buffer
    resnet("decoder.mid_block.resnets.0.")
    a = "decoder.mid_block.attentions.0."
    norm(a + "group_norm")
    out.append(dev(torch.cat([sd[a + "to_q.weight"], sd[a + "to_k.weight"]], 0).to(torch.bfloat16).contiguous()))
    out.append(dev(torch.cat([sd[a + "to_q.bias"], sd[a + "to_k.bias"]], 0).to(torch.bfloat16).contiguous()))
    out.append(dev(sd[a + "to_v.weight"].to(torch.bfloat16).contiguous())); out.append(dev(pad_vec8(sd[a + "to_v.bias"])))
    out.append(dev(sd[a + "to_out.0.weight"].to(torch.bfloat16).contiguous())); out.append(dev(pad_vec8(sd[a + "to_out.0.bias"])))
    resnet("decoder.mid_block.resnets.1.")
    for i in range(len(rev)):
        for j in range(cfg.layers_per_block + 1):
            resnet(f"decoder.up_blocks.{i}.resnets.{j}.")
        if i != len(rev) - 1:
            conv3(f"decoder.up_blocks.{i}.upsamplers.0.conv")
    norm("decoder.conv_norm_out")
    conv3("decoder.conv_out")
    used = expected_weight_count(cfg)
    n_params = sum(1 for k in sd if k.endswith((".weight", ".bias")))
    if len(out) != used or n_params != used + 2:      # to_q / to_k are fused: two tensors fewer than parameters
        raise ValueError(f"state dict does not match the config: packed {len(out)} of {n_params} decoder parameters, expected {used}")
    return out


def postprocess_pt(images: torch.Tensor) -> torch.Tensor:
    """VaeImageProcessor.postprocess(output_type="pt") with do_normalize: (x / 2 + 0.5).clamp(0, 1) (DF/image_processor.py)."""
    return (images / 2 + 0.5).clamp(0, 1)


# ------------------------------------------------------------------------------------------------ binding
_bound = False


def _L() -> C.CDLL:
    global _bound
    L = _lib.lib()
    if not _bound:
        L.ffb200_vae_weight_count.argtypes = [C.POINTER(VaeConfigC)]
        L.ffb200_vae_decoder_create.argtypes = [C.POINTER(VaeConfigC), C.POINTER(vp), ci, ci, ci, ci, C.POINTER(vp)]
        L.ffb200_vae_decoder_destroy.argtypes = [vp]; L.ffb200_vae_decoder_destroy.restype = None
        L.ffb200_vae_decoder_workspace_bytes.argtypes = [vp]; L.ffb200_vae_decoder_workspace_bytes.restype = C.c_longlong
        L.ffb200_vae_decode.argtypes = [vp, vp, vp, vp]
        L.ffb200_conv2d_nhwc.argtypes = [vp, vp, vp, vp, vp, ci, ci, ci, ci, ci, ci, vp]
        L.ffb200_group_norm_nhwc.argtypes = [vp, vp, vp, vp, ci, C.c_longlong, ci, ci, cf, ci, vp, vp]
        _bound = True
    return L


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def conv2d_nhwc(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], residual: Optional[torch.Tensor] = None) -> torch.Tensor:
    """nn.Conv2d(k=3, p=1) or k=1 on an NHWC bf16 tensor through `ffb200_conv2d_nhwc` (weight in the nn.Conv2d layout)."""
    B, H, W, cin = x.shape
    co = weight.shape[0]
    taps = 9 if weight.dim() == 4 and weight.shape[-1] == 3 else 1
    wp = pack_conv3x3(weight) if taps == 9 else pack_conv1x1(weight)
    bp = pad_vec8(bias) if bias is not None else None
    out = torch.empty(B, H, W, co, dtype=torch.bfloat16, device=x.device)
    _lib.check(_L().ffb200_conv2d_nhwc(x.data_ptr(), wp.data_ptr(), bp.data_ptr() if bp is not None else None,
                                       residual.data_ptr() if residual is not None else None, out.data_ptr(), B, H, W, cin, co, taps,
                                       _stream()), "ffb200_conv2d_nhwc")
    return out


def group_norm_nhwc(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, groups: int, eps: float = 1e-6, silu: bool = False) -> torch.Tensor:
    """nn.GroupNorm (+ SiLU) on an NHWC bf16 tensor [B, ..., C] through `ffb200_group_norm_nhwc`."""
    B, Cc = x.shape[0], x.shape[-1]
    P = x.numel() // (B * Cc)
    out = torch.empty_like(x)
    ws = torch.empty(B * Cc * 2, dtype=torch.float64, device=x.device)
    _lib.check(_L().ffb200_group_norm_nhwc(x.data_ptr(), gamma.to(torch.bfloat16).data_ptr(), beta.to(torch.bfloat16).data_ptr(),
                                           out.data_ptr(), B, P, Cc, groups, eps, int(silu), ws.data_ptr(), _stream()),
               "ffb200_group_norm_nhwc")
    return out


class B200VaeDecoder:
    """AutoencoderKL.decode for one latent geometry.  `decode(latents)` takes the rollout's fp16 latents [n, C, h, w] and returns
    bf16 images [n, 3, 8h, 8w]; `decode_latents(latents)` adds the reference's postprocess ("pt": floats in [0, 1])."""

    def __init__(self, config: Any, state_dict: Mapping[str, torch.Tensor], lat_h: int, lat_w: int, batch: int = 4, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("B200VaeDecoder needs a CUDA device (sm_100a); there is no CPU path")
        self.cfg = config if isinstance(config, VaeDecoderConfig) else VaeDecoderConfig.from_config(config)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.batch, self.lat_h, self.lat_w = int(batch), int(lat_h), int(lat_w)
        self.weights = pack_vae_decoder_weights(state_dict, self.cfg, device=self.device)
        L = _L()
        cc = _c_config(self.cfg)
        n = L.ffb200_vae_weight_count(C.byref(cc))
        if n != len(self.weights):
            raise RuntimeError(f"ffb200_vae_weight_count = {n}, packed {len(self.weights)}")
        ptrs = (vp * n)(*[w.data_ptr() for w in self.weights])
        h = vp()
        with torch.cuda.device(self.device):
            _lib.check(L.ffb200_vae_decoder_create(C.byref(cc), ptrs, n, self.batch, self.lat_h, self.lat_w, C.byref(h)), "ffb200_vae_decoder_create")
        self._h = h
        self.upscale = 2 ** (len(self.cfg.block_out_channels) - 1)

    def workspace_bytes(self) -> int:
        return int(_L().ffb200_vae_decoder_workspace_bytes(self._h))

    @staticmethod
    def last_launch_count() -> int:
        return int(_lib.lib().ffb200_last_launch_count())

    def decode(self, latents: torch.Tensor) -> torch.Tensor:
        if latents.dim() != 4 or tuple(latents.shape[1:]) != (self.cfg.latent_channels, self.lat_h, self.lat_w):
            raise ValueError(f"expected latents [n, {self.cfg.latent_channels}, {self.lat_h}, {self.lat_w}], got {tuple(latents.shape)}")
        lat = latents.to(device=self.device, dtype=torch.float16).contiguous()
        n, u = lat.shape[0], self.upscale
        out = torch.empty(n, self.cfg.out_channels, self.lat_h * u, self.lat_w * u, dtype=torch.bfloat16, device=self.device)
        L = _L()
        for i in range(0, n, self.batch):
            chunk = lat[i:i + self.batch]
            m = chunk.shape[0]
            if m < self.batch:                       # ragged tail: decode a full batch, keep the first m images
                chunk = torch.cat([chunk, chunk.new_zeros(self.batch - m, *chunk.shape[1:])], 0)
                tmp = torch.empty(self.batch, *out.shape[1:], dtype=torch.bfloat16, device=self.device)
                _lib.check(L.ffb200_vae_decode(self._h, chunk.data_ptr(), tmp.data_ptr(), _stream()), "ffb200_vae_decode")
                out[i:i + m] = tmp[:m]
            else:
                _lib.check(L.ffb200_vae_decode(self._h, chunk.data_ptr(), out[i:i + m].data_ptr(), _stream()), "ffb200_vae_decode")
        return out

    def decode_latents(self, latents: torch.Tensor, output_type: str = "pt") -> torch.Tensor:
        if output_type != "pt":
            raise NotImplementedError("the native decoder returns tensors (output_type='pt'); PIL / numpy conversion stays in the reference")
        return postprocess_pt(self.decode(latents))

    def close(self) -> None:
        if getattr(self, "_h", None):
            _L().ffb200_vae_decoder_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
