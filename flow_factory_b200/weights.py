"""Pack an SD3Transformer2DModel state dict (diffusers key names, DF/models/transformers/transformer_sd3.py:142-181)
into the flat bf16 device tensors the C ABI borrows (include/ffb200.h: ffb200_weights).  q|k|v projections are
concatenated along out_features; all adaLN projections are stacked into one matrix so that a single skinny GEMV per
denoise step produces every layer's shift/scale/gate vectors (temb is layer-invariant).

Call `pack()` again (or RolloutEngine.refresh_weights) after an optimizer step, an EMA/ref swap or a LoRA merge -
the engine only borrows pointers."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List

import torch

from . import _lib


@dataclass
class EngineConfig:
    num_layers: int
    num_heads: int
    patch_size: int = 2
    in_channels: int = 16
    joint_attention_dim: int = 4096
    pooled_projection_dim: int = 2048
    pos_embed_max_size: int = 384
    num_dual_layers: int = 13

    @property
    def inner_dim(self) -> int:
        return 64 * self.num_heads

    @classmethod
    def from_model_config(cls, cfg) -> "EngineConfig":
        """`cfg`: anything with the SD3Transformer2DModel config attributes (diffusers FrozenDict, oracle SD3Config...)."""
        g = (lambda k: cfg[k]) if isinstance(cfg, dict) else (lambda k: getattr(cfg, k))
        if g("attention_head_dim") != 64:
            raise ValueError("the sm_100a attention kernel is specialised for head_dim 64 (SD3 / SD3.5)")
        dual = tuple(g("dual_attention_layers"))
        if dual != tuple(range(len(dual))):
            raise ValueError("dual_attention_layers must be a prefix range (SD3.5: 0..12)")
        if g("qk_norm") != "rms_norm":
            raise ValueError("only qk_norm='rms_norm' (SD3.5) is implemented")
        if g("caption_projection_dim") != g("num_attention_heads") * 64:
            raise ValueError("caption_projection_dim must equal inner_dim")
        return cls(num_layers=g("num_layers"), num_heads=g("num_attention_heads"), patch_size=g("patch_size"),
                   in_channels=g("in_channels"), joint_attention_dim=g("joint_attention_dim"),
                   pooled_projection_dim=g("pooled_projection_dim"), pos_embed_max_size=g("pos_embed_max_size"),
                   num_dual_layers=len(dual))


class PackedWeights:
    def __init__(self, cfg: EngineConfig, state_dict: Dict[str, torch.Tensor], device: torch.device):
        self.cfg = cfg
        self.device = device
        self.tensors: Dict[str, torch.Tensor] = {}
        self.layer_structs = (_lib.LayerWeights * cfg.num_layers)()
        self.struct = _lib.Weights()
        self.pack(state_dict)

    def _put(self, name: str, t: torch.Tensor, dtype=torch.bfloat16) -> int:
        t = t.detach().to(device=self.device, dtype=dtype).contiguous()
        old = self.tensors.get(name)
        if old is not None and old.shape == t.shape and old.dtype == t.dtype:
            old.copy_(t)            # keep the address stable: plans hold TMA descriptors on it
            t = old
        else:
            self.tensors[name] = t
        assert t.data_ptr() % 16 == 0
        return t.data_ptr()

    def pack(self, sd: Dict[str, torch.Tensor]) -> None:
        cfg, D = self.cfg, self.cfg.inner_dim
        cat = lambda names: torch.cat([sd[n] for n in names], dim=0)
        W = self.struct
        W.pe_w = self._put("pe_w", sd["pos_embed.proj.weight"].reshape(D, -1))
        W.pe_b = self._put("pe_b", sd["pos_embed.proj.bias"])
        W.pos_embed = self._put("pos_embed", sd["pos_embed.pos_embed"].reshape(-1, D), torch.float32)
        for short, key in (("t1", "time_text_embed.timestep_embedder.linear_1"), ("t2", "time_text_embed.timestep_embedder.linear_2"),
                           ("p1", "time_text_embed.text_embedder.linear_1"), ("p2", "time_text_embed.text_embedder.linear_2"),
                           ("ctx", "context_embedder"), ("proj", "proj_out")):
            setattr(W, short + "_w", self._put(short + "_w", sd[key + ".weight"]))
            setattr(W, short + "_b", self._put(short + "_b", sd[key + ".bias"]))
        mod_w: List[torch.Tensor] = []
        mod_b: List[torch.Tensor] = []
        for i in range(cfg.num_layers):
            pre = f"transformer_blocks.{i}."
            last, dual = i == cfg.num_layers - 1, i < cfg.num_dual_layers
            mod_w += [sd[pre + "norm1.linear.weight"], sd[pre + "norm1_context.linear.weight"]]
            mod_b += [sd[pre + "norm1.linear.bias"], sd[pre + "norm1_context.linear.bias"]]
            assert sd[pre + "norm1.linear.weight"].shape[0] == (9 if dual else 6) * D
            assert sd[pre + "norm1_context.linear.weight"].shape[0] == (2 if last else 6) * D
            L = self.layer_structs[i]
            for f in _lib.LAYER_FIELDS:
                setattr(L, f, None)
            put = lambda field, t: setattr(L, field, self._put(f"L{i}.{field}", t))
            a = pre + "attn."
            put("qkv_w", cat([a + "to_q.weight", a + "to_k.weight", a + "to_v.weight"]))
            put("qkv_b", cat([a + "to_q.bias", a + "to_k.bias", a + "to_v.bias"]))
            put("norm_q", sd[a + "norm_q.weight"]); put("norm_k", sd[a + "norm_k.weight"])
            put("add_qkv_w", cat([a + "add_q_proj.weight", a + "add_k_proj.weight", a + "add_v_proj.weight"]))
            put("add_qkv_b", cat([a + "add_q_proj.bias", a + "add_k_proj.bias", a + "add_v_proj.bias"]))
            put("norm_added_q", sd[a + "norm_added_q.weight"]); put("norm_added_k", sd[a + "norm_added_k.weight"])
            put("out_w", sd[a + "to_out.0.weight"]); put("out_b", sd[a + "to_out.0.bias"])
            if not last:
                put("add_out_w", sd[a + "to_add_out.weight"]); put("add_out_b", sd[a + "to_add_out.bias"])
            if dual:
                a2 = pre + "attn2."
                put("qkv2_w", cat([a2 + "to_q.weight", a2 + "to_k.weight", a2 + "to_v.weight"]))
                put("qkv2_b", cat([a2 + "to_q.bias", a2 + "to_k.bias", a2 + "to_v.bias"]))
                put("norm_q2", sd[a2 + "norm_q.weight"]); put("norm_k2", sd[a2 + "norm_k.weight"])
                put("out2_w", sd[a2 + "to_out.0.weight"]); put("out2_b", sd[a2 + "to_out.0.bias"])
            put("ff1_w", sd[pre + "ff.net.0.proj.weight"]); put("ff1_b", sd[pre + "ff.net.0.proj.bias"])
            put("ff2_w", sd[pre + "ff.net.2.weight"]); put("ff2_b", sd[pre + "ff.net.2.bias"])
            if not last:
                put("cff1_w", sd[pre + "ff_context.net.0.proj.weight"]); put("cff1_b", sd[pre + "ff_context.net.0.proj.bias"])
                put("cff2_w", sd[pre + "ff_context.net.2.weight"]); put("cff2_b", sd[pre + "ff_context.net.2.bias"])
        mod_w.append(sd["norm_out.linear.weight"]); mod_b.append(sd["norm_out.linear.bias"])
        W.mod_w = self._put("mod_w", torch.cat(mod_w, dim=0))
        W.mod_b = self._put("mod_b", torch.cat(mod_b, dim=0))
        self.mod_rows = self.tensors["mod_w"].shape[0]
        W.layers = C.cast(self.layer_structs, C.POINTER(_lib.LayerWeights))

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self.tensors.values())


def has_lora_keys(state_dict) -> bool:
    """True for the state dict of a PEFT-wrapped module (`...base_layer.weight`, `...lora_A.<adapter>.weight`, `base_model.model.` prefix)."""
    return any((".lora_A." in k) or (".base_layer." in k) or k.startswith("base_model.model.") for k in state_dict)


def merge_lora_state_dict(state_dict: Dict[str, torch.Tensor], lora_alpha: float, lora_rank: int = 0,
                          adapter: str = "default", scale: float = 1.0) -> Dict[str, torch.Tensor]:
    """Fold PEFT LoRA pairs into the base weights: W' = W + (alpha / r) * B @ A (FF/models/abc.py:859-949 wraps the
    transformer with `get_peft_model` / `add_adapter('default', LoraConfig(r, lora_alpha))`).

    Accepts the state dict of a PEFT-wrapped SD3Transformer2DModel (keys like
    `base_model.model.transformer_blocks.0.attn.to_q.base_layer.weight`, `...to_q.lora_A.default.weight`,
    `...to_q.lora_B.default.weight`) or of a diffusers model with an added adapter (no `base_model.model.` prefix) and
    returns a plain state dict with diffusers key names, ready for PackedWeights.pack / RolloutEngine.refresh_weights.
    `scale` multiplies the folded update: 0.0 returns the BASE weights under plain key names - what the module computes inside PEFT's
    `disable_adapter()` (the LoRA form of `use_ref_parameters`, FF/models/abc.py:556-577).
    Cost: 2*r*out*in FLOP per adapted linear - negligible next to a rollout."""
    strip = lambda k: k[len("base_model.model."):] if k.startswith("base_model.model.") else k
    sd = {strip(k): v for k, v in state_dict.items()}
    out: Dict[str, torch.Tensor] = {}
    lora_a: Dict[str, torch.Tensor] = {}
    lora_b: Dict[str, torch.Tensor] = {}
    for k, v in sd.items():
        if f".lora_A.{adapter}.weight" in k:
            lora_a[k.replace(f".lora_A.{adapter}.weight", "")] = v
        elif f".lora_B.{adapter}.weight" in k:
            lora_b[k.replace(f".lora_B.{adapter}.weight", "")] = v
        elif ".lora_" in k:
            raise NotImplementedError(f"unsupported LoRA tensor {k} (DoRA / embedding LoRA are not merged)")
        else:
            out[k.replace(".base_layer.", ".")] = v
    if set(lora_a) != set(lora_b):
        raise ValueError("unpaired lora_A / lora_B tensors")
    for mod, A in lora_a.items():
        if scale == 0.0:
            if mod + ".weight" not in out:
                raise KeyError(f"LoRA targets {mod} but the base weight {mod}.weight is missing")
            continue
        B = lora_b[mod]
        r = lora_rank or A.shape[0]
        key = mod + ".weight"
        if key not in out:
            raise KeyError(f"LoRA targets {mod} but the base weight {key} is missing")
        W = out[key]
        delta = (B.to(torch.float32) @ A.to(torch.float32)) * (float(lora_alpha) / r * float(scale))
        out[key] = (W.to(torch.float32) + delta.to(W.device)).to(W.dtype)
    return out
