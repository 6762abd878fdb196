"""Output record of the rollout path (mirror of FF/samples/samples.py:68-130 BaseSample and
FF/models/stable_diffusion/sd3_5.py:50-58 SD3_5Sample): tensors carry no batch dimension."""
from __future__ import annotations

import hashlib
from dataclasses import dataclass, field, fields
from typing import Any, ClassVar, Dict, List, Optional

import torch


@dataclass
class SD3_5Sample:
    _shared_fields: ClassVar[frozenset] = frozenset({})
    _id_fields: ClassVar[frozenset] = frozenset({"prompt", "prompt_ids", "negative_prompt", "negative_prompt_ids"})
    # denoising trajectory
    timesteps: Optional[torch.Tensor] = None
    all_latents: Optional[torch.Tensor] = None           # (T', C, H, W) storage dtype
    latent_index_map: Optional[torch.Tensor] = None      # (T+1,) long, -1 = not stored
    log_probs: Optional[torch.Tensor] = None             # (T'',) fp32
    log_prob_index_map: Optional[torch.Tensor] = None    # (T+1,) long
    height: Optional[int] = None
    width: Optional[int] = None
    image: Optional[torch.Tensor] = None
    video: Optional[torch.Tensor] = None                 # (T, C, H, W); filled by the video adapters (FF/samples/samples.py:96-99)
    audio: Optional[torch.Tensor] = None
    audio_sample_rate: Optional[int] = None
    prompt: Optional[str] = None
    prompt_ids: Optional[torch.Tensor] = None
    prompt_embeds: Optional[torch.Tensor] = None
    negative_prompt: Optional[str] = None
    negative_prompt_ids: Optional[torch.Tensor] = None
    negative_prompt_embeds: Optional[torch.Tensor] = None
    pooled_prompt_embeds: Optional[torch.Tensor] = None
    negative_pooled_prompt_embeds: Optional[torch.Tensor] = None
    extra_kwargs: Dict[str, Any] = field(default_factory=dict)
    _unique_id: Optional[int] = field(default=None, repr=False, compare=False)

    @property
    def unique_id(self) -> int:
        """sha256 over the prompt identity fields (FF/samples/samples.py:268-288), cached like the reference's `_unique_id`."""
        if self._unique_id is not None:
            return self._unique_id
        h = hashlib.sha256()
        for name in sorted(self._id_fields):
            v = getattr(self, name)
            if v is None:
                continue
            h.update(name.encode())
            h.update(v.detach().cpu().numpy().tobytes() if isinstance(v, torch.Tensor) else str(v).encode())
        self._unique_id = int.from_bytes(h.digest()[:8], "big", signed=True)
        return self._unique_id

    def to(self, device) -> "SD3_5Sample":
        for f in fields(self):
            v = getattr(self, f.name)
            if isinstance(v, torch.Tensor):
                setattr(self, f.name, v.to(device))
        return self

    def to_dict(self) -> Dict[str, Any]:
        return {f.name: getattr(self, f.name) for f in fields(self)}

    @classmethod
    def stack(cls, samples: List["SD3_5Sample"]) -> Dict[str, Any]:
        """Collate (FF/samples/samples.py:347-375): tensors stacked on a new batch dim, shared fields take element 0."""
        out: Dict[str, Any] = {}
        for f in fields(cls):
            vals = [getattr(s, f.name) for s in samples]
            if f.name in cls._shared_fields:
                out[f.name] = vals[0]
            elif all(isinstance(v, torch.Tensor) for v in vals) and len({tuple(v.shape) for v in vals}) == 1:
                out[f.name] = torch.stack(vals, dim=0)
            else:
                out[f.name] = vals
        return out


@dataclass
class Flux1Sample(SD3_5Sample):
    """Mirror of FF/models/flux/flux1.py:53-59 (T2ISample + pooled_prompt_embeds, img_ids; img_ids is a shared field).
    all_latents rows are PACKED latents (T', Ni, 64)."""
    _shared_fields: ClassVar[frozenset] = frozenset({"img_ids"})
    img_ids: Optional[torch.Tensor] = None


@dataclass
class QwenImageSample(SD3_5Sample):
    """Mirror of the reference's QwenImageSample (FF/models/qwen_image/qwen_image.py:54-61): T2I sample + embedding masks + img_shapes,
    NO shared fields (img_shapes is collated per sample like any other list).  all_latents rows are PACKED latents (T', Ni, 64)."""
    _shared_fields: ClassVar[frozenset] = frozenset({})
    prompt_embeds_mask: Optional[torch.Tensor] = None
    negative_prompt_embeds_mask: Optional[torch.Tensor] = None
    img_shapes: Optional[List] = None


@dataclass
class WanT2VSample(SD3_5Sample):
    """Mirror of FF/models/wan/wan2_t2v.py:47-50 (T2VSample, no shared fields): fills `video` instead of `image`;
    all_latents rows are (T', C, F, H, W)."""
