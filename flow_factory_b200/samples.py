"""Output records of the rollout path (mirrors of FF/samples/samples.py:68-375 BaseSample and the per-model sample classes):
tensors carry no batch dimension.  Behaviour is checked against the REAL reference classes in tests/test_reference_hooks.py."""
from __future__ import annotations

import hashlib
from dataclasses import dataclass, field, fields
from typing import Any, ClassVar, Dict, List, Optional

import torch


@dataclass
class BaseSample:
    """Mirror of FF/samples/samples.py:68-375 (BaseSample): the record the trainers consume through `BaseSample.stack(samples)`
    (grpo.py:215, nft.py:354, dpo.py:536 ...), `sample.unique_id` (dpo.py:339, dgpo.py:392), `sample.to(device)` and attribute / item access
    that falls through to `extra_kwargs`.  Same field names, defaults and behaviour; media fields are expected as tensors already
    (the reference's __post_init__ canonicalises PIL / numpy inputs, which the rollout engine never produces)."""
    _id_fields: ClassVar[frozenset] = frozenset({"prompt", "prompt_ids", "negative_prompt", "negative_prompt_ids"})
    # shared across the batch: stack() keeps the first element only (merged along the inheritance chain, see shared_fields())
    _shared_fields: ClassVar[frozenset] = frozenset({"height", "width", "latent_index_map", "log_prob_index_map"})
    # denoising trajectory
    timesteps: Optional[torch.Tensor] = None
    all_latents: Optional[torch.Tensor] = None           # (T', ...) storage dtype
    latent_index_map: Optional[torch.Tensor] = None      # (T+1,) long, -1 = not stored
    log_probs: Optional[torch.Tensor] = None             # (T'',) fp32
    log_prob_index_map: Optional[torch.Tensor] = None    # (T+1,) long
    height: Optional[int] = None
    width: Optional[int] = None
    image: Optional[torch.Tensor] = None                 # (C, H, W)
    video: Optional[torch.Tensor] = None                 # (T, C, H, W)
    audio: Optional[torch.Tensor] = None
    audio_sample_rate: Optional[int] = None
    prompt: Optional[str] = None
    prompt_ids: Optional[torch.Tensor] = None
    prompt_embeds: Optional[torch.Tensor] = None
    negative_prompt: Optional[str] = None
    negative_prompt_ids: Optional[torch.Tensor] = None
    negative_prompt_embeds: Optional[torch.Tensor] = None
    extra_kwargs: Dict[str, Any] = field(default_factory=dict)
    _unique_id: Optional[int] = field(default=None, repr=False, compare=False)

    # ---- field configuration
    @classmethod
    def shared_fields(cls) -> frozenset:
        out = set()
        for base in cls.__mro__[:-1]:
            out.update(getattr(base, "_shared_fields", ()))
        return frozenset(out)

    # ---- dict views: extra_kwargs are flattened into the top level (samples.py:166-196)
    def to_dict(self) -> Dict[str, Any]:
        d = {f.name: getattr(self, f.name) for f in fields(self)}
        extra = d.pop("extra_kwargs", {})
        d.update(extra)
        return d

    @classmethod
    def from_dict(cls, d: Dict[str, Any]) -> "BaseSample":
        names = {f.name for f in fields(cls)}
        known = {k: v for k, v in d.items() if k in names and k != "extra_kwargs"}
        extra = {k: v for k, v in d.items() if k not in names}
        incoming = d.get("extra_kwargs", {})
        clash = set(incoming) & (names - {"extra_kwargs"})
        if clash:
            raise ValueError(f"extra_kwargs contains reserved field names: {clash}")
        extra.update(incoming)
        return cls(**known, extra_kwargs=extra)

    def __getattr__(self, key: str) -> Any:              # only reached when normal lookup fails: fall through to extra_kwargs
        try:
            extra = object.__getattribute__(self, "extra_kwargs")
        except AttributeError:
            raise AttributeError(f"'{type(self).__name__}' has no attribute '{key}'")
        if key in extra:
            return extra[key]
        raise AttributeError(f"'{type(self).__name__}' has no attribute '{key}'")

    def __setattr__(self, key: str, value: Any) -> None:
        if key in type(self)._id_fields:
            object.__setattr__(self, "_unique_id", None)   # identity changed: drop the cached id
        super().__setattr__(key, value)

    def keys(self):
        return self.to_dict().keys()

    def __getitem__(self, key: str) -> Any:
        try:
            return getattr(self, key)
        except AttributeError:
            raise KeyError(f"Key '{key}' not found in {self.__class__.__name__}")

    def __iter__(self):
        return iter(self.keys())

    def short_rep(self) -> Dict[str, Any]:
        return {k: (f"Tensor{tuple(v.shape)}" if isinstance(v, torch.Tensor) and v.numel() > 16 else v) for k, v in self.to_dict().items()}

    def to(self, device, depth: int = 1) -> "BaseSample":
        assert 0 <= depth <= 1, "Only depth 0 and 1 are supported."
        device = torch.device(device)
        for f in fields(self):
            v = getattr(self, f.name)
            if isinstance(v, torch.Tensor):
                setattr(self, f.name, v.to(device))
            elif depth == 1 and isinstance(v, list) and all(isinstance(t, torch.Tensor) for t in v):
                setattr(self, f.name, [t.to(device) for t in v])
        return self

    # ---- identity (samples.py:250-288): sha256 over the prompt (text, else ids) then the negative prompt (text, else ids)
    def _hash_id_fields(self, hasher) -> None:
        if self.prompt is not None:
            hasher.update(self.prompt.encode("utf-8"))
        elif self.prompt_ids is not None:
            hasher.update(self.prompt_ids.cpu().numpy().tobytes())
        if self.negative_prompt is not None:
            hasher.update(self.negative_prompt.encode("utf-8"))
        elif self.negative_prompt_ids is not None:
            hasher.update(self.negative_prompt_ids.cpu().numpy().tobytes())

    def compute_unique_id(self, num_bytes: int = 8) -> int:
        if not 1 <= num_bytes <= 32:
            raise ValueError(f"num_bytes must be in [1, 32] (sha256 digest), got {num_bytes}")
        h = hashlib.sha256()
        self._hash_id_fields(h)
        return int.from_bytes(h.digest()[:num_bytes], byteorder="big", signed=True)

    @property
    def unique_id(self) -> int:
        if self._unique_id is None:
            self._unique_id = self.compute_unique_id()
        return self._unique_id

    def reset_unique_id(self) -> None:
        self._unique_id = None

    # ---- collate (samples.py:290-375)
    @classmethod
    def _stack_values(cls, key: str, values: List[Any]):
        if not values:
            return values
        if all(v is None for v in values):
            return None
        first = values[0]
        if key in cls.shared_fields():
            return first
        if isinstance(first, torch.Tensor):
            return torch.stack(values) if all(v.shape == first.shape for v in values) else values
        if isinstance(first, dict):
            if all(isinstance(v, dict) for v in values):
                return {k: cls._stack_values(k, [v[k] for v in values]) for k in first.keys()}
            return values
        return values

    @classmethod
    def stack(cls, samples: List["BaseSample"]) -> Dict[str, Any]:
        if not samples:
            raise ValueError("No samples to stack.")
        sample_cls = type(samples[0])
        dicts = [s.to_dict() for s in samples]
        return {k: sample_cls._stack_values(k, [d[k] for d in dicts]) for k in dicts[0].keys()}


@dataclass
class SD3_5Sample(BaseSample):
    """Mirror of FF/models/stable_diffusion/sd3_5.py:50-58 (T2ISample + the pooled embeddings, no extra shared fields)."""
    _shared_fields: ClassVar[frozenset] = frozenset({})
    pooled_prompt_embeds: Optional[torch.Tensor] = None
    negative_pooled_prompt_embeds: Optional[torch.Tensor] = None


@dataclass
class Flux1Sample(BaseSample):
    """Mirror of FF/models/flux/flux1.py:53-59 (T2ISample + pooled_prompt_embeds, img_ids; img_ids is a shared field).
    all_latents rows are PACKED latents (T', Ni, 64)."""
    _shared_fields: ClassVar[frozenset] = frozenset({"img_ids"})
    pooled_prompt_embeds: Optional[torch.Tensor] = None
    img_ids: Optional[torch.Tensor] = None


@dataclass
class QwenImageSample(BaseSample):
    """Mirror of the reference's QwenImageSample (FF/models/qwen_image/qwen_image.py:54-61): T2I sample + embedding masks + img_shapes,
    no extra shared fields (img_shapes is collated per sample like any other list).  all_latents rows are PACKED latents (T', Ni, 64)."""
    _shared_fields: ClassVar[frozenset] = frozenset({})
    prompt_embeds_mask: Optional[torch.Tensor] = None
    negative_prompt_embeds_mask: Optional[torch.Tensor] = None
    img_shapes: Optional[List] = None


@dataclass
class WanT2VSample(BaseSample):
    """Mirror of FF/models/wan/wan2_t2v.py:47-50 (T2VSample, no extra shared fields): fills `video` instead of `image`;
    all_latents rows are (T', C, F, H, W)."""
    _shared_fields: ClassVar[frozenset] = frozenset({})
