"""Data-parallel plumbing of the rollout: prompts shard across ranks (the reference's K-repeat samplers already do
this, FF/data_utils/sampler.py:36-280; rollout itself runs with zero collectives, grpo.py:152-171).  The one
collective on the path is a single all-gather per rollout of the packed {kept latents | log-probs} record
(north-star) - one NCCL call over NVLink/NVSwitch, none inside the denoising loop."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of n_items for `rank` (balanced, first ranks take the remainder)."""
    base, rem = divmod(n_items, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_prompts(items: Sequence, rank: int, world_size: int):
    lo, hi = shard_range(len(items), rank, world_size)
    return items[lo:hi]


def pack_rollout(all_latents: Optional[torch.Tensor], log_probs: Optional[torch.Tensor]) -> Tuple[torch.Tensor, dict]:
    """[B, T', C,H,W] fp16 + [B, T''] fp32 -> one uint8 buffer [B, bytes_per_sample] (single collective payload)."""
    parts, meta = [], {}
    B = (all_latents if all_latents is not None else log_probs).shape[0]
    if all_latents is not None:
        meta["lat_shape"], meta["lat_dtype"] = tuple(all_latents.shape[1:]), all_latents.dtype
        parts.append(all_latents.contiguous().view(B, -1).view(torch.uint8))
    if log_probs is not None:
        meta["lp_shape"], meta["lp_dtype"] = tuple(log_probs.shape[1:]), log_probs.dtype
        parts.append(log_probs.contiguous().view(B, -1).view(torch.uint8))
    meta["sizes"] = [p.shape[1] for p in parts]
    return torch.cat(parts, dim=1).contiguous(), meta


def unpack_rollout(buf: torch.Tensor, meta: dict):
    outs, off = [], 0
    n = buf.shape[0]
    names = [k for k in ("lat", "lp") if f"{k}_shape" in meta]
    for name, size in zip(names, meta["sizes"]):
        chunk = buf[:, off: off + size].contiguous()
        outs.append(chunk.view(meta[f"{name}_dtype"]).view(n, *meta[f"{name}_shape"]))
        off += size
    res = dict(zip(names, outs))
    return res.get("lat"), res.get("lp")


def all_gather_rollout(all_latents: Optional[torch.Tensor], log_probs: Optional[torch.Tensor], group=None):
    """ONE all-gather of the packed per-rank rollout record; returns tensors with leading dim world_size * B."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return all_latents, log_probs
    buf, meta = pack_rollout(all_latents, log_probs)
    world = dist.get_world_size(group)
    out = torch.empty((world * buf.shape[0], buf.shape[1]), dtype=torch.uint8, device=buf.device)
    dist.all_gather_into_tensor(out, buf, group=group)
    return unpack_rollout(out, meta)
