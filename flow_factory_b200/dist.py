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


# ------------------------------------------------------------------------------------------------ FSDP2-sharded weight intake
def _local_shard(t: torch.Tensor):
    """(local dim-0 shard, full shape) of a parameter: a DTensor sharded Shard(0) by FSDP2's fully_shard, or a plain (replicated) tensor."""
    try:
        from torch.distributed.tensor import DTensor, Shard
    except Exception:                      # pragma: no cover - torch without DTensor
        DTensor = None
    if DTensor is not None and isinstance(t, DTensor):
        if len(t.placements) != 1 or not isinstance(t.placements[0], Shard) or t.placements[0].dim != 0:
            raise NotImplementedError(f"only 1-D meshes with Shard(0) placements (FSDP2 fully_shard) are handled, got {t.placements}")
        return t.to_local(), tuple(t.shape)
    return None, tuple(t.shape)


def gather_sharded_state_dict(state_dict, group=None, dtype: Optional[torch.dtype] = torch.bfloat16):
    """FSDP2 (`fully_shard`) keeps every parameter as a DTensor sharded along dim 0 (rank r owns rows [r c, (r+1) c), c = ceil(n / W),
    the tail padded / empty - SURVEY.md section 8(e), BASELINE config 5).  The rollout engine wants replicated weights, and the path
    wants ONE collective per rollout, not one per block: all local shards are packed into a single flat buffer, gathered with one
    all_gather_into_tensor, and unpacked into full tensors (cast to `dtype` BEFORE the gather, so the payload is the bf16 model:
    40.9 GB for Qwen-Image 20 B; measured at N = 2 over NCCL with the first version of the packing - per-parameter staging buffers +
    torch.cat -: 207 ms per rollout, profiles/r02_bench_qwen_n2.json).  Plain tensors in the dict are passed through (already replicated)."""
    if not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("gather_sharded_state_dict needs an initialised process group")
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    plan, locals_ = [], []
    out = {}
    total, dev, dt = 0, None, None
    for name, t in state_dict.items():
        local, full = _local_shard(t)
        if local is None:
            out[name] = t if dtype is None or not t.is_floating_point() else t.to(dtype)
            continue
        n0 = full[0]
        chunk = -(-n0 // world)                                    # ceil: FSDP2's padded shard size along dim 0
        rest = 1
        for d in full[1:]:
            rest *= d
        plan.append((name, full, chunk, rest))
        locals_.append(local)
        total += chunk * rest
        if dtype is None and dt is not None and local.dtype != dt:
            raise NotImplementedError("gather_sharded_state_dict(dtype=None) needs one parameter dtype; pass the dtype the engine wants")
        dev, dt = local.device, (dtype or local.dtype)
    if not plan:
        return out
    # every local shard goes straight into its slice of ONE flat send buffer (cast on the way; no per-parameter staging buffers, no cat):
    # one copy kernel per parameter, and the tail padding is zeroed only where a shard is short (ranks past the tail hold fewer or no rows)
    flat = torch.empty(total, dtype=dt, device=dev)
    off = 0
    for (name, full, chunk, rest), local in zip(plan, locals_):
        sz, n = chunk * rest, local.numel()
        if n:
            flat[off: off + n].copy_(local.reshape(-1))
        if n < sz:
            flat[off + n: off + sz].zero_()
        off += sz
    gathered = torch.empty(world * flat.numel(), dtype=flat.dtype, device=flat.device)
    dist.all_gather_into_tensor(gathered, flat, group=group)       # the one collective
    gathered = gathered.view(world, flat.numel())
    off = 0
    for name, full, chunk, rest in plan:
        sz = chunk * rest
        out[name] = gathered[:, off: off + sz].reshape(world * chunk, *full[1:])[: full[0]].contiguous()
        off += sz
    return out
