"""Per-sample timesteps in the no-grad `forward()` of the adapters.

The reference's `forward` accepts `t` / `t_next` either as scalars (the rollout loop and GRPO's replay, sd3_5.py:273-304, grpo.py:242-263)
or as `(B,)` tensors: NFT / AWM / CRD draw one continuous timestep PER SAMPLE (nft.py:296-304, 366-374) and the reference expands and
uses them per row (sd3_5.py:394; flow_match_euler_discrete.py:302-318 handles 1-D timesteps).  The native step takes one timestep and one
set of step coefficients per launch list, so a batch with several distinct `(t, t_next)` pairs is served as one engine call per distinct
pair over the rows that share it, and the per-row results are scattered back in the caller's order.  A single sample at 1024^2 is already
>= 4429 GEMM rows (x2 under CFG), so the sub-batches stay tensor-core sized; no row is ever evaluated at another row's timestep.
"""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import torch


def _as_rows(v, B: int) -> Optional[torch.Tensor]:
    """`v` as a (B,) fp32 CPU tensor when it holds one value per sample, else None (scalar / 1-element: uniform)."""
    if v is None or not isinstance(v, torch.Tensor) or v.numel() <= 1:
        return None
    flat = v.detach().reshape(-1).float().cpu()
    if flat.numel() != B:
        raise ValueError(f"timestep tensor with {flat.numel()} entries for a batch of {B}")
    return flat


def split_by_timestep(t, t_next, B: int) -> Optional[List[Tuple[torch.Tensor, torch.Tensor, Optional[torch.Tensor]]]]:
    """None when the whole batch shares one (t, t_next); else [(row indices, t scalar, t_next scalar | None)], first-occurrence order."""
    tr, nr = _as_rows(t, B), _as_rows(t_next, B)
    if tr is None and nr is None:
        return None
    if tr is None:
        tr = (t if isinstance(t, torch.Tensor) else torch.tensor(float(t))).detach().reshape(-1)[:1].float().cpu().expand(B)
    if nr is None and t_next is not None:
        nr = (t_next if isinstance(t_next, torch.Tensor) else torch.tensor(float(t_next))).detach().reshape(-1)[:1].float().cpu().expand(B)
    keys: Dict[Tuple[float, Optional[float]], List[int]] = {}
    for b in range(B):
        keys.setdefault((float(tr[b]), None if nr is None else float(nr[b])), []).append(b)
    if len(keys) == 1:
        return None
    return [(torch.tensor(rows, dtype=torch.long), torch.tensor(k[0], dtype=torch.float32),
             None if k[1] is None else torch.tensor(k[1], dtype=torch.float32)) for k, rows in keys.items()]


def forward_grouped(forward: Callable[..., Any], groups, B: int, kwargs: Dict[str, Any], batched: Sequence[str], make_output: Callable[[dict], Any]):
    """Runs `forward` once per group on the rows of that group and reassembles a batch-ordered output.
    `batched`: names of keyword arguments that carry one row per sample (sliced with the group's indices); everything else is passed
    through.  `make_output(dict)` builds the adapter's output type (SDESchedulerOutput.from_dict)."""
    parts: List[Tuple[torch.Tensor, Any]] = []
    for rows, tv, tnv in groups:
        kw = dict(kwargs)
        for name in batched:
            v = kw.get(name)
            if isinstance(v, torch.Tensor) and v.dim() > 0 and v.shape[0] == B:
                kw[name] = v[rows.to(v.device)]
            elif isinstance(v, (list, tuple)) and len(v) == B:
                kw[name] = [v[int(i)] for i in rows]
        kw["t"], kw["t_next"] = tv, tnv
        parts.append((rows, forward(**kw)))
    fields: Dict[str, torch.Tensor] = {}
    first = parts[0][1]
    names = [k for k in first.keys()] if hasattr(first, "keys") else [k for k, v in vars(first).items() if v is not None]
    for name in names:
        sample = first[name] if hasattr(first, "__getitem__") else getattr(first, name)
        if not isinstance(sample, torch.Tensor):
            continue
        out = torch.empty((B,) + tuple(sample.shape[1:]), dtype=sample.dtype, device=sample.device)
        for rows, o in parts:
            v = o[name] if hasattr(o, "__getitem__") else getattr(o, name)
            out[rows.to(out.device)] = v
        fields[name] = out
    return make_output(fields)
