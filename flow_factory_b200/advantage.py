"""Group-normalised advantages - host mirror of AdvantageProcessor.compute_advantages' math
(FF/advantage/advantage_processor.py:314-397 `sum`, 403-481 `gdpo`) in numpy fp64.  O(samples) scalars: no kernel.
The trainers keep calling the reference class; this mirror exists so the rollout output can be consumed stand-alone
(bench / tests) with the same arithmetic."""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np


def _group_index(unique_ids: Sequence[int]):
    ids = np.asarray(unique_ids)
    _, inv = np.unique(ids, return_inverse=True)
    return inv


def advantages_sum(rewards: Dict[str, np.ndarray], weights: Dict[str, float], unique_ids: Sequence[int],
                   global_std: bool = True, eps: float = 1e-6) -> np.ndarray:
    """`sum` aggregation: r = sum_k w_k r_k ; A = (r - mean_group) / std, std global over the batch (>= eps) or per group."""
    keys = list(rewards)
    r = np.zeros(len(unique_ids), dtype=np.float64)
    for k in keys:
        r += float(weights.get(k, 1.0)) * np.asarray(rewards[k], dtype=np.float64)
    g = _group_index(unique_ids)
    adv = np.zeros_like(r)
    gstd = max(float(r.std()), eps)
    for gi in np.unique(g):
        m = g == gi
        mu = r[m].mean()
        sd = gstd if global_std else max(float(r[m].std()), eps)
        adv[m] = (r[m] - mu) / sd
    return adv


def advantages_gdpo(rewards: Dict[str, np.ndarray], weights: Dict[str, float], unique_ids: Sequence[int],
                    eps: float = 1e-6) -> np.ndarray:
    """`gdpo` aggregation: per-reward per-group z-score times weight, summed, then batch-normalised."""
    g = _group_index(unique_ids)
    total = np.zeros(len(unique_ids), dtype=np.float64)
    for k, v in rewards.items():
        v = np.asarray(v, dtype=np.float64)
        z = np.zeros_like(v)
        for gi in np.unique(g):
            m = g == gi
            z[m] = (v[m] - v[m].mean()) / max(float(v[m].std()), eps)
        total += float(weights.get(k, 1.0)) * z
    return (total - total.mean()) / max(float(total.std()), eps)
