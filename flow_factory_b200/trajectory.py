"""Host bookkeeping of which trajectory positions / step outputs are kept (mirror of
FF/utils/trajectory_collector.py:40-180, 344-388).  The engine writes kept latents / log-probs straight into
compact device buffers; these helpers decide the slots and build the dense index maps (-1 = not stored)."""
from __future__ import annotations

from typing import List, Optional, Sequence, Set, Union

import torch

TrajectoryIndicesType = Union[str, None, List[int]]


def normalize_indices(indices: TrajectoryIndicesType, total_steps: int) -> Optional[Set[int]]:
    """None -> collect everything ('all'); empty set -> disabled (indices=None); else normalised positions."""
    if indices is None:
        return set()
    if isinstance(indices, str):
        if indices == "all":
            return None
        raise ValueError(f"unknown trajectory_indices {indices!r}")
    total_positions = total_steps + 1
    out = set()
    for idx in indices:
        idx = int(idx)
        if idx < 0:
            idx += total_positions
        if 0 <= idx < total_positions:
            out.add(idx)
    return out


class TrajectoryCollector:
    """Same contract as the reference collector, but `plan_slots` lets the engine pre-assign compact slots."""

    def __init__(self, indices: TrajectoryIndicesType = "all", total_steps: int = 0):
        self.indices, self.total_steps = indices, total_steps
        self._target = normalize_indices(indices, total_steps)
        self._collected: List[torch.Tensor] = []
        self._collected_indices: List[int] = []

    @property
    def is_disabled(self) -> bool:
        return self._target is not None and len(self._target) == 0

    @property
    def collect_all(self) -> bool:
        return self._target is None

    def should_collect(self, step_idx: int) -> bool:
        if self.is_disabled:
            return False
        return True if self.collect_all else step_idx in self._target

    def collect(self, value, step_idx: int) -> None:
        if self.should_collect(step_idx):
            self._collected.append(value)
            self._collected_indices.append(step_idx)

    def get_result(self):
        return None if self.is_disabled else self._collected

    @property
    def collected_indices(self) -> List[int]:
        return self._collected_indices

    def get_index_map(self) -> Optional[torch.Tensor]:
        if self.is_disabled:
            return None
        total_positions = self.total_steps + 1
        if self.collect_all:
            return torch.arange(total_positions, dtype=torch.long)
        m = torch.full((total_positions,), -1, dtype=torch.long)
        for compact, orig in enumerate(self._collected_indices):
            m[orig] = compact
        return m

    def reset(self) -> None:
        self._collected, self._collected_indices = [], []

    def __len__(self) -> int:
        return len(self._collected)


def compute_trajectory_indices(train_timestep_indices, num_inference_steps: int, include_initial: bool = False) -> List[int]:
    """FF/utils/trajectory_collector.py:344-388."""
    if isinstance(train_timestep_indices, torch.Tensor):
        train_timestep_indices = train_timestep_indices.tolist()
    total_positions = num_inference_steps + 1
    pos = set()
    if include_initial:
        pos.add(0)
    for idx in train_timestep_indices:
        if 0 <= idx < total_positions:
            pos.add(idx)
        if 0 <= idx + 1 < total_positions:
            pos.add(idx + 1)
    return sorted(pos)


def create_trajectory_collector(indices: TrajectoryIndicesType, num_steps: int) -> TrajectoryCollector:
    return TrajectoryCollector(indices=indices, total_steps=num_steps)


class CallbackCollector:
    """Mirror of FF/utils/trajectory_collector.py:187-337: named per-step values (`extra_call_back_kwargs`, e.g. GRPO-Guard's
    `next_latents_mean`, grpo.py:404) recorded at the gated steps; each key is resolved from `capturable` first, then from the step output."""

    def __init__(self, indices: TrajectoryIndicesType = "all", total_steps: int = 0):
        self._gate = TrajectoryCollector(indices=indices, total_steps=total_steps)
        self._data: dict = {}
        self._collected_indices: List[int] = []
        self._collected_set: Set[int] = set()

    @property
    def is_disabled(self) -> bool:
        return self._gate.is_disabled

    def should_collect(self, step_idx: int) -> bool:
        return self._gate.should_collect(step_idx)

    def collect_step(self, step_idx: int, output, keys: Sequence[str], capturable: Optional[dict] = None) -> None:
        if not keys or not self.should_collect(step_idx):
            return
        if step_idx not in self._collected_set:
            self._collected_indices.append(step_idx)
            self._collected_set.add(step_idx)
        for key in keys:
            val = None
            if capturable and key in capturable and capturable[key] is not None:
                val = capturable[key]
            elif hasattr(output, key):
                val = getattr(output, key)
            if val is not None:
                self._data.setdefault(key, []).append(val)

    def get_result(self) -> dict:
        """Tensor lists are stacked batch-first: list of (B, ...) -> (B, T', ...); other values stay lists."""
        return {k: (torch.stack(v, dim=1) if v and isinstance(v[0], torch.Tensor) else v) for k, v in self._data.items()}

    def get_index_map(self) -> Optional[torch.Tensor]:
        if self.is_disabled:
            return None
        total = self._gate.total_steps
        if self._gate.collect_all:
            return torch.arange(total, dtype=torch.long)
        m = torch.full((total,), -1, dtype=torch.long)
        for compact, orig in enumerate(self._collected_indices):
            if 0 <= orig < total:
                m[orig] = compact
        return m

    @property
    def collected_indices(self) -> List[int]:
        return self._collected_indices

    def reset(self) -> None:
        self._data, self._collected_indices, self._collected_set = {}, [], set()

    def __len__(self) -> int:
        return len(self._collected_indices)


def create_callback_collector(indices: TrajectoryIndicesType, num_steps: int) -> CallbackCollector:
    return CallbackCollector(indices=indices, total_steps=num_steps)


def plan_slots(indices: TrajectoryIndicesType, num_steps: int, step_has_logp: Sequence[bool]):
    """Slot assignment for a T-step rollout, identical to what the reference's two collectors would produce
    (sd3_5.py:266-304): latent position p in [0, T] and log-prob of step i are kept iff the index gate admits them.
    Returns (latent_slot[T+1], logp_slot[T], latent_index_map, log_prob_index_map) with -1 = not stored."""
    gate = TrajectoryCollector(indices, num_steps)
    lat_slot, n = [], 0
    for p in range(num_steps + 1):
        if gate.should_collect(p):
            lat_slot.append(n); n += 1
        else:
            lat_slot.append(-1)
    lp_slot, m = [], 0
    for i in range(num_steps):
        if step_has_logp[i] and gate.should_collect(i):
            lp_slot.append(m); m += 1
        else:
            lp_slot.append(-1)
    if gate.is_disabled:
        return lat_slot, lp_slot, None, None
    lat_map = torch.tensor(lat_slot, dtype=torch.long)
    lp_map = torch.full((num_steps + 1,), -1, dtype=torch.long)
    for i, s in enumerate(lp_slot):
        lp_map[i] = s
    if gate.collect_all:
        lp_map = torch.arange(num_steps + 1, dtype=torch.long)   # reference returns the identity map when collect_all
    return lat_slot, lp_slot, lat_map, lp_map
