"""The reference adapters' own denoising loop (e.g. FF/models/stable_diffusion/sd3_5.py:266-304, flux1.py:211-250) around a B200 adapter's
`forward()`: the path `inference()` takes when per-step callback values are requested (`extra_call_back_kwargs`; GRPO-Guard asks for
`next_latents_mean`, grpo.py:404) - the fused T-step rollout keeps only latents and log-probs.  Same kernels, one launch list per step
instead of a replayed CUDA graph; collectors are the mirrored ones of trajectory.py."""
from __future__ import annotations

from typing import Any, Dict, List, Optional

import torch

from .trajectory import TrajectoryIndicesType, create_callback_collector, create_trajectory_collector

SUPPORTED_CALLBACKS = frozenset({"noise_pred", "next_latents_mean", "noise_level", "next_latents", "log_prob", "std_dev_t", "dt"})


def run_stepwise(adapter, timesteps: torch.Tensor, x0: torch.Tensor, trajectory_indices: TrajectoryIndicesType, compute_log_prob: bool,
                 extra_call_back_kwargs: List[str], forward_kwargs: Dict[str, Any], noise: Optional[torch.Tensor] = None,
                 last_t_next: Optional[torch.Tensor] = None) -> Dict[str, Any]:
    """Returns dict(final, all_latents (list of (B, ...) or None), all_log_probs, latent_index_map, log_prob_index_map, extra (dict key ->
    (B, T', ...) tensor or list), callback_index_map)."""
    T = len(timesteps)
    sch = adapter.scheduler
    latent_collector = create_trajectory_collector(trajectory_indices, T)
    lat = adapter.cast_latents(x0)
    latent_collector.collect(lat, step_idx=0)
    log_prob_collector = create_trajectory_collector(trajectory_indices, T) if compute_log_prob else None
    callback_collector = create_callback_collector(trajectory_indices, T)
    zero = last_t_next if last_t_next is not None else torch.zeros((), dtype=timesteps.dtype)
    for i in range(T):
        t = timesteps[i]
        current_noise_level = 0.0 if sch.is_eval else sch.get_noise_level_for_timestep(t)
        t_next = timesteps[i + 1] if i + 1 < T else zero
        current_compute_log_prob = bool(compute_log_prob and current_noise_level > 0)
        return_kwargs = list(set(["next_latents", "log_prob", "noise_pred"] + list(extra_call_back_kwargs)))
        output = adapter.forward(t=t, t_next=t_next, latents=lat, compute_log_prob=current_compute_log_prob, return_kwargs=return_kwargs,
                                 noise_level=current_noise_level, noise=None if noise is None else noise[i], **forward_kwargs)
        lat = adapter.cast_latents(output.next_latents)
        latent_collector.collect(lat, i + 1)
        if current_compute_log_prob:
            log_prob_collector.collect(output.log_prob, i)
        callback_collector.collect_step(step_idx=i, output=output, keys=extra_call_back_kwargs, capturable={"noise_level": current_noise_level})
    return dict(final=lat, all_latents=latent_collector.get_result(), latent_index_map=latent_collector.get_index_map(),
                all_log_probs=log_prob_collector.get_result() if compute_log_prob else None,
                log_prob_index_map=log_prob_collector.get_index_map() if compute_log_prob else None,
                extra=callback_collector.get_result(), callback_index_map=callback_collector.get_index_map())


def per_sample(res: Dict[str, Any], b: int) -> Dict[str, Any]:
    """The per-sample slices the reference adapters put into their sample records (sd3_5.py:318-343)."""
    al, lp = res["all_latents"], res["all_log_probs"]
    return dict(all_latents=torch.stack([x[b] for x in al], dim=0) if al is not None else None,
                log_probs=torch.stack([x[b] for x in lp], dim=0) if lp is not None else None,
                latent_index_map=res["latent_index_map"], log_prob_index_map=res["log_prob_index_map"],
                extra_kwargs={**{k: v[b] for k, v in res["extra"].items()}, "callback_index_map": res["callback_index_map"],
                              "final_latents": res["final"][b]})
