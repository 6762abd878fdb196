"""`extra_call_back_kwargs` path of B200SD3_5Adapter.inference (the reference's step loop over forward(), used by GRPO-Guard for
`next_latents_mean`): with the same per-step noise it must reproduce the fused T-step rollout bit for bit, and the collected means must be
the ones forward() returns.  First green run on a B200: round 2."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from flow_factory_b200.adapter import B200SD3_5Adapter                        # noqa: E402
from flow_factory_b200.scheduler import FlowMatchEulerDiscreteSDEScheduler    # noqa: E402
from flow_factory_b200.trajectory import compute_trajectory_indices           # noqa: E402
from oracle import sd3_oracle as O                                            # noqa: E402


def test_stepwise_equals_fused_rollout_and_collects_means():
    cfg = O.tiny_config()
    w32 = O.make_weights(cfg, seed=0)
    inp = {k: v.cuda() for k, v in O.make_inputs(cfg, 2, 16, 16, 13, seed=2).items()}
    sch = FlowMatchEulerDiscreteSDEScheduler(noise_level=0.7, shift=3.0, num_sde_steps=2, seed=5)
    ad = B200SD3_5Adapter(cfg, w32, scheduler=sch, rng="torch")
    ad.rollout()
    T = 6
    sch.set_timesteps(T, seq_len=64)
    idx = compute_trajectory_indices(sch.train_timesteps.tolist(), T)
    noise = torch.randn(T, 2, 16, 16, 16, generator=torch.Generator().manual_seed(4)).cuda()
    kw = dict(height=128, width=128, num_inference_steps=T, compute_log_prob=True, trajectory_indices=idx, latents=inp["x0"].bfloat16(),
              prompt_embeds=inp["prompt_embeds"], pooled_prompt_embeds=inp["pooled"], negative_prompt_embeds=inp["neg_prompt_embeds"],
              negative_pooled_prompt_embeds=inp["neg_pooled"], guidance_scale=4.5, noise=noise)
    fused = ad.inference(**kw)
    step = ad.inference(extra_call_back_kwargs=["next_latents_mean"], **kw)
    for a, b in zip(fused, step):
        assert torch.equal(a.all_latents, b.all_latents) and torch.equal(a.latent_index_map.cpu(), b.latent_index_map.cpu())
        torch.testing.assert_close(a.log_probs, b.log_probs, rtol=1e-6, atol=1e-6)
        assert torch.equal(a.final_latents, b.final_latents)
    s0 = step[0]
    cmap = s0.callback_index_map
    assert cmap.shape == (T,) and s0.next_latents_mean.dtype == torch.float32
    i = int((cmap >= 0).nonzero()[0])
    out = ad.forward(t=sch.timesteps[i], t_next=sch.timesteps[i + 1], latents=s0.all_latents[int(s0.latent_index_map[i])][None].repeat(2, 1, 1, 1),
                     prompt_embeds=inp["prompt_embeds"], pooled_prompt_embeds=inp["pooled"], negative_prompt_embeds=inp["neg_prompt_embeds"],
                     negative_pooled_prompt_embeds=inp["neg_pooled"], guidance_scale=4.5, noise_level=sch.get_noise_level_for_timestep(sch.timesteps[i]),
                     compute_log_prob=False, noise=noise[i])
    assert out.next_latents_mean.shape[1:] == s0.next_latents_mean.shape[1:]
