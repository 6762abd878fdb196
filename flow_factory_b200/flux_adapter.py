"""B200Flux1Adapter - the drop-in for Flow-Factory's Flux1Adapter on the rollout path.

Mirrors FF/models/flux/flux1.py: `inference()` (152-292) and `forward()` (296-349); parameter names are the ABI
(`filter_kwargs`, FF/utils/base.py:38-63).  No CFG: FLUX.1-dev embeds the guidance scale (flux1.py:318-319).  Latents are the
packed (B, Ni, 64) tensors of the reference (`FluxPipeline.prepare_latents`), stored fp16; the schedule uses the
resolution-dependent shift `mu = calculate_shift(Ni)`.  The T-step loop runs inside the native engine with zero host
synchronisation; `forward()` with autograd stays on the reference's diffusers path; there is no CPU / PyTorch fallback."""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional, Union

import torch

from .flux import FluxRolloutEngine, model_scalar, pack_latents
from .per_sample import forward_grouped, split_by_timestep
from .rng import randn_tensor
from .stepwise import SUPPORTED_CALLBACKS, per_sample, run_stepwise
from .samples import Flux1Sample
from .scheduler import FlowMatchEulerDiscreteSDEScheduler, SDESchedulerOutput, set_scheduler_timesteps
from .trajectory import TrajectoryIndicesType, plan_slots
from .trajectory import create_callback_collector


class B200Flux1Adapter:
    def __init__(self, model_config, state_dict: Dict[str, torch.Tensor], device: Union[str, torch.device] = "cuda",
                 scheduler: Optional[FlowMatchEulerDiscreteSDEScheduler] = None, latent_storage_dtype: str = "fp16",
                 decode_fn: Optional[Callable[..., torch.Tensor]] = None, vae_scale_factor: int = 8, rng: str = "torch",
                 use_graph: bool = True):
        if latent_storage_dtype != "fp16":
            raise ValueError("the step kernel stores latents as fp16 (Flow-Factory's default latent_storage_dtype)")
        if rng not in ("torch", "philox"):
            raise ValueError("rng must be 'torch' (reference-identical noise stream) or 'philox' (in-kernel)")
        self.engine = FluxRolloutEngine(model_config, state_dict, torch.device(device))
        self.device = self.engine.device
        self.model_config = self.engine.cfg
        self.scheduler = scheduler or FlowMatchEulerDiscreteSDEScheduler(noise_level=0.7, use_dynamic_shifting=True,
                                                                         dynamics_type="Flow-SDE")
        self.decode_fn = decode_fn
        self.vae_scale_factor = vae_scale_factor
        self.rng = rng
        self.use_graph = use_graph

    def refresh_weights(self, state_dict: Dict[str, torch.Tensor]) -> None:
        self.engine.refresh_weights(state_dict)

    def rollout(self):
        self.scheduler.rollout()

    def train(self, mode: bool = True):       # FF/models/abc.py:372-378
        self.scheduler.train(mode=mode)

    def eval(self):
        self.scheduler.eval()

    def cast_latents(self, latents: torch.Tensor, default_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
        """FF/models/abc.py:172-182 without the per-call host sync."""
        if latents.dtype == torch.float16:
            return latents
        return latents.clamp(-65504.0, 65504.0).to(torch.float16)

    def decode_latents(self, latents: torch.Tensor, height: int, width: int, output_type: str = "pt"):
        if self.decode_fn is None:
            return None
        return self.decode_fn(latents, height, width)

    # -------------------------------------------------------------- the trajectory sampler (flux1.py:152-292)
    @torch.no_grad()
    def inference(
        self,
        prompt: Optional[Union[str, List[str]]] = None,
        height: int = 512,
        width: int = 512,
        num_inference_steps: int = 28,
        guidance_scale: float = 3.5,
        generator: Optional[torch.Generator] = None,
        prompt_ids: Optional[torch.Tensor] = None,
        prompt_embeds: Optional[torch.Tensor] = None,
        pooled_prompt_embeds: Optional[torch.Tensor] = None,
        joint_attention_kwargs: Optional[Dict[str, Any]] = None,
        compute_log_prob: bool = True,
        extra_call_back_kwargs: List[str] = [],
        trajectory_indices: TrajectoryIndicesType = "all",
        latents: Optional[torch.Tensor] = None,
        noise: Optional[torch.Tensor] = None,
    ) -> List[Flux1Sample]:
        if prompt_embeds is None or pooled_prompt_embeds is None:
            raise ValueError("B200Flux1Adapter.inference needs pre-encoded prompt_embeds / pooled_prompt_embeds")
        if joint_attention_kwargs:
            raise NotImplementedError("joint_attention_kwargs (IP-adapter / LoRA scale) are not on the accelerated path")
        unsupported = set(extra_call_back_kwargs) - SUPPORTED_CALLBACKS
        if unsupported:
            raise NotImplementedError(f"extra_call_back_kwargs {sorted(unsupported)} are not produced by the step kernel")
        dev = self.device
        T = int(num_inference_steps)
        B = len(prompt_embeds)
        # prepare_latents (pipeline_flux.py:596-630): latent grid 2*(h // (vae_scale*2)), packed 2x2
        lh = 2 * (int(height) // (self.vae_scale_factor * 2))
        lw = 2 * (int(width) // (self.vae_scale_factor * 2))
        h2, w2 = lh // 2, lw // 2
        plan = self.engine.plan(B, h2, w2, prompt_embeds.shape[1])
        self.engine.set_prompts(plan, prompt_embeds, pooled_prompt_embeds, float(guidance_scale), latents_dtype=torch.float16)
        if latents is None:
            latents = pack_latents(randn_tensor((B, self.model_config.in_channels // 4, lh, lw), generator=generator, device=dev,
                                               dtype=torch.bfloat16))
        x0 = self.cast_latents(latents.to(dev))
        sch = self.scheduler
        timesteps = set_scheduler_timesteps(sch, T, seq_len=plan.n_img)
        if extra_call_back_kwargs:
            # per-step callback values (GRPO-Guard's next_latents_mean, grpo.py:404): the reference's own loop over forward() (stepwise.py)
            res = run_stepwise(self, timesteps, x0, trajectory_indices, compute_log_prob, list(extra_call_back_kwargs),
                               dict(prompt_embeds=prompt_embeds, pooled_prompt_embeds=pooled_prompt_embeds, img_ids=plan.img_ids.to(dev),
                                    guidance_scale=guidance_scale), noise=noise)
            images = self.decode_latents(res["final"], height, width, output_type="pt")
            return [Flux1Sample(timesteps=timesteps, prompt=prompt[b] if isinstance(prompt, list) else prompt,
                                prompt_ids=prompt_ids[b] if prompt_ids is not None else None, prompt_embeds=prompt_embeds[b],
                                pooled_prompt_embeds=pooled_prompt_embeds[b], height=height, width=width,
                                image=images[b] if images is not None else None, img_ids=plan.img_ids.to(dev), **per_sample(res, b))
                    for b in range(B)]
        sde_now = set(sch.current_sde_steps.tolist())
        nls = [(sch.noise_level if (i in sde_now and not sch.is_eval) else 0.0) for i in range(T)]
        has_lp = [bool(compute_log_prob and nls[i] > 0) for i in range(T)]
        lat_slot, lp_slot, lat_map, lp_map = plan_slots(trajectory_indices, T, has_lp)
        if not compute_log_prob:
            lp_slot, lp_map = [-1] * T, None
        coefs = []
        for i in range(T):
            t, tn = timesteps[i], (timesteps[i + 1] if i + 1 < T else torch.tensor(0.0))
            coefs.append(sch.step_coef(t, tn, nls[i], compute_log_prob=has_lp[i], t_model=model_scalar(float(t) / 1000),
                                       store_slot=lat_slot[i + 1], logp_slot=lp_slot[i]))
        n_lat = sum(1 for s in lat_slot if s >= 0)
        n_lp = sum(1 for s in lp_slot if s >= 0)
        if noise is None and self.rng == "torch" and sch.dynamics_type != "ODE":
            noise = torch.stack([torch.randn(tuple(x0.shape), device=dev, dtype=torch.float32) for _ in range(T)])
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if noise is None else 0
        r = self.engine.rollout(plan, x0, coefs, n_lat, lat_slot[0], n_lp, noise=noise, seed=seed, use_graph=self.use_graph)
        final = r["final_latents"]
        images = self.decode_latents(final, height, width, output_type="pt")
        img_ids = plan.img_ids.to(dev)
        # as the reference: the callback gate's map even when no callback key was requested (all -1, or the identity for 'all')
        callback_index_map = create_callback_collector(trajectory_indices, T).get_index_map()
        samples = []
        for b in range(B):
            samples.append(Flux1Sample(
                timesteps=timesteps,
                all_latents=r["all_latents"][b, :n_lat] if n_lat else None,
                log_probs=(r["log_probs"][b, :n_lp] if n_lp else (torch.zeros(0, device=dev) if compute_log_prob and lp_map is not None else None)),
                latent_index_map=lat_map,
                log_prob_index_map=lp_map if compute_log_prob else None,
                prompt=prompt[b] if isinstance(prompt, list) else prompt,
                prompt_ids=prompt_ids[b] if prompt_ids is not None else None,
                prompt_embeds=prompt_embeds[b],
                pooled_prompt_embeds=pooled_prompt_embeds[b],
                height=height, width=width,
                image=images[b] if images is not None else None,
                img_ids=img_ids,
                extra_kwargs={"callback_index_map": callback_index_map, "final_latents": final[b]},
            ))
        self._last_overflow = r["overflow"]
        return samples

    # -------------------------------------------------------------- one denoise step (flux1.py:296-349)
    def forward(
        self,
        t: torch.Tensor,
        latents: torch.Tensor,
        prompt_embeds: torch.Tensor,
        pooled_prompt_embeds: torch.Tensor,
        img_ids: Optional[torch.Tensor] = None,
        t_next: Optional[torch.Tensor] = None,
        next_latents: Optional[torch.Tensor] = None,
        guidance_scale: Union[float, List[float]] = 3.5,
        noise_level: Optional[float] = None,
        joint_attention_kwargs: Optional[Dict[str, Any]] = None,
        compute_log_prob: bool = True,
        return_kwargs: List[str] = ["noise_pred", "next_latents", "next_latents_mean", "std_dev_t", "dt", "log_prob"],
        noise: Optional[torch.Tensor] = None,
        height: Optional[int] = None,
        width: Optional[int] = None,
    ) -> SDESchedulerOutput:
        if torch.is_grad_enabled() and any(isinstance(x, torch.Tensor) and x.requires_grad for x in (latents, prompt_embeds)):
            raise RuntimeError("B200Flux1Adapter.forward serves the no-grad path; keep the autograd replay on the reference adapter")
        if joint_attention_kwargs:
            raise NotImplementedError("joint_attention_kwargs are not on the accelerated path")
        if isinstance(guidance_scale, (list, tuple)):
            if len(set(float(g) for g in guidance_scale)) != 1:
                raise NotImplementedError("per-sample guidance scales are not on the accelerated path")
            guidance_scale = float(guidance_scale[0])
        B, Ni, _ = latents.shape
        groups = split_by_timestep(t, t_next, B)
        if groups is not None:       # per-sample timesteps (NFT / AWM / CRD): one engine call per distinct (t, t_next) - per_sample.py
            return forward_grouped(
                self.forward, groups, B,
                dict(latents=latents, prompt_embeds=prompt_embeds, pooled_prompt_embeds=pooled_prompt_embeds, img_ids=img_ids,
                     next_latents=next_latents, guidance_scale=guidance_scale, noise_level=noise_level, compute_log_prob=compute_log_prob,
                     return_kwargs=return_kwargs, noise=noise, height=height, width=width),
                batched=("latents", "prompt_embeds", "pooled_prompt_embeds", "next_latents", "noise"),
                make_output=SDESchedulerOutput.from_dict)
        if img_ids is not None:      # recover the token grid from the ids (rows, cols) = max + 1
            h2, w2 = int(img_ids[:, 1].max().item()) + 1, int(img_ids[:, 2].max().item()) + 1
        elif height is not None and width is not None:
            h2, w2 = int(height) // (self.vae_scale_factor * 2), int(width) // (self.vae_scale_factor * 2)
        else:
            raise ValueError("forward needs img_ids (or height / width) to know the latent grid")
        assert h2 * w2 == Ni, (h2, w2, Ni)
        plan = self.engine.plan(B, h2, w2, prompt_embeds.shape[1])
        self.engine.set_prompts(plan, prompt_embeds, pooled_prompt_embeds, float(guidance_scale), latents_dtype=latents.dtype)
        sch = self.scheduler
        t0 = (t if isinstance(t, torch.Tensor) else torch.tensor(float(t))).flatten()[0].detach().cpu().float()
        # t_next omitted: the reference's scheduler.step then reads sigmas[i], sigmas[i + 1] from its tables (step_coef mirrors that)
        tn = None if t_next is None else (t_next if isinstance(t_next, torch.Tensor) else torch.tensor(float(t_next))).flatten()[0].detach().cpu().float()
        coef = sch.step_coef(t0, tn, noise_level, compute_log_prob=compute_log_prob, t_model=model_scalar(float(t0) / 1000))
        if noise is None and next_latents is None and self.rng == "torch" and sch.dynamics_type != "ODE":
            noise = torch.randn(latents.shape, device=self.device, dtype=torch.float32)
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if noise is None else 0
        r = self.engine.step(plan, latents, coef, noise=noise, next_latents=next_latents, seed=seed)
        if next_latents is not None:
            nxt = next_latents.float()
        elif sch.dynamics_type == "ODE":
            nxt = r["next_latents_mean"]
        else:
            nxt = r["next_latents"].float()
        d = dict(next_latents=nxt, next_latents_mean=r["next_latents_mean"], log_prob=r["log_prob"], noise_pred=r["noise_pred"],
                 std_dev_t=torch.full((B, 1, 1), coef.std_dev_t, dtype=torch.float32, device=self.device),
                 dt=torch.full((B, 1, 1), coef.dt, dtype=torch.float32, device=self.device))
        return SDESchedulerOutput.from_dict({k: d[k] for k in return_kwargs if k in d})
