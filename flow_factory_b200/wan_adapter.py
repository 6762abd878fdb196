"""B200Wan21Adapter - the drop-in for Flow-Factory's Wan2_T2V_Adapter on the rollout path (BASELINE config 4).

Mirrors FF/models/wan/wan2_t2v.py: `inference()` (235-420) and `forward()` (425-543); parameter names are the ABI (`filter_kwargs`,
FF/utils/base.py:38-63).  True CFG runs as one forward batch of 2B (negative prompts first) instead of the reference's two sequential
forwards; the schedule is the UniPC flow-sigma schedule with INTEGER timesteps (scheduler.UniPCMultistepSDEScheduler) and the step is
the Euler / SDE arithmetic shared with the other models.  Wan2.2's second transformer / boundary_timestep / expand_timesteps are not
on the accelerated path (errors, not approximations).  STATUS: first GPU run pending (see wan.py)."""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional, Union

import torch

from .per_sample import forward_grouped, split_by_timestep
from .rng import randn_tensor
from .stepwise import SUPPORTED_CALLBACKS, per_sample, run_stepwise
from .samples import WanT2VSample
from .scheduler import SDESchedulerOutput, UniPCMultistepSDEScheduler
from .trajectory import TrajectoryIndicesType, plan_slots
from .trajectory import create_callback_collector
from .wan import WanRolloutEngine


class B200Wan21Adapter:
    def __init__(self, model_config, state_dict: Dict[str, torch.Tensor], device: Union[str, torch.device] = "cuda",
                 scheduler: Optional[UniPCMultistepSDEScheduler] = None, latent_storage_dtype: str = "fp16",
                 decode_fn: Optional[Callable[..., torch.Tensor]] = None, vae_scale_factor_spatial: int = 8,
                 vae_scale_factor_temporal: int = 4, rng: str = "torch", use_graph: bool = True):
        if latent_storage_dtype != "fp16":
            raise ValueError("the step kernel stores latents as fp16 (Flow-Factory's default latent_storage_dtype)")
        if rng not in ("torch", "philox"):
            raise ValueError("rng must be 'torch' (reference-identical noise stream) or 'philox' (in-kernel)")
        self.engine = WanRolloutEngine(model_config, state_dict, torch.device(device))
        self.device = self.engine.device
        self.model_config = self.engine.cfg
        self.scheduler = scheduler or UniPCMultistepSDEScheduler(noise_level=0.7, flow_shift=3.0, dynamics_type="Flow-SDE")
        self.decode_fn = decode_fn
        self.vae_scale_factor_spatial, self.vae_scale_factor_temporal = vae_scale_factor_spatial, vae_scale_factor_temporal
        self.rng, self.use_graph = rng, use_graph

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

    def decode_latents(self, latents: torch.Tensor, output_type: str = "pt"):
        return None if self.decode_fn is None else self.decode_fn(latents)

    def latent_shape(self, batch: int, height: int, width: int, num_frames: int):
        """WanPipeline.prepare_latents: (B, C, (frames - 1) // 4 + 1, H // 8, W // 8)."""
        return (batch, self.model_config.in_channels, (int(num_frames) - 1) // self.vae_scale_factor_temporal + 1,
                int(height) // self.vae_scale_factor_spatial, int(width) // self.vae_scale_factor_spatial)

    # -------------------------------------------------------------- the trajectory sampler (wan2_t2v.py:235-420)
    @torch.no_grad()
    def inference(
        self,
        prompt: Optional[Union[str, List[str]]] = None,
        negative_prompt: Optional[Union[str, List[str]]] = None,
        height: int = 480,
        width: int = 832,
        num_frames: int = 81,
        num_inference_steps: int = 50,
        guidance_scale: float = 5.0,
        guidance_scale_2: Optional[float] = None,
        generator: Optional[torch.Generator] = None,
        prompt_ids: Optional[torch.Tensor] = None,
        prompt_embeds: Optional[torch.Tensor] = None,
        negative_prompt_ids: Optional[torch.Tensor] = None,
        negative_prompt_embeds: Optional[torch.Tensor] = None,
        compute_log_prob: bool = False,
        attention_kwargs: Optional[Dict[str, Any]] = None,
        max_sequence_length: int = 512,
        extra_call_back_kwargs: List[str] = [],
        trajectory_indices: TrajectoryIndicesType = "all",
        latents: Optional[torch.Tensor] = None,
        noise: Optional[torch.Tensor] = None,
    ) -> List[WanT2VSample]:
        if prompt_embeds is None:
            raise ValueError("B200Wan21Adapter.inference needs pre-encoded prompt_embeds")
        if attention_kwargs:
            raise NotImplementedError("attention_kwargs are not on the accelerated path")
        unsupported = set(extra_call_back_kwargs) - SUPPORTED_CALLBACKS
        if unsupported:
            raise NotImplementedError(f"extra_call_back_kwargs {sorted(unsupported)} are not produced by the step kernel")
        if guidance_scale_2 is not None:
            raise NotImplementedError("guidance_scale_2 / the second transformer of Wan2.2 are not on the accelerated path")
        if self.scheduler.is_eval:
            # in eval mode the reference's UniPCMultistepSDEScheduler.step hands over to diffusers' multistep predictor-corrector
            # (unipc_multistep.py:283-285), not the Euler / SDE arithmetic of the rollout: outside this engine
            raise NotImplementedError("evaluation-mode sampling (UniPC multistep solver) is not on the accelerated path; use the reference adapter")
        dev = self.device
        T, B = int(num_inference_steps), len(prompt_embeds)
        do_cfg = guidance_scale > 1.0 and negative_prompt_embeds is not None      # wan2_t2v.py:487-495
        shape = self.latent_shape(B, height, width, num_frames)
        plan = self.engine.plan(B, shape[2], shape[3], shape[4], prompt_embeds.shape[1], cfg=do_cfg)
        self.engine.set_prompts(plan, prompt_embeds, negative_prompt_embeds if do_cfg else None)
        if latents is None:
            latents = randn_tensor(shape, generator=generator, device=dev, dtype=torch.float32)      # prepare_latents(dtype=float32)
        x0 = self.cast_latents(latents.to(dev))
        sch = self.scheduler
        timesteps = sch.set_timesteps(T)
        if extra_call_back_kwargs:
            res = run_stepwise(self, timesteps, x0, trajectory_indices, compute_log_prob, list(extra_call_back_kwargs),
                               dict(prompt_embeds=prompt_embeds, negative_prompt_embeds=negative_prompt_embeds, guidance_scale=guidance_scale),
                               noise=noise, last_t_next=torch.tensor(0))
            videos = self.decode_latents(res["final"], output_type="pt")
            return [WanT2VSample(timesteps=timesteps, video=videos[b] if videos is not None else None, height=height, width=width,
                                 prompt=prompt[b] if isinstance(prompt, list) else prompt,
                                 prompt_ids=prompt_ids[b] if prompt_ids is not None else None, prompt_embeds=prompt_embeds[b],
                                 negative_prompt=negative_prompt[b] if isinstance(negative_prompt, list) else negative_prompt,
                                 negative_prompt_ids=negative_prompt_ids[b] if negative_prompt_ids is not None else None,
                                 negative_prompt_embeds=negative_prompt_embeds[b] if negative_prompt_embeds is not None else None,
                                 **per_sample(res, b)) for b in range(B)]
        sde_now = set(sch.current_sde_steps.tolist())
        nls = [(sch.noise_level if (i in sde_now and not sch.is_eval) else 0.0) for i in range(T)]
        has_lp = [bool(compute_log_prob and nls[i] > 0) for i in range(T)]
        lat_slot, lp_slot, lat_map, lp_map = plan_slots(trajectory_indices, T, has_lp)
        if not compute_log_prob:
            lp_slot, lp_map = [-1] * T, None
        coefs = []
        for i in range(T):
            t, tn = timesteps[i], (timesteps[i + 1] if i + 1 < T else torch.tensor(0))
            coefs.append(sch.step_coef(t, tn, nls[i], compute_log_prob=has_lp[i], t_model=float(t), store_slot=lat_slot[i + 1],
                                       logp_slot=lp_slot[i]))
        n_lat = sum(1 for s in lat_slot if s >= 0)
        n_lp = sum(1 for s in lp_slot if s >= 0)
        if noise is None and self.rng == "torch" and sch.dynamics_type != "ODE":
            noise = torch.stack([torch.randn(tuple(x0.shape), device=dev, dtype=torch.float32) for _ in range(T)])
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if noise is None else 0
        r = self.engine.rollout(plan, x0, coefs, float(guidance_scale) if do_cfg else 1.0, n_lat, lat_slot[0], n_lp, noise=noise, seed=seed,
                                use_graph=self.use_graph)
        final = r["final_latents"]
        videos = self.decode_latents(final, output_type="pt")
        # as the reference: the callback gate's map even when no callback key was requested (all -1, or the identity for 'all')
        callback_index_map = create_callback_collector(trajectory_indices, T).get_index_map()
        samples = []
        for b in range(B):
            samples.append(WanT2VSample(
                timesteps=timesteps,
                all_latents=r["all_latents"][b, :n_lat] if n_lat else None,
                log_probs=(r["log_probs"][b, :n_lp] if n_lp else (torch.zeros(0, device=dev) if compute_log_prob and lp_map is not None else None)),
                latent_index_map=lat_map,
                log_prob_index_map=lp_map if compute_log_prob else None,
                video=videos[b] if videos is not None else None,
                height=height, width=width,
                prompt=prompt[b] if isinstance(prompt, list) else prompt,
                prompt_ids=prompt_ids[b] if prompt_ids is not None else None,
                prompt_embeds=prompt_embeds[b],
                negative_prompt=negative_prompt[b] if isinstance(negative_prompt, list) else negative_prompt,
                negative_prompt_ids=negative_prompt_ids[b] if negative_prompt_ids is not None else None,
                negative_prompt_embeds=negative_prompt_embeds[b] if negative_prompt_embeds is not None else None,
                extra_kwargs={"callback_index_map": callback_index_map, "final_latents": final[b]},
            ))
        self._last_overflow = r["overflow"]
        return samples

    # -------------------------------------------------------------- one denoise step (wan2_t2v.py:425-543)
    def forward(
        self,
        t: torch.Tensor,
        latents: torch.Tensor,
        prompt_embeds: torch.Tensor,
        negative_prompt_embeds: Optional[torch.Tensor] = None,
        guidance_scale: float = 5.0,
        guidance_scale_2: Optional[float] = None,
        t_next: Optional[torch.Tensor] = None,
        next_latents: Optional[torch.Tensor] = None,
        noise_level: Optional[float] = None,
        attention_kwargs: Optional[Dict[str, Any]] = None,
        compute_log_prob: bool = True,
        return_kwargs: List[str] = ["noise_pred", "next_latents", "next_latents_mean", "std_dev_t", "dt", "log_prob"],
        boundary_timestep: Optional[float] = None,
        noise: Optional[torch.Tensor] = None,
    ) -> SDESchedulerOutput:
        if torch.is_grad_enabled() and any(isinstance(x, torch.Tensor) and x.requires_grad for x in (latents, prompt_embeds)):
            raise RuntimeError("B200Wan21Adapter.forward serves the no-grad path; keep the autograd replay on the reference adapter")
        if attention_kwargs:
            raise NotImplementedError("attention_kwargs are not on the accelerated path")
        if boundary_timestep is not None or guidance_scale_2 is not None:
            raise NotImplementedError("boundary_timestep / guidance_scale_2 (Wan2.2) are not on the accelerated path")
        if self.scheduler.is_eval:
            raise NotImplementedError("evaluation-mode stepping (UniPC multistep solver) is not on the accelerated path; use the reference adapter")
        B, _, Fr, H, W = latents.shape
        do_cfg = negative_prompt_embeds is not None and guidance_scale > 1.0
        groups = split_by_timestep(t, t_next, B)
        if groups is not None:       # per-sample timesteps (NFT / AWM / CRD): one engine call per distinct (t, t_next) - per_sample.py
            return forward_grouped(
                self.forward, groups, B,
                dict(latents=latents, prompt_embeds=prompt_embeds, negative_prompt_embeds=negative_prompt_embeds, guidance_scale=guidance_scale,
                     next_latents=next_latents, noise_level=noise_level, compute_log_prob=compute_log_prob, return_kwargs=return_kwargs,
                     noise=noise),
                batched=("latents", "prompt_embeds", "negative_prompt_embeds", "next_latents", "noise"),
                make_output=SDESchedulerOutput.from_dict)
        plan = self.engine.plan(B, Fr, H, W, prompt_embeds.shape[1], cfg=do_cfg)
        self.engine.set_prompts(plan, prompt_embeds, negative_prompt_embeds if do_cfg else None)
        sch = self.scheduler
        t0 = (t if isinstance(t, torch.Tensor) else torch.tensor(t)).flatten()[0].detach().cpu()
        # t_next omitted: the reference's scheduler.step then reads sigmas[i], sigmas[i + 1] from its tables (step_coef mirrors that)
        tn = None if t_next is None else (t_next if isinstance(t_next, torch.Tensor) else torch.tensor(t_next)).flatten()[0].detach().cpu()
        coef = sch.step_coef(t0, tn, noise_level, compute_log_prob=compute_log_prob, t_model=float(t0))
        if noise is None and next_latents is None and self.rng == "torch" and sch.dynamics_type != "ODE":
            noise = torch.randn(latents.shape, device=self.device, dtype=torch.float32)
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if noise is None else 0
        r = self.engine.step(plan, latents, coef, float(guidance_scale) if do_cfg else 1.0, noise=noise, next_latents=next_latents, seed=seed)
        if next_latents is not None:
            nxt = next_latents.float()
        elif sch.dynamics_type == "ODE":
            nxt = r["next_latents_mean"]
        else:
            nxt = r["next_latents"].float()
        d = dict(next_latents=nxt, next_latents_mean=r["next_latents_mean"], log_prob=r["log_prob"], noise_pred=r["noise_pred"],
                 std_dev_t=torch.full((B, 1, 1, 1, 1), coef.std_dev_t, dtype=torch.float32, device=self.device),
                 dt=torch.full((B, 1, 1, 1, 1), coef.dt, dtype=torch.float32, device=self.device))
        return SDESchedulerOutput.from_dict({k: d[k] for k in return_kwargs if k in d})
