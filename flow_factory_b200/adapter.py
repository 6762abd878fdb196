"""B200SD3_5Adapter - the drop-in for Flow-Factory's SD3_5Adapter on the rollout path.

Mirrors FF/models/stable_diffusion/sd3_5.py: `inference()` (175-349, the trajectory sampler; `compute_log_prob`
and `trajectory_indices` are keyword arguments exactly as in the reference) and `forward()` (352-448, one denoise
step, optionally teacher-forced through `next_latents`).  Parameter names are the ABI: the trainers call these through
`filter_kwargs(adapter.inference|forward, **kw)` (FF/utils/base.py:38-63, grpo.py:159-166, 242-263).

What differs from the reference: the T-step loop runs inside the native engine with zero host synchronisation
(the reference syncs >= 4 times per step: SURVEY.md section 3.2); CFG, the Euler/SDE update, the fp16 round trip,
cast_latents' overflow clamp and the Gaussian log-prob are one kernel; kept latents/log-probs land directly in compact
per-sample buffers.  `forward()` with autograd enabled is NOT served here (training replay stays on the reference's
diffusers path - INTEGRATION.md); there is no CPU / PyTorch fallback for the rollout.
"""
from __future__ import annotations

import inspect
from typing import Any, Callable, Dict, List, Optional, Union

import torch

from . import _lib
from .engine import RolloutEngine
from .per_sample import forward_grouped, split_by_timestep
from .rng import randn_tensor
from .stepwise import SUPPORTED_CALLBACKS, per_sample, run_stepwise
from .samples import SD3_5Sample
from .scheduler import FlowMatchEulerDiscreteSDEScheduler, SDESchedulerOutput, set_scheduler_timesteps
from .trajectory import TrajectoryIndicesType, compute_trajectory_indices, plan_slots
from .trajectory import create_callback_collector
from .weights import EngineConfig, has_lora_keys, merge_lora_state_dict


STORAGE_DTYPES = {"fp16": torch.float16, "float16": torch.float16, "bf16": torch.bfloat16, "bfloat16": torch.bfloat16,
                  "fp32": torch.float32, "float32": torch.float32, None: torch.bfloat16}


def filter_kwargs(fn: Callable, **kwargs) -> Dict[str, Any]:
    """FF/utils/base.py:38-63: keep only the keyword arguments `fn` names - everything if `fn` itself takes **kwargs."""
    params = inspect.signature(fn).parameters
    if any(p.kind == inspect.Parameter.VAR_KEYWORD for p in params.values()):
        return kwargs
    return {k: v for k, v in kwargs.items() if k in params}


class B200SD3_5Adapter:
    """Construct from a transformer config + state dict (or `from_reference_adapter`)."""

    def __init__(self, model_config, state_dict: Dict[str, torch.Tensor], device: Union[str, torch.device] = "cuda",
                 scheduler: Optional[FlowMatchEulerDiscreteSDEScheduler] = None, latent_storage_dtype: str = "fp16",
                 decode_fn: Optional[Callable[[torch.Tensor], torch.Tensor]] = None, vae_scale_factor: int = 8,
                 rng: str = "torch", use_graph: bool = True):
        # latent_storage_dtype (FF/hparams/training_args.py:245-252): "fp16" (default), "bf16", "fp32"; None = the transformer's dtype
        # (cast_latents' `default_dtype`, FF/models/abc.py:172-182), which is bf16 on this path
        if latent_storage_dtype not in STORAGE_DTYPES:
            raise ValueError(f"latent_storage_dtype must be one of {list(STORAGE_DTYPES)}, got {latent_storage_dtype!r}")
        self.latent_dtype = STORAGE_DTYPES[latent_storage_dtype]
        if rng not in ("torch", "philox"):
            raise ValueError("rng must be 'torch' (reference-identical noise stream) or 'philox' (in-kernel)")
        self.engine = RolloutEngine(model_config, state_dict, torch.device(device))
        self.device = self.engine.device
        self.model_config = self.engine.cfg
        self.scheduler = scheduler or FlowMatchEulerDiscreteSDEScheduler(noise_level=0.7, shift=3.0, dynamics_type="Flow-SDE")
        self.decode_fn = decode_fn
        self.vae_scale_factor = vae_scale_factor
        self.rng = rng
        self.use_graph = use_graph
        self._mode = "rollout"
        self.overflow_seen = False

    # -------------------------------------------------------------- construction from the reference objects
    @classmethod
    def from_reference_adapter(cls, ref_adapter, state_dict: Optional[Dict[str, torch.Tensor]] = None, **kw) -> "B200SD3_5Adapter":
        """`ref_adapter`: a flow_factory SD3_5Adapter; borrows its transformer weights, scheduler settings and VAE decode.
        BaseAdapter.__init__ applies LoRA BEFORE post_init (FF/models/abc.py:142-148), so the transformer may already be PEFT-wrapped
        (`base_model.model.*`, `*.base_layer.weight`, `*.lora_A.default.weight`): such a state dict is folded to plain diffusers keys
        first (weights.merge_lora_state_dict with the adapter's `model_args.lora_alpha`).  `state_dict`: an already-plain one to use."""
        tr = ref_adapter.transformer
        tr = getattr(tr, "module", tr)
        tr = getattr(tr, "_orig_mod", tr)
        if state_dict is None:
            state_dict = tr.state_dict()
            if has_lora_keys(state_dict):
                state_dict = merge_lora_state_dict(state_dict, lora_alpha=ref_adapter.model_args.lora_alpha)
        cfg = tr.config if not hasattr(tr, "get_base_model") else tr.get_base_model().config
        sched = ref_adapter.scheduler
        mine = FlowMatchEulerDiscreteSDEScheduler(noise_level=sched.noise_level,
                                                  sde_steps=None if sched._sde_steps is None else sched._sde_steps.tolist(),
                                                  num_sde_steps=sched._num_sde_steps, seed=sched.seed,
                                                  dynamics_type=sched.dynamics_type, **dict(sched.config))
        return cls(cfg, state_dict, device=ref_adapter.device, scheduler=mine,
                   decode_fn=lambda lat: ref_adapter.decode_latents(lat, output_type="pt"), **kw)

    def refresh_weights(self, state_dict: Dict[str, torch.Tensor]) -> None:
        self.engine.refresh_weights(state_dict)

    # -------------------------------------------------------------- mode switches (FF/models/abc.py:356-378)
    def rollout(self):
        self._mode = "rollout"; self.scheduler.rollout()

    def train(self, mode: bool = True):
        self._mode = "train" if mode else "eval"; self.scheduler.train(mode=mode)

    def eval(self):
        self._mode = "eval"; self.scheduler.eval()

    def cast_latents(self, latents: torch.Tensor, default_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
        """FF/models/abc.py:172-182, without the per-call host sync: the fp16 clamp is a no-op unless something overflowed."""
        target = self.latent_dtype
        if latents.dtype == target:
            return latents
        if target == torch.float16:
            return latents.clamp(-65504.0, 65504.0).to(torch.float16)
        return latents.to(target)

    def decode_latents(self, latents: torch.Tensor, output_type: str = "pt"):
        if self.decode_fn is None:
            return None
        return self.decode_fn(latents)

    def use_native_vae(self, vae_config, vae_state_dict: Dict[str, torch.Tensor], batch: int = 4) -> None:
        """Route `decode_latents` (sd3_5.py:161-172) through the native decoder (vae.py) instead of a callback into
        the reference's AutoencoderKL; one decoder per latent geometry, built on first use.  (First GPU run of that decoder is
        pending - see flow_factory_b200/vae.py.)"""
        from .vae import B200VaeDecoder, VaeDecoderConfig
        cfg = vae_config if isinstance(vae_config, VaeDecoderConfig) else VaeDecoderConfig.from_config(vae_config)
        cache: Dict[tuple, B200VaeDecoder] = {}

        def decode(lat: torch.Tensor) -> torch.Tensor:
            key = (int(lat.shape[2]), int(lat.shape[3]))
            if key not in cache:
                cache[key] = B200VaeDecoder(cfg, vae_state_dict, key[0], key[1], batch=batch, device=self.device)
            return cache[key].decode_latents(lat, output_type="pt")

        self.decode_fn = decode

    # -------------------------------------------------------------- the trajectory sampler
    @torch.no_grad()
    def inference(
        self,
        prompt: Union[str, List[str], None] = None,
        negative_prompt: Optional[Union[str, List[str]]] = None,
        height: Optional[int] = 1024,
        width: Optional[int] = 1024,
        num_inference_steps: Optional[int] = 50,
        guidance_scale: float = 7.5,
        generator: Optional[torch.Generator] = None,
        joint_attention_kwargs: Optional[Dict[str, Any]] = None,
        prompt_ids: Optional[torch.Tensor] = None,
        prompt_embeds: Optional[torch.Tensor] = None,
        pooled_prompt_embeds: Optional[torch.Tensor] = None,
        negative_prompt_ids: Optional[torch.Tensor] = None,
        negative_prompt_embeds: Optional[torch.Tensor] = None,
        negative_pooled_prompt_embeds: Optional[torch.Tensor] = None,
        compute_log_prob: bool = True,
        extra_call_back_kwargs: List[str] = [],
        trajectory_indices: TrajectoryIndicesType = "all",
        latents: Optional[torch.Tensor] = None,
        noise: Optional[torch.Tensor] = None,
    ) -> List[SD3_5Sample]:
        if prompt_embeds is None or pooled_prompt_embeds is None:
            raise ValueError("B200SD3_5Adapter.inference needs pre-encoded prompt_embeds / pooled_prompt_embeds "
                             "(Flow-Factory's dataloader caches them; text encoders are outside the rollout path)")
        if joint_attention_kwargs:
            raise NotImplementedError("joint_attention_kwargs (IP-adapter / LoRA scale) are not on the accelerated path")
        unsupported = set(extra_call_back_kwargs) - SUPPORTED_CALLBACKS
        if unsupported:
            raise NotImplementedError(f"extra_call_back_kwargs {sorted(unsupported)} are not produced by the fused step")
        if extra_call_back_kwargs:
            # per-step callback values (GRPO-Guard asks for `next_latents_mean`, grpo.py:404) are not kept by the fused T-step rollout:
            # take the reference's own step loop over forward() - same kernels, one launch list per step instead of a replayed graph
            return self._inference_stepwise(prompt, negative_prompt, height, width, num_inference_steps, guidance_scale, generator, prompt_ids,
                                            prompt_embeds, pooled_prompt_embeds, negative_prompt_ids, negative_prompt_embeds,
                                            negative_pooled_prompt_embeds, compute_log_prob, list(extra_call_back_kwargs), trajectory_indices,
                                            latents, noise)
        dev = self.device
        T = int(num_inference_steps)
        do_cfg = guidance_scale > 1.0 and negative_prompt_embeds is not None and negative_pooled_prompt_embeds is not None
        B = len(prompt_embeds)
        C = self.model_config.in_channels
        lh, lw = int(height) // self.vae_scale_factor, int(width) // self.vae_scale_factor
        n_text = prompt_embeds.shape[1]
        plan = self.engine.plan(B, do_cfg, lh, lw, n_text, self.latent_dtype)
        self.engine.set_prompts(plan, prompt_embeds, pooled_prompt_embeds,
                                negative_prompt_embeds if do_cfg else None, negative_pooled_prompt_embeds if do_cfg else None)
        # 3. initial latents: randn in the transformer dtype (pipeline_stable_diffusion_3.py:633-662), then cast_latents
        if latents is None:
            latents = randn_tensor((B, C, lh, lw), generator=generator, device=dev, dtype=torch.bfloat16)
        x0 = self.cast_latents(latents.to(dev))
        # 5. schedule (host only, no device syncs)
        seq_len = (lh // self.model_config.patch_size) * (lw // self.model_config.patch_size)
        sch = self.scheduler
        timesteps = set_scheduler_timesteps(sch, T, seq_len=seq_len)
        sde_now = set(sch.current_sde_steps.tolist())
        nls = [(sch.noise_level if (i in sde_now and not sch.is_eval) else 0.0) for i in range(T)]
        has_lp = [bool(compute_log_prob and nls[i] > 0) for i in range(T)]
        lat_slot, lp_slot, lat_map, lp_map = plan_slots(trajectory_indices, T, has_lp)
        if not compute_log_prob:
            lp_slot, lp_map = [-1] * T, None
        coefs = []
        for i in range(T):
            t, tn = timesteps[i], (timesteps[i + 1] if i + 1 < T else torch.tensor(0.0))
            t_model = float(t.to(self.latent_dtype))                  # sd3_5.py:394: timestep cast to the latents dtype
            coefs.append(sch.step_coef(t, tn, nls[i], compute_log_prob=has_lp[i], t_model=t_model,
                                       store_slot=lat_slot[i + 1], logp_slot=lp_slot[i]))
        n_lat = sum(1 for s in lat_slot if s >= 0)
        n_lp = sum(1 for s in lp_slot if s >= 0)
        # 6. noise: reference draws a full fp32 tensor from the device RNG EVERY step (flow_match...py:350-357)
        if noise is None and self.rng == "torch" and sch.dynamics_type != "ODE":
            noise = torch.stack([torch.randn((B, C, lh, lw), device=dev, dtype=torch.float32) for _ in range(T)])
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if noise is None else 0
        r = self.engine.rollout(plan, x0, coefs, guidance_scale if do_cfg else 1.0, n_lat, lat_slot[0], n_lp,
                                noise=noise, seed=seed, use_graph=self.use_graph)
        final = r["final_latents"]
        images = self.decode_latents(final, output_type="pt")
        # as the reference: the callback gate's map even when no callback key was requested (all -1, or the identity for 'all')
        callback_index_map = create_callback_collector(trajectory_indices, T).get_index_map()
        samples = []
        for b in range(B):
            samples.append(SD3_5Sample(
                timesteps=timesteps,
                all_latents=r["all_latents"][b, :n_lat] if n_lat else None,
                log_probs=(r["log_probs"][b, :n_lp] if n_lp else (torch.zeros(0, device=dev) if compute_log_prob and lp_map is not None else None)),
                latent_index_map=lat_map,
                log_prob_index_map=lp_map if compute_log_prob else None,
                prompt=prompt[b] if isinstance(prompt, list) else prompt,
                prompt_ids=prompt_ids[b] if prompt_ids is not None else None,
                prompt_embeds=prompt_embeds[b],
                pooled_prompt_embeds=pooled_prompt_embeds[b],
                negative_prompt=negative_prompt[b] if isinstance(negative_prompt, list) else negative_prompt,
                negative_prompt_ids=negative_prompt_ids[b] if negative_prompt_ids is not None else None,
                negative_prompt_embeds=negative_prompt_embeds[b] if negative_prompt_embeds is not None else None,
                negative_pooled_prompt_embeds=negative_pooled_prompt_embeds[b] if negative_pooled_prompt_embeds is not None else None,
                height=height, width=width,
                image=images[b] if images is not None else None,
                extra_kwargs={"callback_index_map": callback_index_map, "final_latents": final[b]},
            ))
        self._last_overflow = r["overflow"]
        return samples

    def _inference_stepwise(self, prompt, negative_prompt, height, width, num_inference_steps, guidance_scale, generator, prompt_ids,
                            prompt_embeds, pooled_prompt_embeds, negative_prompt_ids, negative_prompt_embeds, negative_pooled_prompt_embeds,
                            compute_log_prob, extra_call_back_kwargs, trajectory_indices, latents, noise) -> List[SD3_5Sample]:
        """SD3_5Adapter.inference's loop (sd3_5.py:266-349) with the reference's collectors around `forward()` (stepwise.py)."""
        dev = self.device
        T, B, C = int(num_inference_steps), len(prompt_embeds), self.model_config.in_channels
        lh, lw = int(height) // self.vae_scale_factor, int(width) // self.vae_scale_factor
        if latents is None:
            latents = randn_tensor((B, C, lh, lw), generator=generator, device=dev, dtype=torch.bfloat16)
        seq_len = (lh // self.model_config.patch_size) * (lw // self.model_config.patch_size)
        timesteps = set_scheduler_timesteps(self.scheduler, T, seq_len=seq_len)
        res = run_stepwise(self, timesteps, latents.to(dev), trajectory_indices, compute_log_prob, extra_call_back_kwargs,
                           dict(prompt_embeds=prompt_embeds, pooled_prompt_embeds=pooled_prompt_embeds, negative_prompt_embeds=negative_prompt_embeds,
                                negative_pooled_prompt_embeds=negative_pooled_prompt_embeds, guidance_scale=guidance_scale), noise=noise)
        images = self.decode_latents(res["final"], output_type="pt")
        samples = []
        for b in range(B):
            samples.append(SD3_5Sample(
                timesteps=timesteps,
                prompt=prompt[b] if isinstance(prompt, list) else prompt,
                prompt_ids=prompt_ids[b] if prompt_ids is not None else None,
                prompt_embeds=prompt_embeds[b],
                pooled_prompt_embeds=pooled_prompt_embeds[b],
                negative_prompt=negative_prompt[b] if isinstance(negative_prompt, list) else negative_prompt,
                negative_prompt_ids=negative_prompt_ids[b] if negative_prompt_ids is not None else None,
                negative_prompt_embeds=negative_prompt_embeds[b] if negative_prompt_embeds is not None else None,
                negative_pooled_prompt_embeds=negative_pooled_prompt_embeds[b] if negative_pooled_prompt_embeds is not None else None,
                height=height, width=width,
                image=images[b] if images is not None else None,
                **per_sample(res, b),
            ))
        return samples

    # -------------------------------------------------------------- one denoise step
    def forward(
        self,
        t: torch.Tensor,
        latents: torch.Tensor,
        prompt_embeds: torch.Tensor,
        pooled_prompt_embeds: torch.Tensor,
        negative_prompt_embeds: Optional[torch.Tensor] = None,
        negative_pooled_prompt_embeds: Optional[torch.Tensor] = None,
        guidance_scale: float = 7.5,
        t_next: Optional[torch.Tensor] = None,
        next_latents: Optional[torch.Tensor] = None,
        noise_level: Optional[float] = None,
        joint_attention_kwargs: Optional[Dict[str, Any]] = None,
        compute_log_prob: bool = True,
        return_kwargs: List[str] = ["noise_pred", "next_latents", "next_latents_mean", "std_dev_t", "dt", "log_prob"],
        noise: Optional[torch.Tensor] = None,
    ) -> SDESchedulerOutput:
        if torch.is_grad_enabled() and any(isinstance(x, torch.Tensor) and x.requires_grad for x in (latents, prompt_embeds)):
            raise RuntimeError("B200SD3_5Adapter.forward serves the no-grad path (rollout step, teacher-forced old-log-prob / "
                               "KL-reference recompute). Keep the autograd replay on the reference adapter (INTEGRATION.md).")
        if joint_attention_kwargs:
            raise NotImplementedError("joint_attention_kwargs are not on the accelerated path")
        do_cfg = negative_prompt_embeds is not None and negative_pooled_prompt_embeds is not None and guidance_scale > 1.0
        B, C, lh, lw = latents.shape
        groups = split_by_timestep(t, t_next, B)
        if groups is not None:
            # per-sample timesteps (NFT / AWM / CRD no-grad forward, nft.py:366-374): one engine call per distinct (t, t_next)
            return forward_grouped(
                self.forward, groups, B,
                dict(latents=latents, prompt_embeds=prompt_embeds, pooled_prompt_embeds=pooled_prompt_embeds,
                     negative_prompt_embeds=negative_prompt_embeds, negative_pooled_prompt_embeds=negative_pooled_prompt_embeds,
                     guidance_scale=guidance_scale, next_latents=next_latents, noise_level=noise_level, compute_log_prob=compute_log_prob,
                     return_kwargs=return_kwargs, noise=noise),
                batched=("latents", "prompt_embeds", "pooled_prompt_embeds", "negative_prompt_embeds", "negative_pooled_prompt_embeds",
                         "next_latents", "noise"),
                make_output=SDESchedulerOutput.from_dict)
        plan = self.engine.plan(B, do_cfg, lh, lw, prompt_embeds.shape[1], self.latent_dtype)
        self.engine.set_prompts(plan, prompt_embeds, pooled_prompt_embeds, negative_prompt_embeds if do_cfg else None,
                                negative_pooled_prompt_embeds if do_cfg else None)
        sch = self.scheduler
        t0 = t if isinstance(t, torch.Tensor) else torch.tensor(float(t))
        t0 = t0.flatten()[0].detach().cpu().float()
        # t_next omitted: the reference's scheduler.step then reads sigmas[i], sigmas[i + 1] from its tables (step_coef mirrors that)
        tn = None if t_next is None else (t_next if isinstance(t_next, torch.Tensor) else torch.tensor(float(t_next))).flatten()[0].detach().cpu().float()
        coef = sch.step_coef(t0, tn, noise_level, compute_log_prob=compute_log_prob, t_model=float(t0.to(self.latent_dtype)))
        if noise is None and next_latents is None and self.rng == "torch" and sch.dynamics_type != "ODE":
            noise = torch.randn(latents.shape, device=self.device, dtype=torch.float32)
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if noise is None else 0
        r = self.engine.step(plan, latents, coef, guidance_scale if do_cfg else 1.0, noise=noise, next_latents=next_latents,
                             seed=seed, want_mean=("next_latents_mean" in return_kwargs) or sch.dynamics_type == "ODE")
        if next_latents is not None:
            nxt = next_latents.float()
        elif sch.dynamics_type == "ODE" and r["next_latents_mean"] is not None:
            nxt = r["next_latents_mean"]            # ODE: next_latents IS the mean, no storage round trip (flow_match...py:333-334)
        else:
            nxt = r["next_latents"].float()
        d = dict(next_latents=nxt, next_latents_mean=r["next_latents_mean"], log_prob=r["log_prob"], noise_pred=r["noise_pred"],
                 std_dev_t=torch.full((B, 1, 1, 1), coef.std_dev_t, dtype=torch.float32, device=self.device),
                 dt=torch.full((B, 1, 1, 1), coef.dt, dtype=torch.float32, device=self.device))
        return SDESchedulerOutput.from_dict({k: d[k] for k in return_kwargs if k in d})
