"""B200QwenImageAdapter - drop-in for Flow-Factory's QwenImageAdapter on the rollout path (first cut).

Mirrors FF/models/qwen_image/qwen_image.py: `inference()` (290-470) and `forward()` (476-600); parameter names are the ABI.
True CFG (two prompt sets, per-token norm rescale) runs as one forward batch of 2B; prompts of different lengths are right-padded
to one length and the padding is masked as attention keys (the reference's encoder_hidden_states_mask).  Only prefix masks are
accepted; text encoding, VAE decode and the autograd replay stay on the reference."""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional, Union

import torch

from .qwen import QwenRolloutEngine
from .per_sample import forward_grouped, split_by_timestep
from .rng import randn_tensor
from .stepwise import SUPPORTED_CALLBACKS, per_sample, run_stepwise
from .samples import QwenImageSample
from .scheduler import FlowMatchEulerDiscreteSDEScheduler, SDESchedulerOutput, set_scheduler_timesteps
from .trajectory import TrajectoryIndicesType, plan_slots
from .trajectory import create_callback_collector


def _dense(embeds, mask, what: str):
    """List / padded tensor + mask -> (right-padded dense [B, Nt, J], valid lengths).  The reference's masks are prefix masks
    (tokenizer padding on the right, `_pad_batch_prompt`); anything else is refused."""
    if isinstance(embeds, (list, tuple)):
        nt = max(e.shape[0] for e in embeds)
        lens = [int(e.shape[0]) for e in embeds]
        embeds = torch.stack([torch.nn.functional.pad(e, (0, 0, 0, nt - e.shape[0])) for e in embeds], dim=0)
        if mask is not None:
            lens = [min(l, int(torch.as_tensor(m).sum())) for l, m in zip(lens, mask)]
            mask = None
    else:
        lens = [int(embeds.shape[1])] * len(embeds)
    if mask is not None:
        m = (torch.stack(list(mask), dim=0) if isinstance(mask, (list, tuple)) else mask).to(torch.bool)
        lens = [int(v) for v in m.sum(dim=1).tolist()]
        ar = torch.arange(m.shape[1], device=m.device)[None, :]
        if not bool((m == (ar < torch.as_tensor(lens, device=m.device)[:, None])).all()):
            raise NotImplementedError(f"{what}: only right-padded (prefix) masks are on the accelerated path")
    if min(lens) < 1:
        raise ValueError(f"{what}: empty prompt")
    return embeds, lens


def _pad_to(embeds: torch.Tensor, nt: int) -> torch.Tensor:
    return embeds if embeds.shape[1] == nt else torch.nn.functional.pad(embeds, (0, 0, 0, nt - embeds.shape[1]))


class B200QwenImageAdapter:
    def __init__(self, model_config, state_dict: Dict[str, torch.Tensor], device: Union[str, torch.device] = "cuda",
                 scheduler: Optional[FlowMatchEulerDiscreteSDEScheduler] = None, decode_fn: Optional[Callable[..., torch.Tensor]] = None,
                 vae_scale_factor: int = 8, rng: str = "torch", use_graph: bool = True):
        if rng not in ("torch", "philox"):
            raise ValueError("rng must be 'torch' or 'philox'")
        self.engine = QwenRolloutEngine(model_config, state_dict, torch.device(device))
        self.device = self.engine.device
        self.model_config = self.engine.cfg
        self.scheduler = scheduler or FlowMatchEulerDiscreteSDEScheduler(noise_level=0.7, use_dynamic_shifting=True, dynamics_type="ODE")
        self.decode_fn, self.vae_scale_factor, self.rng, self.use_graph = decode_fn, vae_scale_factor, rng, use_graph

    def refresh_weights(self, state_dict: Dict[str, torch.Tensor]) -> None:
        self.engine.refresh_weights(state_dict)

    def rollout(self):
        self.scheduler.rollout()

    def train(self, mode: bool = True):       # FF/models/abc.py:372-378
        self.scheduler.train(mode=mode)

    def eval(self):
        self.scheduler.eval()

    def cast_latents(self, latents: torch.Tensor, default_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
        return latents if latents.dtype == torch.float16 else latents.clamp(-65504.0, 65504.0).to(torch.float16)

    def _plan(self, B: int, h2: int, w2: int, pos, neg, guidance_scale: float):
        """pos / neg: (dense embeds, lengths).  Both sets are right-padded to one common text length (multiple of 8 rows keeps the
        plan cache small); the padded keys are masked in the attention kernel."""
        (pe, plens), do_cfg = pos, guidance_scale > 1.0 and neg is not None
        npe, nlens = neg if do_cfg else (None, None)
        nt = max(pe.shape[1], npe.shape[1] if do_cfg else 0)
        nt = (nt + 7) // 8 * 8
        plan = self.engine.plan(B, h2, w2, nt, cfg=do_cfg)
        self.engine.set_prompts(plan, _pad_to(pe, nt), _pad_to(npe, nt) if do_cfg else None, float(guidance_scale) if do_cfg else 1.0,
                                prompt_lengths=plens, negative_lengths=nlens)
        return plan

    @torch.no_grad()
    def inference(
        self,
        prompt: Union[str, List[str]] = None,
        negative_prompt: Union[str, List[str]] = None,
        num_inference_steps: int = 50,
        guidance_scale: float = 4.0,
        height: int = 1024,
        width: int = 1024,
        generator: Optional[torch.Generator] = None,
        prompt_ids=None,
        prompt_embeds=None,
        prompt_embeds_mask=None,
        negative_prompt_ids=None,
        negative_prompt_embeds=None,
        negative_prompt_embeds_mask=None,
        attention_kwargs: Optional[Dict[str, Any]] = None,
        max_sequence_length: int = 1024,
        compute_log_prob: bool = False,
        extra_call_back_kwargs: List[str] = [],
        trajectory_indices: TrajectoryIndicesType = "all",
        latents: Optional[torch.Tensor] = None,
        noise: Optional[torch.Tensor] = None,
    ) -> List[QwenImageSample]:
        if prompt_embeds is None:
            raise ValueError("B200QwenImageAdapter.inference needs pre-encoded prompt_embeds (+ masks)")
        if attention_kwargs:
            raise NotImplementedError("attention_kwargs are not on the accelerated path")
        unsupported = set(extra_call_back_kwargs) - SUPPORTED_CALLBACKS
        if unsupported:
            raise NotImplementedError(f"extra_call_back_kwargs {sorted(unsupported)} are not produced by the step kernel")
        dev, T = self.device, int(num_inference_steps)
        pos = _dense(prompt_embeds, prompt_embeds_mask, "prompt_embeds")
        neg = _dense(negative_prompt_embeds, negative_prompt_embeds_mask, "negative_prompt_embeds") if negative_prompt_embeds is not None else None
        pe, npe = pos[0], (neg[0] if neg is not None else None)
        B = len(pe)
        h2, w2 = int(height) // self.vae_scale_factor // 2, int(width) // self.vae_scale_factor // 2
        plan = self._plan(B, h2, w2, pos, neg, guidance_scale)
        if latents is None:   # prepare_latents: randn (B, 1, 16, 2*h2, 2*w2) packed 2x2 -> (B, h2*w2, 64)
            z = randn_tensor((B, 16, 2 * h2, 2 * w2), generator=generator, device=dev, dtype=torch.bfloat16)
            latents = z.view(B, 16, h2, 2, w2, 2).permute(0, 2, 4, 1, 3, 5).reshape(B, h2 * w2, 64)
        x0 = self.cast_latents(latents.to(dev))
        sch = self.scheduler
        timesteps = set_scheduler_timesteps(sch, T, seq_len=plan.n_img)
        mk = lambda e, n: (torch.arange(e.shape[0], device=e.device) < n).long()
        if extra_call_back_kwargs:
            # per-step callback values: the reference's own loop over forward() (stepwise.py)
            res = run_stepwise(self, timesteps, x0, trajectory_indices, compute_log_prob, list(extra_call_back_kwargs),
                               dict(prompt_embeds=prompt_embeds, prompt_embeds_mask=prompt_embeds_mask, img_shapes=[[(1, h2, w2)]] * B,
                                    negative_prompt_embeds=negative_prompt_embeds, negative_prompt_embeds_mask=negative_prompt_embeds_mask,
                                    guidance_scale=guidance_scale), noise=noise)
            images = self.decode_fn(res["final"], height, width) if self.decode_fn is not None else None
            return [QwenImageSample(timesteps=timesteps, height=height, width=width, image=images[b] if images is not None else None,
                                    img_shapes=[(1, h2, w2)], prompt=prompt[b] if isinstance(prompt, list) else prompt,
                                    prompt_ids=prompt_ids[b] if prompt_ids is not None else None,
                                    prompt_embeds=pe[b], prompt_embeds_mask=mk(pe[b], pos[1][b]),
                                    negative_prompt=negative_prompt[b] if isinstance(negative_prompt, list) else negative_prompt,
                                    negative_prompt_ids=negative_prompt_ids[b] if negative_prompt_ids is not None else None,
                                    negative_prompt_embeds=npe[b] if npe is not None else None,
                                    negative_prompt_embeds_mask=mk(npe[b], neg[1][b]) if npe is not None else None,
                                    **per_sample(res, b)) for b in range(B)]
        sde_now = set(sch.current_sde_steps.tolist())
        nls = [(sch.noise_level if (i in sde_now and not sch.is_eval and sch.dynamics_type != "ODE") else 0.0) for i in range(T)]
        has_lp = [bool(compute_log_prob and nls[i] > 0) for i in range(T)]
        lat_slot, lp_slot, lat_map, lp_map = plan_slots(trajectory_indices, T, has_lp)
        if not compute_log_prob:
            lp_slot, lp_map = [-1] * T, None
        coefs = [sch.step_coef(timesteps[i], timesteps[i + 1] if i + 1 < T else torch.tensor(0.0), nls[i], compute_log_prob=has_lp[i],
                               t_model=self.engine.t_model(float(timesteps[i])), store_slot=lat_slot[i + 1], logp_slot=lp_slot[i])
                 for i in range(T)]
        n_lat = sum(1 for s in lat_slot if s >= 0)
        n_lp = sum(1 for s in lp_slot if s >= 0)
        if noise is None and self.rng == "torch" and sch.dynamics_type != "ODE":
            noise = torch.stack([torch.randn(tuple(x0.shape), device=dev, dtype=torch.float32) for _ in range(T)])
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if noise is None else 0
        r = self.engine.rollout(plan, x0, coefs, n_lat, lat_slot[0], n_lp, noise=noise, seed=seed, use_graph=self.use_graph)
        final = r["final_latents"]
        images = self.decode_fn(final, height, width) if self.decode_fn is not None else None
        # as the reference: the callback gate's map even when no callback key was requested (all -1, or the identity for 'all')
        callback_index_map = create_callback_collector(trajectory_indices, T).get_index_map()
        samples = []
        for b in range(B):
            samples.append(QwenImageSample(
                timesteps=timesteps,
                all_latents=r["all_latents"][b, :n_lat] if n_lat else None,
                log_probs=(r["log_probs"][b, :n_lp] if n_lp else (torch.zeros(0, device=dev) if compute_log_prob and lp_map is not None else None)),
                latent_index_map=lat_map, log_prob_index_map=lp_map if compute_log_prob else None,
                height=height, width=width, image=images[b] if images is not None else None, img_shapes=[(1, h2, w2)],
                prompt=prompt[b] if isinstance(prompt, list) else prompt,
                prompt_ids=prompt_ids[b] if prompt_ids is not None else None,
                prompt_embeds=pe[b], prompt_embeds_mask=mk(pe[b], pos[1][b]),
                negative_prompt=negative_prompt[b] if isinstance(negative_prompt, list) else negative_prompt,
                negative_prompt_ids=negative_prompt_ids[b] if negative_prompt_ids is not None else None,
                negative_prompt_embeds=npe[b] if npe is not None else None,
                negative_prompt_embeds_mask=mk(npe[b], neg[1][b]) if npe is not None else None,
                extra_kwargs={"callback_index_map": callback_index_map, "final_latents": final[b]},
            ))
        self._last_overflow = r["overflow"]
        return samples

    def forward(
        self,
        t: torch.Tensor,
        latents: torch.Tensor,
        prompt_embeds,
        prompt_embeds_mask,
        img_shapes,
        negative_prompt_embeds=None,
        negative_prompt_embeds_mask=None,
        guidance_scale: float = 4.0,
        t_next: Optional[torch.Tensor] = None,
        next_latents: Optional[torch.Tensor] = None,
        noise_level: Optional[float] = None,
        attention_kwargs: Optional[Dict[str, Any]] = None,
        compute_log_prob: bool = True,
        return_kwargs: List[str] = ["noise_pred", "next_latents", "next_latents_mean", "std_dev_t", "dt", "log_prob"],
        noise: Optional[torch.Tensor] = None,
    ) -> SDESchedulerOutput:
        if torch.is_grad_enabled() and any(isinstance(x, torch.Tensor) and x.requires_grad for x in (latents,)):
            raise RuntimeError("B200QwenImageAdapter.forward serves the no-grad path; keep the autograd replay on the reference adapter")
        if attention_kwargs:
            raise NotImplementedError("attention_kwargs are not on the accelerated path")
        B, Ni, _ = latents.shape
        groups = split_by_timestep(t, t_next, B)
        if groups is not None:       # per-sample timesteps (NFT / AWM / CRD): one engine call per distinct (t, t_next) - per_sample.py
            return forward_grouped(
                self.forward, groups, B,
                dict(latents=latents, prompt_embeds=prompt_embeds, prompt_embeds_mask=prompt_embeds_mask, img_shapes=img_shapes,
                     negative_prompt_embeds=negative_prompt_embeds, negative_prompt_embeds_mask=negative_prompt_embeds_mask,
                     guidance_scale=guidance_scale, next_latents=next_latents, noise_level=noise_level, compute_log_prob=compute_log_prob,
                     return_kwargs=return_kwargs, noise=noise),
                batched=("latents", "prompt_embeds", "prompt_embeds_mask", "img_shapes", "negative_prompt_embeds", "negative_prompt_embeds_mask",
                         "next_latents", "noise"),
                make_output=SDESchedulerOutput.from_dict)
        pos = _dense(prompt_embeds, prompt_embeds_mask, "prompt_embeds")
        neg = _dense(negative_prompt_embeds, negative_prompt_embeds_mask, "negative_prompt_embeds") if negative_prompt_embeds is not None else None
        shp = img_shapes[0]
        shp = shp[0] if isinstance(shp, (list, tuple)) and isinstance(shp[0], (list, tuple)) else shp
        _, h2, w2 = shp
        assert h2 * w2 == Ni, (h2, w2, Ni)
        plan = self._plan(B, int(h2), int(w2), pos, neg, guidance_scale)
        sch = self.scheduler
        t0 = (t if isinstance(t, torch.Tensor) else torch.tensor(float(t))).flatten()[0].detach().cpu().float()
        # t_next omitted: the reference's scheduler.step then reads sigmas[i], sigmas[i + 1] from its tables (step_coef mirrors that)
        tn = None if t_next is None else (t_next if isinstance(t_next, torch.Tensor) else torch.tensor(float(t_next))).flatten()[0].detach().cpu().float()
        coef = sch.step_coef(t0, tn, noise_level, compute_log_prob=compute_log_prob, t_model=self.engine.t_model(float(t0), latents.dtype))
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
