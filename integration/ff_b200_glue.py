"""Reference-side glue (the ONLY file a Flow-Factory user adds next to the training config): selected with
`model.model_type: "ff_b200_glue.B200GlueSD3_5Adapter"` - see INTEGRATION.md section 1, which quotes this file.  It imports the reference
(`flow_factory`), so it is not part of the `flow_factory_b200` package; tests/test_reference_hooks.py imports it where the reference exists.

Which weights the engine holds.  The engine works on a PACKED COPY of the transformer weights, so every place where the reference changes
what `self.transformer` computes has to reach that copy before the next no-grad call:
  * optimizer steps between rollouts                    -> `rollout()` re-packs (and any no-grad forward while `mode == 'train'` re-packs);
  * `use_ema_parameters()` / `use_named_parameters()`   (FF/models/abc.py:523-531, 660-682; trainers' `sampling_context()`, nft.py:75-82)
    copy other values into the parameters on enter and the originals back on exit -> both edges mark the copy stale;
  * `use_ref_parameters()` (abc.py:556-585): full fine-tuning swaps parameter values (same as above); LoRA enters PEFT's
    `disable_adapter()`, which changes NO tensor - inside it the engine must hold the BASE weights, i.e. the LoRA pair is not folded.
The copy is refreshed lazily, at the next `inference()` / no-grad `forward()`, so nested contexts cost one re-pack, not one per edge.

Rollout / replay consistency.  GRPO's importance ratio exp(new_log_prob - old_log_prob) is clipped at ~1e-4, and the two
log-probs come from two different bf16 implementations here: the engine's kernels (rollout) and diffusers + autograd (replay).  Measured on
a B200 at the benchmarked size (tests/test_gpu_parity_c2.py -> profiles/r02_parity_c2_rollout.json): the engine's own log-probs agree with
the reference numerics to 6e-7 .. 2e-5 relative, but |ratio - 1| of a cross-path replay reaches 1e-4 .. 2e-4 at late steps (bf16 noise of two
independent forwards, amplified by CFG) - the size of the clip range.  So, as SURVEY 7.2 #1 option (b) prescribes, `inference()` re-evaluates
`old_log_prob` of the stored SDE transitions (<= num_sde_steps per rollout, 1 by default) with the REFERENCE forward, teacher-forced on
the engine's own latents: the ratio of the first optimisation step is then exactly 1, as in the reference, at the cost of num_sde_steps of
T transformer forwards on the diffusers path.  `B200GlueSD3_5Adapter.recompute_old_log_probs = False` keeps the engine's values.
"""
from contextlib import contextmanager

import torch
from flow_factory.models.stable_diffusion.sd3_5 import SD3_5Adapter           # reference adapter (unchanged)
from flow_factory_b200.adapter import B200SD3_5Adapter                         # this repo
from flow_factory_b200.weights import has_lora_keys, merge_lora_state_dict


class B200GlueSD3_5Adapter(SD3_5Adapter):
    """Rollout + no-grad steps on the B200 engine; everything else (load_pipeline, LoRA, EMA, checkpointing,
    decode_latents, the autograd replay in optimize()) is inherited from the reference."""

    recompute_old_log_probs = True   # see "Rollout / replay consistency" above
    _engine_stale = True          # the packed copy may differ from what self.transformer computes right now
    _lora_off = 0                 # depth of active use_ref_parameters() contexts under LoRA (adapter disabled)

    def post_init(self):
        super().post_init()        # BaseAdapter.__init__ has already applied LoRA (abc.py:142-148): the state dict may carry PEFT keys
        self._b200 = B200SD3_5Adapter.from_reference_adapter(self, rng="torch", state_dict=self._engine_state_dict())
        self._engine_stale = False

    # ------------------------------------------------------------------ keeping the packed copy in step with the module
    def _engine_state_dict(self):
        tr = getattr(self.transformer, "module", self.transformer)
        tr = getattr(tr, "_orig_mod", tr)
        sd = tr.state_dict()
        if has_lora_keys(sd):       # PEFT-wrapped transformer: fold W + (alpha/r) B A, or only strip the wrapper while the adapter is disabled
            sd = merge_lora_state_dict(sd, lora_alpha=self.model_args.lora_alpha, scale=0.0 if self._lora_off > 0 else 1.0)
        return sd

    def _sync_engine(self):
        self._b200.refresh_weights(self._engine_state_dict())   # re-pack in place: addresses (plans, TMA descriptors, graphs) stay valid
        self._engine_stale = False

    def _sync_engine_if_stale(self):
        self._b200.scheduler.set_seed(self.scheduler.seed)      # grpo.py:63 sets the per-epoch SDE-window seed on the reference scheduler
        if self._engine_stale or self.mode == "train":          # in train mode the optimizer may have stepped since the last call
            self._sync_engine()

    @contextmanager
    def _swapped(self, ctx, lora_off=False):
        with ctx:
            self._engine_stale = True
            self._lora_off += int(lora_off)
            try:
                yield
            finally:
                self._lora_off -= int(lora_off)
                self._engine_stale = True

    def use_ema_parameters(self):
        return self._swapped(super().use_ema_parameters())

    def use_named_parameters(self, name):
        return self._swapped(super().use_named_parameters(name))

    def use_ref_parameters(self):
        return self._swapped(super().use_ref_parameters(), lora_off=self.model_args.finetune_type == "lora")

    # ------------------------------------------------------------------ mode switches reach BOTH schedulers (FF/models/abc.py:356-378)
    def rollout(self, *a, **kw):
        super().rollout(*a, **kw)
        self._engine_stale = True                               # the optimizer ran since the last rollout
        self._b200.rollout()

    def eval(self):
        super().eval()
        self._engine_stale = True
        self._b200.eval()

    def train(self, mode: bool = True):
        super().train(mode)
        self._b200.train(mode)

    # ------------------------------------------------------------------ the accelerated calls
    @torch.no_grad()
    def inference(self, *args, **kwargs):                   # GRPOTrainer.sample() / evaluate(), grpo.py:159-166, 110-119
        if kwargs.get("prompt_embeds") is None or kwargs.get("pooled_prompt_embeds") is None:
            # raw prompts (no cached embeddings): the text encoders stay in the reference (sd3_5.py:216-231)
            enc = self.encode_prompt(kwargs.get("prompt"), kwargs.get("negative_prompt"), guidance_scale=kwargs.get("guidance_scale", 7.5),
                                     device=self.device)
            kwargs.update({k: v for k, v in enc.items() if v is not None})
        self._sync_engine_if_stale()
        samples = self._b200.inference(*args, **kwargs)
        if self.recompute_old_log_probs and kwargs.get("compute_log_prob", True) and self.mode != "eval":
            self._recompute_old_log_probs(samples, kwargs)
        return samples

    def _recompute_old_log_probs(self, samples, kw):
        """old_log_prob of every stored SDE transition through the reference's own forward (teacher-forced on the engine's latents)."""
        if not samples or samples[0].log_probs is None or samples[0].log_probs.numel() == 0:
            return
        s0 = samples[0]
        lmap, pmap, ts = s0.latent_index_map.tolist(), s0.log_prob_index_map.tolist(), s0.timesteps
        for i, slot in enumerate(pmap):
            if slot < 0 or lmap[i] < 0 or lmap[i + 1] < 0:
                continue                                        # no log-prob kept here, or the transition's end points are not both stored
            x_t = torch.stack([s.all_latents[lmap[i]] for s in samples]).to(self.device)
            x_n = torch.stack([s.all_latents[lmap[i + 1]] for s in samples]).to(self.device)
            t = ts[i].to(self.device)
            t_next = (ts[i + 1] if i + 1 < len(ts) else torch.zeros((), dtype=ts.dtype)).to(self.device)
            out = SD3_5Adapter.forward(                          # the reference path, NOT self.forward (which serves no-grad calls natively)
                self, t=t, t_next=t_next, latents=x_t, next_latents=x_n, prompt_embeds=kw["prompt_embeds"],
                pooled_prompt_embeds=kw["pooled_prompt_embeds"], negative_prompt_embeds=kw.get("negative_prompt_embeds"),
                negative_pooled_prompt_embeds=kw.get("negative_pooled_prompt_embeds"), guidance_scale=kw.get("guidance_scale", 7.5),
                noise_level=self.scheduler.get_noise_level_for_timestep(t), compute_log_prob=True, return_kwargs=["log_prob"])
            for b, s in enumerate(samples):
                s.log_probs[slot] = out.log_prob[b].to(s.log_probs.dtype)

    def forward(self, *args, **kwargs):                     # grpo.py:242-263 (with grad) / 282-292 (no grad, inside use_ref_parameters)
        if torch.is_grad_enabled():
            return super().forward(*args, **kwargs)         # training replay stays on diffusers + autograd
        self._sync_engine_if_stale()
        return self._b200.forward(*args, **kwargs)
