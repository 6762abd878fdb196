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
"""
from contextlib import contextmanager

import torch
from flow_factory.models.stable_diffusion.sd3_5 import SD3_5Adapter           # reference adapter (unchanged)
from flow_factory_b200.adapter import B200SD3_5Adapter                         # this repo
from flow_factory_b200.weights import has_lora_keys, merge_lora_state_dict


class B200GlueSD3_5Adapter(SD3_5Adapter):
    """Rollout + no-grad steps on the B200 engine; everything else (load_pipeline, LoRA, EMA, checkpointing,
    decode_latents, the autograd replay in optimize()) is inherited from the reference."""

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
        return self._b200.inference(*args, **kwargs)

    def forward(self, *args, **kwargs):                     # grpo.py:242-263 (with grad) / 282-292 (no grad, inside use_ref_parameters)
        if torch.is_grad_enabled():
            return super().forward(*args, **kwargs)         # training replay stays on diffusers + autograd
        self._sync_engine_if_stale()
        return self._b200.forward(*args, **kwargs)
