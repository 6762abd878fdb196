"""Reference-side glue (the ONLY file a Flow-Factory user adds next to the training config): selected with
`model.model_type: "ff_b200_glue.B200GlueSD3_5Adapter"` - see INTEGRATION.md section 1, which quotes this file.  It imports the reference
(`flow_factory`), so it is not part of the `flow_factory_b200` package; tests/test_reference_hooks.py imports it where the reference exists."""
import torch
from flow_factory.models.stable_diffusion.sd3_5 import SD3_5Adapter           # reference adapter (unchanged)
from flow_factory_b200.adapter import B200SD3_5Adapter                         # this repo

class B200GlueSD3_5Adapter(SD3_5Adapter):
    """Rollout + no-grad steps on the B200 engine; everything else (load_pipeline, LoRA, EMA, checkpointing,
    decode_latents, the autograd replay in optimize()) is inherited from the reference."""

    def post_init(self):
        super().post_init()
        self._b200 = B200SD3_5Adapter.from_reference_adapter(self, rng="torch")   # borrows transformer weights + scheduler

    def _sync_engine(self):
        tr = getattr(self.transformer, "module", self.transformer)
        sd = tr.state_dict()
        if any(".lora_A." in k for k in sd):                # PEFT-wrapped transformer (FF/models/abc.py:859-949): fold W + (alpha/r) B A
            from flow_factory_b200.weights import merge_lora_state_dict
            sd = merge_lora_state_dict(sd, lora_alpha=self.model_args.lora_alpha)
        self._b200.refresh_weights(sd)                     # weights moved (optimizer / EMA / LoRA): re-pack, addresses stay stable
        self._b200.scheduler.set_seed(self.scheduler.seed)

    def rollout(self, *a, **kw):                            # called inside `use_ema_parameters()` where a trainer samples with EMA weights
        super().rollout(*a, **kw)
        self._sync_engine()
        self._b200.rollout()

    def eval(self):                                         # mode switches reach BOTH schedulers (FF/models/abc.py:356-378)
        super().eval()
        self._b200.eval()

    def train(self, mode: bool = True):
        super().train(mode)
        self._b200.train(mode)

    @torch.no_grad()
    def inference(self, *args, **kwargs):                   # GRPOTrainer.sample() / evaluate(), grpo.py:159-166, 110-119
        if kwargs.get("prompt_embeds") is None or kwargs.get("pooled_prompt_embeds") is None:
            # raw prompts (no cached embeddings): the text encoders stay in the reference (sd3_5.py:216-231)
            enc = self.encode_prompt(kwargs.get("prompt"), kwargs.get("negative_prompt"), guidance_scale=kwargs.get("guidance_scale", 7.5),
                                     device=self.device)
            kwargs.update({k: v for k, v in enc.items() if v is not None})
        if self.mode == "eval":                             # evaluate() swaps the EMA weights in AFTER eval(): pick them up here
            self._sync_engine()
        return self._b200.inference(*args, **kwargs)

    def forward(self, *args, **kwargs):                     # grpo.py:242-263 (with grad) / 282-292 (no grad)
        if torch.is_grad_enabled():
            return super().forward(*args, **kwargs)         # training replay stays on diffusers + autograd
        return self._b200.forward(*args, **kwargs)
