"""RolloutEngine: Python face of the C ABI (include/ffb200.h).  Device memory, streams and tensors are torch's
(plumbing); every kernel on the path is ours.  There is no fallback: constructing the engine without a B200 or
without libffb200.so raises."""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from .scheduler import FlowMatchEulerDiscreteSDEScheduler, make_step_coef
from .weights import EngineConfig, PackedWeights


LATENT_DTYPES = {torch.float16: 0, torch.bfloat16: 1, torch.float32: 2}      # FFB200_LAT_* of include/ffb200.h


class Plan:
    def __init__(self, engine: "RolloutEngine", batch: int, cfg: bool, lat_h: int, lat_w: int, n_text: int,
                 latent_dtype: torch.dtype = torch.float16):
        if latent_dtype not in LATENT_DTYPES:
            raise ValueError(f"latent storage dtype must be one of {list(LATENT_DTYPES)}, got {latent_dtype}")
        self.engine, self.batch, self.cfg, self.lat_h, self.lat_w, self.n_text = engine, batch, cfg, lat_h, lat_w, n_text
        self.latent_dtype = latent_dtype
        self.handle = C.c_void_p()
        _lib.check(_lib.lib().ffb200_plan_create(engine.handle, batch, int(cfg), lat_h, lat_w, n_text, C.byref(self.handle)),
                   "ffb200_plan_create")
        _lib.check(_lib.lib().ffb200_plan_set_latent_dtype(self.handle, LATENT_DTYPES[latent_dtype]), "ffb200_plan_set_latent_dtype")
        self._keep: List[torch.Tensor] = []

    @property
    def workspace_bytes(self) -> int:
        return int(_lib.lib().ffb200_plan_workspace_bytes(self.handle))

    def __del__(self):
        try:
            if self.handle:
                _lib.lib().ffb200_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class RolloutEngine:
    """Owns the packed weights + the native engine; hands out geometry-specific plans."""

    def __init__(self, model_config, state_dict: Dict[str, torch.Tensor], device: Optional[torch.device] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("flow_factory_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device(device if device is not None else "cuda")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        torch.cuda.set_device(self.device)
        self.cfg = model_config if isinstance(model_config, EngineConfig) else EngineConfig.from_model_config(model_config)
        self.weights = PackedWeights(self.cfg, state_dict, self.device)
        mc = _lib.ModelConfig(self.cfg.num_layers, self.cfg.num_heads, self.cfg.patch_size, self.cfg.in_channels,
                              self.cfg.joint_attention_dim, self.cfg.pooled_projection_dim, self.cfg.pos_embed_max_size,
                              self.cfg.num_dual_layers)
        self.handle = C.c_void_p()
        _lib.check(_lib.lib().ffb200_engine_create(C.byref(mc), C.byref(self.weights.struct), C.byref(self.handle)),
                   "ffb200_engine_create")
        assert _lib.lib().ffb200_engine_mod_rows(self.handle) == self.weights.mod_rows
        self._plans: Dict[Tuple, Plan] = {}
        self._stream: Optional[torch.cuda.Stream] = None

    def refresh_weights(self, state_dict: Dict[str, torch.Tensor]) -> None:
        """Re-pack after the trainer changed the weights (optimizer step, EMA/ref swap, LoRA merge).  Addresses are
        kept stable, so existing plans (and their TMA descriptors) stay valid."""
        self.weights.pack(state_dict)
        _lib.check(_lib.lib().ffb200_engine_set_weights(self.handle, C.byref(self.weights.struct)), "ffb200_engine_set_weights")

    def plan(self, batch: int, cfg: bool, lat_h: int, lat_w: int, n_text: int, latent_dtype: torch.dtype = torch.float16) -> Plan:
        """`latent_dtype`: Flow-Factory's latent_storage_dtype (fp16 default, bf16, fp32) - the element type of every latents tensor that
        goes into / comes out of this plan, and the dtype a freshly sampled next_latents is rounded through before its log-prob."""
        key = (batch, bool(cfg), lat_h, lat_w, n_text, latent_dtype)
        if key not in self._plans:
            self._plans[key] = Plan(self, batch, bool(cfg), lat_h, lat_w, n_text, latent_dtype)
        return self._plans[key]

    def stream(self) -> torch.cuda.Stream:
        if self._stream is None:
            self._stream = torch.cuda.Stream(self.device)
        return self._stream

    # ------------------------------------------------------------------ conditioning
    def set_prompts(self, plan: Plan, prompt_embeds: torch.Tensor, pooled: torch.Tensor,
                    neg_prompt_embeds: Optional[torch.Tensor] = None, neg_pooled: Optional[torch.Tensor] = None) -> None:
        if plan.cfg:
            assert neg_prompt_embeds is not None and neg_pooled is not None
            pe = torch.cat([neg_prompt_embeds, prompt_embeds], dim=0)     # negative half first (sd3_5.py:409-413)
            pp = torch.cat([neg_pooled, pooled], dim=0)
        else:
            pe, pp = prompt_embeds, pooled
        pe = pe.to(device=self.device, dtype=torch.bfloat16).contiguous()
        pp = pp.to(device=self.device, dtype=torch.bfloat16).contiguous()
        assert pe.shape == (plan.batch * (2 if plan.cfg else 1), plan.n_text, self.cfg.joint_attention_dim), pe.shape
        plan._keep = [pe, pp]
        st = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(_lib.lib().ffb200_plan_set_prompts(plan.handle, pe.data_ptr(), pp.data_ptr(), st), "ffb200_plan_set_prompts")

    # ------------------------------------------------------------------ single forward / step
    def transformer_forward(self, plan: Plan, latents: torch.Tensor, t_model: float) -> torch.Tensor:
        """-> noise_pred bf16 [Bp, C, H, W] (both CFG halves, uncond first)."""
        x = latents.to(device=self.device, dtype=plan.latent_dtype).contiguous()
        bp = plan.batch * (2 if plan.cfg else 1)
        out = torch.empty((bp, self.cfg.in_channels, plan.lat_h, plan.lat_w), dtype=torch.bfloat16, device=self.device)
        st = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(_lib.lib().ffb200_transformer_forward(plan.handle, x.data_ptr(), float(t_model), out.data_ptr(), st),
                   "ffb200_transformer_forward")
        return out

    def step(self, plan: Plan, latents: torch.Tensor, coef: "_lib.StepCoef", guidance_scale: float,
             noise: Optional[torch.Tensor] = None, next_latents: Optional[torch.Tensor] = None, seed: int = 0,
             step_index: int = 0, want_mean: bool = True, want_noise_pred: bool = True) -> Dict[str, torch.Tensor]:
        x = latents.to(device=self.device, dtype=plan.latent_dtype).contiguous()
        shp = tuple(x.shape)
        nz = noise.to(device=self.device, dtype=torch.float32).contiguous() if noise is not None else None
        ng = next_latents.to(device=self.device, dtype=plan.latent_dtype).contiguous() if next_latents is not None else None
        o_next = torch.empty(shp, dtype=plan.latent_dtype, device=self.device)
        o_mean = torch.empty(shp, dtype=torch.float32, device=self.device) if want_mean else None
        o_lp = torch.zeros(shp[0], dtype=torch.float32, device=self.device) if coef.compute_log_prob else None
        o_v = torch.empty(shp, dtype=torch.bfloat16, device=self.device) if want_noise_pred else None
        flag = torch.zeros(1, dtype=torch.int32, device=self.device)
        a = _lib.StepArgs()
        a.latents = x.data_ptr(); a.coef = coef; a.guidance_scale = float(guidance_scale)
        a.noise = nz.data_ptr() if nz is not None else None
        a.seed = int(seed); a.step_index = int(step_index)
        a.next_latents = ng.data_ptr() if ng is not None else None
        a.out_next_latents = o_next.data_ptr()
        a.out_mean = o_mean.data_ptr() if o_mean is not None else None
        a.out_log_prob = o_lp.data_ptr() if o_lp is not None else None
        a.out_noise_pred = o_v.data_ptr() if o_v is not None else None
        a.overflow_flag = flag.data_ptr()
        st = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(_lib.lib().ffb200_step(plan.handle, C.byref(a), st), "ffb200_step")
        return dict(next_latents=o_next, next_latents_mean=o_mean, log_prob=o_lp, noise_pred=o_v, overflow=flag)

    # ------------------------------------------------------------------ the trajectory sampler
    def rollout(self, plan: Plan, x0: torch.Tensor, coefs: Sequence["_lib.StepCoef"], guidance_scale: float,
                n_latent_slots: int, store_initial_slot: int, n_logp_slots: int, noise: Optional[torch.Tensor] = None,
                seed: int = 0, use_graph: bool = True) -> Dict[str, torch.Tensor]:
        """Runs all steps on the device without host synchronisation.  Returns all_latents [B, slots, C,H,W] and final latents [B,C,H,W] in
        the plan's latent storage dtype, log_probs fp32 [B, logp_slots], overflow flag."""
        T = len(coefs)
        B, Cc, H, W = plan.batch, self.cfg.in_channels, plan.lat_h, plan.lat_w
        x = x0.to(device=self.device, dtype=plan.latent_dtype).contiguous()
        assert tuple(x.shape) == (B, Cc, H, W)
        arr = (_lib.StepCoef * T)(*coefs)
        traj = torch.empty((B, max(n_latent_slots, 1), Cc, H, W), dtype=plan.latent_dtype, device=self.device) if n_latent_slots else None
        lps = torch.zeros((B, max(n_logp_slots, 1)), dtype=torch.float32, device=self.device) if n_logp_slots else None
        final = torch.empty((B, Cc, H, W), dtype=plan.latent_dtype, device=self.device)
        flag = torch.zeros(1, dtype=torch.int32, device=self.device)
        nz = None
        if noise is not None:
            nz = noise.to(device=self.device, dtype=torch.float32).contiguous()
            assert tuple(nz.shape) == (T, B, Cc, H, W)
        a = _lib.RolloutArgs()
        a.num_steps = T; a.coefs = C.cast(arr, C.POINTER(_lib.StepCoef)); a.guidance_scale = float(guidance_scale)
        a.x0 = x.data_ptr(); a.noise = nz.data_ptr() if nz is not None else None; a.seed = int(seed)
        a.all_latents = traj.data_ptr() if traj is not None else None
        a.n_latent_slots = n_latent_slots; a.store_initial_slot = store_initial_slot
        a.log_probs = lps.data_ptr() if lps is not None else None; a.n_logp_slots = n_logp_slots
        a.final_latents = final.data_ptr(); a.overflow_flag = flag.data_ptr(); a.use_graph = int(use_graph)
        cur = torch.cuda.current_stream(self.device)
        if use_graph:
            s = self.stream()
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                _lib.check(_lib.lib().ffb200_rollout(plan.handle, C.byref(a), s.cuda_stream), "ffb200_rollout")
            cur.wait_stream(s)
            for t in (x, traj, lps, final, flag, nz):
                if t is not None:
                    t.record_stream(s)
        else:
            _lib.check(_lib.lib().ffb200_rollout(plan.handle, C.byref(a), cur.cuda_stream), "ffb200_rollout")
        return dict(all_latents=traj, log_probs=lps, final_latents=final, overflow=flag)

    def rollout_host(self, plan: Plan, x0: torch.Tensor, prompt_embeds: torch.Tensor, pooled: torch.Tensor,
                     coefs: Sequence["_lib.StepCoef"], guidance_scale: float, n_latent_slots: int, store_initial_slot: int,
                     n_logp_slots: int, seed: int = 0, use_graph: bool = True) -> Dict[str, torch.Tensor]:
        """End-to-end entry with HOST (pinned) buffers in and out: H2D / D2H copies are inside the call."""
        T = len(coefs)
        B, Cc, H, W = plan.batch, self.cfg.in_channels, plan.lat_h, plan.lat_w
        for t in (x0, prompt_embeds, pooled):
            assert not t.is_cuda and t.is_contiguous()
        assert x0.dtype == plan.latent_dtype and prompt_embeds.dtype == torch.bfloat16 and pooled.dtype == torch.bfloat16
        pin = lambda *shape, dtype: torch.empty(shape, dtype=dtype, pin_memory=True)
        traj = pin(B, max(n_latent_slots, 1), Cc, H, W, dtype=plan.latent_dtype) if n_latent_slots else None
        lps = pin(B, max(n_logp_slots, 1), dtype=torch.float32) if n_logp_slots else None
        final = pin(B, Cc, H, W, dtype=plan.latent_dtype)
        flag = torch.zeros(1, dtype=torch.int32).pin_memory()
        arr = (_lib.StepCoef * T)(*coefs)
        a = _lib.RolloutArgs()
        a.num_steps = T; a.coefs = C.cast(arr, C.POINTER(_lib.StepCoef)); a.guidance_scale = float(guidance_scale)
        a.x0 = x0.data_ptr(); a.noise = None; a.seed = int(seed)
        a.all_latents = traj.data_ptr() if traj is not None else None
        a.n_latent_slots = n_latent_slots; a.store_initial_slot = store_initial_slot
        a.log_probs = lps.data_ptr() if lps is not None else None; a.n_logp_slots = n_logp_slots
        a.final_latents = final.data_ptr(); a.overflow_flag = flag.data_ptr(); a.use_graph = int(use_graph)
        s = self.stream()
        _lib.check(_lib.lib().ffb200_rollout_host(plan.handle, C.byref(a), prompt_embeds.data_ptr(), pooled.data_ptr(),
                                                  s.cuda_stream), "ffb200_rollout_host")
        return dict(all_latents=traj, log_probs=lps, final_latents=final, overflow=flag)

    @staticmethod
    def last_launch_count() -> int:
        return int(_lib.lib().ffb200_last_launch_count())

    def __del__(self):
        try:
            self._plans.clear()
            if self.handle:
                _lib.lib().ffb200_engine_destroy(self.handle)
                self.handle = None
        except Exception:
            pass
