"""FLUX.1 rollout engine (SURVEY.md 8f row 2 / BASELINE config 3): Python face of the `ffb200_flux_*` C ABI.

Mirrors, for the no-grad rollout path only:
  FluxTransformer2DModel.forward            DF/models/transformers/transformer_flux.py:676-778
  Flux1Adapter.inference / forward          FF/models/flux/flux1.py:152-292 / 296-349
  FluxPipeline._pack_latents / _prepare_latent_image_ids / FluxPosEmbed (host-side index + table helpers)
Device memory, streams and tensors are torch's (plumbing); every kernel on the path is ours; no fallback."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .scheduler import calculate_shift, make_step_coef


@dataclass
class FluxEngineConfig:
    num_layers: int = 19
    num_single_layers: int = 38
    num_heads: int = 24
    in_channels: int = 64
    joint_attention_dim: int = 4096
    pooled_projection_dim: int = 768
    guidance_embeds: bool = True
    axes_dims_rope: Tuple[int, int, int] = (16, 56, 56)
    variant: int = 0     # 0 = FLUX.1 ; 1 = Qwen-Image (flow_factory_b200/qwen.py)

    @property
    def inner_dim(self) -> int:
        return 128 * self.num_heads

    @classmethod
    def from_model_config(cls, cfg) -> "FluxEngineConfig":
        """`cfg`: anything with the FluxTransformer2DModel config attributes (diffusers FrozenDict, oracle FluxConfig...)."""
        g = (lambda k: cfg[k]) if isinstance(cfg, dict) else (lambda k: getattr(cfg, k))
        if g("attention_head_dim") != 128:
            raise ValueError("the FLUX path is specialised for head_dim 128")
        if g("patch_size") != 1 or g("in_channels") != 64:
            raise ValueError("FLUX.1 packed latents: patch_size 1, in_channels 64")
        if sum(g("axes_dims_rope")) != 128:
            raise ValueError("axes_dims_rope must sum to the head dim (128)")
        return cls(num_layers=g("num_layers"), num_single_layers=g("num_single_layers"), num_heads=g("num_attention_heads"),
                   in_channels=g("in_channels"), joint_attention_dim=g("joint_attention_dim"),
                   pooled_projection_dim=g("pooled_projection_dim"), guidance_embeds=bool(g("guidance_embeds")),
                   axes_dims_rope=tuple(g("axes_dims_rope")))


# ---------------------------------------------------------------------------------------------- host-side helpers
def pack_latents(lat: torch.Tensor) -> torch.Tensor:
    """FluxPipeline._pack_latents (DF/pipelines/flux/pipeline_flux.py:521-526): [B, C, H, W] -> [B, (H/2)(W/2), 4C]."""
    B, Cc, H, W = lat.shape
    return lat.view(B, Cc, H // 2, 2, W // 2, 2).permute(0, 2, 4, 1, 3, 5).reshape(B, (H // 2) * (W // 2), Cc * 4)


def latent_image_ids(h2: int, w2: int) -> torch.Tensor:
    """FluxPipeline._prepare_latent_image_ids (pipeline_flux.py:507-518): [h2*w2, 3] = (0, row, col)."""
    ids = torch.zeros(h2, w2, 3)
    ids[..., 1] = ids[..., 1] + torch.arange(h2)[:, None]
    ids[..., 2] = ids[..., 2] + torch.arange(w2)[None, :]
    return ids.reshape(h2 * w2, 3)


def rope_tables(ids: torch.Tensor, axes_dim: Sequence[int], theta: float = 10000.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """FluxPosEmbed.forward (transformer_flux.py:500-522) + get_1d_rotary_pos_embed(use_real, repeat_interleave_real)
    (embeddings.py:1155-1174): float64 frequencies, cos / sin fp32 [S, sum(axes_dim)] with every value repeated twice."""
    pos = ids.float().cpu()
    cos_out, sin_out = [], []
    for i, dim in enumerate(axes_dim):
        freqs = 1.0 / (theta ** (torch.arange(0, dim, 2, dtype=torch.float64) / dim))
        freqs = torch.outer(pos[:, i], freqs)
        cos_out.append(freqs.cos().repeat_interleave(2, dim=1).float())
        sin_out.append(freqs.sin().repeat_interleave(2, dim=1).float())
    return torch.cat(cos_out, dim=-1).contiguous(), torch.cat(sin_out, dim=-1).contiguous()


def model_scalar(x: float, in_dtype: torch.dtype = torch.float32) -> float:
    """What the sinusoid sees for a scalar conditioning input: `x.to(hidden_states.dtype) * 1000` in bf16
    (transformer_flux.py:679-682), `x` first held in `in_dtype` by the caller (flux1.py:318, 325)."""
    t = torch.as_tensor(x, dtype=torch.float32).to(in_dtype).to(torch.bfloat16) * 1000
    return float(t.float())


def flux_make_schedule(num_inference_steps: int, image_seq_len: int, num_train_timesteps: int = 1000):
    """set_scheduler_timesteps with dynamic shifting (FF/scheduler/flow_match_euler_discrete.py:49-77 ->
    DF/schedulers/scheduling_flow_match_euler_discrete.py:346-348, 648-649)."""
    import math
    mu = calculate_shift(image_seq_len)
    sig = np.linspace(1.0, 1 / num_inference_steps, num_inference_steps).astype(np.float32)
    sig = math.exp(mu) / (math.exp(mu) + (1 / sig - 1) ** 1.0)
    sigmas = torch.from_numpy(sig).to(dtype=torch.float32)
    return sigmas * num_train_timesteps, torch.cat([sigmas, torch.zeros(1)])


# ---------------------------------------------------------------------------------------------- ctypes mirrors
DUAL_FIELDS = ("qkv_w", "qkv_b", "norm_q", "norm_k", "add_qkv_w", "add_qkv_b", "norm_added_q", "norm_added_k", "out_w", "out_b",
               "add_out_w", "add_out_b", "ff1_w", "ff1_b", "ff2_w", "ff2_b", "cff1_w", "cff1_b", "cff2_w", "cff2_b")
SINGLE_FIELDS = ("qkv_w", "qkv_b", "norm_q", "norm_k", "mlp_w", "mlp_b", "out_w", "out_b")
GLOBAL_FIELDS = ("x_w", "x_b", "ctx_w", "ctx_b", "t1_w", "t1_b", "t2_w", "t2_b", "g1_w", "g1_b", "g2_w", "g2_b",
                 "p1_w", "p1_b", "p2_w", "p2_b", "mod_w", "mod_b", "proj_w", "proj_b", "ctxn_w")


class FluxConfigC(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("num_layers", "num_single_layers", "num_heads", "in_channels", "joint_attention_dim",
                                       "pooled_projection_dim", "guidance_embeds", "variant")]


class FluxDualWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in DUAL_FIELDS]


class FluxSingleWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in SINGLE_FIELDS]


class FluxWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in GLOBAL_FIELDS] + [("dual", C.POINTER(FluxDualWeights)), ("single", C.POINTER(FluxSingleWeights))]


_bound = False


def _L() -> C.CDLL:
    global _bound
    L = _lib.lib()
    if not _bound:
        vp, ci, cf = C.c_void_p, C.c_int, C.c_float
        L.ffb200_flux_engine_create.argtypes = [C.POINTER(FluxConfigC), C.POINTER(FluxWeights), C.POINTER(vp)]
        L.ffb200_flux_engine_set_weights.argtypes = [vp, C.POINTER(FluxWeights)]
        L.ffb200_flux_engine_destroy.argtypes = [vp]; L.ffb200_flux_engine_destroy.restype = None
        L.ffb200_flux_engine_mod_rows.argtypes = [vp]
        L.ffb200_flux_plan_create.argtypes = [vp, ci, ci, ci, vp, vp, C.POINTER(vp)]
        L.ffb200_flux_plan_create_ex.argtypes = [vp, ci, ci, ci, ci, vp, vp, C.POINTER(vp)]
        L.ffb200_flux_set_text_lengths.argtypes = [vp, C.POINTER(ci), vp]
        L.ffb200_flux_plan_destroy.argtypes = [vp]; L.ffb200_flux_plan_destroy.restype = None
        L.ffb200_flux_plan_workspace_bytes.argtypes = [vp]; L.ffb200_flux_plan_workspace_bytes.restype = C.c_longlong
        L.ffb200_flux_set_prompts.argtypes = [vp, vp, vp, cf, vp]
        L.ffb200_flux_forward.argtypes = [vp, vp, cf, vp, vp]
        L.ffb200_flux_step.argtypes = [vp, C.POINTER(_lib.StepArgs), vp]
        L.ffb200_flux_rollout.argtypes = [vp, C.POINTER(_lib.RolloutArgs), vp]
        _bound = True
    return L


class FluxPackedWeights:
    """FluxTransformer2DModel.state_dict() (diffusers key names) -> the flat bf16 device tensors the C ABI borrows:
    q|k|v concatenated along out_features, all adaLN projections stacked ([norm1 ; norm1_context] per dual block, norm per
    single block, norm_out).  Addresses stay stable across pack() calls (plans hold TMA descriptors on them)."""

    def __init__(self, cfg: FluxEngineConfig, state_dict: Dict[str, torch.Tensor], device: torch.device):
        self.cfg, self.device = cfg, device
        self.tensors: Dict[str, torch.Tensor] = {}
        self.dual_structs = (FluxDualWeights * max(cfg.num_layers, 1))()
        self.single_structs = (FluxSingleWeights * max(cfg.num_single_layers, 1))()
        self.struct = FluxWeights()
        self.pack(state_dict)

    def _put(self, name: str, t: torch.Tensor) -> int:
        t = t.detach().to(device=self.device, dtype=torch.bfloat16).contiguous()
        old = self.tensors.get(name)
        if old is not None and old.shape == t.shape:
            old.copy_(t)
            t = old
        else:
            self.tensors[name] = t
        assert t.data_ptr() % 16 == 0
        return t.data_ptr()

    def pack(self, sd: Dict[str, torch.Tensor]) -> None:
        cfg, D = self.cfg, self.cfg.inner_dim
        cat = lambda names: torch.cat([sd[n] for n in names], dim=0)
        W = self.struct
        for f in GLOBAL_FIELDS:
            setattr(W, f, None)
        qwen = cfg.variant == 1
        # Qwen-Image module names (transformer_qwenimage.py:829-846, 622-650) for the same roles
        k_norm1, k_norm1c, k_ff, k_ffc = (("img_mod.1", "txt_mod.1", "img_mlp", "txt_mlp") if qwen else
                                          ("norm1.linear", "norm1_context.linear", "ff", "ff_context"))
        names = [("x", "img_in" if qwen else "x_embedder"), ("ctx", "txt_in" if qwen else "context_embedder"),
                 ("t1", "time_text_embed.timestep_embedder.linear_1"), ("t2", "time_text_embed.timestep_embedder.linear_2"),
                 ("proj", "proj_out")]
        if qwen:
            W.ctxn_w = self._put("ctxn_w", sd["txt_norm.weight"])
        else:
            names += [("p1", "time_text_embed.text_embedder.linear_1"), ("p2", "time_text_embed.text_embedder.linear_2")]
        if cfg.guidance_embeds:
            names += [("g1", "time_text_embed.guidance_embedder.linear_1"), ("g2", "time_text_embed.guidance_embedder.linear_2")]
        for short, key in names:
            setattr(W, short + "_w", self._put(short + "_w", sd[key + ".weight"]))
            setattr(W, short + "_b", self._put(short + "_b", sd[key + ".bias"]))
        mod_w: List[torch.Tensor] = []
        mod_b: List[torch.Tensor] = []
        for i in range(cfg.num_layers):
            pre, a = f"transformer_blocks.{i}.", f"transformer_blocks.{i}.attn."
            mod_w += [sd[pre + k_norm1 + ".weight"], sd[pre + k_norm1c + ".weight"]]
            mod_b += [sd[pre + k_norm1 + ".bias"], sd[pre + k_norm1c + ".bias"]]
            L = self.dual_structs[i]
            put = lambda field, t: setattr(L, field, self._put(f"D{i}.{field}", t))
            put("qkv_w", cat([a + "to_q.weight", a + "to_k.weight", a + "to_v.weight"]))
            put("qkv_b", cat([a + "to_q.bias", a + "to_k.bias", a + "to_v.bias"]))
            put("norm_q", sd[a + "norm_q.weight"]); put("norm_k", sd[a + "norm_k.weight"])
            put("add_qkv_w", cat([a + "add_q_proj.weight", a + "add_k_proj.weight", a + "add_v_proj.weight"]))
            put("add_qkv_b", cat([a + "add_q_proj.bias", a + "add_k_proj.bias", a + "add_v_proj.bias"]))
            put("norm_added_q", sd[a + "norm_added_q.weight"]); put("norm_added_k", sd[a + "norm_added_k.weight"])
            put("out_w", sd[a + "to_out.0.weight"]); put("out_b", sd[a + "to_out.0.bias"])
            put("add_out_w", sd[a + "to_add_out.weight"]); put("add_out_b", sd[a + "to_add_out.bias"])
            put("ff1_w", sd[pre + k_ff + ".net.0.proj.weight"]); put("ff1_b", sd[pre + k_ff + ".net.0.proj.bias"])
            put("ff2_w", sd[pre + k_ff + ".net.2.weight"]); put("ff2_b", sd[pre + k_ff + ".net.2.bias"])
            put("cff1_w", sd[pre + k_ffc + ".net.0.proj.weight"]); put("cff1_b", sd[pre + k_ffc + ".net.0.proj.bias"])
            put("cff2_w", sd[pre + k_ffc + ".net.2.weight"]); put("cff2_b", sd[pre + k_ffc + ".net.2.bias"])
        for i in range(cfg.num_single_layers):
            pre, a = f"single_transformer_blocks.{i}.", f"single_transformer_blocks.{i}.attn."
            mod_w.append(sd[pre + "norm.linear.weight"]); mod_b.append(sd[pre + "norm.linear.bias"])
            L = self.single_structs[i]
            put = lambda field, t: setattr(L, field, self._put(f"S{i}.{field}", t))
            put("qkv_w", cat([a + "to_q.weight", a + "to_k.weight", a + "to_v.weight"]))
            put("qkv_b", cat([a + "to_q.bias", a + "to_k.bias", a + "to_v.bias"]))
            put("norm_q", sd[a + "norm_q.weight"]); put("norm_k", sd[a + "norm_k.weight"])
            put("mlp_w", sd[pre + "proj_mlp.weight"]); put("mlp_b", sd[pre + "proj_mlp.bias"])
            put("out_w", sd[pre + "proj_out.weight"]); put("out_b", sd[pre + "proj_out.bias"])
        mod_w.append(sd["norm_out.linear.weight"]); mod_b.append(sd["norm_out.linear.bias"])
        W.mod_w = self._put("mod_w", torch.cat(mod_w, dim=0))
        W.mod_b = self._put("mod_b", torch.cat(mod_b, dim=0))
        self.mod_rows = self.tensors["mod_w"].shape[0]
        W.dual = C.cast(self.dual_structs, C.POINTER(FluxDualWeights))
        W.single = C.cast(self.single_structs, C.POINTER(FluxSingleWeights))

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self.tensors.values())


class FluxPlan:
    def __init__(self, engine: "FluxRolloutEngine", batch: int, h2: int, w2: int, n_text: int, cfg: bool = False):
        self.engine, self.batch, self.h2, self.w2, self.n_text, self.cfg = engine, batch, h2, w2, n_text, bool(cfg)
        self.n_img = h2 * w2
        self.img_ids = latent_image_ids(h2, w2)
        cos, sin = engine.rope_tables(h2, w2, n_text)                        # fp32 [n_text + n_img, 128], text rows first
        self.handle = C.c_void_p()
        _lib.check(_L().ffb200_flux_plan_create_ex(engine.handle, batch, int(self.cfg), self.n_img, n_text, cos.data_ptr(), sin.data_ptr(),
                                                   C.byref(self.handle)), "ffb200_flux_plan_create_ex")
        self._keep: List[torch.Tensor] = []

    @property
    def workspace_bytes(self) -> int:
        return int(_L().ffb200_flux_plan_workspace_bytes(self.handle))

    def __del__(self):
        try:
            if self.handle:
                _L().ffb200_flux_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class FluxRolloutEngine:
    """Owns the packed FLUX.1 weights + the native engine; hands out geometry-specific plans."""

    def __init__(self, model_config, state_dict: Dict[str, torch.Tensor], device: Optional[torch.device] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("flow_factory_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device(device if device is not None else "cuda")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        torch.cuda.set_device(self.device)
        self.cfg = model_config if isinstance(model_config, FluxEngineConfig) else FluxEngineConfig.from_model_config(model_config)
        self.weights = FluxPackedWeights(self.cfg, state_dict, self.device)
        mc = FluxConfigC(self.cfg.num_layers, self.cfg.num_single_layers, self.cfg.num_heads, self.cfg.in_channels,
                         self.cfg.joint_attention_dim, self.cfg.pooled_projection_dim, int(self.cfg.guidance_embeds), int(self.cfg.variant))
        self.handle = C.c_void_p()
        _lib.check(_L().ffb200_flux_engine_create(C.byref(mc), C.byref(self.weights.struct), C.byref(self.handle)),
                   "ffb200_flux_engine_create")
        assert _L().ffb200_flux_engine_mod_rows(self.handle) == self.weights.mod_rows
        self._plans: Dict[Tuple, FluxPlan] = {}
        self._stream: Optional[torch.cuda.Stream] = None

    def refresh_weights(self, state_dict: Dict[str, torch.Tensor]) -> None:
        self.weights.pack(state_dict)
        _lib.check(_L().ffb200_flux_engine_set_weights(self.handle, C.byref(self.weights.struct)), "ffb200_flux_engine_set_weights")

    def rope_tables(self, h2: int, w2: int, n_text: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """cos / sin fp32 [n_text + h2*w2, 128] for ids = cat(txt_ids = 0 (flux1.py:330), img_ids)."""
        ids = torch.cat([torch.zeros(n_text, 3), latent_image_ids(h2, w2)], dim=0)
        return rope_tables(ids, self.cfg.axes_dims_rope)

    def plan(self, batch: int, h2: int, w2: int, n_text: int, cfg: bool = False) -> FluxPlan:
        key = (batch, h2, w2, n_text, bool(cfg))
        if key not in self._plans:
            self._plans[key] = FluxPlan(self, batch, h2, w2, n_text, cfg)
        return self._plans[key]

    def stream(self) -> torch.cuda.Stream:
        if self._stream is None:
            self._stream = torch.cuda.Stream(self.device)
        return self._stream

    def set_prompts(self, plan: FluxPlan, prompt_embeds: torch.Tensor, pooled: torch.Tensor, guidance_scale: float,
                    latents_dtype: torch.dtype = torch.float16) -> None:
        pe = prompt_embeds.to(device=self.device, dtype=torch.bfloat16).contiguous()
        pp = pooled.to(device=self.device, dtype=torch.bfloat16).contiguous()
        assert pe.shape == (plan.batch, plan.n_text, self.cfg.joint_attention_dim), pe.shape
        plan._keep = [pe, pp]
        st = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(_L().ffb200_flux_set_prompts(plan.handle, pe.data_ptr(), pp.data_ptr(), model_scalar(guidance_scale, latents_dtype), st),
                   "ffb200_flux_set_prompts")

    def transformer_forward(self, plan: FluxPlan, packed_latents: torch.Tensor, timestep: float) -> torch.Tensor:
        """`timestep` on the scheduler's 0..1000 scale; -> noise prediction bf16 [B, Ni, 64]."""
        x = packed_latents.to(device=self.device, dtype=torch.float16).contiguous()
        assert tuple(x.shape) == (plan.batch, plan.n_img, 64)
        out = torch.empty_like(x, dtype=torch.bfloat16)
        st = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(_L().ffb200_flux_forward(plan.handle, x.data_ptr(), model_scalar(float(timestep) / 1000), out.data_ptr(), st),
                   "ffb200_flux_forward")
        return out

    def step(self, plan: FluxPlan, latents: torch.Tensor, coef: "_lib.StepCoef", noise: Optional[torch.Tensor] = None,
             next_latents: Optional[torch.Tensor] = None, seed: int = 0) -> Dict[str, torch.Tensor]:
        x = latents.to(device=self.device, dtype=torch.float16).contiguous()
        shp = tuple(x.shape)
        nz = noise.to(device=self.device, dtype=torch.float32).contiguous() if noise is not None else None
        ng = next_latents.to(device=self.device, dtype=torch.float16).contiguous() if next_latents is not None else None
        o_next = torch.empty(shp, dtype=torch.float16, device=self.device)
        o_mean = torch.empty(shp, dtype=torch.float32, device=self.device)
        o_lp = torch.zeros(shp[0], dtype=torch.float32, device=self.device) if coef.compute_log_prob else None
        o_v = torch.empty(shp, dtype=torch.bfloat16, device=self.device)
        flag = torch.zeros(1, dtype=torch.int32, device=self.device)
        a = _lib.StepArgs()
        a.latents = x.data_ptr(); a.coef = coef; a.guidance_scale = 1.0
        a.noise = nz.data_ptr() if nz is not None else None
        a.seed = int(seed); a.step_index = 0
        a.next_latents = ng.data_ptr() if ng is not None else None
        a.out_next_latents = o_next.data_ptr(); a.out_mean = o_mean.data_ptr()
        a.out_log_prob = o_lp.data_ptr() if o_lp is not None else None
        a.out_noise_pred = o_v.data_ptr(); a.overflow_flag = flag.data_ptr()
        st = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(_L().ffb200_flux_step(plan.handle, C.byref(a), st), "ffb200_flux_step")
        return dict(next_latents=o_next, next_latents_mean=o_mean, log_prob=o_lp, noise_pred=o_v, overflow=flag)

    def make_coefs(self, plan: FluxPlan, num_steps: int, noise_level: float, sde_steps: Sequence[int], dynamics: str = "Flow-SDE",
                   compute_log_prob: bool = True, store_slots: Optional[Sequence[int]] = None,
                   logp_slots: Optional[Sequence[int]] = None):
        """Per-step scalar blocks for a rollout with resolution-dependent shifting (mu from the image sequence length)."""
        timesteps, sigmas = flux_make_schedule(num_steps, plan.n_img)
        sde = set(int(i) for i in sde_steps)
        coefs = []
        for i in range(num_steps):
            nl = noise_level if i in sde else 0.0
            coefs.append(make_step_coef(float(sigmas[i]), float(sigmas[i + 1]), nl, float(sigmas[1]), dynamics,
                                        t_model=model_scalar(float(timesteps[i]) / 1000),
                                        compute_log_prob=compute_log_prob and nl > 0,
                                        store_slot=-1 if store_slots is None else int(store_slots[i]),
                                        logp_slot=-1 if logp_slots is None else int(logp_slots[i])))
        return timesteps, sigmas, coefs

    def rollout(self, plan: FluxPlan, x0: torch.Tensor, coefs: Sequence["_lib.StepCoef"], n_latent_slots: int,
                store_initial_slot: int, n_logp_slots: int, noise: Optional[torch.Tensor] = None, seed: int = 0,
                use_graph: bool = True) -> Dict[str, torch.Tensor]:
        """All steps on the device without host synchronisation: all_latents fp16 [B, slots, Ni, 64], log_probs fp32
        [B, logp_slots], final latents fp16 [B, Ni, 64]."""
        T = len(coefs)
        B, Ni = plan.batch, plan.n_img
        x = x0.to(device=self.device, dtype=torch.float16).contiguous()
        assert tuple(x.shape) == (B, Ni, 64)
        arr = (_lib.StepCoef * T)(*coefs)
        traj = torch.empty((B, max(n_latent_slots, 1), Ni, 64), dtype=torch.float16, device=self.device) if n_latent_slots else None
        lps = torch.zeros((B, max(n_logp_slots, 1)), dtype=torch.float32, device=self.device) if n_logp_slots else None
        final = torch.empty((B, Ni, 64), dtype=torch.float16, device=self.device)
        flag = torch.zeros(1, dtype=torch.int32, device=self.device)
        nz = None
        if noise is not None:
            nz = noise.to(device=self.device, dtype=torch.float32).contiguous()
            assert tuple(nz.shape) == (T, B, Ni, 64)
        a = _lib.RolloutArgs()
        a.num_steps = T; a.coefs = C.cast(arr, C.POINTER(_lib.StepCoef)); a.guidance_scale = 1.0
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
                _lib.check(_L().ffb200_flux_rollout(plan.handle, C.byref(a), s.cuda_stream), "ffb200_flux_rollout")
            cur.wait_stream(s)
            for t in (x, traj, lps, final, flag, nz):
                if t is not None:
                    t.record_stream(s)
        else:
            _lib.check(_L().ffb200_flux_rollout(plan.handle, C.byref(a), cur.cuda_stream), "ffb200_flux_rollout")
        return dict(all_latents=traj, log_probs=lps, final_latents=final, overflow=flag)

    @staticmethod
    def last_launch_count() -> int:
        return int(_lib.lib().ffb200_last_launch_count())

    def __del__(self):
        try:
            self._plans.clear()
            if self.handle:
                _L().ffb200_flux_engine_destroy(self.handle)
                self.handle = None
        except Exception:
            pass
