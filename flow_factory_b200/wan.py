"""Host side of the Wan2.1 T2V rollout path (SURVEY.md section 8f row 4 / BASELINE config 4): weight packing, RoPE tables, the ctypes
binding of `ffb200_wan_*` and a rollout-level engine class in the style of `flux.py`.

STATUS: validated on B200 in round 2 (tests/test_gpu_wan.py against the pinned oracle; rollout lines under profiles/r02_*wan21*).
Packing, RoPE tables and the UniPC schedule are also unit-tested on CPU against the pinned oracle / reference-minted fixtures
(tests/test_host_logic_wan.py).

Reference: WanTransformer3DModel (DF/models/transformers/transformer_wan.py:507-740) behind Wan2_T2V_Adapter
(FF/models/wan/wan2_t2v.py:235-543)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Any, Dict, List, Mapping, Optional, Sequence, Tuple

import torch

from . import _lib
from .scheduler import UniPCMultistepSDEScheduler

vp, ci, cf = C.c_void_p, C.c_int, C.c_float


@dataclass
class WanEngineConfig:
    num_layers: int = 30                 # Wan2.1-T2V-1.3B (HF model card; the in-tree defaults are the 14 B model)
    num_attention_heads: int = 12
    attention_head_dim: int = 128
    in_channels: int = 16
    out_channels: int = 16
    text_dim: int = 4096
    freq_dim: int = 256
    ffn_dim: int = 8960
    patch_size: Tuple[int, int, int] = (1, 2, 2)
    eps: float = 1e-6
    rope_max_seq_len: int = 1024
    cross_attn_norm: bool = True

    @property
    def inner_dim(self) -> int:
        return self.num_attention_heads * self.attention_head_dim

    @classmethod
    def from_model_config(cls, cfg: Any) -> "WanEngineConfig":
        get = (lambda k, d=None: cfg.get(k, d)) if isinstance(cfg, Mapping) else (lambda k, d=None: getattr(cfg, k, d))
        if get("image_dim") is not None or get("added_kv_proj_dim") is not None:
            raise NotImplementedError("image-conditioned Wan checkpoints (I2V) are not on the accelerated path")
        if get("qk_norm", "rms_norm_across_heads") != "rms_norm_across_heads":
            raise NotImplementedError("only qk_norm='rms_norm_across_heads' is implemented")
        if not get("cross_attn_norm", True):
            raise NotImplementedError("cross_attn_norm=False is not implemented")
        if int(get("attention_head_dim", 128)) != 128:
            raise NotImplementedError("head_dim must be 128")
        return cls(num_layers=int(get("num_layers")), num_attention_heads=int(get("num_attention_heads")),
                   in_channels=int(get("in_channels", 16)), out_channels=int(get("out_channels", 16)), text_dim=int(get("text_dim", 4096)),
                   freq_dim=int(get("freq_dim", 256)), ffn_dim=int(get("ffn_dim")), patch_size=tuple(get("patch_size", (1, 2, 2))),
                   eps=float(get("eps", 1e-6)), rope_max_seq_len=int(get("rope_max_seq_len", 1024)))


def wan_rope_tables(cfg: WanEngineConfig, ppf: int, pph: int, ppw: int, table_dtype: torch.dtype = torch.bfloat16,
                    theta: float = 10000.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """WanRotaryPosEmbed (transformer_wan.py:354-417): fp32 [ppf*pph*ppw, head_dim] cos / sin with float64 frequencies, every value
    repeated twice, head_dim split (t, h, w) = (d - 4 (d // 6), 2 (d // 6), 2 (d // 6)).  The module keeps them as non-persistent BUFFERS,
    so `model.to(bf16)` rounds them with the weights: `table_dtype` applies that rounding (the values are returned as fp32)."""
    d = cfg.attention_head_dim
    h_dim = w_dim = 2 * (d // 6)
    t_dim = d - h_dim - w_dim
    cos_l, sin_l = [], []
    for dim in (t_dim, h_dim, w_dim):
        freqs = 1.0 / (theta ** (torch.arange(0, dim, 2, dtype=torch.float64) / dim))
        freqs = torch.outer(torch.arange(cfg.rope_max_seq_len), freqs)
        cos_l.append(freqs.cos().repeat_interleave(2, dim=1).float())
        sin_l.append(freqs.sin().repeat_interleave(2, dim=1).float())
    if max(ppf, pph, ppw) > cfg.rope_max_seq_len:
        raise ValueError("token grid exceeds rope_max_seq_len")
    ex = lambda t, n, shape: t[:n].view(*shape, -1).expand(ppf, pph, ppw, -1)
    cos = torch.cat([ex(cos_l[0], ppf, (ppf, 1, 1)), ex(cos_l[1], pph, (1, pph, 1)), ex(cos_l[2], ppw, (1, 1, ppw))], dim=-1)
    sin = torch.cat([ex(sin_l[0], ppf, (ppf, 1, 1)), ex(sin_l[1], pph, (1, pph, 1)), ex(sin_l[2], ppw, (1, 1, ppw))], dim=-1)
    S = ppf * pph * ppw
    return (cos.reshape(S, d).to(table_dtype).float().contiguous(), sin.reshape(S, d).to(table_dtype).float().contiguous())


# ---------------------------------------------------------------------------------------------- ctypes mirrors
LAYER_FIELDS = ("table", "qkv_w", "qkv_b", "norm_q", "norm_k", "out_w", "out_b", "norm2_w", "norm2_b", "q2_w", "q2_b", "kv2_w", "kv2_b",
                "norm_q2", "norm_k2", "out2_w", "out2_b", "ff1_w", "ff1_b", "ff2_w", "ff2_b")
GLOBAL_FIELDS = ("pe_w", "pe_b", "t1_w", "t1_b", "t2_w", "t2_b", "tp_w", "tp_b", "x1_w", "x1_b", "x2_w", "x2_b", "table", "proj_w", "proj_b")


class WanConfigC(C.Structure):
    _fields_ = [(n, ci) for n in ("num_layers", "num_heads", "in_channels", "text_dim", "freq_dim", "ffn_dim", "patch_t", "patch_h", "patch_w")] + \
               [("eps", cf)]


class WanLayerWeights(C.Structure):
    _fields_ = [(n, vp) for n in LAYER_FIELDS]


class WanWeights(C.Structure):
    _fields_ = [(n, vp) for n in GLOBAL_FIELDS] + [("layers", C.POINTER(WanLayerWeights))]


_bound = False


def _L() -> C.CDLL:
    global _bound
    L = _lib.lib()
    if not _bound:
        L.ffb200_wan_engine_create.argtypes = [C.POINTER(WanConfigC), C.POINTER(WanWeights), C.POINTER(vp)]
        L.ffb200_wan_engine_set_weights.argtypes = [vp, C.POINTER(WanWeights)]
        L.ffb200_wan_engine_destroy.argtypes = [vp]; L.ffb200_wan_engine_destroy.restype = None
        L.ffb200_wan_plan_create.argtypes = [vp, ci, ci, ci, ci, ci, ci, vp, vp, C.POINTER(vp)]
        L.ffb200_wan_plan_destroy.argtypes = [vp]; L.ffb200_wan_plan_destroy.restype = None
        L.ffb200_wan_plan_workspace_bytes.argtypes = [vp]; L.ffb200_wan_plan_workspace_bytes.restype = C.c_longlong
        L.ffb200_wan_set_prompts.argtypes = [vp, vp, vp]
        L.ffb200_wan_forward.argtypes = [vp, vp, cf, cf, vp, vp]
        L.ffb200_wan_step.argtypes = [vp, C.POINTER(_lib.StepArgs), vp]
        L.ffb200_wan_rollout.argtypes = [vp, C.POINTER(_lib.RolloutArgs), vp]
        cll = C.c_longlong
        L.ffb200_wan_rms_rope.argtypes = [vp, cll, ci, ci, ci, vp, cf, vp, vp, vp]
        L.ffb200_wan_layer_norm.argtypes = [vp, vp, ci, ci, ci, cf, ci, vp, vp, cll, vp, vp, vp]
        L.ffb200_wan_gate_residual.argtypes = [vp, vp, vp, cll, ci, cll, ci, vp]
        L.ffb200_wan_patchify.argtypes = [vp, ci, ci, ci, ci, ci, ci, ci, ci, ci, vp, vp]
        L.ffb200_attention_cross.argtypes = [vp, ci, vp, ci, ci, ci, ci, ci, ci, vp, vp]
        _bound = True
    return L


def pack_wan_state_dict(cfg: WanEngineConfig, sd: Mapping[str, torch.Tensor]) -> Tuple[Dict[str, torch.Tensor], List[Dict[str, torch.Tensor]]]:
    """WanTransformer3DModel.state_dict() (diffusers key names) -> ({global field: tensor}, [{layer field: tensor}]) in the layout the C ABI
    expects: attn1 q|k|v and attn2 k|v concatenated along out_features, the Conv3d patch embedding flattened to [D, C*pt*ph*pw], the
    scale_shift_tables flattened to [6 D] / [2 D].  Pure torch (CPU-testable); dtype conversion happens in WanPackedWeights."""
    D = cfg.inner_dim
    cat = lambda names: torch.cat([sd[n] for n in names], dim=0)
    g = {
        "pe_w": sd["patch_embedding.weight"].reshape(D, -1), "pe_b": sd["patch_embedding.bias"],
        "t1_w": sd["condition_embedder.time_embedder.linear_1.weight"], "t1_b": sd["condition_embedder.time_embedder.linear_1.bias"],
        "t2_w": sd["condition_embedder.time_embedder.linear_2.weight"], "t2_b": sd["condition_embedder.time_embedder.linear_2.bias"],
        "tp_w": sd["condition_embedder.time_proj.weight"], "tp_b": sd["condition_embedder.time_proj.bias"],
        "x1_w": sd["condition_embedder.text_embedder.linear_1.weight"], "x1_b": sd["condition_embedder.text_embedder.linear_1.bias"],
        "x2_w": sd["condition_embedder.text_embedder.linear_2.weight"], "x2_b": sd["condition_embedder.text_embedder.linear_2.bias"],
        "table": sd["scale_shift_table"].reshape(2 * D), "proj_w": sd["proj_out.weight"], "proj_b": sd["proj_out.bias"],
    }
    if any(k.startswith("condition_embedder.image_embedder") for k in sd):
        raise NotImplementedError("image-conditioned Wan checkpoints (I2V) are not on the accelerated path")
    layers = []
    for i in range(cfg.num_layers):
        p, a1, a2 = f"blocks.{i}.", f"blocks.{i}.attn1.", f"blocks.{i}.attn2."
        layers.append({
            "table": sd[p + "scale_shift_table"].reshape(6 * D),
            "qkv_w": cat([a1 + "to_q.weight", a1 + "to_k.weight", a1 + "to_v.weight"]),
            "qkv_b": cat([a1 + "to_q.bias", a1 + "to_k.bias", a1 + "to_v.bias"]),
            "norm_q": sd[a1 + "norm_q.weight"], "norm_k": sd[a1 + "norm_k.weight"],
            "out_w": sd[a1 + "to_out.0.weight"], "out_b": sd[a1 + "to_out.0.bias"],
            "norm2_w": sd[p + "norm2.weight"], "norm2_b": sd[p + "norm2.bias"],
            "q2_w": sd[a2 + "to_q.weight"], "q2_b": sd[a2 + "to_q.bias"],
            "kv2_w": cat([a2 + "to_k.weight", a2 + "to_v.weight"]), "kv2_b": cat([a2 + "to_k.bias", a2 + "to_v.bias"]),
            "norm_q2": sd[a2 + "norm_q.weight"], "norm_k2": sd[a2 + "norm_k.weight"],
            "out2_w": sd[a2 + "to_out.0.weight"], "out2_b": sd[a2 + "to_out.0.bias"],
            "ff1_w": sd[p + "ffn.net.0.proj.weight"], "ff1_b": sd[p + "ffn.net.0.proj.bias"],
            "ff2_w": sd[p + "ffn.net.2.weight"], "ff2_b": sd[p + "ffn.net.2.bias"],
        })
    return g, layers


class WanPackedWeights:
    """Flat bf16 device tensors the C ABI borrows; addresses stay stable across pack() calls (plans hold TMA descriptors on them)."""

    def __init__(self, cfg: WanEngineConfig, state_dict: Mapping[str, torch.Tensor], device: torch.device):
        self.cfg, self.device = cfg, device
        self.tensors: Dict[str, torch.Tensor] = {}
        self.layer_structs = (WanLayerWeights * cfg.num_layers)()
        self.struct = WanWeights()
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

    def pack(self, sd: Mapping[str, torch.Tensor]) -> None:
        g, layers = pack_wan_state_dict(self.cfg, sd)
        for f in GLOBAL_FIELDS:
            setattr(self.struct, f, self._put(f, g[f]))
        for i, lw in enumerate(layers):
            for f in LAYER_FIELDS:
                setattr(self.layer_structs[i], f, self._put(f"L{i}.{f}", lw[f]))
        self.struct.layers = C.cast(self.layer_structs, C.POINTER(WanLayerWeights))

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self.tensors.values())


class WanPlan:
    def __init__(self, engine: "WanRolloutEngine", batch: int, frames: int, height: int, width: int, n_text: int, cfg: bool):
        self.engine, self.batch, self.frames, self.height, self.width, self.n_text, self.cfg = engine, batch, frames, height, width, n_text, bool(cfg)
        pt, ph, pw = engine.cfg.patch_size
        self.grid = (frames // pt, height // ph, width // pw)
        self.n_tokens = self.grid[0] * self.grid[1] * self.grid[2]
        cos, sin = wan_rope_tables(engine.cfg, *self.grid, table_dtype=torch.bfloat16)
        self.handle = vp()
        _lib.check(_L().ffb200_wan_plan_create(engine.handle, batch, int(self.cfg), frames, height, width, n_text, cos.data_ptr(), sin.data_ptr(),
                                               C.byref(self.handle)), "ffb200_wan_plan_create")
        self._keep: List[torch.Tensor] = []

    @property
    def latent_shape(self) -> Tuple[int, ...]:
        return (self.batch, self.engine.cfg.in_channels, self.frames, self.height, self.width)

    @property
    def workspace_bytes(self) -> int:
        return int(_L().ffb200_wan_plan_workspace_bytes(self.handle))

    def __del__(self):
        try:
            if self.handle:
                _L().ffb200_wan_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class WanRolloutEngine:
    """Owns the packed Wan2.1 weights + the native engine; hands out geometry-specific plans."""

    def __init__(self, model_config, state_dict: Mapping[str, torch.Tensor], device: Optional[torch.device] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("flow_factory_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device(device if device is not None else "cuda")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        torch.cuda.set_device(self.device)
        self.cfg = model_config if isinstance(model_config, WanEngineConfig) else WanEngineConfig.from_model_config(model_config)
        self.weights = WanPackedWeights(self.cfg, state_dict, self.device)
        c = self.cfg
        mc = WanConfigC(c.num_layers, c.num_attention_heads, c.in_channels, c.text_dim, c.freq_dim, c.ffn_dim, *c.patch_size, c.eps)
        self.handle = vp()
        _lib.check(_L().ffb200_wan_engine_create(C.byref(mc), C.byref(self.weights.struct), C.byref(self.handle)), "ffb200_wan_engine_create")
        self._plans: Dict[Tuple, WanPlan] = {}
        self._stream: Optional[torch.cuda.Stream] = None

    def refresh_weights(self, state_dict: Mapping[str, torch.Tensor]) -> None:
        self.weights.pack(state_dict)
        _lib.check(_L().ffb200_wan_engine_set_weights(self.handle, C.byref(self.weights.struct)), "ffb200_wan_engine_set_weights")

    def plan(self, batch: int, frames: int, height: int, width: int, n_text: int, cfg: bool = True) -> WanPlan:
        key = (batch, frames, height, width, n_text, bool(cfg))
        if key not in self._plans:
            self._plans[key] = WanPlan(self, batch, frames, height, width, n_text, cfg)
        return self._plans[key]

    def stream(self) -> torch.cuda.Stream:
        if self._stream is None:
            self._stream = torch.cuda.Stream(self.device)
        return self._stream

    def set_prompts(self, plan: WanPlan, prompt_embeds: torch.Tensor, negative_prompt_embeds: Optional[torch.Tensor] = None) -> None:
        """Forward batch = [negative ; positive] with CFG (the uncond half first, as the step kernel expects)."""
        pe = prompt_embeds.to(device=self.device, dtype=torch.bfloat16)
        if plan.cfg:
            if negative_prompt_embeds is None:
                raise ValueError("a CFG plan needs negative_prompt_embeds")
            pe = torch.cat([negative_prompt_embeds.to(device=self.device, dtype=torch.bfloat16), pe], dim=0)
        pe = pe.contiguous()
        assert tuple(pe.shape) == (plan.batch * (2 if plan.cfg else 1), plan.n_text, self.cfg.text_dim), pe.shape
        plan._keep = [pe]
        st = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(_L().ffb200_wan_set_prompts(plan.handle, pe.data_ptr(), st), "ffb200_wan_set_prompts")

    def transformer_forward(self, plan: WanPlan, latents: torch.Tensor, timestep: float, guidance_scale: float = 1.0) -> torch.Tensor:
        """`timestep` on the 0..1000 scale (the integer UniPC timesteps); -> CFG-combined noise prediction bf16 [B, C, F, H, W]."""
        x = latents.to(device=self.device, dtype=torch.float16).contiguous()
        assert tuple(x.shape) == plan.latent_shape, (tuple(x.shape), plan.latent_shape)
        out = torch.empty_like(x, dtype=torch.bfloat16)
        st = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(_L().ffb200_wan_forward(plan.handle, x.data_ptr(), float(timestep), float(guidance_scale), out.data_ptr(), st), "ffb200_wan_forward")
        return out

    def step(self, plan: WanPlan, latents: torch.Tensor, coef: "_lib.StepCoef", guidance_scale: float, noise: Optional[torch.Tensor] = None,
             next_latents: Optional[torch.Tensor] = None, seed: int = 0) -> Dict[str, torch.Tensor]:
        x = latents.to(device=self.device, dtype=torch.float16).contiguous()
        shp = tuple(x.shape)
        assert shp == plan.latent_shape
        nz = noise.to(device=self.device, dtype=torch.float32).contiguous() if noise is not None else None
        ng = next_latents.to(device=self.device, dtype=torch.float16).contiguous() if next_latents is not None else None
        o_next = torch.empty(shp, dtype=torch.float16, device=self.device)
        o_mean = torch.empty(shp, dtype=torch.float32, device=self.device)
        o_lp = torch.zeros(shp[0], dtype=torch.float32, device=self.device) if coef.compute_log_prob else None
        o_v = torch.empty(shp, dtype=torch.bfloat16, device=self.device)
        flag = torch.zeros(1, dtype=torch.int32, device=self.device)
        a = _lib.StepArgs()
        a.latents = x.data_ptr(); a.coef = coef; a.guidance_scale = float(guidance_scale)
        a.noise = nz.data_ptr() if nz is not None else None
        a.seed = int(seed); a.step_index = 0
        a.next_latents = ng.data_ptr() if ng is not None else None
        a.out_next_latents = o_next.data_ptr(); a.out_mean = o_mean.data_ptr()
        a.out_log_prob = o_lp.data_ptr() if o_lp is not None else None
        a.out_noise_pred = o_v.data_ptr(); a.overflow_flag = flag.data_ptr()
        st = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(_L().ffb200_wan_step(plan.handle, C.byref(a), st), "ffb200_wan_step")
        return dict(next_latents=o_next, next_latents_mean=o_mean, log_prob=o_lp, noise_pred=o_v, overflow=flag)

    def make_coefs(self, scheduler: UniPCMultistepSDEScheduler, num_steps: int, compute_log_prob: bool = True,
                   store_slots: Optional[Sequence[int]] = None, logp_slots: Optional[Sequence[int]] = None):
        """Per-step scalar blocks of a rollout on the scheduler's integer timesteps (wan2_t2v.py:346-358: t, t_next -> sigma = t / 1000)."""
        ts = scheduler.set_timesteps(num_steps)
        sde = set(int(i) for i in scheduler.current_sde_steps.tolist())
        coefs = []
        for i in range(num_steps):
            nl = scheduler.noise_level if (i in sde and not scheduler.is_eval) else 0.0
            tn = ts[i + 1] if i + 1 < num_steps else torch.tensor(0)
            coefs.append(scheduler.step_coef(ts[i], tn, nl, compute_log_prob=compute_log_prob and nl > 0, t_model=float(ts[i]),
                                             store_slot=-1 if store_slots is None else int(store_slots[i]),
                                             logp_slot=-1 if logp_slots is None else int(logp_slots[i])))
        return ts, coefs

    def rollout(self, plan: WanPlan, x0: torch.Tensor, coefs: Sequence["_lib.StepCoef"], guidance_scale: float, n_latent_slots: int,
                store_initial_slot: int, n_logp_slots: int, noise: Optional[torch.Tensor] = None, seed: int = 0,
                use_graph: bool = True) -> Dict[str, torch.Tensor]:
        """All steps on the device without host synchronisation: all_latents fp16 [B, slots, C, F, H, W], log_probs fp32 [B, logp_slots],
        final latents fp16 [B, C, F, H, W]."""
        T = len(coefs)
        shp = plan.latent_shape
        x = x0.to(device=self.device, dtype=torch.float16).contiguous()
        assert tuple(x.shape) == shp
        arr = (_lib.StepCoef * T)(*coefs)
        traj = torch.empty((shp[0], max(n_latent_slots, 1), *shp[1:]), dtype=torch.float16, device=self.device) if n_latent_slots else None
        lps = torch.zeros((shp[0], max(n_logp_slots, 1)), dtype=torch.float32, device=self.device) if n_logp_slots else None
        final = torch.empty(shp, dtype=torch.float16, device=self.device)
        flag = torch.zeros(1, dtype=torch.int32, device=self.device)
        nz = None
        if noise is not None:
            nz = noise.to(device=self.device, dtype=torch.float32).contiguous()
            assert tuple(nz.shape) == (T, *shp)
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
                _lib.check(_L().ffb200_wan_rollout(plan.handle, C.byref(a), s.cuda_stream), "ffb200_wan_rollout")
            cur.wait_stream(s)
            for t in (x, traj, lps, final, flag, nz):
                if t is not None:
                    t.record_stream(s)
        else:
            _lib.check(_L().ffb200_wan_rollout(plan.handle, C.byref(a), cur.cuda_stream), "ffb200_wan_rollout")
        return dict(all_latents=traj, log_probs=lps, final_latents=final, overflow=flag)

    @staticmethod
    def last_launch_count() -> int:
        return int(_lib.lib().ffb200_last_launch_count())

    def __del__(self):
        try:
            self._plans.clear()
            if self.handle:
                _L().ffb200_wan_engine_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


# ---------------------------------------------------------------------------------------------- op-level wrappers (parity tests)
def _st(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def rms_rope_(x: torch.Tensor, width: int, weight: torch.Tensor, eps: float = 1e-6, cos: Optional[torch.Tensor] = None,
              sin: Optional[torch.Tensor] = None, col: int = 0) -> torch.Tensor:
    """In place on columns [col, col + width) of a bf16 [B, S, ld] tensor: RMSNorm across the width (+ RoPE with fp32 [S, 128] tables)."""
    B, S, ld = x.shape
    w = weight.to(device=x.device, dtype=torch.bfloat16).contiguous()
    _lib.check(_L().ffb200_wan_rms_rope(x.data_ptr() + 2 * col, B * S, S, ld, width, w.data_ptr(), eps,
                                        cos.data_ptr() if cos is not None else None, sin.data_ptr() if sin is not None else None, _st(x)),
               "ffb200_wan_rms_rope")
    return x


def layer_norm(x: torch.Tensor, eps: float = 1e-6, scale: Optional[torch.Tensor] = None, shift: Optional[torch.Tensor] = None,
               weight: Optional[torch.Tensor] = None, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x bf16 [B, S, D]; (scale, shift) fp32 [B, D] -> FP32LayerNorm + fp32 modulate; (weight, bias) -> affine FP32LayerNorm."""
    B, S, D = x.shape
    out = torch.empty_like(x)
    if scale is not None:
        sc, sh = scale.float().contiguous(), shift.float().contiguous()
        _lib.check(_L().ffb200_wan_layer_norm(x.data_ptr(), out.data_ptr(), B, S, D, eps, 0, sc.data_ptr(), sh.data_ptr(), D, None, None, _st(x)),
                   "ffb200_wan_layer_norm")
    else:
        wb, bb = weight.to(x.device, torch.bfloat16).contiguous(), bias.to(x.device, torch.bfloat16).contiguous()
        _lib.check(_L().ffb200_wan_layer_norm(x.data_ptr(), out.data_ptr(), B, S, D, eps, 1, None, None, 0, wb.data_ptr(), bb.data_ptr(), _st(x)),
                   "ffb200_wan_layer_norm")
    return out


def gate_residual_(h: torch.Tensor, y: torch.Tensor, gate: torch.Tensor) -> torch.Tensor:
    """In place: h = bf16(float(h) + float(y) * gate[b]); h, y bf16 [B, S, D], gate fp32 [B, D]."""
    B, S, D = h.shape
    g = gate.float().contiguous()
    _lib.check(_L().ffb200_wan_gate_residual(h.data_ptr(), y.data_ptr(), g.data_ptr(), D, B, S, D, _st(h)), "ffb200_wan_gate_residual")
    return h


def patchify(x: torch.Tensor, patch: Tuple[int, int, int] = (1, 2, 2), reps: int = 1) -> torch.Tensor:
    """fp16 [B, C, F, H, W] -> bf16 [reps * B * tokens, C * pt * ph * pw] (rows (b, f', h', w'), columns (c, dt, dh, dw))."""
    B, Cc, Fr, H, Wd = x.shape
    pt, ph, pw = patch
    out = torch.empty(reps * B * (Fr // pt) * (H // ph) * (Wd // pw), Cc * pt * ph * pw, dtype=torch.bfloat16, device=x.device)
    _lib.check(_L().ffb200_wan_patchify(x.contiguous().data_ptr(), B, reps, Cc, Fr, H, Wd, pt, ph, pw, out.data_ptr(), _st(x)), "ffb200_wan_patchify")
    return out


def attention_cross(q: torch.Tensor, kv: torch.Tensor, num_heads: int) -> torch.Tensor:
    """q bf16 [B, Sq, 128 H]; kv bf16 [B, Skv, 2 * 128 H] (k | v) -> bf16 [B, Sq, 128 H]."""
    B, Sq, D = q.shape
    out = torch.empty_like(q)
    _lib.check(_L().ffb200_attention_cross(q.data_ptr(), q.shape[2], kv.data_ptr(), kv.shape[2], D, B, Sq, kv.shape[1], num_heads, out.data_ptr(), _st(q)),
               "ffb200_attention_cross")
    return out
