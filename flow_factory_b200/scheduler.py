"""Host mirror of Flow-Factory's FlowMatchEulerDiscreteSDEScheduler for the rollout path.

Same attribute / method names and semantics as FF/scheduler/flow_match_euler_discrete.py:86-438 (and the pieces of
diffusers' FlowMatchEulerDiscreteScheduler.set_timesteps it relies on, DF/schedulers/scheduling_flow_match_euler_discrete.py:282-384),
but `step()` runs the fused sm_100a kernel (csrc/elementwise.cu: sde_step_kernel) through the C ABI instead of ~25 ATen ops,
and all per-step scalars are produced on the host (no `.item()` device syncs inside the denoising loop).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, fields
from typing import Any, Dict, List, Optional, Sequence, Union

import numpy as np
import torch

from . import _lib
from .rng import randn_tensor

DYNAMICS = {"Flow-SDE": 0, "Dance-SDE": 1, "CPS": 2, "ODE": 3}


@dataclass
class SDESchedulerOutput:
    """FF/scheduler/abc.py:24-40."""
    next_latents: Optional[torch.Tensor] = None
    next_latents_mean: Optional[torch.Tensor] = None
    std_dev_t: Optional[torch.Tensor] = None
    dt: Optional[torch.Tensor] = None
    log_prob: Optional[torch.Tensor] = None
    noise_pred: Optional[torch.Tensor] = None

    def to_dict(self) -> Dict[str, Any]:
        return {f.name: getattr(self, f.name) for f in fields(self)}

    @classmethod
    def from_dict(cls, data: Dict[str, Any]) -> "SDESchedulerOutput":
        names = {f.name for f in fields(cls)}
        return cls(**{k: v for k, v in data.items() if k in names})

    # The reference class derives from diffusers' BaseOutput (an ordered mapping over the non-None fields): the NFT / AWM / CRD trainers read
    # the no-grad forward as `output['noise_pred']` (nft.py:375, awm.py:461, crd.py:699), pipelines as `output[0]`.
    def keys(self):
        return [f.name for f in fields(self) if getattr(self, f.name) is not None]

    def values(self):
        return [getattr(self, k) for k in self.keys()]

    def items(self):
        return [(k, getattr(self, k)) for k in self.keys()]

    def to_tuple(self):
        return tuple(self.values())

    def __getitem__(self, k):
        if isinstance(k, str):
            if k not in self.keys():
                raise KeyError(k)
            return getattr(self, k)
        return self.to_tuple()[k]

    def __iter__(self):
        return iter(self.keys())

    def __len__(self) -> int:
        return len(self.keys())

    def __contains__(self, k) -> bool:
        return k in self.keys()


def calculate_shift(image_seq_len: int, base_seq_len: int = 256, max_seq_len: int = 4096, base_shift: float = 0.5,
                    max_shift: float = 1.15) -> float:
    """FF/scheduler/flow_match_euler_discrete.py:37-47."""
    m = (max_shift - base_shift) / (max_seq_len - base_seq_len)
    b = base_shift - m * base_seq_len
    return image_seq_len * m + b


def _f32(x) -> torch.Tensor:
    return torch.tensor(float(x), dtype=torch.float32)


def make_step_coef(sigma: float, sigma_prev: float, noise_level: float, sigma_max: float, dynamics_type: str,
                   t_model: float = 0.0, compute_log_prob: bool = True, store_slot: int = -1, logp_slot: int = -1
                   ) -> "_lib.StepCoef":
    """The (B,1,1,1) fp32 scalars of scheduler.step (flow_match_euler_discrete.py:322-420), evaluated with the same
    fp32 torch ops in the same order, packed for the kernel."""
    s, sp, eta = _f32(sigma), _f32(sigma_prev), _f32(noise_level)
    dt = sp - s
    c = _lib.StepCoef()
    c.t_model = float(t_model)
    c.sigma, c.sigma_prev, c.dt, c.noise_level = float(s), float(sp), float(dt), float(eta)
    c.dynamics = DYNAMICS[dynamics_type]
    c.compute_log_prob = int(bool(compute_log_prob))
    c.store_slot, c.logp_slot = int(store_slot), int(logp_slot)
    c.two_var, c.log_norm = 1.0, 0.0
    log_sqrt_2pi = torch.log(torch.sqrt(2 * torch.as_tensor(math.pi)))
    if dynamics_type == "ODE":
        c.std_dev_t = 0.0
    elif dynamics_type == "Flow-SDE":
        smax = _f32(sigma_max)
        std = torch.sqrt(s / (1 - torch.where(s == 1.0, smax, s))) * eta
        c.std_dev_t = float(std)
        c.c_x = float(1 + std ** 2 / (2 * s) * dt)
        c.c_v = float(1 + std ** 2 * (1 - s) / (2 * s))
        sv = std * torch.sqrt(-1 * dt)
        c.noise_scale = float(sv)
        if noise_level > 0:
            c.two_var = float(2 * sv ** 2)
            c.log_norm = float(torch.log(sv) + log_sqrt_2pi)
    elif dynamics_type == "Dance-SDE":
        std = eta
        c.std_dev_t = float(std)
        c.c_x = float(0.5 * eta ** 2)
        c.c_v = float(1 - s)
        sv = std * torch.sqrt(-1 * dt)
        c.noise_scale = float(sv)
        if noise_level > 0:
            c.two_var = float(2 * sv ** 2)
            c.log_norm = float(torch.log(sv) + log_sqrt_2pi)
    elif dynamics_type == "CPS":
        std = sp * torch.sin(eta * torch.pi / 2)
        c.std_dev_t = float(std)
        c.c_v = float(1 - s)
        c.cps_a = float(1 - sp)
        c.cps_b = float(torch.sqrt(sp ** 2 - std ** 2))
        c.noise_scale = float(std)
    else:
        raise ValueError(f"unknown dynamics_type {dynamics_type}")
    return c


class FlowMatchEulerDiscreteSDEScheduler:
    """Drop-in for the reference class on the rollout path (same constructor keywords, flow_match...py:90-110)."""

    order = 1

    def __init__(self, noise_level: float = 0.7, sde_steps: Optional[Union[int, list, torch.Tensor]] = None,
                 num_sde_steps: Optional[int] = None, seed: int = 42, dynamics_type: str = "Flow-SDE",
                 num_train_timesteps: int = 1000, shift: float = 1.0, use_dynamic_shifting: bool = False,
                 base_shift: float = 0.5, max_shift: float = 1.15, base_image_seq_len: int = 256,
                 max_image_seq_len: int = 4096, **kwargs):
        assert noise_level >= 0, "Noise level must be non-negative."
        self.noise_level = noise_level
        self._sde_steps = torch.tensor(sde_steps, dtype=torch.int64) if sde_steps is not None else None
        self._num_sde_steps = num_sde_steps
        self.seed = seed
        self.dynamics_type = dynamics_type
        self._is_eval = False
        self.config = dict(num_train_timesteps=num_train_timesteps, shift=shift, use_dynamic_shifting=use_dynamic_shifting,
                           base_shift=base_shift, max_shift=max_shift, base_image_seq_len=base_image_seq_len,
                           max_image_seq_len=max_image_seq_len, **kwargs)
        self.timesteps = torch.zeros(0)
        self.sigmas = torch.zeros(1)

    # ---- mode switches (flow_match...py:112-124)
    @property
    def is_eval(self):
        return self._is_eval

    def eval(self):
        self._is_eval = True

    def train(self, mode: bool = True):
        self._is_eval = not mode

    def rollout(self, mode: bool = True):
        self.train(mode=mode)

    def set_seed(self, seed: int):
        self.seed = seed

    # ---- schedule (set_scheduler_timesteps, flow_match...py:49-77 ; diffusers set_timesteps :282-384)
    def set_timesteps(self, num_inference_steps: int, seq_len: Optional[int] = None, device=None,
                      sigmas: Optional[Sequence[float]] = None, mu: Optional[float] = None) -> torch.Tensor:
        sig = np.linspace(1.0, 1 / num_inference_steps, num_inference_steps) if sigmas is None else np.array(sigmas)
        sig = sig.astype(np.float32)
        if self.config["use_dynamic_shifting"]:
            if mu is None:
                assert seq_len is not None, "`seq_len` must be provided if `mu` is not given."
                mu = calculate_shift(seq_len, self.config["base_image_seq_len"], self.config["max_image_seq_len"],
                                     self.config["base_shift"], self.config["max_shift"])
            sig = (math.exp(mu) / (math.exp(mu) + (1 / sig - 1) ** 1.0)).astype(np.float32)   # time_shift, exponential
        else:
            shift = self.config["shift"]
            sig = shift * sig / (1 + (shift - 1) * sig)
        s = torch.from_numpy(np.asarray(sig)).to(dtype=torch.float32)
        self.timesteps = s * self.config["num_train_timesteps"]
        self.sigmas = torch.cat([s, torch.zeros(1)])
        self.num_inference_steps = num_inference_steps
        return self.timesteps

    # ---- SDE step selection (flow_match...py:126-198)
    @property
    def sde_steps(self) -> torch.Tensor:
        if self._sde_steps is not None:
            return self._sde_steps
        return torch.arange(0, len(self.timesteps) - 1, dtype=torch.int64)

    @property
    def num_sde_steps(self) -> int:
        return self._num_sde_steps if self._num_sde_steps is not None else len(self.sde_steps)

    @property
    def current_sde_steps(self) -> torch.Tensor:
        if self.num_sde_steps >= len(self.sde_steps):
            return self.sde_steps
        g = torch.Generator().manual_seed(self.seed)
        sel = torch.randperm(len(self.sde_steps), generator=g)[: self.num_sde_steps]
        return self.sde_steps[sel]

    @property
    def train_timesteps(self) -> torch.Tensor:
        return self.current_sde_steps

    def get_train_timesteps(self) -> torch.Tensor:
        return self.timesteps[self.train_timesteps]

    def get_train_sigmas(self) -> torch.Tensor:
        return self.sigmas[self.train_timesteps]

    def get_noise_levels(self) -> torch.Tensor:
        nl = torch.zeros_like(self.timesteps, dtype=torch.float32)
        nl[self.current_sde_steps] = self.noise_level
        return nl

    def index_for_timestep(self, timestep) -> int:
        t = float(timestep)
        idx = (self.timesteps == torch.tensor(t, dtype=self.timesteps.dtype)).nonzero()
        if len(idx) == 0:
            raise ValueError(f"timestep {t} not in schedule")
        return int(idx[1 if len(idx) > 1 else 0])

    def get_noise_level_for_timestep(self, timestep) -> float:
        return self.noise_level if self.index_for_timestep(timestep) in self.current_sde_steps.tolist() else 0.0

    def get_noise_level_for_sigma(self, sigma) -> float:
        idx = (self.sigmas == torch.tensor(float(sigma), dtype=torch.float32)).nonzero()
        if len(idx) == 0:
            raise ValueError(f"Sigmas {sigma} not found in scheduler sigmas.")
        return self.noise_level if int(idx[0]) in self.current_sde_steps.tolist() else 0.0

    def step_coef(self, timestep, timestep_next, noise_level: Optional[float], dynamics_type: Optional[str] = None,
                  sigma_max: Optional[float] = None, compute_log_prob: bool = True, t_model: float = 0.0,
                  store_slot: int = -1, logp_slot: int = -1):
        """Host scalars of one step; `timestep`/`timestep_next` in [0, 1000] as the adapters pass them.  With `timestep_next=None` the
        reference looks the step up in its tables and uses `sigmas[i]`, `sigmas[i + 1]` themselves (flow_match...py:262-299), which can
        differ from `timesteps / 1000` in the last bit - mirrored."""
        dyn = dynamics_type or self.dynamics_type
        if timestep_next is None:
            i = self.index_for_timestep(timestep)
            sigma, sigma_prev = float(self.sigmas[i]), float(self.sigmas[i + 1])
        else:
            sigma = (torch.as_tensor(timestep, dtype=torch.float32) / 1000).item()
            sigma_prev = (torch.as_tensor(timestep_next, dtype=torch.float32) / 1000).item()
        if self.is_eval or dyn == "ODE":
            noise_level = 0.0
        elif noise_level is None:
            noise_level = self.get_noise_level_for_sigma(sigma)
        smax = sigma_max if sigma_max is not None else (float(self.sigmas[1]) if len(self.sigmas) > 1 else 1.0)
        return make_step_coef(sigma, sigma_prev, float(noise_level), smax, dyn, t_model=t_model,
                              compute_log_prob=compute_log_prob, store_slot=store_slot, logp_slot=logp_slot)

    # ---- the step itself (flow_match...py:243-438), fused kernel
    def step(self, noise_pred: torch.Tensor, timestep, latents: torch.Tensor, next_latents: Optional[torch.Tensor] = None,
             timestep_next=None, generator=None, noise_level=None, compute_log_prob: bool = True, return_dict: bool = True,
             return_kwargs: List[str] = ["next_latents", "next_latents_mean", "std_dev_t", "dt", "log_prob", "noise_pred"],
             dynamics_type: Optional[str] = None, sigma_max: Optional[float] = None, noise: Optional[torch.Tensor] = None,
             seed: Optional[int] = None, step_index: int = 0):
        """`noise` / `seed` / `step_index` are extensions: caller-provided noise, or (seed given, noise not) the in-kernel Philox stream
        keyed by (seed, step_index).  With neither, the noise is drawn exactly as the reference draws it."""
        if not latents.is_cuda:
            raise RuntimeError("flow_factory_b200 scheduler.step needs CUDA tensors (no CPU fallback)")
        c = self.step_coef(timestep, timestep_next, noise_level, dynamics_type, sigma_max, compute_log_prob)
        dyn = dynamics_type or self.dynamics_type
        L = _lib.lib()
        B, Cc, H, W = latents.shape
        in_dtype = latents.dtype
        x16 = latents.contiguous()                 # storage dtype as given
        v16 = noise_pred.to(torch.bfloat16).contiguous()
        # the reference rounds fresh next_latents through the INPUT latents dtype (flow_match...py:309, 359-362) - the latents' storage dtype
        codes = {torch.float16: 0, torch.bfloat16: 1, torch.float32: 2}
        if latents.dtype not in codes:
            raise NotImplementedError(f"scheduler.step: latents are {latents.dtype}; supported storage dtypes: fp16, bf16, fp32")
        if latents.dim() != 4:
            raise NotImplementedError("scheduler.step mirror: (B, C, H, W) latents (the packed / video adapters call the engine step directly)")
        if noise is None and seed is None and next_latents is None and dyn != "ODE":
            # as the reference (flow_match...py:350-357): a fresh fp32 draw from `generator` (one, a per-sample list, or None = the device
            # RNG stream) on EVERY call - never a fixed in-kernel seed, which would correlate the exploration noise across steps
            noise = randn_tensor(tuple(noise_pred.shape), generator=generator, device=latents.device, dtype=torch.float32)
        nz = noise.to(torch.float32).contiguous() if noise is not None else None
        ng = next_latents.to(latents.dtype).contiguous() if next_latents is not None else None
        out_next = torch.empty_like(x16)
        out_mean = torch.empty(latents.shape, dtype=torch.float32, device=latents.device)
        out_lp = torch.zeros(B, dtype=torch.float32, device=latents.device) if compute_log_prob else None
        flag = torch.zeros(1, dtype=torch.int32, device=latents.device)
        st = torch.cuda.current_stream(latents.device).cuda_stream
        _lib.check(L.ffb200_sde_step_ex(v16.data_ptr(), x16.data_ptr(), B, Cc, H, W, c,
                                     nz.data_ptr() if nz is not None else None, int(seed or 0), step_index,
                                     ng.data_ptr() if ng is not None else None, out_next.data_ptr(), out_mean.data_ptr(),
                                     out_lp.data_ptr() if out_lp is not None else None, flag.data_ptr(), codes[latents.dtype], st), "ffb200_sde_step_ex")
        if next_latents is not None:
            nxt = next_latents.float()
        elif dyn == "ODE":
            nxt = out_mean                        # ODE: next_latents IS the mean, no storage round trip (flow_match...py:333-334)
        else:
            nxt = out_next.float()
        d = dict(next_latents=nxt,
                 next_latents_mean=out_mean,
                 std_dev_t=torch.full((B, 1, 1, 1), c.std_dev_t, dtype=torch.float32, device=latents.device),
                 dt=torch.full((B, 1, 1, 1), c.dt, dtype=torch.float32, device=latents.device),
                 log_prob=out_lp, noise_pred=noise_pred.float())
        if not return_dict:
            return (d["next_latents"], d["next_latents_mean"], d["noise_pred"], d["log_prob"], d["std_dev_t"], d["dt"])
        return SDESchedulerOutput.from_dict({k: d[k] for k in return_kwargs if k in d})


def set_scheduler_timesteps(scheduler: FlowMatchEulerDiscreteSDEScheduler, num_inference_steps: int,
                            seq_len: Optional[int] = None, sigmas=None, device=None, mu: Optional[float] = None):
    """FF/scheduler/flow_match_euler_discrete.py:49-77."""
    return scheduler.set_timesteps(num_inference_steps, seq_len=seq_len, device=device, sigmas=sigmas, mu=mu)


class UniPCMultistepSDEScheduler(FlowMatchEulerDiscreteSDEScheduler):
    """Host mirror of FF/scheduler/unipc_multistep.py (Wan2.x): the rollout uses the SAME Euler / SDE step arithmetic as
    FlowMatchEulerDiscreteSDEScheduler.step (unipc_multistep.py:290-421); only the schedule differs - diffusers'
    UniPCMultistepScheduler.set_timesteps with `use_flow_sigmas` (DF/schedulers/scheduling_unipc_multistep.py:428-466): INTEGER
    timesteps (`sigmas * 1000` truncated to int64) next to fp32 sigmas, so the step's sigma = int_timestep / 1000 differs from
    `sigmas[i]` while `sigma_max` stays `sigmas[1]`."""

    def __init__(self, noise_level: float = 0.7, sde_steps=None, num_sde_steps=None, seed: int = 42, dynamics_type: str = "Flow-SDE",
                 num_train_timesteps: int = 1000, flow_shift: float = 3.0, **kwargs):
        kwargs.pop("shift", None)
        if kwargs.get("use_dynamic_shifting") or kwargs.get("shift_terminal"):
            raise NotImplementedError("UniPC mirror: only the static flow_shift schedule of the Wan configs is implemented")
        super().__init__(noise_level=noise_level, sde_steps=sde_steps, num_sde_steps=num_sde_steps, seed=seed, dynamics_type=dynamics_type,
                         num_train_timesteps=num_train_timesteps, shift=flow_shift, **kwargs)
        self.config["flow_shift"] = flow_shift
        self.order = 1

    def set_timesteps(self, num_inference_steps: int, seq_len=None, device=None, sigmas=None, mu=None) -> torch.Tensor:
        n = self.config["num_train_timesteps"]
        sig = np.linspace(1, 1 / n, num_inference_steps + 1)[:-1] if sigmas is None else np.array(sigmas, dtype=np.float64)
        shift = self.config["flow_shift"]
        sig = shift * sig / (1 + (shift - 1) * sig)
        if np.fabs(sig[0] - 1) < 1e-6:
            sig[0] -= 1e-6
        timesteps = (sig * n).copy()
        self.sigmas = torch.from_numpy(np.concatenate([sig, [0]]).astype(np.float32))
        self.timesteps = torch.from_numpy(timesteps).to(dtype=torch.int64)
        self.num_inference_steps = len(timesteps)
        return self.timesteps

    def get_noise_levels(self) -> torch.Tensor:
        nl = torch.zeros(len(self.timesteps), dtype=torch.float32)
        nl[self.current_sde_steps] = self.noise_level
        return nl

    def step_coef(self, timestep, timestep_next, noise_level, *args, **kwargs):
        if timestep_next is None:
            # the reference's step() takes an INTEGER timestep without timestep_next for a step index (unipc_multistep.py:246-256) - with the
            # integer UniPC schedule that path cannot address a timestep; Wan2_T2V_Adapter always passes t_next (wan2_t2v.py:349-355)
            raise ValueError("UniPC schedule: pass timestep_next (integer timesteps without it are read as step indices by the reference)")
        return super().step_coef(timestep, timestep_next, noise_level, *args, **kwargs)
