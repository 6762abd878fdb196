"""TEST / BENCH INFRASTRUCTURE - never imported by the product (flow_factory_b200/).

Drives the reference's OWN code of the hot path on the host CPU, from `oracle/_ref/` (unmodified copies made by tools/make_oracle_ref.py; see
its header): diffusers' `SD3Transformer2DModel.forward` and Flow-Factory's `FlowMatchEulerDiscreteSDEScheduler.step`, called the way
`SD3_5Adapter.inference` / `.forward` call them (FF/models/stable_diffusion/sd3_5.py:273-304, 392-445): CFG batch `cat([latents] * 2)`,
timestep cast to the latents dtype, `uncond + g * (text - uncond)`, `scheduler.step(..., timestep_next, noise_level, compute_log_prob)`,
all under `torch.autocast('cpu', bf16)` with bf16 weights (trainers/abc.py:72-76).  The adapter object itself needs accelerate / a pipeline
and is not constructed - the two calls above ARE its hot path.

Used by `bench.py --impl reference` and by the `cpu_baseline` leg of the default bench line (`kind: "reference"`).
"""
from __future__ import annotations

import os
import sys
import time
from typing import Dict, List, Optional, Tuple

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_DIR, "diffusers")) and os.path.isdir(os.path.join(REF_DIR, "flow_factory"))


def load():
    """Imports the reference classes from oracle/_ref (raises if the recipe has not been run)."""
    if not available():
        raise RuntimeError("oracle/_ref is missing: run `python tools/make_oracle_ref.py` in the build container (needs /root/reference)")
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    from diffusers.models.transformers.transformer_sd3 import SD3Transformer2DModel
    from flow_factory.scheduler import FlowMatchEulerDiscreteSDEScheduler, set_scheduler_timesteps
    import diffusers
    import flow_factory
    assert os.path.abspath(diffusers.__file__).startswith(REF_DIR) and os.path.abspath(flow_factory.__file__).startswith(REF_DIR), \
        "another diffusers / flow_factory shadows oracle/_ref"
    return SD3Transformer2DModel, FlowMatchEulerDiscreteSDEScheduler, set_scheduler_timesteps


# ------------------------------------------------------------------------------------------------ host topology
def _parse_cpulist(s: str) -> List[int]:
    out: List[int] = []
    for part in s.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            out.extend(range(int(a), int(b) + 1))
        else:
            out.append(int(part))
    return out


def pick_cores() -> Dict:
    """One hardware thread per physical core of ONE NUMA node, restricted to this process's affinity mask: a stable, stated core set
    (all logical CPUs of a two-socket box oversubscribe the memory system and swing the timing by an order of magnitude - round 1)."""
    allowed = sorted(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else list(range(os.cpu_count() or 1))
    node_cpus = allowed
    try:
        nodes = sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit())
        best: List[int] = []
        for nd in nodes:
            with open(f"/sys/devices/system/node/{nd}/cpulist") as f:
                cs = [c for c in _parse_cpulist(f.read()) if c in allowed]
            if len(cs) > len(best):
                best = cs
        if best:
            node_cpus = best
    except OSError:
        pass
    seen, cores = set(), []
    for c in node_cpus:
        try:
            with open(f"/sys/devices/system/cpu/cpu{c}/topology/thread_siblings_list") as f:
                sib = tuple(_parse_cpulist(f.read()))
        except OSError:
            sib = (c,)
        key = min(sib)
        if key not in seen:
            seen.add(key)
            cores.append(c)
    model = "unknown"
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    model = line.split(":", 1)[1].strip()
                    break
    except OSError:
        pass
    return {"cpus": cores, "n": len(cores), "logical_allowed": len(allowed), "model": model}


def pin(cores: Dict) -> None:
    if hasattr(os, "sched_setaffinity") and cores["cpus"]:
        try:
            os.sched_setaffinity(0, set(cores["cpus"]))
        except OSError:
            pass
    torch.set_num_threads(max(1, cores["n"]))


# ------------------------------------------------------------------------------------------------ the reference objects
def build_model(cfg, weights: Optional[Dict[str, torch.Tensor]] = None, num_layers: Optional[int] = None, seed: int = 0):
    """A REAL SD3Transformer2DModel in bf16.  With `weights` (a complete state dict, e.g. sd3_oracle.make_weights) the module is built on
    the meta device and the tensors are assigned (no 2.2 B-element CPU random init); without, seeded default init.
    `num_layers` keeps only the first blocks (bounded timing samples; the kept blocks are untouched reference modules)."""
    SD3, _, _ = load()
    kw = cfg.ref_kwargs()
    if weights is not None:
        with torch.device("meta"):
            m = SD3(**kw)
        m.load_state_dict({k: v.to(torch.bfloat16) if v.is_floating_point() else v for k, v in weights.items()}, strict=True, assign=True)
        left = [n for n, t in list(m.named_parameters()) + list(m.named_buffers()) if t.is_meta]
        assert not left, f"state dict does not cover {left[:3]}"
    else:
        torch.manual_seed(seed)
        m = SD3(**kw)
    m = m.to(torch.bfloat16).eval()
    if num_layers is not None and num_layers < len(m.transformer_blocks):
        m.transformer_blocks = m.transformer_blocks[:num_layers]
    return m


def truncated(model, num_layers: int):
    """A shallow copy of `model` running only its first `num_layers` blocks (shares every module)."""
    import copy
    t = copy.copy(model)
    t._modules = dict(model._modules)
    t.transformer_blocks = model.transformer_blocks[:num_layers]
    return t


def make_scheduler(T: int, seq_len: int, noise_level: float = 0.7, shift: float = 3.0, num_sde_steps: Optional[int] = None, seed: int = 42):
    _, Sched, set_ts = load()
    s = Sched(noise_level=noise_level, shift=shift, num_sde_steps=num_sde_steps, seed=seed, dynamics_type="Flow-SDE")
    ts = set_ts(s, T, seq_len=seq_len, device="cpu")
    return s, ts


@torch.no_grad()
def reference_step(model, sched, timesteps, i: int, latents, prompt_embeds, pooled, neg_embeds, neg_pooled, guidance: float,
                   compute_log_prob: bool = True):
    """SD3_5Adapter.forward (sd3_5.py:392-445) for step i of the loop at 273-304."""
    t = timesteps[i]
    t_next = timesteps[i + 1] if i + 1 < len(timesteps) else torch.tensor(0.0)
    do_cfg = guidance > 1.0 and neg_embeds is not None
    # the batch assembly sits outside the autocast region: CPU autocast cannot promote the fp16 latents in torch.cat (a CPU-only quirk;
    # under CUDA autocast, where the reference runs, the same calls are fine) - values are identical either way
    if do_cfg:
        x = torch.cat([latents] * 2)
        pe, pp = torch.cat([neg_embeds, prompt_embeds]), torch.cat([neg_pooled, pooled])
    else:
        x, pe, pp = latents, prompt_embeds, pooled
    timestep = t.expand(x.shape[0]).to(latents.dtype)
    with torch.autocast("cpu", dtype=torch.bfloat16):
        v = model(hidden_states=x, timestep=timestep, encoder_hidden_states=pe, pooled_projections=pp, return_dict=False)[0]
    if do_cfg:
        vu, vt = v.chunk(2)
        v = vu + guidance * (vt - vu)
    out = sched.step(noise_pred=v, timestep=t, latents=latents, timestep_next=t_next, next_latents=None,
                     compute_log_prob=compute_log_prob, return_dict=True,
                     return_kwargs=["next_latents", "log_prob", "noise_pred"], noise_level=sched.get_noise_level_for_timestep(t))
    return out


def synthetic_inputs(cfg, B: int, res: int, n_text: int, seed: int = 1):
    lat = res // 8
    g = torch.Generator().manual_seed(seed)
    bf = lambda *s: torch.randn(*s, generator=g).bfloat16()
    return dict(prompt_embeds=bf(B, n_text, cfg.joint_attention_dim), pooled=bf(B, cfg.pooled_projection_dim),
                neg_embeds=bf(B, n_text, cfg.joint_attention_dim), neg_pooled=bf(B, cfg.pooled_projection_dim),
                x0=torch.randn(B, cfg.in_channels, lat, lat, generator=g).half())


def time_c1_full(cfg, model) -> Tuple[float, float]:
    """BASELINE config C1 in full: 256^2, 4-step Euler/SDE rollout, B=1, guidance 1.0 (no CFG), random-init weights.
    Returns (seconds, latents/s)."""
    res, T, n_text = 256, 4, 333
    ni = (res // 16) ** 2
    sched, ts = make_scheduler(T, ni)
    inp = synthetic_inputs(cfg, 1, res, n_text)
    x = inp["x0"]
    t0 = time.perf_counter()
    for i in range(T):
        out = reference_step(model, sched, ts, i, x, inp["prompt_embeds"], inp["pooled"], None, None, 1.0)
        x = out.next_latents.to(torch.float16)
    dt = time.perf_counter() - t0
    return dt, 1.0 / dt


def time_c2_step(cfg, model, guidance: float, n_text: int = 333, res: int = 1024, T: int = 30, step_index: int = 3) -> float:
    """Seconds of ONE step of BASELINE config C2 at B=1 through `model` (which may be layer-truncated for a bounded sample)."""
    ni = (res // 16) ** 2
    sched, ts = make_scheduler(T, ni)
    inp = synthetic_inputs(cfg, 1, res, n_text)
    t0 = time.perf_counter()
    reference_step(model, sched, ts, step_index, inp["x0"], inp["prompt_embeds"], inp["pooled"], inp["neg_embeds"], inp["neg_pooled"], guidance)
    return time.perf_counter() - t0


def block_flops(cfg, ni: int, nt: int, n_blocks: int) -> float:
    """FLOPs of the first `n_blocks` blocks of one forward (same model as bench.py's flops_per_forward, truncated)."""
    D = cfg.inner_dim
    S = ni + nt
    total = 0.0
    for i in range(n_blocks):
        last = i == cfg.num_layers - 1
        dual = i in cfg.dual_attention_layers
        total += 24 * ni * D * D + (6 if last else 24) * nt * D * D + 4 * S * S * D
        if dual:
            total += 8 * ni * D * D + 4 * ni * ni * D
    return total
