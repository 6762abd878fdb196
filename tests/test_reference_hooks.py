"""The drop-in boundary checked against the REAL reference, imported read-only in the build container:
* the reference's own plug points resolve this repo's classes - `get_model_adapter_class(<python path>)` (FF/models/registry.py:73-79, what
  `model.model_type: "<path>"` in the YAML goes through) and `register_scheduler` / `get_sde_scheduler_class` (FF/scheduler/registry.py);
* "parameter names are the ABI" (`filter_kwargs`, FF/utils/base.py:38-63): every keyword the reference adapters' `inference()` /
  `forward()` accept is accepted by the B200 adapters, so whatever a trainer passes through `filter_kwargs` arrives.
Runs only where /root/reference exists (the build container; CPU suite) - the GPU box never reads the reference."""
import json
import os
import subprocess
import sys

import pytest

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src", "flow_factory")), reason="reference tree not present (GPU box)")

_SCRIPT = r'''
import inspect, json, sys
sys.path[:0] = [sys.argv[1] + "/tests/golden/ref_stubs", "/root/reference/diffusers/src", "/root/reference/src", sys.argv[1]]
from flow_factory.models.registry import get_model_adapter_class
from flow_factory.scheduler.registry import register_scheduler, get_sde_scheduler_class
from flow_factory.utils.base import filter_kwargs
out = {}
pairs = {"sd3_5": ("flow_factory.models.stable_diffusion.sd3_5.SD3_5Adapter", "flow_factory_b200.adapter.B200SD3_5Adapter"),
         "flux1": ("flow_factory.models.flux.flux1.Flux1Adapter", "flow_factory_b200.flux_adapter.B200Flux1Adapter"),
         "qwen": ("flow_factory.models.qwen_image.qwen_image.QwenImageAdapter", "flow_factory_b200.qwen_adapter.B200QwenImageAdapter"),
         "wan": ("flow_factory.models.wan.wan2_t2v.Wan2_T2V_Adapter", "flow_factory_b200.wan_adapter.B200Wan21Adapter")}
for name, (ref_path, my_path) in pairs.items():
    mine = get_model_adapter_class(my_path)                       # the registry's importlib fall-through
    ref = get_model_adapter_class(ref_path)
    rec = {"resolved": mine.__module__ + "." + mine.__name__}
    for fn in ("inference", "forward"):
        rp = [p for p in inspect.signature(getattr(ref, fn)).parameters if p != "self"]
        mp = [p for p in inspect.signature(getattr(mine, fn)).parameters if p != "self"]
        rec[fn + "_missing"] = [p for p in rp if p not in mp]
        kw = {p: 0 for p in rp}
        rec[fn + "_filtered"] = sorted(filter_kwargs(getattr(mine, fn), **kw)) == sorted(p for p in rp if p in mp)
    out[name] = rec
import dataclasses
samples = {"sd3_5": ("flow_factory.models.stable_diffusion.sd3_5", "SD3_5Sample", "SD3_5Sample"),
           "flux1": ("flow_factory.models.flux.flux1", "Flux1Sample", "Flux1Sample"),
           "qwen": ("flow_factory.models.qwen_image.qwen_image", "QwenImageSample", "QwenImageSample"),
           "wan": ("flow_factory.models.wan.wan2_t2v", "WanT2VSample", "WanT2VSample")}
import importlib
mine_mod = importlib.import_module("flow_factory_b200.samples")
for name, (mod, ref_cls, my_cls) in samples.items():
    rf = {f.name for f in dataclasses.fields(getattr(importlib.import_module(mod), ref_cls))}
    mf = {f.name for f in dataclasses.fields(getattr(mine_mod, my_cls))}
    out[name]["sample_fields_missing"] = sorted(rf ^ mf)          # symmetric difference: identical field sets
    R, M = getattr(importlib.import_module(mod), ref_cls), getattr(mine_mod, my_cls)
    out[name]["shared_fields"] = [sorted(R.shared_fields()), sorted(M.shared_fields())]
    # behaviour on identical data: unique ids, dict views, attribute fall-through, collate - also the REFERENCE's stack() fed with B200 records
    import torch
    def mk(cls, i, prompt):
        kw = dict(timesteps=torch.arange(4.0), all_latents=torch.full((2, 3, 4), float(i)), log_probs=torch.tensor([0.5 * i]),
                  latent_index_map=torch.tensor([0, -1, 1, -1, -1]), log_prob_index_map=torch.tensor([-1, 0, -1, -1]), height=64, width=32,
                  prompt=prompt, prompt_ids=torch.tensor([1, 2, i]), prompt_embeds=torch.ones(5, 2) * i,
                  extra_kwargs={"final_latents": torch.full((3, 4), float(i)), "callback_index_map": None, "tag": {"a": torch.tensor([float(i)])}})
        if "img_ids" in {f.name for f in dataclasses.fields(cls)}:
            kw["img_ids"] = torch.arange(6.0).reshape(2, 3)
        if "img_shapes" in {f.name for f in dataclasses.fields(cls)}:
            kw["img_shapes"] = [(1, 4, 2)]
        return cls(**kw)
    def norm(v):
        if isinstance(v, torch.Tensor): return ["T", list(v.shape), v.flatten().tolist()]
        if isinstance(v, dict): return {k: norm(x) for k, x in v.items()}
        if isinstance(v, (list, tuple)): return [norm(x) for x in v]
        return v
    rs = [mk(R, 1, "a cat"), mk(R, 2, "a dog")]
    ms = [mk(M, 1, "a cat"), mk(M, 2, "a dog")]
    beh = {}
    beh["unique_id"] = [[int(x.unique_id) for x in rs], [int(x.unique_id) for x in ms]]
    ids_only_r, ids_only_m = R(prompt_ids=torch.tensor([7, 8]), negative_prompt="n"), M(prompt_ids=torch.tensor([7, 8]), negative_prompt="n")
    beh["unique_id_ids"] = [int(ids_only_r.unique_id), int(ids_only_m.unique_id)]
    beh["to_dict_keys"] = [list(rs[0].to_dict().keys()), list(ms[0].to_dict().keys())]
    beh["getattr_extra"] = [norm(rs[0].final_latents), norm(ms[0].final_latents), norm(rs[0]["tag"]), norm(ms[0]["tag"])]
    beh["stack"] = [norm(R.stack(rs)), norm(M.stack(ms))]
    from flow_factory.samples import BaseSample as RefBase
    beh["ref_stack_on_mine"] = norm(RefBase.stack(ms)) == beh["stack"][0]
    rt_r, rt_m = R.from_dict(rs[0].to_dict()), M.from_dict(ms[0].to_dict())
    beh["from_dict"] = [norm(rt_r.to_dict()), norm(rt_m.to_dict())]
    ms[0].prompt = "changed"; rs[0].prompt = "changed"
    beh["reset_on_set"] = [int(rs[0].unique_id), int(ms[0].unique_id)]
    out[name]["behaviour"] = beh
# ---- trajectory bookkeeping: the reference's two collectors driven as in sd3_5.py:266-304 vs plan_slots, on randomised index specs
import random, torch
from flow_factory.utils.trajectory_collector import create_trajectory_collector as ref_collector, compute_trajectory_indices as ref_cti
from flow_factory_b200.trajectory import plan_slots, compute_trajectory_indices as my_cti
rng = random.Random(0)
mism = []
for case in range(300):
    T = rng.choice([1, 2, 4, 7, 10, 30])
    kind = rng.choice(["all", "none", "list", "neg", "dups", "cti"])
    if kind == "all": idx = "all"
    elif kind == "none": idx = None
    elif kind == "list": idx = sorted(rng.sample(range(T + 1), rng.randint(1, T + 1)))
    elif kind == "neg": idx = [rng.randint(-(T + 1), T) for _ in range(rng.randint(1, 5))]
    elif kind == "dups": idx = [rng.randint(0, T) for _ in range(rng.randint(1, 8))]
    else:
        tr = sorted(rng.sample(range(T), rng.randint(1, T)))
        inc = rng.random() < 0.5
        idx = ref_cti(tr, T, include_initial=inc)
        if idx != my_cti(tr, T, include_initial=inc): mism.append(("cti", T, tr, inc))
    has_lp = [rng.random() < 0.5 for _ in range(T)]
    try:
        lc, pc = ref_collector(idx, T), ref_collector(idx, T)
    except Exception as e:
        try:
            plan_slots(idx, T, has_lp); mism.append(("ref raised, mine did not", T, idx, repr(e)))
        except Exception:
            pass
        continue
    lc.collect(torch.zeros(1), step_idx=0)
    for i in range(T):
        lc.collect(torch.zeros(1), i + 1)
        if has_lp[i]: pc.collect(torch.zeros(1), i)
    lat_slot, lp_slot, lat_map, lp_map = plan_slots(idx, T, has_lp)
    rl, rp = lc.get_result(), pc.get_result()
    n_lat, n_lp = sum(1 for x in lat_slot if x >= 0), sum(1 for x in lp_slot if x >= 0)
    ok = (0 if rl is None else len(rl)) == n_lat and (0 if rp is None else len(rp)) == n_lp
    for mine, ref in ((lat_map, lc.get_index_map()), (lp_map, pc.get_index_map())):
        ok = ok and ((mine is None) == (ref is None)) and (mine is None or torch.equal(mine, ref))
    if not ok: mism.append((kind, T, idx, has_lp))
from flow_factory.utils.trajectory_collector import create_callback_collector as ref_cb
from flow_factory_b200.trajectory import create_callback_collector as my_cb
class _O:
    def __init__(self, i): self.next_latents_mean = torch.full((2, 3), float(i)); self.log_prob = None; self.tag = "s%d" % i
for case in range(120):
    T = rng.choice([1, 3, 6, 10])
    idx = rng.choice(["all", None, sorted(rng.sample(range(T + 1), rng.randint(1, T + 1))), [rng.randint(-(T + 1), T) for _ in range(3)]])
    keys = rng.choice([[], ["next_latents_mean"], ["next_latents_mean", "noise_level", "tag", "log_prob", "missing"]])
    a, b_ = ref_cb(idx, T), my_cb(idx, T)
    for i in range(T):
        cap = {"noise_level": 0.7 if i % 2 else None}
        a.collect_step(i, _O(i), keys, cap); b_.collect_step(i, _O(i), keys, cap)
    ra, rb = a.get_result(), b_.get_result()
    ok = list(ra.keys()) == list(rb.keys()) and len(a) == len(b_) and a.collected_indices == b_.collected_indices and a.is_disabled == b_.is_disabled
    for k in ra:
        ok = ok and (torch.equal(ra[k], rb[k]) if isinstance(ra[k], torch.Tensor) else ra[k] == rb[k])
    ma, mb = a.get_index_map(), b_.get_index_map()
    ok = ok and ((ma is None) == (mb is None)) and (ma is None or torch.equal(ma, mb))
    if not ok: mism.append(("callback", T, idx, keys))
out["trajectory_mismatches"] = [repr(m) for m in mism[:5]]
# ---- scheduler mirrors: schedule, SDE-step selection and per-step scalars vs the reference on randomised settings
from flow_factory.scheduler import FlowMatchEulerDiscreteSDEScheduler as RefFM, UniPCMultistepSDEScheduler as RefUP, set_scheduler_timesteps as ref_set
from flow_factory_b200.scheduler import FlowMatchEulerDiscreteSDEScheduler as MyFM, UniPCMultistepSDEScheduler as MyUP
smis = []
x = torch.randn(2, 4, 4, 4, generator=torch.Generator().manual_seed(0)).half()
v = torch.randn(2, 4, 4, 4, generator=torch.Generator().manual_seed(1)).bfloat16()
for case in range(60):
    T = rng.choice([4, 10, 28, 30])
    dyn = rng.choice(["Flow-SDE", "Dance-SDE", "CPS", "ODE"])
    nsde = rng.choice([None, 1, 2, 3])
    sde = rng.choice([None, sorted(rng.sample(range(T), rng.randint(1, T)))])
    seed = rng.randint(0, 1000)
    nl = rng.choice([0.3, 0.7, 1.0])
    if rng.random() < 0.5:
        dynshift = rng.random() < 0.5
        kw = dict(use_dynamic_shifting=True) if dynshift else dict(shift=rng.choice([1.0, 3.0, 5.0]))
        r = RefFM(noise_level=nl, sde_steps=sde, num_sde_steps=nsde, seed=seed, dynamics_type=dyn, **kw)
        m = MyFM(noise_level=nl, sde_steps=sde, num_sde_steps=nsde, seed=seed, dynamics_type=dyn, **kw)
        seq = rng.choice([256, 1024, 4096])
        rt = ref_set(r, T, seq_len=seq, device="cpu"); mt = m.set_timesteps(T, seq_len=seq)
    else:
        fs = rng.choice([3.0, 5.0])
        r = RefUP(noise_level=nl, sde_steps=sde, num_sde_steps=nsde, seed=seed, dynamics_type=dyn, prediction_type="flow_prediction",
                  use_flow_sigmas=True, flow_shift=fs, num_train_timesteps=1000)
        m = MyUP(noise_level=nl, sde_steps=sde, num_sde_steps=nsde, seed=seed, dynamics_type=dyn, flow_shift=fs)
        r.set_timesteps(T, device="cpu"); rt = r.timesteps; mt = m.set_timesteps(T)
    r.rollout(); m.rollout()
    ok = torch.equal(rt.cpu(), mt) and torch.equal(r.sigmas.cpu(), m.sigmas)
    ok = ok and torch.equal(r.sde_steps, m.sde_steps) and r.num_sde_steps == m.num_sde_steps and torch.equal(r.current_sde_steps, m.current_sde_steps)
    ok = ok and torch.equal(r.get_noise_levels().float(), m.get_noise_levels().float()) and torch.equal(torch.as_tensor(r.train_timesteps), torch.as_tensor(m.train_timesteps))
    i = rng.randrange(T)
    t, tn = rt[i], (rt[i + 1] if i + 1 < T else torch.tensor(0, dtype=rt.dtype))
    ok = ok and r.get_noise_level_for_timestep(t) == m.get_noise_level_for_timestep(t) and r.index_for_timestep(t) == m.index_for_timestep(t)
    cur = r.get_noise_level_for_timestep(t)
    import logging; logging.disable(logging.WARNING)
    o = r.step(noise_pred=v, timestep=t, latents=x, timestep_next=tn, noise_level=cur, compute_log_prob=False, return_dict=True)
    c = m.step_coef(t, tn, cur, compute_log_prob=False)
    ok = ok and float(o.dt.flatten()[0]) == c.dt and float(torch.as_tensor(o.std_dev_t).flatten()[0]) == c.std_dev_t
    # timestep_next omitted: the reference reads sigmas[i], sigmas[i + 1] from its tables (FlowMatch only: UniPC reads an integer timestep as an index)
    if isinstance(r, RefFM):
        o2 = r.step(noise_pred=v, timestep=t, latents=x, noise_level=cur, compute_log_prob=False, return_dict=True)
        c2 = m.step_coef(t, None, cur, compute_log_prob=False)
        ok = ok and float(o2.dt.flatten()[0]) == c2.dt and float(torch.as_tensor(o2.std_dev_t).flatten()[0]) == c2.std_dev_t
    if not ok: smis.append((case, T, dyn, nsde, sde, seed, type(r).__name__))
out["scheduler_mismatches"] = [repr(m) for m in smis[:5]]
# ---- positional tables and latent packing vs the real diffusers modules
from diffusers.models.transformers.transformer_flux import FluxPosEmbed
from diffusers.models.transformers.transformer_qwenimage import QwenEmbedRope
from diffusers.models.transformers.transformer_wan import WanRotaryPosEmbed
from diffusers.pipelines.flux.pipeline_flux import FluxPipeline
from flow_factory.scheduler.flow_match_euler_discrete import calculate_shift as ref_shift
from flow_factory_b200 import flux as MF, qwen as MQ, wan as MW
from flow_factory_b200.scheduler import calculate_shift as my_shift
pos = {}
h2, w2, nt = 6, 10, 7
ids = torch.cat([torch.zeros(nt, 3), FluxPipeline._prepare_latent_image_ids(1, h2, w2, "cpu", torch.float32)], 0)
rc, rs_ = FluxPosEmbed(theta=10000, axes_dim=[16, 56, 56])(ids)
mc, ms_ = MF.rope_tables(torch.cat([torch.zeros(nt, 3), MF.latent_image_ids(h2, w2)], 0), (16, 56, 56))
pos["flux_rope"] = bool(torch.equal(rc, mc) and torch.equal(rs_, ms_))
lat = torch.randn(2, 16, 2 * h2, 2 * w2)
pos["flux_pack"] = bool(torch.equal(FluxPipeline._pack_latents(lat, 2, 16, 2 * h2, 2 * w2), MF.pack_latents(lat)))
pos["shift"] = [ref_shift(n) == my_shift(n) for n in (256, 1024, 4096, 4429)]
qr = QwenEmbedRope(theta=10000, axes_dim=[16, 56, 56], scale_rope=True)
vid, txt = qr([(1, h2, w2)], max_txt_seq_len=nt, device=torch.device("cpu"))
qc, qs = MQ.qwen_rope_tables(h2, w2, nt, (16, 56, 56))
ref_c = torch.cat([txt, vid], 0)
pos["qwen_rope"] = bool(torch.equal(ref_c.real.float().repeat_interleave(2, 1), qc) and torch.equal(ref_c.imag.float().repeat_interleave(2, 1), qs))
wr = WanRotaryPosEmbed(attention_head_dim=128, patch_size=(1, 2, 2), max_seq_len=64)
wc, ws = wr(torch.zeros(1, 16, 3, 8, 10))
mwc, mws = MW.wan_rope_tables(MW.WanEngineConfig(rope_max_seq_len=64), 3, 4, 5, table_dtype=torch.float32)
pos["wan_rope"] = bool(torch.equal(wc.reshape(60, 128), mwc) and torch.equal(ws.reshape(60, 128), mws))
wrb = wr.to(torch.bfloat16)                                    # the buffers follow the module dtype
wcb, wsb = wrb(torch.zeros(1, 16, 3, 8, 10))
mwcb, mwsb = MW.wan_rope_tables(MW.WanEngineConfig(rope_max_seq_len=64), 3, 4, 5)
pos["wan_rope_bf16"] = bool(wcb.dtype == torch.bfloat16 and torch.equal(wcb.float().reshape(60, 128), mwcb) and torch.equal(wsb.float().reshape(60, 128), mwsb))
out["positional"] = pos
# ---- VAE host side against the real AutoencoderKL / VaeImageProcessor
from diffusers.models.autoencoders.autoencoder_kl import AutoencoderKL
from diffusers.image_processor import VaeImageProcessor
from flow_factory_b200 import vae as MV
vm = AutoencoderKL(in_channels=3, out_channels=3, down_block_types=("DownEncoderBlock2D",) * 2, up_block_types=("UpDecoderBlock2D",) * 2,
                   block_out_channels=(32, 64), layers_per_block=1, latent_channels=4, norm_num_groups=8, use_quant_conv=False,
                   use_post_quant_conv=False, scaling_factor=1.5305, shift_factor=0.0609)
vcfg = MV.VaeDecoderConfig.from_config(vm.config)
packed = MV.pack_vae_decoder_weights(vm.state_dict(), vcfg)
img = torch.randn(2, 3, 8, 8) * 2
vae_rec = {"cfg": [vcfg.latent_channels, list(vcfg.block_out_channels), vcfg.layers_per_block, vcfg.norm_num_groups, vcfg.scaling_factor, vcfg.shift_factor],
           "n_packed": [len(packed), MV.expected_weight_count(vcfg)],
           "postprocess": bool(torch.equal(VaeImageProcessor(vae_scale_factor=8).postprocess(img, output_type="pt"), MV.postprocess_pt(img)))}
try:
    MV.VaeDecoderConfig.from_config(AutoencoderKL(block_out_channels=(32,), norm_num_groups=8).config); vae_rec["post_quant_guard"] = "no error"
except NotImplementedError:
    vae_rec["post_quant_guard"] = "NotImplementedError"
out["vae_host"] = vae_rec
# ---- config intake + weight packing from REAL (tiny) diffusers transformers: every key the packers read exists in the real state dicts
from diffusers.models.transformers.transformer_sd3 import SD3Transformer2DModel
from diffusers.models.transformers.transformer_flux import FluxTransformer2DModel
from diffusers.models.transformers.transformer_qwenimage import QwenImageTransformer2DModel
from diffusers.models.transformers.transformer_wan import WanTransformer3DModel
from flow_factory_b200.weights import EngineConfig, PackedWeights
cpu = torch.device("cpu")
intake = {}
m = SD3Transformer2DModel(sample_size=16, patch_size=2, in_channels=16, num_layers=2, attention_head_dim=64, num_attention_heads=2,
                          joint_attention_dim=64, caption_projection_dim=128, pooled_projection_dim=32, out_channels=16, pos_embed_max_size=16,
                          dual_attention_layers=(0,), qk_norm="rms_norm")
c = EngineConfig.from_model_config(m.config)
pw = PackedWeights(c, m.state_dict(), cpu)
intake["sd3"] = [c.num_layers, c.num_heads, c.num_dual_layers, len(pw.tensors) > 0]
m = FluxTransformer2DModel(patch_size=1, in_channels=64, num_layers=1, num_single_layers=1, attention_head_dim=128, num_attention_heads=2,
                           joint_attention_dim=64, pooled_projection_dim=32, guidance_embeds=True, axes_dims_rope=(16, 56, 56))
c = MF.FluxEngineConfig.from_model_config(m.config)
pw = MF.FluxPackedWeights(c, m.state_dict(), cpu)
intake["flux"] = [c.num_layers, c.num_single_layers, c.num_heads, bool(c.guidance_embeds), pw.mod_rows == (12 + 3 + 2) * 256]
m = QwenImageTransformer2DModel(patch_size=2, in_channels=64, out_channels=16, num_layers=1, attention_head_dim=128, num_attention_heads=2,
                                joint_attention_dim=64, axes_dims_rope=(16, 56, 56))
c = MQ.qwen_engine_config(m.config)
pw = MF.FluxPackedWeights(c, m.state_dict(), cpu)
intake["qwen"] = [c.num_layers, c.variant, c.num_heads, len(pw.tensors) > 0]
m = WanTransformer3DModel(patch_size=(1, 2, 2), num_attention_heads=2, attention_head_dim=128, in_channels=16, out_channels=16, text_dim=64,
                          freq_dim=256, ffn_dim=320, num_layers=1, cross_attn_norm=True, qk_norm="rms_norm_across_heads", eps=1e-6,
                          rope_max_seq_len=64)
c = MW.WanEngineConfig.from_model_config(m.config)
g_, layers_ = MW.pack_wan_state_dict(c, m.state_dict())
intake["wan"] = [c.num_layers, c.num_attention_heads, c.ffn_dim, list(g_["pe_w"].shape), len(layers_)]
out["intake"] = intake
# ---- B200SD3_5Adapter.from_reference_adapter fed with real reference objects (engine stubbed: no GPU here)
import flow_factory_b200.adapter as MA
class _Eng:
    def __init__(self, model_config, state_dict, device):
        self.device, self.cfg, self.n_keys = torch.device("cpu"), EngineConfig.from_model_config(model_config), len(state_dict)
MA.RolloutEngine = _Eng
class _RefAdapter:                                   # the attributes from_reference_adapter reads off a flow_factory SD3_5Adapter
    pass
ra = _RefAdapter()
ra.transformer = SD3Transformer2DModel(sample_size=16, patch_size=2, in_channels=16, num_layers=2, attention_head_dim=64, num_attention_heads=2,
                                       joint_attention_dim=64, caption_projection_dim=128, pooled_projection_dim=32, out_channels=16,
                                       pos_embed_max_size=16, dual_attention_layers=(0,), qk_norm="rms_norm")
ra.scheduler = RefFM(noise_level=0.6, sde_steps=[1, 2, 5], num_sde_steps=2, seed=7, dynamics_type="Dance-SDE", shift=3.0)
ra.device = "cpu"
ra.decode_latents = lambda lat, output_type="pt": lat * 0 + 0.5
b = MA.B200SD3_5Adapter.from_reference_adapter(ra, rng="torch")
ref_set(ra.scheduler, 10, seq_len=64, device="cpu"); b.scheduler.set_timesteps(10, seq_len=64)
out["from_reference"] = {"cfg": [b.model_config.num_layers, b.model_config.num_heads], "n_keys": b.engine.n_keys == len(ra.transformer.state_dict()),
                         "sched": [b.scheduler.noise_level, b.scheduler.dynamics_type, b.scheduler.seed, b.scheduler.num_sde_steps, b.scheduler.sde_steps.tolist()],
                         "timesteps_equal": bool(torch.equal(ra.scheduler.timesteps, b.scheduler.timesteps)),
                         "sde_equal": bool(torch.equal(ra.scheduler.current_sde_steps, b.scheduler.current_sde_steps)),
                         "decode": float(b.decode_latents(torch.zeros(1, 2)).mean())}
# ---- the reference-side glue file (integration/ff_b200_glue.py, quoted by INTEGRATION.md) driven with the heavy base-class parts stubbed
sys.path.insert(0, sys.argv[1] + "/integration")
import ff_b200_glue as GL
import flow_factory.models.abc as ABC
glue = GL.B200GlueSD3_5Adapter.__new__(GL.B200GlueSD3_5Adapter)
GL.B200GlueSD3_5Adapter.transformer = property(lambda self: ra.transformer)
GL.B200GlueSD3_5Adapter.device = property(lambda self: "cpu")
import types
glue.pipeline = types.SimpleNamespace(scheduler=ra.scheduler)        # BaseAdapter.scheduler is a property over pipeline.scheduler
glue.decode_latents = ra.decode_latents
glue._mode = "train"
GL.B200GlueSD3_5Adapter.trainable_component_names = property(lambda self: [])   # BaseAdapter.eval / rollout / train loop over these
glue._b200 = MA.B200SD3_5Adapter.from_reference_adapter(glue, rng="torch")
ra.scheduler.set_seed(123)
def _refresh(self, sd):
    self.refreshed = getattr(self, "refreshed", 0) + 1
    self.last_sd = {k: v.detach().clone() for k, v in sd.items()}
_Eng.refresh_weights = _refresh
glue.rollout()
stale_after_rollout = glue._engine_stale
glue._sync_engine_if_stale()                          # what inference() / the no-grad forward() do first
g1 = [glue.mode, glue._b200.scheduler.seed, glue._b200.scheduler.is_eval, ra.scheduler.is_eval, glue._b200.engine.refreshed]
# ---- weight-swapping contexts (ADVICE r1: the KL-reference / EMA forward must run on the swapped weights, grpo.py:282, nft.py:360)
from flow_factory.ema import EMAModuleWrapper
params = list(ra.transformer.parameters())
GL.B200GlueSD3_5Adapter.get_trainable_parameters = lambda self: params
GL.B200GlueSD3_5Adapter.target_module_map = property(lambda self: {})
glue.model_args = types.SimpleNamespace(lora_alpha=8.0, finetune_type="full")
glue.ema_wrapper = EMAModuleWrapper(parameters=params, decay=0.5, update_step_interval=1, device="cpu")
glue._ref_ema = EMAModuleWrapper(parameters=params, decay=0.0, update_step_interval=0, device="cpu")
key = "proj_out.bias"
with torch.no_grad():
    for e in glue.ema_wrapper.ema_parameters: e.fill_(0.25)          # the EMA policy
    for e in glue._ref_ema.ema_parameters: e.fill_(-0.5)             # theta_ref
    for q in params: q.fill_(1.0)                                     # the current policy (an optimizer step happened)
glue._b200.forward = lambda *a, **k: float(glue._b200.engine.last_sd[key].mean())   # stand-in: reports which weights the engine holds
glue.rollout()
ctx = {}
with torch.no_grad():
    ctx["policy"] = glue.forward()
    with glue.use_ref_parameters():
        ctx["ref"] = glue.forward()
        ctx["ref_again_refreshes"] = glue._b200.engine.refreshed
        glue.forward()
        ctx["ref_again_refreshes"] = glue._b200.engine.refreshed - ctx["ref_again_refreshes"]
    ctx["after_ref"] = glue.forward()
    with glue.use_ema_parameters():
        ctx["ema"] = glue.forward()
        with glue.use_ref_parameters():
            ctx["ema_then_ref"] = glue.forward()
        ctx["back_in_ema"] = glue.forward()
    ctx["after_ema"] = glue.forward()
    ctx["module_restored"] = float(dict(ra.transformer.named_parameters())[key].mean())
# train mode: the optimizer steps between no-grad calls without any context edge
glue.train(True)
with torch.no_grad():
    for q in params: q.fill_(2.0)
    ctx["train_mode_tracks_optimizer"] = glue.forward()
    for q in params: q.fill_(1.0)
# ---- LoRA (finetune_type='lora'): PEFT key names at post_init time, adapter folded for the policy, NOT folded inside use_ref_parameters
import torch.nn as nn
class _LoraLinear(nn.Module):                           # PEFT's lora.Linear layout: base_layer + lora_A / lora_B ModuleDicts keyed by adapter name
    def __init__(self, base, r):
        super().__init__()
        self.base_layer = base
        self.lora_A = nn.ModuleDict({"default": nn.Linear(base.in_features, r, bias=False)})
        self.lora_B = nn.ModuleDict({"default": nn.Linear(r, base.out_features, bias=False)})
class _PeftLike(nn.Module):                             # PeftModel(base_model=LoraModel(model=transformer)) -> keys 'base_model.model.<...>'
    def __init__(self, model):
        super().__init__()
        self.base_model = nn.Module(); self.base_model.model = model
        self.config = model.config
tr2 = SD3Transformer2DModel(sample_size=16, patch_size=2, in_channels=16, num_layers=2, attention_head_dim=64, num_attention_heads=2,
                            joint_attention_dim=64, caption_projection_dim=128, pooled_projection_dim=32, out_channels=16,
                            pos_embed_max_size=16, dual_attention_layers=(0,), qk_norm="rms_norm")
base_q = tr2.transformer_blocks[0].attn.to_q.weight.detach().clone()
tr2.transformer_blocks[0].attn.to_q = _LoraLinear(tr2.transformer_blocks[0].attn.to_q, r=4)
with torch.no_grad():
    tr2.transformer_blocks[0].attn.to_q.lora_A["default"].weight.fill_(0.5)
    tr2.transformer_blocks[0].attn.to_q.lora_B["default"].weight.fill_(0.25)
peft_like = _PeftLike(tr2)
GL.B200GlueSD3_5Adapter.transformer = property(lambda self: peft_like)
glue.model_args = types.SimpleNamespace(lora_alpha=8.0, finetune_type="lora")
lora = {"peft_keys": sorted(k for k in peft_like.state_dict() if "to_q" in k and "transformer_blocks.0.attn." in k)}
built = MA.B200SD3_5Adapter.from_reference_adapter(glue, rng="torch")          # what post_init does; raised KeyError before the fix
lora["post_init_plain_keys"] = built.engine.n_keys == len(ra.transformer.state_dict())
qk = "transformer_blocks.0.attn.to_q.weight"
glue._b200.forward = lambda *a, **k: float((glue._b200.engine.last_sd[qk] - base_q).abs().max())
glue.rollout()
with torch.no_grad():
    lora["policy_delta"] = glue.forward()              # (alpha / r) * B A = 2 * (0.25 * 0.5 * 4) = 1.0 on every entry
    with glue.use_ref_parameters():
        lora["ref_delta"] = glue.forward()             # adapter disabled: base weights
    lora["after_delta"] = glue.forward()
lora["plain_keys_only"] = not any(("lora_" in k) or ("base_layer" in k) or k.startswith("base_model.") for k in glue._b200.engine.last_sd)
GL.B200GlueSD3_5Adapter.transformer = property(lambda self: ra.transformer)
glue.model_args = types.SimpleNamespace(lora_alpha=8.0, finetune_type="full")
glue.rollout(); glue._sync_engine_if_stale()
# ---- option (b) of SURVEY 7.2 #1: old_log_prob of the stored SDE transitions re-evaluated through the REFERENCE forward
from flow_factory_b200.samples import SD3_5Sample as MySample
from flow_factory_b200.scheduler import SDESchedulerOutput as MyOut
def _mk(b):
    # T = 4: latents kept at positions 0, 1, 2, 4 (slots 0..3); log-probs kept for steps 0, 1 (slots 0, 1) and step 3 (slot 2), whose start
    # latent (position 3) is NOT stored
    return MySample(timesteps=torch.tensor([1000.0, 900.0, 750.0, 500.0]), all_latents=torch.arange(4.0).view(4, 1, 1, 1) + 10 * b,
                    log_probs=torch.full((3,), -1.0), latent_index_map=torch.tensor([0, 1, 2, -1, 3]), log_prob_index_map=torch.tensor([0, 1, -1, 2]))
fake = [_mk(0), _mk(1)]
glue._b200.inference = lambda *a, **k: fake
calls_b = []
def _ref_forward(self, **kw):
    calls_b.append({"t": float(kw["t"]), "t_next": float(kw["t_next"]), "x_t": kw["latents"].flatten().tolist(), "x_n": kw["next_latents"].flatten().tolist(),
                    "g": kw["guidance_scale"], "rk": kw["return_kwargs"], "nl": kw["noise_level"], "pe": kw["prompt_embeds"].shape[0]})
    return MyOut(log_prob=torch.tensor([100.0 + kw["latents"].flatten()[0].item(), 200.0 + kw["latents"].flatten()[1].item()]))
GL.SD3_5Adapter.forward = _ref_forward
ra.scheduler.get_noise_level_for_timestep = lambda t: 0.7
glue._mode = "rollout"
res = glue.inference(prompt_embeds=torch.zeros(2, 3, 4), pooled_prompt_embeds=torch.zeros(2, 4), guidance_scale=4.5, compute_log_prob=True)
rec = {"calls": calls_b, "log_probs": [s_.log_probs.tolist() for s_ in res]}
glue.recompute_old_log_probs = False
fake[0].log_probs.fill_(-1.0); n_before = len(calls_b)
glue.inference(prompt_embeds=torch.zeros(2, 3, 4), pooled_prompt_embeds=torch.zeros(2, 4), guidance_scale=4.5, compute_log_prob=True)
rec["disabled_keeps_engine_values"] = [len(calls_b) == n_before, fake[0].log_probs.tolist()]
out["glue_contexts"] = {"stale_after_rollout": stale_after_rollout, "ctx": ctx, "lora": lora, "recompute": rec}
glue.eval()
g2 = [glue.mode, glue._b200.scheduler.is_eval, ra.scheduler.is_eval]
glue.train(True)
g3 = [glue.mode, glue._b200.scheduler.is_eval]
out["glue"] = [g1, g2, g3, [c.__name__ for c in GL.B200GlueSD3_5Adapter.__mro__[:3]]]
# ---- GRPO / GDPO advantage arithmetic (FF/advantage/advantage_processor.py:314-481) on randomised groups, incl. constant-reward groups
import numpy as np
from flow_factory.advantage.advantage_processor import AdvantageProcessor
from flow_factory_b200 import advantage as MyAdv
nrng = np.random.default_rng(1)
amis = []
for case in range(40):
    n_groups, k = int(nrng.integers(1, 7)), int(nrng.integers(1, 6))
    gid = nrng.permutation(np.repeat(np.arange(n_groups), k))
    keys = ["pick", "ocr", "aes"][: int(nrng.integers(1, 4))]
    rewards = {kk: (nrng.normal(size=len(gid)) if nrng.random() < 0.8 else np.full(len(gid), 0.5)) for kk in keys}
    if nrng.random() < 0.3:
        rewards[keys[0]][gid == 0] = 1.25                       # one group with identical rewards: std clamps
    weights = {kk: float(nrng.choice([1.0, 0.5, 2.0])) for kk in keys}
    for global_std in (True, False):
        ap = AdvantageProcessor.__new__(AdvantageProcessor)
        ap.reward_weights, ap.global_std, ap.group_size, ap.group_on_same_rank = weights, global_std, k, False
        ap.collect_group_rewards = lambda samples, rw, gid=gid: ({kk: np.asarray(v, dtype=np.float64) for kk, v in rw.items()}, gid)
        ap._to_local = lambda a: a
        ap._build_weighted_sum_log_data = lambda *a, **kw: {}
        ap._build_gdpo_log_data = lambda *a, **kw: {}
        r_sum = np.asarray(ap.compute_weighted_sum([], rewards, False))
        m_sum = np.asarray(MyAdv.advantages_sum(rewards, weights, gid, global_std=global_std))
        if not np.allclose(r_sum, m_sum, rtol=1e-12, atol=1e-12): amis.append(("sum", case, global_std))
    ap.global_std = True
    r_g = np.asarray(ap.compute_gdpo([], rewards, False))
    m_g = np.asarray(MyAdv.advantages_gdpo(rewards, weights, gid))
    if not np.allclose(r_g, m_g, rtol=1e-12, atol=1e-12): amis.append(("gdpo", case))
out["advantage_mismatches"] = [repr(m) for m in amis[:5]]
from flow_factory.scheduler.abc import SDESchedulerOutput as RefOut
from flow_factory_b200.scheduler import SDESchedulerOutput as MyOut
def probe(cls):
    o = cls.from_dict(dict(next_latents=torch.ones(2), log_prob=torch.zeros(2), noise_pred=torch.full((2,), 3.0), junk=1))
    res = {"keys": list(o.keys()), "len": len(o), "by_key": o["noise_pred"].tolist(), "by_index": o[0].tolist(), "tuple_len": len(o.to_tuple()),
           "iter": list(iter(o)), "contains": ["log_prob" in o, "dt" in o], "attr_none": o.dt is None}
    try:
        o["dt"]; res["missing_key"] = "no error"
    except KeyError:
        res["missing_key"] = "KeyError"
    return res
out["scheduler_output"] = [probe(RefOut), probe(MyOut)]
from diffusers.utils.torch_utils import randn_tensor as ref_randn
from flow_factory_b200.rng import randn_tensor as my_randn
from flow_factory.utils.base import create_generator_by_prompt
gens = lambda: create_generator_by_prompt(["a cat", "a dog", "a cat"], 42)          # what GRPOTrainer.evaluate passes (grpo.py:110)
rr = [bool(torch.equal(ref_randn((3, 4, 2, 2), generator=gens(), device=torch.device("cpu"), dtype=torch.bfloat16),
                       my_randn((3, 4, 2, 2), generator=gens(), device="cpu", dtype=torch.bfloat16))),
      bool(torch.equal(ref_randn((2, 5), generator=torch.Generator().manual_seed(3), device=torch.device("cpu"), dtype=torch.float32),
                       my_randn((2, 5), generator=torch.Generator().manual_seed(3), device="cpu", dtype=torch.float32))),
      bool(torch.equal(ref_randn((1, 5), generator=[torch.Generator().manual_seed(4)], device=torch.device("cpu"), dtype=torch.float32),
                       my_randn((1, 5), generator=[torch.Generator().manual_seed(4)], device="cpu", dtype=torch.float32)))]
g3 = gens()
rr.append(bool(torch.equal(my_randn((3, 2), generator=g3, dtype=torch.float32)[0], my_randn((3, 2), generator=gens(), dtype=torch.float32)[2])))  # same prompt, same noise
out["randn_tensor"] = rr
from flow_factory_b200.adapter import filter_kwargs as my_filter
def f1(a, b=1): pass
def f2(a, **kw): pass
def f3(*args, c=3): pass
out["filter_kwargs"] = [[sorted(filter_kwargs(f, a=1, b=2, c=3, z=4)) for f in (f1, f2, f3)], [sorted(my_filter(f, a=1, b=2, c=3, z=4)) for f in (f1, f2, f3)]]
register_scheduler("FlowMatchEulerDiscreteScheduler", "flow_factory_b200.scheduler.FlowMatchEulerDiscreteSDEScheduler")
register_scheduler("UniPCMultistepScheduler", "flow_factory_b200.scheduler.UniPCMultistepSDEScheduler")
class FlowMatchEulerDiscreteScheduler: pass
class UniPCMultistepScheduler: pass
out["sched"] = [get_sde_scheduler_class(FlowMatchEulerDiscreteScheduler()).__module__, get_sde_scheduler_class(UniPCMultistepScheduler).__module__]
# ---- diffusers' own attention plug points (SURVEY 8b ii / iii) with the real vendored classes; the kernels need a GPU, so the native forward
# is replaced by an SDPA stand-in HERE and everything around it (registry slot, dispatch, (B,S,H,D) layout, scale, the SD3 processor's joint
# [image ; text] order, output split, to_out / to_add_out, autograd) is the shipped code
import flow_factory_b200.diffusers_hooks as DH
from diffusers.models.attention_dispatch import AttentionBackendName, _AttentionBackendRegistry as REG, dispatch_attention_fn
import torch.nn.functional as F
hk = {}
try:
    dispatch_attention_fn(torch.zeros(1, 4, 2, 64), torch.zeros(1, 4, 2, 64), torch.zeros(1, 4, 2, 64), backend=DH.install_attention_backend("_native_flash"))
    hk["cpu_call_raises"] = False
except RuntimeError as e:
    hk["cpu_call_raises"] = "CUDA tensors only" in str(e)                  # dispatch reached OUR function, which has no CPU path
hk["slot"] = [REG._backends[AttentionBackendName._NATIVE_FLASH] is DH.b200_attention_backend,
              sorted(REG._supported_arg_names[AttentionBackendName._NATIVE_FLASH]) == sorted(REG._supported_arg_names[AttentionBackendName.NATIVE]),
              REG._constraints[AttentionBackendName._NATIVE_FLASH] == []]
calls = []
def _stand_in(q, k, v, scale):
    calls.append((tuple(q.shape), tuple(k.shape), scale))
    return F.scaled_dot_product_attention(q.permute(0, 2, 1, 3), k.permute(0, 2, 1, 3), v.permute(0, 2, 1, 3), scale=scale).permute(0, 2, 1, 3)
DH._forward_native = _stand_in
DH._check_device_dtype = lambda q, k, v: None
torch.manual_seed(0)
fl = FluxTransformer2DModel(patch_size=1, in_channels=16, num_layers=1, num_single_layers=1, attention_head_dim=64, num_attention_heads=2,
                            joint_attention_dim=32, pooled_projection_dim=16, guidance_embeds=True, axes_dims_rope=(8, 28, 28)).eval()
fin = dict(hidden_states=torch.randn(2, 12, 16), encoder_hidden_states=torch.randn(2, 5, 32), pooled_projections=torch.randn(2, 16),
           timestep=torch.tensor([0.5, 0.5]), img_ids=torch.zeros(12, 3), txt_ids=torch.zeros(5, 3), guidance=torch.tensor([3.5, 3.5]))
with torch.no_grad():
    base = fl(**fin).sample
    fl.set_attention_backend("_native_flash")            # what model.attn_backend: "_native_flash" does (FF/models/abc.py:782-798)
    n0 = len(calls)
    hooked = fl(**fin).sample
hk["flux_calls"] = len(calls) - n0
hk["flux_shapes"] = [list(calls[-1][0]), calls[-1][2]]
hk["flux_max_diff"] = float((base - hooked).abs().max())
# autograd through the backend function: native forward, gradients from the SDPA recomputation
q = torch.randn(1, 6, 2, 64, requires_grad=True); k = torch.randn(1, 6, 2, 64, requires_grad=True); v = torch.randn(1, 6, 2, 64, requires_grad=True)
o = DH.b200_attention_backend(q, k, v, scale=0.2); o.square().sum().backward()
q2, k2, v2 = (t.detach().clone().requires_grad_(True) for t in (q, k, v))
o2 = F.scaled_dot_product_attention(q2.permute(0, 2, 1, 3), k2.permute(0, 2, 1, 3), v2.permute(0, 2, 1, 3), scale=0.2).permute(0, 2, 1, 3); o2.square().sum().backward()
hk["grad_max_diff"] = max(float((a.grad - b.grad).abs().max()) for a, b in ((q, q2), (k, k2), (v, v2)))
rej = []
for kw in (dict(attn_mask=torch.ones(1, 6, dtype=torch.bool)), dict(is_causal=True), dict(dropout_p=0.1), dict(return_lse=True), dict(enable_gqa=True)):
    try:
        DH.b200_attention_backend(q.detach(), k.detach(), v.detach(), **kw); rej.append(False)
    except (NotImplementedError, ValueError):
        rej.append(True)
hk["unsupported_arguments_raise"] = rej
DH.uninstall_attention_backend("_native_flash")
hk["uninstalled"] = REG._backends[AttentionBackendName._NATIVE_FLASH] is not DH.b200_attention_backend
fl.reset_attention_backend()
# the SD3 processor on a real SD3Transformer2DModel (dual-attention block 0 + a context_pre_only last block)
sd = SD3Transformer2DModel(sample_size=16, patch_size=2, in_channels=16, num_layers=2, attention_head_dim=64, num_attention_heads=2,
                           joint_attention_dim=64, caption_projection_dim=128, pooled_projection_dim=32, out_channels=16,
                           pos_embed_max_size=16, dual_attention_layers=(0,), qk_norm="rms_norm").eval()
sin = dict(hidden_states=torch.randn(2, 16, 8, 8), encoder_hidden_states=torch.randn(2, 7, 64), pooled_projections=torch.randn(2, 32),
           timestep=torch.tensor([500.0, 500.0]))
with torch.no_grad():
    base = sd(**sin).sample
    proc = DH.install_sd3_attn_processor(sd)
    n0 = len(calls)
    hooked = sd(**sin).sample
hk["sd3_processors"] = sorted({type(p_).__name__ for p_ in sd.attn_processors.values()})
hk["sd3_calls"] = len(calls) - n0                       # block 0: joint + attn2, block 1: joint
hk["sd3_joint_len"] = sorted({c[0][1] for c in calls[n0:]})
hk["sd3_max_diff"] = float((base - hooked).abs().max())
import inspect as _insp
from diffusers.models.attention_processor import JointAttnProcessor2_0
hk["sd3_signature"] = list(_insp.signature(DH.B200JointAttnProcessor.__call__).parameters)[:5] == list(_insp.signature(JointAttnProcessor2_0.__call__).parameters)[:5]
out["diffusers_hooks"] = hk
print("RESULT " + json.dumps(out))
'''


@pytest.fixture(scope="module")
def hooks():
    r = subprocess.run([sys.executable, "-c", _SCRIPT, ROOT], capture_output=True, text=True, timeout=600)
    lines = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
    assert lines, r.stdout[-2000:] + r.stderr[-4000:]
    return json.loads(lines[-1][7:])


def test_registry_resolves_b200_adapters_by_python_path(hooks):
    assert hooks["sd3_5"]["resolved"] == "flow_factory_b200.adapter.B200SD3_5Adapter"
    assert hooks["flux1"]["resolved"] == "flow_factory_b200.flux_adapter.B200Flux1Adapter"
    assert hooks["qwen"]["resolved"] == "flow_factory_b200.qwen_adapter.B200QwenImageAdapter"
    assert hooks["wan"]["resolved"] == "flow_factory_b200.wan_adapter.B200Wan21Adapter"
    assert hooks["sched"] == ["flow_factory_b200.scheduler", "flow_factory_b200.scheduler"]


@pytest.mark.parametrize("model", ["sd3_5", "flux1", "qwen", "wan"])
def test_keyword_abi_is_a_superset_of_the_reference(hooks, model):
    rec = hooks[model]
    assert rec["inference_missing"] == [], rec
    assert rec["forward_missing"] == [], rec
    assert rec["inference_filtered"] and rec["forward_filtered"]


@pytest.mark.parametrize("model", ["sd3_5", "flux1", "qwen", "wan"])
def test_sample_records_carry_the_reference_fields(hooks, model):
    rec = hooks[model]
    assert rec["shared_fields"][0] == rec["shared_fields"][1], rec["shared_fields"]      # what stack() collates as one value per batch
    assert rec["sample_fields_missing"] == [], rec["sample_fields_missing"]            # every field of the reference record exists


@pytest.mark.parametrize("model", ["sd3_5", "flux1", "qwen", "wan"])
def test_sample_records_behave_like_the_reference(hooks, model):
    """Same data into the reference record and the B200 record: identical ids, dict views, extra_kwargs fall-through and collate; and the
    reference's own `BaseSample.stack` (what the trainers call) accepts B200 records and returns the same batch."""
    beh = hooks[model]["behaviour"]
    assert beh["unique_id"][0] == beh["unique_id"][1] and beh["unique_id"][0][0] != beh["unique_id"][0][1]
    assert beh["unique_id_ids"][0] == beh["unique_id_ids"][1]
    assert beh["to_dict_keys"][0] == beh["to_dict_keys"][1]
    assert beh["getattr_extra"][0] == beh["getattr_extra"][1] and beh["getattr_extra"][2] == beh["getattr_extra"][3]
    assert beh["stack"][0] == beh["stack"][1]
    assert beh["ref_stack_on_mine"] is True
    assert beh["from_dict"][0] == beh["from_dict"][1]
    assert beh["reset_on_set"][0] == beh["reset_on_set"][1]


def test_trajectory_bookkeeping_matches_the_reference_collectors(hooks):
    """300 randomised (T, trajectory_indices, SDE-step) cases: stored-latent / log-prob counts and both index maps equal what the
    reference's TrajectoryCollector pair produces when driven like SD3_5Adapter.inference."""
    assert hooks["trajectory_mismatches"] == []


def test_scheduler_mirrors_match_the_reference_on_random_settings(hooks):
    """60 randomised settings of both scheduler mirrors: timesteps, sigmas, SDE-step selection under the seed, noise levels, index lookup
    and the per-step scalars (dt, std_dev_t of the reference's step) - all bit-exact."""
    assert hooks["scheduler_mismatches"] == []


def test_filter_kwargs_mirror(hooks):
    assert hooks["filter_kwargs"][0] == hooks["filter_kwargs"][1]


def test_scheduler_output_mapping_protocol(hooks):
    """`output['noise_pred']` (NFT / AWM / CRD trainers), `output[0]`, keys / iteration over the non-None fields - as diffusers' BaseOutput."""
    assert hooks["scheduler_output"][0] == hooks["scheduler_output"][1]


def test_advantage_arithmetic_matches_the_reference_on_random_groups(hooks):
    assert hooks["advantage_mismatches"] == []


def test_positional_tables_and_packing_match_diffusers(hooks):
    """FluxPosEmbed / QwenEmbedRope (scale_rope) / WanRotaryPosEmbed (fp32 and bf16-cast buffers), FLUX latent packing + image ids and
    calculate_shift: the host tables the kernels read are bit-identical to what the real modules produce."""
    pos = hooks["positional"]
    assert all(pos["shift"])
    assert {k: v for k, v in pos.items() if k != "shift"} == {"flux_rope": True, "flux_pack": True, "qwen_rope": True, "wan_rope": True, "wan_rope_bf16": True}


def test_vae_host_side_against_the_real_autoencoder(hooks):
    """Config intake from a real AutoencoderKL.config, packing of its real state_dict, the 'pt' postprocess, and the post_quant_conv guard
    (diffusers' default AutoencoderKL has one)."""
    v = hooks["vae_host"]
    assert v["cfg"] == [4, [32, 64], 1, 8, 1.5305, 0.0609]
    assert v["n_packed"][0] == v["n_packed"][1]
    assert v["postprocess"] is True and v["post_quant_guard"] == "NotImplementedError"


def test_config_intake_and_packing_from_real_models(hooks):
    """`transformer.config` (FrozenDict) and `transformer.state_dict()` of real (tiny) diffusers models go through every engine's config
    intake and weight packer - the key names the packers read are the real ones."""
    it = hooks["intake"]
    assert it["sd3"] == [2, 2, 1, True]
    assert it["flux"] == [1, 1, 2, True, True]
    assert it["qwen"] == [1, 1, 2, True]
    assert it["wan"] == [1, 2, 320, [256, 64], 1]


def test_from_reference_adapter_reads_real_reference_objects(hooks):
    """The constructor INTEGRATION.md's glue uses: a real SD3Transformer2DModel and a real reference scheduler instance in, the engine config,
    every weight, the scheduler settings (incl. SDE-step list / seed / dynamics) and the decode callback out."""
    r = hooks["from_reference"]
    assert r["cfg"] == [2, 2] and r["n_keys"] is True
    assert r["sched"] == [0.6, "Dance-SDE", 7, 2, [1, 2, 5]]
    assert r["timesteps_equal"] and r["sde_equal"] and r["decode"] == 0.5


def test_randn_tensor_mirror(hooks):
    """Initial latents from per-prompt CPU generator lists (GRPOTrainer.evaluate) are the same numbers diffusers' randn_tensor draws."""
    assert hooks["randn_tensor"] == [True, True, True, True]


def test_reference_side_glue_file(hooks):
    """integration/ff_b200_glue.py (subclass of the REAL SD3_5Adapter): rollout() re-packs the weights and syncs the seed, eval() / train()
    reach both schedulers."""
    g1, g2, g3, mro = hooks["glue"]
    assert g1 == ["rollout", 123, False, False, 1]
    assert g2 == ["eval", True, True] and g3 == ["train", False]
    assert mro[:2] == ["B200GlueSD3_5Adapter", "SD3_5Adapter"]


def test_glue_engine_follows_weight_swapping_contexts(hooks):
    """The engine's packed copy follows use_ref_parameters / use_ema_parameters (nested too) and the optimizer in train mode: the no-grad
    KL-reference forward (grpo.py:282-292) and NFT's sampling_context forward (nft.py:360) see theta_ref / the EMA policy, not the weights
    of the last rollout.  The stand-in forward reports the mean of the packed proj_out.bias: policy 1.0, EMA 0.25, theta_ref -0.5."""
    g = hooks["glue_contexts"]
    assert g["stale_after_rollout"] is True
    c = g["ctx"]
    assert c["policy"] == 1.0 and c["ref"] == -0.5 and c["after_ref"] == 1.0
    assert c["ref_again_refreshes"] == 0                      # same context, nothing changed: no second re-pack
    assert c["ema"] == 0.25 and c["ema_then_ref"] == -0.5 and c["back_in_ema"] == 0.25 and c["after_ema"] == 1.0
    assert c["module_restored"] == 1.0
    assert c["train_mode_tracks_optimizer"] == 2.0


def test_glue_handles_a_peft_wrapped_transformer(hooks):
    """finetune_type='lora' (examples/grpo/lora/sd3_5/default.yaml): post_init sees PEFT key names; the policy engine holds W + (alpha/r) B A,
    inside use_ref_parameters() (PEFT disable_adapter: no tensor changes) it holds the base W."""
    l = hooks["glue_contexts"]["lora"]
    assert any(k.endswith("to_q.base_layer.weight") for k in l["peft_keys"]) and any("lora_A.default" in k for k in l["peft_keys"])
    assert all(k.startswith("base_model.model.") for k in l["peft_keys"])
    assert l["post_init_plain_keys"] and l["plain_keys_only"]
    assert abs(l["policy_delta"] - 1.0) < 1e-6 and l["ref_delta"] == 0.0 and abs(l["after_delta"] - 1.0) < 1e-6


def test_attention_backend_slot_routes_to_the_b200_function(hooks):
    """`install_attention_backend` + the reference's own `set_attention_backend` / `dispatch_attention_fn`: the registry slot holds our
    function with the registry's argument names; a FLUX forward through it (kernel replaced by an SDPA stand-in on this GPU-less box) is
    the default backend's output - layout (B,S,H,D), scale and call count (1 dual + 1 single block) are right; gradients come from the
    SDPA recomputation; unsupported arguments raise instead of being ignored; uninstall restores the slot."""
    h = hooks["diffusers_hooks"]
    assert h["cpu_call_raises"] is True and h["slot"] == [True, True, True]
    assert h["flux_calls"] == 2 and h["flux_shapes"][0] == [2, 17, 2, 64] and h["flux_max_diff"] <= 1e-5
    assert h["grad_max_diff"] <= 1e-5
    assert all(h["unsupported_arguments_raise"]) and h["uninstalled"] is True


def test_sd3_attn_processor_is_a_drop_in_for_joint_attn_processor(hooks):
    """`B200JointAttnProcessor` installed through `set_attn_processor` on a real SD3Transformer2DModel: same call signature as
    JointAttnProcessor2_0, three attention calls (joint + attn2 in the dual block, joint in the context_pre_only block) over
    [image ; text] (16 + 7 tokens) and image-only (16) sequences, and the model output of the stock processor."""
    h = hooks["diffusers_hooks"]
    assert h["sd3_processors"] == ["B200JointAttnProcessor"] and h["sd3_signature"] is True
    assert h["sd3_calls"] == 3 and h["sd3_joint_len"] == [16, 23]
    assert h["sd3_max_diff"] <= 2e-3          # q / k / v pass through bf16 (the kernel's input type); a wrong token order would be O(1)


def test_glue_recomputes_old_log_probs_through_the_reference_forward(hooks):
    """SURVEY 7.2 #1 option (b): the stored SDE transitions (both end points kept) are replayed teacher-forced through SD3_5Adapter.forward
    (the reference path, not the engine) and their log-probs replace the engine's; a log-prob whose start latent is not stored keeps the
    engine's value; the switch turns it off."""
    r = hooks["glue_contexts"]["recompute"]
    assert [(c["t"], c["t_next"]) for c in r["calls"]] == [(1000.0, 900.0), (900.0, 750.0)]
    assert r["calls"][0]["x_t"] == [0.0, 10.0] and r["calls"][0]["x_n"] == [1.0, 11.0] and r["calls"][1]["x_t"] == [1.0, 11.0]
    assert all(c["g"] == 4.5 and c["rk"] == ["log_prob"] and c["nl"] == 0.7 and c["pe"] == 2 for c in r["calls"])
    assert r["log_probs"] == [[100.0, 101.0, -1.0], [210.0, 211.0, -1.0]]
    assert r["disabled_keeps_engine_values"] == [True, [-1.0, -1.0, -1.0]]
