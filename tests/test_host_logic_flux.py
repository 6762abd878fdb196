"""CPU tests of the FLUX.1 host-side logic (no GPU): index / table helpers against the oracle (itself pinned against the reference),
the dynamic-shift schedule against the golden fixture, the adapter's keyword ABI and the C struct layouts."""
import ctypes as C
import inspect
import os

import pytest
import torch

from flow_factory_b200 import flux as F
from flow_factory_b200.scheduler import FlowMatchEulerDiscreteSDEScheduler, set_scheduler_timesteps
from oracle import flux_oracle as FO


def test_pack_ids_and_rope_tables_match_the_oracle():
    lat = torch.randn(2, 16, 12, 8)
    assert torch.equal(F.pack_latents(lat), FO.pack_latents(lat))
    assert torch.equal(F.latent_image_ids(6, 4), FO.latent_image_ids(12, 8))
    ids = torch.cat([torch.zeros(7, 3), F.latent_image_ids(6, 4)])
    c1, s1 = F.rope_tables(ids, (16, 56, 56))
    c2, s2 = FO.rope_tables(ids, (16, 56, 56))
    assert torch.equal(c1, c2) and torch.equal(s1, s2) and c1.shape == (31, 128) and c1.dtype == torch.float32
    # interleaved repetition: column 2i == column 2i+1 (get_1d_rotary_pos_embed, repeat_interleave_real)
    assert torch.equal(c1[:, 0::2], c1[:, 1::2]) and torch.equal(s1[:, 0::2], s1[:, 1::2])


def test_dynamic_shift_schedule_matches_the_golden_fixture(golden_dir):
    g = torch.load(os.path.join(golden_dir, "flux_tiny.pt"), weights_only=False)["rollout_fp32"]
    ts, sig = F.flux_make_schedule(4, 16)
    assert torch.equal(ts, g["timesteps"]) and torch.equal(sig, g["sigmas"])
    sch = FlowMatchEulerDiscreteSDEScheduler(noise_level=0.7, use_dynamic_shifting=True)
    ts2 = set_scheduler_timesteps(sch, 4, seq_len=16)
    assert torch.equal(ts2, g["timesteps"]) and torch.equal(sch.sigmas, g["sigmas"])
    gs = torch.load(os.path.join(golden_dir, "schedule.pt"), weights_only=False)
    ts3, sig3 = F.flux_make_schedule(28, 4096)
    assert torch.equal(ts3, gs["dyn_T28"]["timesteps"]) and torch.equal(sig3, gs["dyn_T28"]["sigmas"])


def test_model_scalar_follows_the_bf16_casts():
    # transformer_flux.py:679: timestep.to(bf16) * 1000 in bf16
    assert F.model_scalar(0.5) == 500.0
    t = 0.98863
    expect = float((torch.tensor(t).to(torch.bfloat16) * 1000).float())
    assert F.model_scalar(t) == expect
    # guidance 3.5 is exact in fp16 and bf16 -> 3500 rounds to the bf16 grid (ulp 16 at 2048..4096)
    assert F.model_scalar(3.5, torch.float16) == float((torch.tensor(3.5).bfloat16() * 1000).float()) == 3504.0


def test_flux_adapter_keywords_are_the_reference_abi():
    from flow_factory_b200.flux_adapter import B200Flux1Adapter
    # FF/models/flux/flux1.py:153-172 and 296-312
    ref_inf = ["prompt", "height", "width", "num_inference_steps", "guidance_scale", "generator", "prompt_ids", "prompt_embeds",
               "pooled_prompt_embeds", "joint_attention_kwargs", "compute_log_prob", "extra_call_back_kwargs", "trajectory_indices"]
    ref_fwd = ["t", "latents", "prompt_embeds", "pooled_prompt_embeds", "img_ids", "t_next", "next_latents", "guidance_scale",
               "noise_level", "joint_attention_kwargs", "compute_log_prob", "return_kwargs"]
    assert set(ref_inf) <= set(inspect.signature(B200Flux1Adapter.inference).parameters)
    assert set(ref_fwd) <= set(inspect.signature(B200Flux1Adapter.forward).parameters)


def test_flux_struct_layouts_and_config():
    assert C.sizeof(F.FluxConfigC) == 8 * 4
    assert C.sizeof(F.FluxDualWeights) == 20 * 8 and C.sizeof(F.FluxSingleWeights) == 8 * 8
    assert C.sizeof(F.FluxWeights) == 21 * 8 + 2 * 8
    cfg = F.FluxEngineConfig.from_model_config(FO.flux1_dev())
    assert (cfg.num_layers, cfg.num_single_layers, cfg.inner_dim, cfg.guidance_embeds) == (19, 38, 3072, True)
    with pytest.raises(ValueError):
        F.FluxEngineConfig.from_model_config(dict(FO.flux1_dev().ref_kwargs(), attention_head_dim=64))


def test_flux_sample_collate_shares_img_ids():
    from flow_factory_b200.samples import Flux1Sample
    ids = torch.zeros(4, 3)
    a = Flux1Sample(all_latents=torch.zeros(3, 4, 64), img_ids=ids, height=32, width=32)
    b = Flux1Sample(all_latents=torch.ones(3, 4, 64), img_ids=ids, height=32, width=32)
    out = Flux1Sample.stack([a, b])
    assert out["img_ids"] is ids and tuple(out["all_latents"].shape) == (2, 3, 4, 64)


def test_qwen_config_and_rope_tables():
    from flow_factory_b200.qwen import qwen_engine_config, qwen_rope_tables
    from oracle import qwen_oracle as QO
    cfg = qwen_engine_config(QO.qwen_image_20b())
    assert (cfg.variant, cfg.num_layers, cfg.num_single_layers, cfg.inner_dim, cfg.joint_attention_dim) == (1, 60, 0, 3072, 3584)
    with pytest.raises(NotImplementedError):
        qwen_engine_config(dict(QO.qwen_image_20b().ref_kwargs(), zero_cond_t=True))
    cos, sin = qwen_rope_tables(6, 4, 9, (16, 56, 56))
    vid, txt = QO.qwen_rope(1, 6, 4, 9, (16, 56, 56))
    f = torch.cat([txt, vid])
    assert torch.equal(cos[:, 0::2], f.real) and torch.equal(cos[:, 1::2], f.real)
    assert torch.equal(sin[:, 0::2], f.imag) and torch.equal(sin[:, 1::2], f.imag)


def test_qwen_adapter_keywords_are_the_reference_abi():
    from flow_factory_b200.qwen_adapter import B200QwenImageAdapter
    # FF/models/qwen_image/qwen_image.py:290-314 and 476-496
    ref_inf = ["prompt", "negative_prompt", "num_inference_steps", "guidance_scale", "height", "width", "generator", "prompt_ids",
               "prompt_embeds", "prompt_embeds_mask", "negative_prompt_ids", "negative_prompt_embeds", "negative_prompt_embeds_mask",
               "attention_kwargs", "max_sequence_length", "compute_log_prob", "extra_call_back_kwargs", "trajectory_indices"]
    ref_fwd = ["t", "latents", "prompt_embeds", "prompt_embeds_mask", "img_shapes", "negative_prompt_embeds", "negative_prompt_embeds_mask",
               "guidance_scale", "t_next", "next_latents", "noise_level", "attention_kwargs", "compute_log_prob", "return_kwargs"]
    assert set(ref_inf) <= set(inspect.signature(B200QwenImageAdapter.inference).parameters)
    assert set(ref_fwd) <= set(inspect.signature(B200QwenImageAdapter.forward).parameters)


def test_qwen_prompt_densify_and_mask_rules():
    """qwen_adapter._dense: ragged lists are right-padded with their lengths, prefix masks become lengths, anything else is refused."""
    from flow_factory_b200.qwen_adapter import _dense, _pad_to
    a, b = torch.randn(5, 8), torch.randn(3, 8)
    e, lens = _dense([a, b], None, "p")
    assert tuple(e.shape) == (2, 5, 8) and lens == [5, 3] and torch.equal(e[1, :3], b) and bool((e[1, 3:] == 0).all())
    e2, lens2 = _dense([a, b], [torch.ones(5), torch.tensor([1.0, 1.0, 0.0])], "p")
    assert lens2 == [5, 2]
    x = torch.randn(2, 6, 8)
    m = torch.tensor([[1, 1, 1, 1, 1, 1], [1, 1, 1, 0, 0, 0]])
    e3, lens3 = _dense(x, m, "p")
    assert e3 is x and lens3 == [6, 3]
    with pytest.raises(NotImplementedError):
        _dense(x, torch.tensor([[1, 1, 1, 1, 1, 1], [1, 0, 1, 0, 0, 0]]), "p")
    with pytest.raises(ValueError):
        _dense(x, torch.tensor([[1, 1, 1, 1, 1, 1], [0, 0, 0, 0, 0, 0]]), "p")
    assert tuple(_pad_to(x, 8).shape) == (2, 8, 8) and _pad_to(x, 6) is x


def test_flux_adapter_stepwise_callbacks_with_stub_engine(monkeypatch):
    """`extra_call_back_kwargs` on the FLUX.1 adapter: the step loop over forward() (stepwise.py) with a stub engine."""
    import pytest
    import torch
    from flow_factory_b200 import flux_adapter as FA
    from flow_factory_b200.flux import FluxEngineConfig, latent_image_ids
    from flow_factory_b200.scheduler import FlowMatchEulerDiscreteSDEScheduler
    from flow_factory_b200.trajectory import compute_trajectory_indices

    class Plan:
        def __init__(self, batch, h2, w2, n_text):
            self.batch, self.h2, self.w2, self.n_text, self.n_img, self.img_ids = batch, h2, w2, n_text, h2 * w2, latent_image_ids(h2, w2)

    class Eng:
        def __init__(self, model_config, state_dict, device):
            self.device, self.cfg, self.steps = torch.device("cpu"), FluxEngineConfig(num_layers=1, num_single_layers=1, num_heads=1), []

        def plan(self, batch, h2, w2, n_text, cfg=False):
            return Plan(batch, h2, w2, n_text)

        def set_prompts(self, *a, **k):
            pass

        def step(self, plan, latents, coef, noise=None, next_latents=None, seed=0):
            self.steps.append((coef.sigma, coef.sigma_prev, coef.noise_level))
            z = latents.float() * 0.5
            return dict(next_latents=z.half(), next_latents_mean=z, log_prob=torch.zeros(latents.shape[0]) if coef.compute_log_prob else None,
                        noise_pred=torch.zeros_like(latents, dtype=torch.bfloat16), overflow=torch.zeros(1))

    monkeypatch.setattr(FA, "FluxRolloutEngine", Eng)
    sch = FlowMatchEulerDiscreteSDEScheduler(noise_level=0.7, use_dynamic_shifting=True, num_sde_steps=1, seed=2)
    ad = FA.B200Flux1Adapter(None, {}, device="cpu", scheduler=sch, rng="philox")
    ad.rollout()
    T = 5
    sch.set_timesteps(T, seq_len=16)
    idx = compute_trajectory_indices(sch.train_timesteps.tolist(), T)
    pe, pp = torch.zeros(2, 3, 8), torch.zeros(2, 8)
    out = ad.inference(height=64, width=64, num_inference_steps=T, prompt_embeds=pe, pooled_prompt_embeds=pp, compute_log_prob=True,
                       trajectory_indices=idx, extra_call_back_kwargs=["next_latents_mean"], latents=torch.ones(2, 16, 64))
    assert len(ad.engine.steps) == T and ad.engine.steps[-1][1] == 0.0
    s0 = out[0]
    assert s0.all_latents.shape == (len(idx), 16, 64) and s0.img_ids.shape == (16, 3) and s0.callback_index_map.shape == (T,)
    assert s0.next_latents_mean.shape[1:] == (16, 64) and torch.equal(s0.final_latents, (torch.ones(16, 64) * 0.5 ** T).half())
    with pytest.raises(NotImplementedError):
        ad.inference(height=64, width=64, num_inference_steps=2, prompt_embeds=pe, pooled_prompt_embeds=pp, extra_call_back_kwargs=["img_ids"])

def test_qwen_adapter_stepwise_callbacks_with_stub_engine(monkeypatch):
    """`extra_call_back_kwargs` on the Qwen-Image adapter: the step loop over forward() with a stub engine (fast path untouched)."""
    from flow_factory_b200 import qwen_adapter as QA
    from flow_factory_b200.flux import FluxEngineConfig
    from flow_factory_b200.scheduler import FlowMatchEulerDiscreteSDEScheduler
    from flow_factory_b200.trajectory import compute_trajectory_indices

    class Plan:
        def __init__(self, batch, h2, w2, n_text, cfg):
            self.batch, self.h2, self.w2, self.n_text, self.n_img, self.cfg = batch, h2, w2, n_text, h2 * w2, cfg

    class Eng:
        def __init__(self, model_config, state_dict, device):
            self.device, self.cfg, self.steps, self.prompts = torch.device("cpu"), FluxEngineConfig(num_layers=1, num_single_layers=0, num_heads=1, variant=1), [], []

        def plan(self, batch, h2, w2, n_text, cfg=False):
            return Plan(batch, h2, w2, n_text, cfg)

        def set_prompts(self, plan, pe, npe, g, prompt_lengths=None, negative_lengths=None):
            self.prompts.append((tuple(pe.shape), None if npe is None else tuple(npe.shape), g, prompt_lengths, negative_lengths))

        def t_model(self, t, dtype=torch.float16):
            return float(t) / 1000

        def step(self, plan, latents, coef, noise=None, next_latents=None, seed=0):
            self.steps.append((coef.sigma, coef.sigma_prev, coef.noise_level))
            z = latents.float() * 0.5
            return dict(next_latents=z.half(), next_latents_mean=z, log_prob=torch.zeros(latents.shape[0]) if coef.compute_log_prob else None,
                        noise_pred=torch.zeros_like(latents, dtype=torch.bfloat16), overflow=torch.zeros(1))

    monkeypatch.setattr(QA, "QwenRolloutEngine", Eng)
    sch = FlowMatchEulerDiscreteSDEScheduler(noise_level=0.7, use_dynamic_shifting=True, num_sde_steps=1, seed=2, dynamics_type="Flow-SDE")
    ad = QA.B200QwenImageAdapter(None, {}, device="cpu", scheduler=sch, rng="philox")
    ad.rollout()
    T = 4
    sch.set_timesteps(T, seq_len=16)
    idx = compute_trajectory_indices(sch.train_timesteps.tolist(), T)
    pe = [torch.zeros(5, 8), torch.zeros(3, 8)]                      # ragged prompts, as the reference's list form
    npe = [torch.zeros(2, 8), torch.zeros(2, 8)]
    out = ad.inference(prompt=["a", "b"], height=64, width=64, num_inference_steps=T, guidance_scale=4.0, prompt_embeds=pe, negative_prompt_embeds=npe,
                       compute_log_prob=True, trajectory_indices=idx, extra_call_back_kwargs=["next_latents_mean"], latents=torch.ones(2, 16, 64))
    eng = ad.engine
    assert len(eng.steps) == T and eng.steps[-1][1] == 0.0
    assert eng.prompts[-1][3] == [5, 3] and eng.prompts[-1][4] == [2, 2] and eng.prompts[-1][2] == 4.0
    s0, s1 = out
    assert s0.all_latents.shape == (len(idx), 16, 64) and s0.img_shapes == [(1, 4, 4)] and s0.callback_index_map.shape == (T,)
    assert s0.prompt_embeds_mask.tolist() == [1, 1, 1, 1, 1] and s1.prompt_embeds_mask.tolist() == [1, 1, 1, 0, 0]
    assert s0.next_latents_mean.shape[1:] == (16, 64) and s1.prompt == "b"
    with pytest.raises(NotImplementedError):
        ad.inference(height=64, width=64, num_inference_steps=2, prompt_embeds=pe, extra_call_back_kwargs=["img_shapes"])
