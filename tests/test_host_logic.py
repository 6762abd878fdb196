"""CPU tests of the host-side mirror of the reference interface (scheduler bookkeeping, trajectory slots, samples,
C-ABI symbols).  No GPU, no compute calls."""
import ctypes as C
import json
import os
import re

import pytest
import torch

import flow_factory_b200 as F
from flow_factory_b200 import _lib, trajectory as TR
from flow_factory_b200.samples import SD3_5Sample
from oracle import sd3_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), weights_only=False)


def test_schedule_and_sde_window_match_reference(golden_dir):
    g = _load(golden_dir, "schedule.pt")
    for T, seq in ((4, 256), (30, 4096), (10, 1024)):
        s = F.FlowMatchEulerDiscreteSDEScheduler(noise_level=0.7, shift=3.0)
        ts = F.set_scheduler_timesteps(s, T, seq_len=seq)
        assert torch.equal(ts, g[f"T{T}"]["timesteps"]) and torch.equal(s.sigmas, g[f"T{T}"]["sigmas"])
        assert torch.equal(s.current_sde_steps, g[f"T{T}"]["sde"])
    for seed in (0, 42, 43):
        for n in (1, 3):
            s = F.FlowMatchEulerDiscreteSDEScheduler(noise_level=0.7, shift=3.0, num_sde_steps=n, seed=seed)
            F.set_scheduler_timesteps(s, 30, seq_len=4096)
            assert torch.equal(s.current_sde_steps, g[f"sde_seed{seed}_n{n}"])
            assert torch.equal(s.get_noise_levels(), g[f"noise_levels_seed{seed}_n{n}"])
            assert torch.equal(s.train_timesteps, s.current_sde_steps)
    for T, seq in ((28, 4096), (10, 1024)):   # dynamic shifting (mu from calculate_shift; FLUX-style configs)
        d = F.FlowMatchEulerDiscreteSDEScheduler(noise_level=0.7, use_dynamic_shifting=True)
        ts = F.set_scheduler_timesteps(d, T, seq_len=seq)
        assert torch.equal(ts, g[f"dyn_T{T}"]["timesteps"]) and torch.equal(d.sigmas, g[f"dyn_T{T}"]["sigmas"])
    s.set_seed(7); assert s.seed == 7
    s.eval(); assert s.is_eval
    s.rollout(); assert not s.is_eval


@pytest.mark.parametrize("dyn", ["Flow-SDE", "Dance-SDE", "CPS", "ODE"])
def test_step_coefficients_match_reference_scalars(golden_dir, dyn):
    """std_dev_t and dt of the packed coefficient block equal the reference's (B,1,1,1) tensors bit for bit."""
    g = _load(golden_dir, "sde_step.pt")
    s = F.FlowMatchEulerDiscreteSDEScheduler(noise_level=0.7, shift=3.0, dynamics_type=dyn)
    s.set_timesteps(30, seq_len=4096)
    for i in (0, 5, 28, 29):
        e = g[f"{dyn}_{i}"]
        c = s.step_coef(e["t"], e["tn"], e["nl"])
        assert c.std_dev_t == float(e["std"].flatten()[0])
        assert c.dt == float(e["dt"].flatten()[0])
        assert c.dynamics == {"Flow-SDE": 0, "Dance-SDE": 1, "CPS": 2, "ODE": 3}[dyn]
    assert s.get_noise_level_for_timestep(s.timesteps[3]) == 0.7
    assert s.get_noise_level_for_timestep(s.timesteps[29]) == 0.0     # default window excludes the last step


def test_trajectory_indices_and_maps_match_reference(golden_dir):
    with open(os.path.join(golden_dir, "trajectory.json")) as f:
        g = json.load(f)
    for c in g["cases"]:
        assert TR.compute_trajectory_indices(c["idx"], c["T"], c["inc"]) == c["out"]
    for m in g["maps"]:
        col = TR.create_trajectory_collector(m["ind"], m["T"])
        for p in range(m["T"] + 1):
            col.collect(torch.tensor(float(p)), p)
        im = col.get_index_map()
        assert (None if im is None else im.tolist()) == m["index_map"]
        res = col.get_result()
        assert (None if res is None else [float(x) for x in res]) == m["collected"]
        # slot planner == what the collector does when every position is offered
        lat_slot, _, lat_map, _ = TR.plan_slots(m["ind"], m["T"], [True] * m["T"])
        assert (None if lat_map is None else lat_map.tolist()) == m["index_map"]
        assert [p for p, s in enumerate(lat_slot) if s >= 0] == ([] if m["collected"] is None else [int(x) for x in m["collected"]])


def test_plan_slots_grpo_default():
    # num_sde_steps=1 at step 7 of 30: trainer needs positions 7 and 8 and the log-prob of step 7
    idx = TR.compute_trajectory_indices([7], 30)
    has = [i == 7 for i in range(30)]
    lat_slot, lp_slot, lat_map, lp_map = TR.plan_slots(idx, 30, has)
    assert idx == [7, 8] and lat_slot[7] == 0 and lat_slot[8] == 1 and sum(s >= 0 for s in lat_slot) == 2
    assert lp_slot[7] == 0 and sum(s >= 0 for s in lp_slot) == 1
    assert lat_map[7] == 0 and lat_map[8] == 1 and lp_map[7] == 0 and int((lp_map >= 0).sum()) == 1


def test_sample_record_stack_and_unique_id():
    a = SD3_5Sample(prompt="a cat", all_latents=torch.zeros(2, 16, 4, 4), log_probs=torch.zeros(1), height=32, width=32)
    b = SD3_5Sample(prompt="a cat", all_latents=torch.ones(2, 16, 4, 4), log_probs=torch.ones(1), height=32, width=32)
    c = SD3_5Sample(prompt="a dog")
    assert a.unique_id == b.unique_id != c.unique_id
    st = SD3_5Sample.stack([a, b])
    assert st["all_latents"].shape == (2, 2, 16, 4, 4) and st["log_probs"].shape == (2, 1) and st["prompt"] == ["a cat", "a cat"]


def test_filter_kwargs_is_the_abi():
    from flow_factory_b200.adapter import B200SD3_5Adapter, filter_kwargs
    kw = filter_kwargs(B200SD3_5Adapter.inference, height=512, width=512, num_inference_steps=10, guidance_scale=4.5,
                       compute_log_prob=True, trajectory_indices=[1, 2], prompt_embeds=1, per_device_batch_size=8, lr=1e-4)
    assert set(kw) == {"height", "width", "num_inference_steps", "guidance_scale", "compute_log_prob", "trajectory_indices", "prompt_embeds"}
    # every parameter of the reference's inference()/forward() (sd3_5.py:176-199, 352-370) is accepted by name
    ref_inf = ["prompt", "negative_prompt", "height", "width", "num_inference_steps", "guidance_scale", "generator",
               "joint_attention_kwargs", "prompt_ids", "prompt_embeds", "pooled_prompt_embeds", "negative_prompt_ids",
               "negative_prompt_embeds", "negative_pooled_prompt_embeds", "compute_log_prob", "extra_call_back_kwargs",
               "trajectory_indices"]
    ref_fwd = ["t", "latents", "prompt_embeds", "pooled_prompt_embeds", "negative_prompt_embeds", "negative_pooled_prompt_embeds",
               "guidance_scale", "t_next", "next_latents", "noise_level", "joint_attention_kwargs", "compute_log_prob", "return_kwargs"]
    import inspect
    assert set(ref_inf) <= set(inspect.signature(B200SD3_5Adapter.inference).parameters)
    assert set(ref_fwd) <= set(inspect.signature(B200SD3_5Adapter.forward).parameters)


def test_c_abi_library_loads_and_exports_every_declared_symbol():
    L = _lib.lib()
    assert L.ffb200_abi_version() == 3
    with open(os.path.join(ROOT, "include", "ffb200.h")) as f:
        hdr = f.read()
    declared = set(re.findall(r"\b(ffb200_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    for sym in declared:
        assert hasattr(L, sym), sym
    # struct layouts the ABI depends on
    assert C.sizeof(_lib.StepCoef) == 13 * 4 + 4 * 4
    assert C.sizeof(_lib.ModelConfig) == 8 * 4
    assert C.sizeof(_lib.LayerWeights) == 26 * 8


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "flow_factory_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), fn


def test_engine_refuses_to_run_without_a_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    cfg = O.tiny_config()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        F.RolloutEngine(cfg, O.make_weights(cfg))


def test_advantage_mirror_matches_reference(golden_dir):
    import numpy as np
    from flow_factory_b200 import advantage as A
    g = _load(golden_dir, "advantage.pt")
    for gs in (True, False):
        a = A.advantages_sum(g["rewards"], g["weights"], g["gid"], global_std=gs)
        np.testing.assert_allclose(a, g[f"sum_global{int(gs)}"], rtol=1e-12, atol=1e-12)
    a = A.advantages_gdpo(g["rewards"], g["weights"], g["gid"])
    np.testing.assert_allclose(a, g["gdpo_global1"], rtol=1e-12, atol=1e-12)


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    from flow_factory_b200 import dist as D
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        B = 3
        lat = (torch.arange(B * 2 * 4, dtype=torch.float32).reshape(B, 2, 4) + 100 * rank).half()
        lp = torch.arange(B * 2, dtype=torch.float32).reshape(B, 2) - 10 * rank
        gl, glp = D.all_gather_rollout(lat, lp)
        ok = gl.shape == (world * B, 2, 4) and glp.shape == (world * B, 2)
        for r in range(world):
            ok &= torch.equal(gl[r * B:(r + 1) * B], (torch.arange(B * 2 * 4, dtype=torch.float32).reshape(B, 2, 4) + 100 * r).half())
            ok &= torch.equal(glp[r * B:(r + 1) * B], torch.arange(B * 2, dtype=torch.float32).reshape(B, 2) - 10 * r)
        lo, hi = D.shard_range(10, rank, world)
        q.put((rank, bool(ok), lo, hi))
    finally:
        dist.destroy_process_group()


def test_prompt_sharding_and_single_allgather_world2_gloo():
    import torch.multiprocessing as mp
    from flow_factory_b200 import dist as D
    assert [D.shard_range(10, r, 3) for r in range(3)] == [(0, 4), (4, 7), (7, 10)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29611
    ps = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(timeout=60) for p in ps]
    assert res == [(0, True, 0, 5), (1, True, 5, 10)]


def test_merge_lora_state_dict_matches_explicit_lora_forward():
    from flow_factory_b200.weights import merge_lora_state_dict
    g = torch.Generator().manual_seed(0)
    W = torch.randn(24, 16, generator=g); b = torch.randn(24, generator=g)
    A = torch.randn(4, 16, generator=g) * 0.1; Bm = torch.randn(24, 4, generator=g) * 0.1
    x = torch.randn(5, 16, generator=g)
    sd = {"base_model.model.blk.to_q.base_layer.weight": W, "base_model.model.blk.to_q.base_layer.bias": b,
          "base_model.model.blk.to_q.lora_A.default.weight": A, "base_model.model.blk.to_q.lora_B.default.weight": Bm,
          "base_model.model.norm.weight": torch.ones(3)}
    m = merge_lora_state_dict(sd, lora_alpha=8.0)
    assert set(m) == {"blk.to_q.weight", "blk.to_q.bias", "norm.weight"}
    ref = torch.nn.functional.linear(x, W, b) + (x @ A.t() @ Bm.t()) * (8.0 / 4)      # peft: base(x) + scale * B(A(x))
    torch.testing.assert_close(torch.nn.functional.linear(x, m["blk.to_q.weight"], m["blk.to_q.bias"]), ref, rtol=1e-5, atol=1e-5)
    with pytest.raises(ValueError):
        merge_lora_state_dict({"a.lora_A.default.weight": A, "a.weight": W}, 8.0)


def _fsdp_worker(rank, world, port, q):
    import torch.distributed as dist
    from torch.distributed.device_mesh import init_device_mesh
    from torch.distributed.tensor import Shard, distribute_tensor
    from flow_factory_b200 import dist as D
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        mesh = init_device_mesh("cpu", (world,))
        g = torch.Generator().manual_seed(0)
        full = {"a.weight": torch.randn(7, 5, generator=g), "b.bias": torch.randn(3, generator=g), "c.weight": torch.randn(8, 2, 3, generator=g),
                "d.weight": torch.randn(1, 4, generator=g)}                     # 7 and 3 rows: padded tails; 1 row: rank 1 holds nothing
        sharded = {k: distribute_tensor(v, mesh, [Shard(0)]) for k, v in full.items()}
        sharded["plain"] = torch.arange(4.0)                                    # replicated entries pass through
        calls = []
        orig = dist.all_gather_into_tensor
        dist.all_gather_into_tensor = lambda *a, **k: (calls.append(1), orig(*a, **k))[1]
        out = D.gather_sharded_state_dict(sharded, dtype=torch.bfloat16)
        dist.all_gather_into_tensor = orig
        ok = len(calls) == 1 and set(out) == set(sharded)
        for k, v in full.items():
            ok &= out[k].dtype == torch.bfloat16 and torch.equal(out[k], v.bfloat16())
        ok &= torch.equal(out["plain"], torch.arange(4.0).bfloat16())
        out32 = D.gather_sharded_state_dict({"a.weight": sharded["a.weight"]}, dtype=None)
        ok &= torch.equal(out32["a.weight"], full["a.weight"])
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_fsdp2_sharded_weight_intake_one_allgather_world2_gloo():
    """SURVEY 8(e) / BASELINE config 5: DTensor Shard(0) parameters (what FSDP2's fully_shard holds) -> replicated bf16 tensors with ONE
    collective, including padded and empty tail shards."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_fsdp_worker, args=(r, 2, 29613, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=180) for _ in range(2))
    [p.join(timeout=60) for p in ps]
    assert res == [(0, True), (1, True)]


def test_callback_collector_and_stepwise_inference_with_stub_engine(monkeypatch):
    """`extra_call_back_kwargs` (GRPO-Guard: next_latents_mean, grpo.py:404) routes inference() through the reference's step loop over
    forward(); bookkeeping checked with a stub engine (the kernels behind forward() are the validated ones)."""
    from flow_factory_b200 import adapter as A
    from flow_factory_b200.trajectory import CallbackCollector, compute_trajectory_indices
    from flow_factory_b200.weights import EngineConfig

    class Plan:
        pass

    class Eng:
        def __init__(self, model_config, state_dict, device):
            self.device, self.steps = torch.device("cpu"), []
            self.cfg = EngineConfig(num_layers=1, num_heads=1, patch_size=2, in_channels=16, joint_attention_dim=8, pooled_projection_dim=8,
                                    pos_embed_max_size=8, num_dual_layers=0)

        def plan(self, *a):
            return Plan()

        def set_prompts(self, *a):
            pass

        def step(self, plan, latents, coef, guidance, noise=None, next_latents=None, seed=0, want_mean=True):
            self.steps.append((coef.sigma, coef.noise_level, bool(coef.compute_log_prob), want_mean))
            B = latents.shape[0]
            nxt = (latents.float() * 0.5).half()
            return dict(next_latents=nxt, next_latents_mean=latents.float() * 0.5 if want_mean else None,
                        log_prob=torch.full((B,), float(len(self.steps))) if coef.compute_log_prob else None,
                        noise_pred=torch.zeros_like(latents, dtype=torch.bfloat16), overflow=torch.zeros(1))

    monkeypatch.setattr(A, "RolloutEngine", Eng)
    sch = A.FlowMatchEulerDiscreteSDEScheduler(noise_level=0.7, shift=3.0, num_sde_steps=2, seed=3)
    ad = A.B200SD3_5Adapter(None, {}, device="cpu", scheduler=sch, rng="philox")
    ad.rollout()
    T = 6
    sch.set_timesteps(T, seq_len=16)
    sde = sorted(sch.current_sde_steps.tolist())
    idx = compute_trajectory_indices(sde, T)
    pe, pp = torch.zeros(2, 3, 8), torch.zeros(2, 8)
    out = ad.inference(prompt=["a", "b"], height=64, width=64, num_inference_steps=T, guidance_scale=1.0, prompt_embeds=pe, pooled_prompt_embeds=pp,
                       compute_log_prob=True, extra_call_back_kwargs=["next_latents_mean", "noise_level"], trajectory_indices=idx,
                       latents=torch.ones(2, 16, 8, 8))
    eng = ad.engine
    assert len(eng.steps) == T and [s[2] for s in eng.steps] == [i in sde for i in range(T)] and all(s[3] for s in eng.steps)
    s0 = out[0]
    n_lat = len(idx)
    gated = [i for i in range(T) if i in set(idx)]                         # callback steps use the same gate as the latents (step index)
    assert s0.all_latents.shape == (n_lat, 16, 8, 8) and s0.log_probs.shape == (len([i for i in sde if i in set(idx)]),)
    assert s0.next_latents_mean.shape == (len(gated), 16, 8, 8)            # extra_kwargs fall-through, (T', C, H, W) per sample
    cmap = s0.callback_index_map
    assert cmap.shape == (T,) and [int(i) for i in (cmap >= 0).nonzero().flatten()] == gated
    # latents halve every step in the stub: position p holds 0.5 ** p
    for pos in idx:
        assert float(s0.all_latents[int(s0.latent_index_map[pos])].float().mean()) == pytest.approx(0.5 ** pos, rel=1e-3)
    for step in gated:
        assert float(s0.next_latents_mean[int(cmap[step])].mean()) == pytest.approx(0.5 ** (step + 1), rel=1e-3)
    assert torch.equal(s0.final_latents, (torch.ones(16, 8, 8) * 0.5 ** T).half())
    c = CallbackCollector(None, 4)
    c.collect_step(0, object(), ["x"], {"x": 1})
    assert c.is_disabled and c.get_result() == {} and c.get_index_map() is None
    with pytest.raises(NotImplementedError):
        ad.inference(height=64, width=64, num_inference_steps=2, prompt_embeds=pe, pooled_prompt_embeds=pp, extra_call_back_kwargs=["prompt_embeds"])
