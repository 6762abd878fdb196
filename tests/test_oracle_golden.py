"""The oracle (oracle/sd3_oracle.py) pinned against fixtures minted from the REAL reference
(tests/golden/make_golden.py; Flow-Factory a0b2bc5 + vendored diffusers f7fd76a, CPU)."""
import os, json
import torch, pytest
from oracle import sd3_oracle as O


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), weights_only=False)


def test_schedule_matches_reference(golden_dir):
    g = _load(golden_dir, "schedule.pt")
    for T in (4, 30, 10):
        ts, sig = O.make_schedule(T, shift=3.0)
        assert torch.equal(ts, g[f"T{T}"]["timesteps"])
        assert torch.equal(sig, g[f"T{T}"]["sigmas"])
        assert O.current_sde_steps(T, None, None, 42) == g[f"T{T}"]["sde"].tolist()
    for seed in (0, 42, 43):
        for n in (1, 3):
            assert O.current_sde_steps(30, None, n, seed) == g[f"sde_seed{seed}_n{n}"].tolist()


@pytest.mark.parametrize("dyn", ["Flow-SDE", "Dance-SDE", "CPS", "ODE"])
def test_sde_step_bit_exact(golden_dir, dyn):
    g = _load(golden_dir, "sde_step.pt")
    _, sigmas = O.make_schedule(30, 3.0)
    for i in (0, 5, 28, 29):
        e = g[f"{dyn}_{i}"]
        r = O.sde_step(g["v"], g["x"], (e["t"] / 1000).item(), (e["tn"] / 1000).item(), e["nl"], float(sigmas[1]),
                       dyn, noise=e["noise"], compute_log_prob=True)
        assert torch.equal(r["next_latents_mean"], e["mean"]), (dyn, i)
        assert torch.equal(r["next_latents"], e["next_latents"]), (dyn, i)
        assert torch.equal(r["std_dev_t"].flatten()[:1], e["std"].flatten()[:1])
        if e["nl"] > 0 or dyn == "ODE":
            torch.testing.assert_close(r["log_prob"], e["log_prob"], rtol=1e-6, atol=1e-6)
        if "tf_next" in e:
            r2 = O.sde_step(g["v"], g["x"], (e["t"] / 1000).item(), (e["tn"] / 1000).item(), e["nl"],
                            float(sigmas[1]), dyn, next_latents=e["tf_next"], compute_log_prob=True)
            torch.testing.assert_close(r2["log_prob"], e["tf_log_prob"], rtol=1e-6, atol=1e-6)


def test_forward_tiny_fp32_and_autocast(golden_dir):
    g = _load(golden_dir, "forward_tiny.pt")["tiny"]
    cfg = O.tiny_config()
    w = O.make_weights(cfg, seed=0)
    assert sorted(w.keys()) == g["keys"]                      # key parity with SD3Transformer2DModel.state_dict()
    assert torch.equal(w["pos_embed.pos_embed"], g["pos_embed"])
    inp = O.make_inputs(cfg, batch=2, lat_h=16, lat_w=16, n_text=13, seed=1)
    with torch.no_grad():
        y = O.transformer_forward(w, cfg, inp["x0"], inp["prompt_embeds"], inp["pooled"], g["t"])
    torch.testing.assert_close(y, g["y32"], rtol=1e-5, atol=1e-5)
    wb = {k: v.bfloat16() for k, v in w.items()}
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        yb = O.transformer_forward(wb, cfg, inp["x0"].half(), inp["prompt_embeds"].bfloat16(),
                                   inp["pooled"].bfloat16(), g["t"].half())
    assert yb.dtype == g["y_bf16_cpu_autocast"].dtype
    assert torch.equal(yb, g["y_bf16_cpu_autocast"])          # same ops, same order -> bit exact on CPU


def test_forward_tiny3_context_pre_only_nonsquare(golden_dir):
    g = _load(golden_dir, "forward_tiny.pt")["tiny3"]
    cfg = O.tiny_config(num_layers=3, heads=3, dual=(0, 1), joint_dim=96, pooled_dim=48, pos_max=24, sample_size=32)
    w = O.make_weights(cfg, seed=5)
    inp = O.make_inputs(cfg, batch=1, lat_h=24, lat_w=16, n_text=21, seed=6)
    with torch.no_grad():
        y = O.transformer_forward(w, cfg, inp["x0"], inp["prompt_embeds"], inp["pooled"], g["t"])
    torch.testing.assert_close(y, g["y32"], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_rollout_loop_matches_reference(golden_dir, mode):
    g = _load(golden_dir, "rollout_tiny.pt")[mode]
    cfg = O.tiny_config()
    dt = torch.float32 if mode == "fp32" else torch.bfloat16
    w = {k: v.to(dt) for k, v in O.make_weights(cfg, seed=0).items()}
    inp = O.make_inputs(cfg, batch=2, lat_h=16, lat_w=16, n_text=13, seed=1)
    noises = O.make_noises(4, (2, 16, 16, 16), seed=123)
    with torch.no_grad():
        r = O.rollout(w, cfg, inp["x0"].to(dt), inp["prompt_embeds"].to(dt), inp["pooled"].to(dt),
                      inp["neg_prompt_embeds"].to(dt), inp["neg_pooled"].to(dt), 4, 4.5, noises=noises,
                      autocast="cpu" if mode == "bf16" else None)
    assert torch.equal(r["timesteps"], g["timesteps"])
    for a, b in zip(r["latents"], g["latents"]):
        if mode == "bf16":
            assert torch.equal(a, b)
        else:
            torch.testing.assert_close(a.float(), b.float(), rtol=2e-3, atol=2e-3)
    assert sorted(r["log_probs"]) == sorted(g["log_probs"])
    for i in r["log_probs"]:
        torch.testing.assert_close(r["log_probs"][i], g["log_probs"][i], rtol=1e-5, atol=1e-6)
