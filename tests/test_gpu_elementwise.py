"""-m gpu parity: HBM-bound kernels (fused Euler/SDE + log-prob, LayerNorm+modulate, skinny linear) vs the oracle."""
import os
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import sd3_oracle as O


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), weights_only=False)


@pytest.mark.parametrize("dyn", ["Flow-SDE", "Dance-SDE", "CPS", "ODE"])
def test_sde_step_vs_reference_golden(golden_dir, dyn):
    """scheduler.step (FF flow_match_euler_discrete.py:243-438): fixtures minted from the reference itself.
    Tolerances: mean <= 1e-6 rel, fp16-rounded next_latents exact given the same noise, log_prob <= 1e-5 rel."""
    from flow_factory_b200 import FlowMatchEulerDiscreteSDEScheduler
    g = _load(golden_dir, "sde_step.pt")
    s = FlowMatchEulerDiscreteSDEScheduler(noise_level=0.7, shift=3.0, dynamics_type=dyn)
    s.set_timesteps(30, seq_len=4096)
    x, v = g["x"].cuda(), g["v"].cuda()
    for i in (0, 5, 28, 29):
        e = g[f"{dyn}_{i}"]
        r = s.step(noise_pred=v, timestep=e["t"], latents=x, timestep_next=e["tn"], noise_level=e["nl"],
                   compute_log_prob=True, noise=e["noise"].cuda())
        torch.testing.assert_close(r.next_latents_mean.cpu(), e["mean"], rtol=1e-6, atol=1e-6)
        assert torch.equal(r.next_latents.cpu(), e["next_latents"]), (dyn, i)
        assert abs(float(r.std_dev_t.flatten()[0]) - float(e["std"].flatten()[0])) <= 1e-6 * max(1.0, abs(float(e["std"].flatten()[0])))
        if e["nl"] > 0 or dyn == "ODE":
            torch.testing.assert_close(r.log_prob.cpu(), e["log_prob"], rtol=1e-5, atol=1e-6)
        if "tf_next" in e:   # teacher-forced replay (GRPO optimize path, grpo.py:242-271)
            r2 = s.step(noise_pred=v, timestep=e["t"], latents=x, timestep_next=e["tn"], noise_level=e["nl"],
                        next_latents=e["tf_next"].cuda(), compute_log_prob=True)
            torch.testing.assert_close(r2.log_prob.cpu(), e["tf_log_prob"], rtol=1e-5, atol=1e-6)


def test_sde_step_full_size_properties():
    """BASELINE size (B=8, 16x128x128): in-kernel Philox noise is N(0,1); log-prob of the drawn sample equals the
    closed form -mean(eps_q^2)/2 - log(s) - log(sqrt(2pi)); ODE step is idempotent wrt noise."""
    from flow_factory_b200 import FlowMatchEulerDiscreteSDEScheduler
    torch.manual_seed(0)
    s = FlowMatchEulerDiscreteSDEScheduler(noise_level=0.7, shift=3.0)
    ts = s.set_timesteps(30, seq_len=4096)
    x = torch.randn(8, 16, 128, 128, device="cuda").half()
    v = torch.randn(8, 16, 128, 128, device="cuda").bfloat16()
    r = s.step(noise_pred=v, timestep=ts[3], latents=x, timestep_next=ts[4], noise_level=0.7, seed=1234, step_index=3)
    c = s.step_coef(ts[3], ts[4], 0.7)
    eps = (r.next_latents - r.next_latents_mean) / c.noise_scale
    assert abs(float(eps.mean())) < 5e-3 and abs(float(eps.std()) - 1.0) < 5e-3
    assert abs(float((eps ** 4).mean()) - 3.0) < 0.05          # Gaussian kurtosis
    lp = -((r.next_latents - r.next_latents_mean) ** 2).mean(dim=(1, 2, 3)) / c.two_var - c.log_norm
    torch.testing.assert_close(r.log_prob, lp, rtol=1e-5, atol=1e-6)
    r2 = s.step(noise_pred=v, timestep=ts[3], latents=x, timestep_next=ts[4], noise_level=0.7, seed=1234, step_index=3)
    assert torch.equal(r.next_latents, r2.next_latents) and torch.equal(r.log_prob, r2.log_prob)   # deterministic
    r3 = s.step(noise_pred=v, timestep=ts[3], latents=x, timestep_next=ts[4], noise_level=0.7, seed=1235, step_index=3)
    assert not torch.equal(r.next_latents, r3.next_latents)
    # oracle at full size with the kernel's own noise
    ro = O.sde_step(v.cpu(), x.cpu(), (ts[3] / 1000).item(), (ts[4] / 1000).item(), 0.7, float(s.sigmas[1]),
                    noise=eps.cpu())
    torch.testing.assert_close(r.next_latents_mean.cpu(), ro["next_latents_mean"], rtol=1e-6, atol=1e-6)


def test_ln_modulate_matches_torch():
    from tests.gpu_util import ptr, stream
    from flow_factory_b200 import _lib
    torch.manual_seed(1)
    for (B, R, D) in ((2, 77, 128), (2, 333, 1536), (1, 4096, 1536), (3, 50, 192)):
        x = (torch.randn(B, R, D, device="cuda") * 2 + 0.3).bfloat16()
        mod = (torch.randn(B, 4 * D, device="cuda") * 0.5).bfloat16()
        o1 = torch.empty_like(x); o2 = torch.empty_like(x)
        _lib.check(_lib.lib().ffb200_ln_modulate(ptr(x), B, R, D, 1e-6, ptr(mod), ptr(mod[:, D:]), ptr(o1), ptr(mod[:, 2 * D:]),
                                                 ptr(mod[:, 3 * D:]), ptr(o2), mod.stride(0), stream()))
        xn = torch.nn.functional.layer_norm(x.float(), (D,), None, None, 1e-6)
        for o, sh, sc in ((o1, mod[:, :D], mod[:, D:2 * D]), (o2, mod[:, 2 * D:3 * D], mod[:, 3 * D:])):
            ref = (xn * (1 + sc[:, None]).float() + sh[:, None].float()).bfloat16()     # CUDA-autocast semantics
            d = (o.float() - ref.float()).abs()
            assert float(d.max()) <= 2 ** -6 * max(1.0, float(ref.float().abs().max())), (B, R, D, float(d.max()))
            assert float((d > 0).float().mean()) < 0.02       # at most rare 1-ulp bf16 flips


def test_small_linear_matches_torch():
    from tests.gpu_util import ptr, stream
    from flow_factory_b200 import _lib
    torch.manual_seed(2)
    for (B, K, N, silu, add) in ((2, 256, 128, 0, 0), (16, 1536, 4000, 1, 1), (3, 2048, 1536, 0, 0), (9, 32, 128, 1, 0), (20, 3072, 1003, 1, 1), (5, 48, 64, 0, 0),
                                 (16, 96, 4099, 1, 0)):
        x = torch.randn(B, K, device="cuda").bfloat16()
        W = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
        b = (torch.randn(N, device="cuda") * 0.1).bfloat16()
        addend = torch.randn(B, N, device="cuda").bfloat16() if add else None
        out = torch.empty(B, N, device="cuda", dtype=torch.bfloat16)
        _lib.check(_lib.lib().ffb200_small_linear(ptr(x), B, K, K, ptr(W), ptr(b), N, ptr(out), N, ptr(addend), N, silu, stream()))
        xi = torch.nn.functional.silu(x) if silu else x
        ref = torch.nn.functional.linear(xi.float(), W.float(), b.float()).bfloat16()
        if add:
            ref = ref + addend
        torch.testing.assert_close(out.float(), ref.float(), rtol=2e-2, atol=2e-2)


def test_sde_step_default_noise_is_a_fresh_device_draw():
    """Called as the reference adapters call it (no generator / noise / seed, sd3_5.py:435-445): every call draws new fp32 noise from the
    device RNG (flow_match...py:350-357) - two calls differ, and the draw is torch.randn's stream; a generator list works per sample."""
    from flow_factory_b200 import FlowMatchEulerDiscreteSDEScheduler
    s = FlowMatchEulerDiscreteSDEScheduler(noise_level=0.7, shift=3.0)
    ts = s.set_timesteps(30, seq_len=4096)
    x = torch.randn(2, 16, 32, 32, device="cuda").half()
    v = torch.randn(2, 16, 32, 32, device="cuda").bfloat16()
    torch.manual_seed(11)
    r1 = s.step(noise_pred=v, timestep=ts[3], latents=x, timestep_next=ts[4], noise_level=0.7)
    r2 = s.step(noise_pred=v, timestep=ts[3], latents=x, timestep_next=ts[4], noise_level=0.7)
    assert not torch.equal(r1.next_latents, r2.next_latents)
    torch.manual_seed(11)
    z = torch.randn(v.shape, device="cuda", dtype=torch.float32)
    r3 = s.step(noise_pred=v, timestep=ts[3], latents=x, timestep_next=ts[4], noise_level=0.7, noise=z)
    assert torch.equal(r1.next_latents, r3.next_latents) and torch.equal(r1.log_prob, r3.log_prob)
    gens = [torch.Generator(device="cuda").manual_seed(5), torch.Generator(device="cuda").manual_seed(6)]
    r4 = s.step(noise_pred=v, timestep=ts[3], latents=x, timestep_next=ts[4], noise_level=0.7, generator=gens)
    z4 = torch.cat([torch.randn((1,) + v.shape[1:], device="cuda", generator=torch.Generator(device="cuda").manual_seed(sd)) for sd in (5, 6)])
    r5 = s.step(noise_pred=v, timestep=ts[3], latents=x, timestep_next=ts[4], noise_level=0.7, noise=z4)
    assert torch.equal(r4.next_latents, r5.next_latents)
    with pytest.raises(NotImplementedError):
        s.step(noise_pred=v, timestep=ts[3], latents=x.double(), timestep_next=ts[4], noise_level=0.7)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32, torch.float16])
@pytest.mark.parametrize("dyn", ["Flow-SDE", "Dance-SDE", "CPS"])
def test_sde_step_latent_storage_dtypes(dtype, dyn):
    """latent_storage_dtype fp16 / bf16 / fp32 (FF/hparams/training_args.py:245-252): the sampled next_latents is rounded through the INPUT
    latents dtype before its log-prob (flow_match...py:309, 359-362).  Against the oracle step with the same noise: next_latents bit for
    bit in that dtype, mean 1e-6, log-prob 1e-5; teacher-forced replay of the stored value gives the same log-prob."""
    from flow_factory_b200 import FlowMatchEulerDiscreteSDEScheduler
    s = FlowMatchEulerDiscreteSDEScheduler(noise_level=0.7, shift=3.0, dynamics_type=dyn)
    ts = s.set_timesteps(30, seq_len=4096)
    g = torch.Generator(device="cuda").manual_seed(17)
    x = torch.randn(2, 16, 24, 32, device="cuda", generator=g).to(dtype)
    v = torch.randn(2, 16, 24, 32, device="cuda", generator=g).bfloat16()
    z = torch.randn(2, 16, 24, 32, device="cuda", generator=g)
    r = s.step(noise_pred=v, timestep=ts[5], latents=x, timestep_next=ts[6], noise_level=0.7, noise=z)
    ro = O.sde_step(v, x, (ts[5] / 1000).item(), (ts[6] / 1000).item(), 0.7, float(s.sigmas[1]), dyn, noise=z)
    assert torch.equal(r.next_latents, ro["next_latents"])                         # fp32 views of the storage-rounded values
    assert torch.equal(r.next_latents, r.next_latents.to(dtype).float())           # representable in the storage dtype
    torch.testing.assert_close(r.next_latents_mean, ro["next_latents_mean"], rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(r.log_prob, ro["log_prob"], rtol=1e-5, atol=1e-6)
    stored = r.next_latents.to(dtype)
    r2 = s.step(noise_pred=v, timestep=ts[5], latents=x, timestep_next=ts[6], noise_level=0.7, next_latents=stored)
    torch.testing.assert_close(r2.log_prob, r.log_prob, rtol=1e-6, atol=1e-7)
