"""Generate golden fixtures from the REAL reference (imported read-only from /root/reference).

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py
Reference commits: Flow-Factory a0b2bc5, diffusers submodule f7fd76a.  torch 2.11.0 CPU.
Writes tests/golden/*.pt (small).  The oracle (oracle/sd3_oracle.py) and the CUDA engine are both
checked against these files by the test-suite; nothing here is imported by the product.
"""
import os, sys, json, math
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(HERE, "ref_stubs"), "/root/reference/diffusers/src", "/root/reference/src", ROOT]

import torch
from diffusers.models.transformers.transformer_sd3 import SD3Transformer2DModel
from flow_factory.scheduler import FlowMatchEulerDiscreteSDEScheduler, set_scheduler_timesteps
from flow_factory.utils.trajectory_collector import (compute_trajectory_indices, create_trajectory_collector)
from oracle import sd3_oracle as O

torch.set_num_threads(8)


def ref_model(cfg, weights, dtype):
    m = SD3Transformer2DModel(**cfg.ref_kwargs())
    missing, unexpected = m.load_state_dict(weights, strict=True), None
    return m.to(dtype).eval()


def golden_schedule():
    out = {}
    for T, seq in ((4, 256), (30, 4096), (10, 1024)):
        s = FlowMatchEulerDiscreteSDEScheduler(noise_level=0.7, shift=3.0, num_sde_steps=None, dynamics_type="Flow-SDE")
        ts = set_scheduler_timesteps(s, T, seq_len=seq, device="cpu")
        out[f"T{T}"] = dict(timesteps=ts.clone(), sigmas=s.sigmas.clone(), sde=s.current_sde_steps.clone())
    # dynamic shifting (FLUX-style: mu = calculate_shift(seq_len); diffusers time_shift 'exponential')
    for T, seq in ((28, 4096), (10, 1024)):
        s = FlowMatchEulerDiscreteSDEScheduler(noise_level=0.7, use_dynamic_shifting=True, base_shift=0.5, max_shift=1.15,
                                               base_image_seq_len=256, max_image_seq_len=4096)
        ts = set_scheduler_timesteps(s, T, seq_len=seq, device="cpu")
        out[f"dyn_T{T}"] = dict(timesteps=ts.clone(), sigmas=s.sigmas.clone())
    for seed in (0, 42, 43):
        for n in (1, 3):
            s = FlowMatchEulerDiscreteSDEScheduler(noise_level=0.7, shift=3.0, num_sde_steps=n, seed=seed)
            set_scheduler_timesteps(s, 30, seq_len=4096, device="cpu")
            out[f"sde_seed{seed}_n{n}"] = s.current_sde_steps.clone()
            out[f"noise_levels_seed{seed}_n{n}"] = s.get_noise_levels().clone()
    return out


def golden_step():
    """scheduler.step for the 4 dynamics, sampled + teacher-forced, incl. the sigma==1 first step."""
    out = {}
    g = torch.Generator().manual_seed(7)
    x = torch.randn(2, 16, 8, 8, generator=g).half()
    v = torch.randn(2, 16, 8, 8, generator=g).bfloat16()
    out["x"], out["v"] = x, v
    for dyn in ("Flow-SDE", "Dance-SDE", "CPS", "ODE"):
        s = FlowMatchEulerDiscreteSDEScheduler(noise_level=0.7, shift=3.0, num_sde_steps=None, dynamics_type=dyn)
        ts = set_scheduler_timesteps(s, 30, seq_len=4096, device="cpu")
        for i in (0, 5, 28, 29):
            t = ts[i]
            tn = ts[i + 1] if i + 1 < len(ts) else torch.tensor(0.0)
            nl = s.get_noise_level_for_timestep(t)
            torch.manual_seed(123 + i)
            r = s.step(noise_pred=v, timestep=t, latents=x, timestep_next=tn, noise_level=nl, compute_log_prob=True)
            noise = torch.randn(v.shape, generator=torch.Generator().manual_seed(123 + i), dtype=torch.float32)
            key = f"{dyn}_{i}"
            out[key] = dict(t=t.clone(), tn=tn.clone(), nl=float(nl), noise=noise,
                            next_latents=r.next_latents, mean=r.next_latents_mean, std=r.std_dev_t, dt=r.dt,
                            log_prob=r.log_prob)
            if dyn != "ODE" and nl > 0:  # teacher forced replay of a perturbed stored sample
                stored = (r.next_latents + 0.01).half()
                r2 = s.step(noise_pred=v, timestep=t, latents=x, timestep_next=tn, noise_level=nl,
                            next_latents=stored, compute_log_prob=True)
                out[key]["tf_next"] = stored
                out[key]["tf_log_prob"] = r2.log_prob
    return out


def golden_forward():
    """Tiny SD3.5-style transformer (dual layer 0, qk rms_norm): fp32 truth and bf16 CPU-autocast outputs."""
    out = {}
    cfg = O.tiny_config()
    w32 = O.make_weights(cfg, seed=0)
    inp = O.make_inputs(cfg, batch=2, lat_h=16, lat_w=16, n_text=13, seed=1)
    t = torch.tensor([988.5, 250.0])
    m32 = ref_model(cfg, w32, torch.float32)
    with torch.no_grad():
        y32 = m32(hidden_states=inp["x0"], timestep=t, encoder_hidden_states=inp["prompt_embeds"],
                  pooled_projections=inp["pooled"], return_dict=False)[0]
        mb = ref_model(cfg, w32, torch.bfloat16)
        with torch.autocast("cpu", dtype=torch.bfloat16):
            yb = mb(hidden_states=inp["x0"].half(), timestep=t.half(),
                    encoder_hidden_states=inp["prompt_embeds"].bfloat16(),
                    pooled_projections=inp["pooled"].bfloat16(), return_dict=False)[0]
    out["tiny"] = dict(t=t, y32=y32, y_bf16_cpu_autocast=yb, pos_embed=m32.pos_embed.pos_embed.clone(),
                       keys=sorted(m32.state_dict().keys()))
    # 3-layer config: 2 dual layers, last layer context_pre_only, non-square latent, 3 heads
    cfg2 = O.tiny_config(num_layers=3, heads=3, dual=(0, 1), joint_dim=96, pooled_dim=48, pos_max=24, sample_size=32)
    w2 = O.make_weights(cfg2, seed=5)
    inp2 = O.make_inputs(cfg2, batch=1, lat_h=24, lat_w=16, n_text=21, seed=6)
    t2 = torch.tensor([612.0])
    m2 = ref_model(cfg2, w2, torch.float32)
    with torch.no_grad():
        y2 = m2(hidden_states=inp2["x0"], timestep=t2, encoder_hidden_states=inp2["prompt_embeds"],
                pooled_projections=inp2["pooled"], return_dict=False)[0]
    out["tiny3"] = dict(t=t2, y32=y2)
    return out


def golden_rollout():
    """The loop body of SD3_5Adapter.inference/forward (sd3_5.py:266-304, 392-446) driven with the REAL
    reference transformer + scheduler: T=4, CFG on, Flow-SDE, fp16 latent storage, bf16 CPU autocast and fp32."""
    out = {}
    cfg = O.tiny_config()
    w32 = O.make_weights(cfg, seed=0)
    inp = O.make_inputs(cfg, batch=2, lat_h=16, lat_w=16, n_text=13, seed=1)
    T, gs = 4, 4.5
    for mode in ("fp32", "bf16"):
        dt = torch.float32 if mode == "fp32" else torch.bfloat16
        m = ref_model(cfg, w32, dt)
        s = FlowMatchEulerDiscreteSDEScheduler(noise_level=0.7, shift=3.0, num_sde_steps=None, dynamics_type="Flow-SDE")
        ts = set_scheduler_timesteps(s, T, seq_len=64, device="cpu")
        s.rollout()
        pe, pp = inp["prompt_embeds"].to(dt), inp["pooled"].to(dt)
        npe, npp = inp["neg_prompt_embeds"].to(dt), inp["neg_pooled"].to(dt)
        latents = inp["x0"].to(dt).half()                      # cast_latents -> fp16 storage
        lat_list, lps, vps = [latents], {}, []
        torch.manual_seed(123)
        with torch.no_grad():
            for i, t in enumerate(ts):
                nl = s.get_noise_level_for_timestep(t)
                tn = ts[i + 1] if i + 1 < len(ts) else torch.tensor(0.0)
                timestep = t.expand(2).to(latents.dtype)
                li = torch.cat([latents, latents]); ti = timestep.repeat(2)
                pei = torch.cat([npe, pe]); ppi = torch.cat([npp, pp])
                if mode == "bf16":
                    with torch.autocast("cpu", dtype=torch.bfloat16):
                        v = m(hidden_states=li, timestep=ti, encoder_hidden_states=pei, pooled_projections=ppi,
                              return_dict=False)[0]
                else:
                    v = m(hidden_states=li.float(), timestep=ti, encoder_hidden_states=pei, pooled_projections=ppi,
                          return_dict=False)[0]
                vu, vc = v.chunk(2)
                v = vu + gs * (vc - vu)
                clp = nl > 0
                r = s.step(noise_pred=v, timestep=t, latents=latents, timestep_next=tn, noise_level=nl,
                           compute_log_prob=clp)
                latents = r.next_latents.half()
                lat_list.append(latents); vps.append(v)
                if clp:
                    lps[i] = r.log_prob
        out[mode] = dict(latents=lat_list, log_probs=lps, noise_preds=vps, timesteps=ts.clone())
    return out


def golden_traj():
    cases = []
    for T in (4, 10, 30):
        for idx in ([0], [2, 5, 8], [0, 1, 2], [T - 2], list(range(T - 1)), [T - 1]):
            idx = [i for i in idx if i < T]
            for inc in (False, True):
                cases.append(dict(T=T, idx=idx, inc=inc, out=compute_trajectory_indices(idx, T, inc)))
    maps = []
    for T, ind in ((4, "all"), (4, None), (6, [0, -1]), (6, [1, 2, 5]), (30, [3, 4, 29, 30])):
        c = create_trajectory_collector(ind, T)
        for p in range(T + 1):
            c.collect(torch.tensor(float(p)), p)
        im = c.get_index_map()
        res = c.get_result()
        maps.append(dict(T=T, ind=ind, index_map=None if im is None else im.tolist(),
                         collected=None if res is None else [float(x) for x in res]))
    return dict(cases=cases, maps=maps)


def golden_advantage():
    """AdvantageProcessor.compute_weighted_sum / compute_gdpo (FF/advantage/advantage_processor.py:314-481) driven with a
    minimal fake self (no accelerator): the numpy fp64 group-normalisation math only."""
    import numpy as np
    from flow_factory.advantage.advantage_processor import AdvantageProcessor
    rng = np.random.default_rng(0)
    n_groups, k = 6, 4
    gid = np.repeat(np.arange(n_groups), k)
    perm = rng.permutation(len(gid))
    gid = gid[perm]
    rewards = {"pick": rng.normal(size=len(gid)), "ocr": rng.uniform(size=len(gid))}
    weights = {"pick": 1.0, "ocr": 0.5}
    out = dict(gid=gid, rewards=rewards, weights=weights)
    for global_std in (True, False):
        ap = AdvantageProcessor.__new__(AdvantageProcessor)
        ap.reward_weights, ap.global_std, ap.group_size, ap.group_on_same_rank = weights, global_std, k, False
        ap.collect_group_rewards = lambda samples, rw: ({kk: np.asarray(v, dtype=np.float64) for kk, v in rw.items()}, gid)
        ap._to_local = lambda a: a
        ap._build_weighted_sum_log_data = lambda *a, **kw: {}
        ap._build_gdpo_log_data = lambda *a, **kw: {}
        out[f"sum_global{int(global_std)}"] = np.asarray(ap.compute_weighted_sum([], rewards, False))
        out[f"gdpo_global{int(global_std)}"] = np.asarray(ap.compute_gdpo([], rewards, False))
    return out


def golden_flux():
    """FLUX.1: the REAL FluxTransformer2DModel on tiny configs (fp32 truth + bf16 CPU autocast), and the loop
    body of Flux1Adapter.inference/forward (FF/models/flux/flux1.py:211-250, 310-349) with the real scheduler (dynamic shift)."""
    from diffusers.models.transformers.transformer_flux import FluxTransformer2DModel
    from oracle import flux_oracle as FO
    out = {}
    for name, cfg, (B, lh, lw, nt), seed in (("tiny", FO.tiny_flux_config(), (2, 8, 8, 7), 0),
                                              ("tiny3", FO.tiny_flux_config(num_layers=2, num_single_layers=1, heads=3, joint_dim=96, pooled_dim=48), (1, 12, 8, 21), 5)):
        w32 = FO.make_flux_weights(cfg, seed=seed)
        lat, pe, pooled, img_ids, txt_ids = FO.make_flux_inputs(cfg, B, lh, lw, nt, seed=seed + 1)
        t = torch.tensor([0.9885, 0.25][:B])
        gd = torch.full((B,), 3.5)
        m = FluxTransformer2DModel(**cfg.ref_kwargs())
        m.load_state_dict(w32, strict=True)
        m = m.eval()
        with torch.no_grad():
            y32 = m(hidden_states=lat, timestep=t, guidance=gd, pooled_projections=pooled, encoder_hidden_states=pe,
                    txt_ids=txt_ids, img_ids=img_ids, return_dict=False)[0]
            mb = FluxTransformer2DModel(**cfg.ref_kwargs()); mb.load_state_dict(w32, strict=True); mb = mb.to(torch.bfloat16).eval()
            with torch.autocast("cpu", dtype=torch.bfloat16):
                yb = mb(hidden_states=lat.bfloat16(), timestep=t.bfloat16(), guidance=gd.bfloat16(), pooled_projections=pooled.bfloat16(),
                        encoder_hidden_states=pe.bfloat16(), txt_ids=txt_ids.bfloat16(), img_ids=img_ids.bfloat16(), return_dict=False)[0]
        out[name] = dict(t=t, guidance=gd, y32=y32, y_bf16_cpu_autocast=yb, keys=sorted(m.state_dict().keys()),
                         shape=(B, lh, lw, nt), seed=seed)
    # rollout: T=4, Flow-SDE, dynamic shift, packed fp32 latents round-tripped through the storage dtype like cast_latents
    cfg = FO.tiny_flux_config()
    w32 = FO.make_flux_weights(cfg, seed=0)
    lat, pe, pooled, img_ids, txt_ids = FO.make_flux_inputs(cfg, 2, 8, 8, 7, seed=1)
    m = FluxTransformer2DModel(**cfg.ref_kwargs()); m.load_state_dict(w32, strict=True); m = m.eval()
    s = FlowMatchEulerDiscreteSDEScheduler(noise_level=0.7, use_dynamic_shifting=True, base_shift=0.5, max_shift=1.15,
                                           base_image_seq_len=256, max_image_seq_len=4096, num_sde_steps=None, dynamics_type="Flow-SDE")
    ts = set_scheduler_timesteps(s, 4, seq_len=lat.shape[1], device="cpu")
    s.rollout()
    latents = lat.half()
    lats, lps, vps = [latents], {}, []
    torch.manual_seed(123)
    with torch.no_grad():
        for i, t in enumerate(ts):
            nl = s.get_noise_level_for_timestep(t)
            tn = ts[i + 1] if i + 1 < len(ts) else torch.tensor(0.0)
            guidance = torch.as_tensor(3.5, dtype=latents.dtype).expand(2)
            v = m(hidden_states=latents.float(), timestep=t.expand(2) / 1000, guidance=guidance.float(), pooled_projections=pooled,
                  encoder_hidden_states=pe, txt_ids=torch.zeros(pe.shape[1], 3), img_ids=img_ids, return_dict=False)[0]
            r = s.step(noise_pred=v, timestep=t, latents=latents, timestep_next=tn, noise_level=nl, compute_log_prob=nl > 0)
            latents = r.next_latents.half()
            lats.append(latents); vps.append(v)
            if nl > 0:
                lps[i] = r.log_prob
    out["rollout_fp32"] = dict(latents=lats, log_probs=lps, noise_preds=vps, timesteps=ts.clone(), sigmas=s.sigmas.clone())
    return out


def golden_qwen():
    """Qwen-Image (groundwork): the REAL QwenImageTransformer2DModel on a tiny config, fp32 + bf16 CPU autocast,
    with and without a text padding mask, and the norm-rescaled true-CFG combine of FF/models/qwen_image/qwen_image.py:580-587."""
    from diffusers.models.transformers.transformer_qwenimage import QwenImageTransformer2DModel
    from oracle import qwen_oracle as QO
    out = {}
    cfg = QO.tiny_qwen_config()
    w32 = QO.make_qwen_weights(cfg, seed=0)
    B, h2, w2, nt = 2, 6, 4, 9
    lat, pe = QO.make_qwen_inputs(cfg, B, h2, w2, nt, seed=1)
    t = torch.tensor([0.9885, 0.25])
    m = QwenImageTransformer2DModel(**cfg.ref_kwargs()); m.load_state_dict(w32, strict=True); m = m.eval()
    mask = torch.ones(B, nt); mask[1, 6:] = 0
    with torch.no_grad():
        y32 = m(hidden_states=lat, timestep=t, encoder_hidden_states=pe, encoder_hidden_states_mask=None, img_shapes=[[(1, h2, w2)]] * B, return_dict=False)[0]
        y32m = m(hidden_states=lat, timestep=t, encoder_hidden_states=pe, encoder_hidden_states_mask=mask, img_shapes=[[(1, h2, w2)]] * B, return_dict=False)[0]
        mb = QwenImageTransformer2DModel(**cfg.ref_kwargs()); mb.load_state_dict(w32, strict=True); mb = mb.to(torch.bfloat16).eval()
        with torch.autocast("cpu", dtype=torch.bfloat16):
            yb = mb(hidden_states=lat.bfloat16(), timestep=t.bfloat16(), encoder_hidden_states=pe.bfloat16(), encoder_hidden_states_mask=None,
                    img_shapes=[[(1, h2, w2)]] * B, return_dict=False)[0]
    out["tiny"] = dict(t=t, y32=y32, y32_masked=y32m, mask=mask, y_bf16_cpu_autocast=yb, keys=sorted(m.state_dict().keys()), shape=(B, h2, w2, nt))
    # true CFG with norm rescale (qwen_image.py:580-587) on the two predictions above
    comb = y32m + 4.0 * (y32 - y32m)
    out["cfg"] = dict(gs=4.0, pred=comb * (torch.norm(y32, dim=-1, keepdim=True) / torch.norm(comb, dim=-1, keepdim=True)))
    return out


def golden_wan():
    """Wan2.1 T2V (groundwork): the REAL WanTransformer3DModel on a tiny config, fp32 + bf16 CPU autocast."""
    from diffusers.models.transformers.transformer_wan import WanTransformer3DModel
    from oracle import wan_oracle as WO
    cfg = WO.tiny_wan_config()
    w32 = WO.make_wan_weights(cfg, seed=0)
    B, fr, lh, lw, nt = 2, 3, 8, 12, 11
    lat, pe = WO.make_wan_inputs(cfg, B, fr, lh, lw, nt, seed=1)
    t = torch.tensor([988.5, 250.0])
    m = WanTransformer3DModel(**cfg.ref_kwargs()); m.load_state_dict(w32, strict=True); m = m.eval()
    with torch.no_grad():
        y32 = m(hidden_states=lat, timestep=t, encoder_hidden_states=pe, return_dict=False)[0]
        mb = WanTransformer3DModel(**cfg.ref_kwargs()); mb.load_state_dict(w32, strict=True); mb = mb.to(torch.bfloat16).eval()
        with torch.autocast("cpu", dtype=torch.bfloat16):
            yb = mb(hidden_states=lat.bfloat16(), timestep=t, encoder_hidden_states=pe.bfloat16(), return_dict=False)[0]
    return {"tiny": dict(t=t, y32=y32, y_bf16_cpu_autocast=yb, keys=sorted(m.state_dict().keys()), shape=(B, fr, lh, lw, nt))}


def golden_vae():
    """VAE decode (groundwork): the REAL AutoencoderKL.decode on a tiny decoder, through the latent rescale of
    SD3_5Adapter.decode_latents (sd3_5.py:166-169)."""
    from diffusers.models.autoencoders.autoencoder_kl import AutoencoderKL
    from oracle import vae_oracle as VO
    cfg = VO.tiny_vae_config()
    w = VO.make_vae_decoder_weights(cfg, seed=0)
    m = AutoencoderKL(**cfg.ref_kwargs())
    missing, unexpected = m.load_state_dict(w, strict=False)
    assert not unexpected and all(k.startswith("encoder.") for k in missing), (missing, unexpected)
    m = m.eval()
    lat = torch.randn(2, cfg.latent_channels, 6, 10, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        z = (lat / m.config.scaling_factor) + m.config.shift_factor
        img = m.decode(z, return_dict=False)[0]
    dec_keys = sorted(k for k in m.state_dict().keys() if k.startswith("decoder."))
    mb = AutoencoderKL(**cfg.ref_kwargs())
    mb.load_state_dict(w, strict=False)
    mb = mb.eval().to(torch.bfloat16)                       # the trainer's frozen-VAE dtype under mixed_precision bf16 (FF/models/abc.py:826-853)
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        lb = lat.to(torch.bfloat16)
        img_bf16 = mb.decode((lb / mb.config.scaling_factor) + mb.config.shift_factor, return_dict=False)[0]
    return {"tiny": dict(lat=lat, img=img, keys=dec_keys, img_bf16=img_bf16)}


def golden_wan_schedule():
    """UniPCMultistepSDEScheduler (FF/scheduler/unipc_multistep.py; diffusers use_flow_sigmas branch): integer timesteps, fp32 sigmas, SDE
    step selection, and scheduler.step through the `timestep_next` path Wan2_T2V_Adapter uses (sigma = int timestep / 1000)."""
    from flow_factory.scheduler import UniPCMultistepSDEScheduler
    out = {}
    for T, shift in ((10, 3.0), (50, 3.0), (20, 5.0)):
        s = UniPCMultistepSDEScheduler(noise_level=0.7, num_sde_steps=2, seed=5, dynamics_type="Flow-SDE", prediction_type="flow_prediction",
                                       use_flow_sigmas=True, flow_shift=shift, num_train_timesteps=1000)
        s.set_timesteps(T, device="cpu")
        out[f"T{T}_s{shift}"] = dict(timesteps=s.timesteps.clone(), sigmas=s.sigmas.clone(), sde=s.current_sde_steps.clone(),
                                      noise_levels=s.get_noise_levels().clone())
    g = torch.Generator().manual_seed(11)
    x = torch.randn(2, 16, 2, 4, 4, generator=g).half()
    v = torch.randn(2, 16, 2, 4, 4, generator=g).bfloat16()
    out["x"], out["v"] = x, v
    for dyn in ("Flow-SDE", "Dance-SDE", "CPS", "ODE"):
        s = UniPCMultistepSDEScheduler(noise_level=0.7, dynamics_type=dyn, prediction_type="flow_prediction", use_flow_sigmas=True,
                                       flow_shift=3.0, num_train_timesteps=1000)
        s.set_timesteps(10, device="cpu")
        s.rollout()
        for i in (0, 4, 9):
            t = s.timesteps[i]
            tn = s.timesteps[i + 1] if i + 1 < len(s.timesteps) else torch.tensor(0)
            torch.manual_seed(100 + i)
            o = s.step(noise_pred=v, timestep=t, latents=x, timestep_next=tn, noise_level=0.7, compute_log_prob=True, return_dict=True)
            torch.manual_seed(100 + i)
            noise = torch.randn(v.shape, dtype=torch.float32)
            out[f"{dyn}_{i}"] = dict(t=t.clone(), tn=tn.clone(), next_latents=o.next_latents.clone(), mean=o.next_latents_mean.clone(),
                                     log_prob=None if o.log_prob is None else o.log_prob.clone(), std_dev_t=o.std_dev_t.clone(),
                                     dt=o.dt.clone(), noise=noise, sigma_max=float(s.sigmas[1]))
    return out


if __name__ == "__main__":
    torch.save(golden_wan_schedule(), os.path.join(HERE, "wan_schedule.pt"))
    torch.save(golden_vae(), os.path.join(HERE, "vae_tiny.pt"))
    torch.save(golden_wan(), os.path.join(HERE, "wan_tiny.pt"))
    torch.save(golden_qwen(), os.path.join(HERE, "qwen_tiny.pt"))
    torch.save(golden_flux(), os.path.join(HERE, "flux_tiny.pt"))
    torch.save(golden_schedule(), os.path.join(HERE, "schedule.pt"))
    torch.save(golden_step(), os.path.join(HERE, "sde_step.pt"))
    torch.save(golden_forward(), os.path.join(HERE, "forward_tiny.pt"))
    torch.save(golden_rollout(), os.path.join(HERE, "rollout_tiny.pt"))
    torch.save(golden_advantage(), os.path.join(HERE, "advantage.pt"))
    with open(os.path.join(HERE, "trajectory.json"), "w") as f:
        json.dump(golden_traj(), f)
    print("golden written:", sorted(os.listdir(HERE)))
