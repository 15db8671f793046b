"""GPU parity of the denoising loops (C-ABI kernels driven by acestep_b200.sampler) against
  (a) the golden outputs of the REAL reference samplers (fp32, tests/golden/*.npz), and
  (b) the fp32 oracle on identical seeds / noise.

Tolerance: the loop compounds bf16 rounding over steps x layers, so the bound is measured in the
test: an all-bf16 torch run of the oracle (what the reference executes on a GPU) is compared with
the fp32 oracle, and the CUDA result must be no further from fp32 than
max(1.5 x that spread, 3e-2) in rel-L2 on the final latents.
"""
import pytest
import torch

from helpers import golden, rel_l2

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("CUDA device required", allow_module_level=True)

from acestep_b200.dit import B200DiT, DiTShape  # noqa: E402
from acestep_b200.sampler import B200Sampler, prepare_noise  # noqa: E402
from oracle import sampler as osamp  # noqa: E402
from oracle.dit import CrossCache, DiTConfig, dit_forward  # noqa: E402
from oracle.weights import bf16_round_, make_dit_weights  # noqa: E402

DEV = torch.device("cuda:0")


@pytest.fixture(scope="module")
def env():
    cfg = DiTConfig.tiny()
    w = bf16_round_(make_dit_weights(cfg, seed=0))
    wb = {k: v.to(torch.bfloat16) for k, v in w.items()}
    dit = B200DiT(w, DiTShape.from_config(cfg), DEV)
    vel32 = lambda xt, t, ctx, enc, cache: dit_forward(w, cfg, xt, t, ctx, enc, cache, bf16_time=True)
    vel16 = lambda xt, t, ctx, enc, cache: dit_forward(wb, cfg, xt, t, ctx, enc, cache)
    return cfg, dit, vel32, vel16


def _b(x):
    return x.to(torch.bfloat16)


def _check(got, want32, run16, gold=None):
    got = got.cpu().float()
    floor = rel_l2(run16().float(), want32)
    tol = max(1.5 * floor, 3e-2)
    err = rel_l2(got, want32)
    assert torch.isfinite(got).all()
    assert err <= tol, (err, floor)
    if gold is not None:  # real reference fp32 output (fp32 weights): one more bf16 rounding away
        assert rel_l2(got, gold) <= tol + 2e-2, (rel_l2(got, gold), floor)


def test_turbo_ode_matches_reference(env):
    cfg, dit, vel32, vel16 = env
    g = golden("turbo_ode_shift3")
    seeds = [int(s) for s in g["seeds"]]
    noise = prepare_noise((2, 40, 64), seeds, "cpu", torch.float32)  # CPU generator == the golden's
    s = B200Sampler(dit)
    out = s.generate_turbo(_b(g["enc"]), _b(g["ctx"]), _b(g["src"]), seeds, shift=3.0, noise=_b(noise))
    for k in ("diffusion_time_cost", "diffusion_per_step_time_cost", "total_time_cost"):
        assert k in out["time_costs"]
    args = (_b(g["enc"]).float(), _b(g["ctx"]).float(), _b(g["src"]).float(), seeds)
    want = osamp.sample_turbo(vel32, *args, shift=3.0, new_cache=CrossCache, noise=_b(noise).float())
    run16 = lambda: osamp.sample_turbo(vel16, _b(g["enc"]), _b(g["ctx"]), _b(g["src"]), seeds, shift=3.0,
                                       new_cache=CrossCache, noise=_b(noise))
    _check(out["target_latents"], want, run16, g["out"])


def test_turbo_sde_cover(env):
    cfg, dit, vel32, vel16 = env
    g = golden("turbo_cover")
    gen = torch.Generator().manual_seed(99)
    noise = torch.randn(2, 40, 64, generator=gen)
    sde = [torch.randn(2, 40, 64, generator=gen) for _ in range(8)]
    kw = dict(shift=2.0, infer_method="sde", cover_noise_strength=0.4, audio_cover_strength=0.5)
    s = B200Sampler(dit)
    out = s.generate_turbo(_b(g["enc"]), _b(g["ctx"]), _b(g["src"]), 0, encoder_hidden_states_non_cover=_b(g["enc_nc"]),
                           context_latents_non_cover=_b(g["ctx_nc"]), noise=_b(noise), sde_noise=[_b(x) for x in sde], **kw)
    f = lambda x: _b(x).float()
    want = osamp.sample_turbo(vel32, f(g["enc"]), f(g["ctx"]), f(g["src"]), 0, enc_non_cover=f(g["enc_nc"]),
                              ctx_non_cover=f(g["ctx_nc"]), noise=f(noise), sde_noise=[f(x) for x in sde],
                              new_cache=CrossCache, **kw)
    run16 = lambda: osamp.sample_turbo(vel16, _b(g["enc"]), _b(g["ctx"]), _b(g["src"]), 0, enc_non_cover=_b(g["enc_nc"]),
                                       ctx_non_cover=_b(g["ctx_nc"]), noise=_b(noise), sde_noise=[_b(x) for x in sde],
                                       new_cache=CrossCache, **kw)
    _check(out["target_latents"], want, run16)


@pytest.mark.parametrize("name,kw", [
    ("base_apg_shift3", dict(infer_steps=6, guidance=7.0, shift=3.0)),
    ("base_apg_interval", dict(infer_steps=5, guidance=4.0, shift=1.0, cfg_interval_start=0.3, cfg_interval_end=0.85)),
    ("base_adg", dict(infer_steps=4, guidance=5.0, shift=2.0, use_adg=True)),
])
def test_base_cfg_matches_reference(env, name, kw):
    cfg, dit, vel32, vel16 = env
    g = golden(name)
    seed = [int(s) for s in g["seeds"]] if "seeds" in g else g["seed"]
    B = g["enc"].shape[0]
    noise = prepare_noise((B, 40, 64), seed, "cpu", torch.float32)
    kw = dict(kw)
    guidance = kw.pop("guidance")
    s = B200Sampler(dit, g["null_emb"])
    out = s.generate_base(_b(g["enc"]), _b(g["ctx"]), _b(g["src"]), seed, diffusion_guidance_sale=guidance,
                          noise=_b(noise), **kw)
    f = lambda x: _b(x).float()
    okw = dict(kw, guidance_scale=guidance)
    want = osamp.sample_base(vel32, f(g["enc"]), f(g["ctx"]), f(g["src"]), seed, null_emb=f(g["null_emb"]),
                             noise=f(noise), new_cache=CrossCache, **okw)
    run16 = lambda: osamp.sample_base(vel16, _b(g["enc"]), _b(g["ctx"]), _b(g["src"]), seed, null_emb=_b(g["null_emb"]),
                                      noise=_b(noise), new_cache=CrossCache, **okw)
    _check(out["target_latents"], want, run16, g["out"])


def test_sft_explicit_timesteps_matches_reference(env):
    """SFT variant: explicit `timesteps` replace the linspace/shift schedule and infer_steps
    (sft/modeling_acestep_v15_base.py:1864-1875); golden from the real sft module (tools/make_golden_sft.py)."""
    cfg, dit, vel32, vel16 = env
    g = golden("sft_timesteps")
    seed = [int(s) for s in g["seeds"]]
    noise = prepare_noise((2, 40, 64), seed, "cpu", torch.float32)
    ts = [float(x) for x in g["timesteps"]]
    kw = dict(infer_steps=99, shift=2.0, timesteps=ts)
    s = B200Sampler(dit, g["null_emb"])
    out = s.generate_base(_b(g["enc"]), _b(g["ctx"]), _b(g["src"]), seed, diffusion_guidance_sale=6.0, noise=_b(noise), **kw)
    f = lambda x: _b(x).float()
    want = osamp.sample_base(vel32, f(g["enc"]), f(g["ctx"]), f(g["src"]), seed, null_emb=f(g["null_emb"]),
                             guidance_scale=6.0, noise=f(noise), new_cache=CrossCache, **kw)
    run16 = lambda: osamp.sample_base(vel16, _b(g["enc"]), _b(g["ctx"]), _b(g["src"]), seed, null_emb=_b(g["null_emb"]),
                                      guidance_scale=6.0, noise=_b(noise), new_cache=CrossCache, **kw)
    _check(out["target_latents"], want, run16, g["out"])


def test_base_nocfg_sde_and_cover(env):
    cfg, dit, vel32, vel16 = env
    g = golden("base_cover")
    gen = torch.Generator().manual_seed(5)
    noise = torch.randn(2, 40, 64, generator=gen)
    sde = [torch.randn(2, 40, 64, generator=gen) for _ in range(6)]
    f = lambda x: _b(x).float()
    s = B200Sampler(dit, g["null_emb"])
    # (a) no CFG + SDE
    out = s.generate_base(_b(g["enc"]), _b(g["ctx"]), _b(g["src"]), 0, infer_steps=4, diffusion_guidance_sale=1.0,
                          infer_method="sde", noise=_b(noise), sde_noise=[_b(x) for x in sde])
    want = osamp.sample_base(vel32, f(g["enc"]), f(g["ctx"]), f(g["src"]), 0, null_emb=f(g["null_emb"]), infer_steps=4,
                             guidance_scale=1.0, infer_method="sde", noise=f(noise), sde_noise=[f(x) for x in sde],
                             new_cache=CrossCache)
    run16 = lambda: osamp.sample_base(vel16, _b(g["enc"]), _b(g["ctx"]), _b(g["src"]), 0, null_emb=_b(g["null_emb"]),
                                      infer_steps=4, guidance_scale=1.0, infer_method="sde", noise=_b(noise),
                                      sde_noise=[_b(x) for x in sde], new_cache=CrossCache)
    _check(out["target_latents"], want, run16)
    # (b) CFG + cover-noise start + cover->non-cover switch
    kw = dict(infer_steps=6, shift=3.0, cover_noise_strength=0.3, audio_cover_strength=0.5)
    out = s.generate_base(_b(g["enc"]), _b(g["ctx"]), _b(g["src"]), 0, diffusion_guidance_sale=3.0,
                          encoder_hidden_states_non_cover=_b(g["enc_nc"]), context_latents_non_cover=_b(g["ctx_nc"]),
                          noise=_b(noise), **kw)
    want = osamp.sample_base(vel32, f(g["enc"]), f(g["ctx"]), f(g["src"]), 0, null_emb=f(g["null_emb"]),
                             guidance_scale=3.0, enc_non_cover=f(g["enc_nc"]), ctx_non_cover=f(g["ctx_nc"]),
                             noise=f(noise), new_cache=CrossCache, **kw)
    run16 = lambda: osamp.sample_base(vel16, _b(g["enc"]), _b(g["ctx"]), _b(g["src"]), 0, null_emb=_b(g["null_emb"]),
                                      guidance_scale=3.0, enc_non_cover=_b(g["enc_nc"]), ctx_non_cover=_b(g["ctx_nc"]),
                                      noise=_b(noise), new_cache=CrossCache, **kw)
    _check(out["target_latents"], want, run16)


def test_validation_errors_match_reference_seam(env):
    """Same exception types as DiffusionMixin._mlx_run_diffusion (pinned by handler/diffusion_test.py)."""
    cfg, dit, _, _ = env
    s = B200Sampler(dit)
    enc, ctx, src = torch.zeros(2, 5, 256), torch.zeros(2, 10, 128), torch.zeros(2, 10, 64)
    with pytest.raises(ValueError):
        s.generate_turbo(enc, ctx, src, 0, infer_method="euler")
    with pytest.raises(TypeError):
        s.generate_turbo(enc, ctx, src, 0, timesteps=3)
    with pytest.raises(ValueError):
        s.generate_turbo(enc[:1], ctx, src, 0)
    with pytest.raises(ValueError):
        s.generate_base(enc, ctx, src[:1], 0)


def test_apg_adg_kernels_vs_reference_functions():
    """Kernel-level check against the golden outputs of apg_forward / adg_forward."""
    from acestep_b200 import _lib
    from oracle import guidance as og

    lib = _lib.load()
    g = golden("guidance")
    st = torch.cuda.current_stream().cuda_stream
    pc, pu, pc2, pu2 = (_b(g[k]).to(DEV) for k in ("pc", "pu", "pc2", "pu2"))
    mom = torch.zeros_like(pc)
    o1, o2 = torch.empty_like(pc), torch.empty_like(pc)
    _lib.check(lib.ace_apg(pc.data_ptr(), pu.data_ptr(), mom.data_ptr(), 1, -0.75, 2.5, 7.0, o1.data_ptr(), 2, 50, st))
    _lib.check(lib.ace_apg(pc2.data_ptr(), pu2.data_ptr(), mom.data_ptr(), 0, -0.75, 2.5, 7.0, o2.data_ptr(), 2, 50, st))
    torch.cuda.synchronize()
    m = og.Momentum()
    w1 = og.apg(pc.cpu().float(), pu.cpu().float(), 7.0, m)
    w2 = og.apg(pc2.cpu().float(), pu2.cpu().float(), 7.0, m)
    assert rel_l2(o1.cpu().float(), w1) <= 1e-2 and rel_l2(o2.cpu().float(), w2) <= 1e-2
    assert rel_l2(o1.cpu().float(), g["apg1"]) <= 2e-2 and rel_l2(o2.cpu().float(), g["apg2"]) <= 2e-2
    lat = _b(g["lat"]).to(DEV)
    od = torch.empty_like(lat)
    _lib.check(lib.ace_adg(lat.data_ptr(), pc[:1].contiguous().data_ptr(), pu[:1].contiguous().data_ptr(), 0.7, 5.0,
                           3.14 / 6, od.data_ptr(), 1, 50, st))
    torch.cuda.synchronize()
    wd = og.adg(lat.cpu().float(), pc[:1].cpu().float(), pu[:1].cpu().float(), 0.7, 5.0)
    assert rel_l2(od.cpu().float(), wd) <= 2e-2
