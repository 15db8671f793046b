"""GPU parity of the standalone pipeline's repaint path (BASELINE config 5: VAE encode of the source
audio -> src latents / chunk mask -> base sampler with CFG + APG -> VAE decode) against the fp32 oracle
chain on identical audio, posterior noise, starting noise and weights.

Tolerances: the source latents come out of a bf16 codec pass (bound: the measured spread of an
all-bf16 torch run of the oracle encoder, floor 2e-2, like tests/test_gpu_kernels.py); the denoised
latents and the waveform use the bounds of __graft_entry__.smoke() (3e-2 / 2e-2 rel-L2, low-gain codec init), with the
oracle loop started from the CUDA path's own source latents so the comparison isolates the loop.
"""
import pytest
import torch

from helpers import rel_l2

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("CUDA device required", allow_module_level=True)

from acestep_b200.dit import DiTShape  # noqa: E402
from acestep_b200.pack import folded_vae_state  # noqa: E402
from acestep_b200.pipeline import B200Pipeline  # noqa: E402
from acestep_b200.vae import VaeShape  # noqa: E402
from oracle import sampler as osamp  # noqa: E402
from oracle import vae as ovae  # noqa: E402
from oracle.dit import CrossCache, DiTConfig, dit_forward  # noqa: E402
from oracle.weights import bf16_round_, make_dit_weights, make_null_condition_emb, make_vae_weights  # noqa: E402

DEV = torch.device("cuda:0")


def test_repaint_matches_oracle_chain():
    cfg, vcfg = DiTConfig.tiny(), ovae.VaeConfig.tiny()
    w = bf16_round_(make_dit_weights(cfg, seed=0))
    vsd = make_vae_weights(vcfg, seed=3, gain=0.5)  # low-gain init: tight codec bound
    wf = folded_vae_state(vsd)
    null = make_null_condition_emb(cfg).to(torch.bfloat16)
    vshape = VaeShape(encoder_hidden_size=vcfg.encoder_hidden_size, downsampling_ratios=vcfg.downsampling_ratios,
                      channel_multiples=vcfg.channel_multiples, decoder_channels=vcfg.decoder_channels)
    pipe = B200Pipeline(w, vsd, DiTShape.from_config(cfg), vshape, null, DEV, turbo=False)
    g = torch.Generator().manual_seed(11)
    B, T, E = 2, 48, 20
    s0, s1 = 12, 36
    audio = torch.rand(B, 2, T * vcfg.hop, generator=g) - 0.5
    eps = torch.randn(B, T, 64, generator=g).to(torch.bfloat16)
    sil = torch.randn(1, T, 64, generator=g).to(torch.bfloat16)
    enc = torch.randn(B, E, cfg.hidden_size, generator=g).to(torch.bfloat16)
    noise = torch.randn(B, T, 64, generator=g).to(torch.bfloat16)
    out = pipe.repaint(enc, audio, s0, s1, sil, None, posterior_eps=eps, noise=noise, infer_steps=3,
                       diffusion_guidance_sale=5.0, shift=3.0)
    src_lat = out["src_latents"].cpu().float()  # [B, T, 64]
    lat, wav = out["target_latents"].cpu().float(), out["audio"].cpu()
    assert wav.shape == (B, 2, T * vcfg.hop) and torch.isfinite(wav).all()
    for k in ("vae_encode_time_cost", "diffusion_time_cost", "vae_decode_time_cost", "total_time_cost"):
        assert k in out["time_costs"]

    # (1) source latents vs the oracle encoder + posterior sample
    a16 = audio.to(torch.bfloat16).float()
    mean, scale = ovae.encode_moments(wf, vcfg, a16)
    want_src = (mean + (torch.nn.functional.softplus(scale) + 1e-4) * eps.float().transpose(1, 2)).transpose(1, 2)
    wb = {k: v.to(torch.bfloat16) for k, v in wf.items()}
    floor = rel_l2(ovae.encode_moments(wb, vcfg, audio.to(torch.bfloat16))[0].float(), mean)
    assert rel_l2(src_lat, want_src) <= max(1.1 * floor, 2e-2), floor

    # (2) the loop, started from the CUDA path's own source latents: silence inside [s0, s1), mask 1 inside
    src = out["src_latents"].clone().cpu()
    src[:, s0:s1] = sil[:, s0:s1]
    mask = torch.zeros(B, T, 64, dtype=torch.bfloat16)
    mask[:, s0:s1] = 1.0
    ctx = torch.cat([src, mask], -1)
    vel = lambda xt, t, c, e, cache: dit_forward(w, cfg, xt, t, c, e, cache, bf16_time=True)
    want = osamp.sample_base(vel, enc.float(), ctx.float(), src.float(), None, null_emb=null.float(), infer_steps=3,
                             guidance_scale=5.0, shift=3.0, noise=noise.float(), new_cache=CrossCache)
    err = rel_l2(lat, want)
    assert err < 3e-2, err

    # (3) decode + per-sample peak normalisation (generate_music_decode.py:191-195)
    want_wav = ovae.decode(wf, vcfg, out["target_latents"].cpu().float().transpose(1, 2))
    peak = want_wav.abs().amax(dim=[1, 2], keepdim=True).clamp(min=1.0)
    werr = rel_l2(wav, want_wav / peak)
    assert werr < 2e-2, werr
    pipe.close()


def test_encode_seam_keeps_latents_on_device_and_caches_posterior_moments():
    """SURVEY §8f row 2 through `install()`: the wrapped `tiled_encode` is called the way the reference's callers
    call it (`infer_refer_latent`, handler/conditioning_embed.py:52-62: tiled_encode(audio, offload_latent_to_cpu=True)
    then .to(device).to(dtype), transpose) on a host whose VAE is the real engine.  (a) The result lives on the
    device although the caller asked for the CPU round trip; (b) a clip seen before skips the encoder — the
    launch counter shows only the posterior kernel — and (c) with the same torch seed the cached path returns
    exactly what a fresh encode returns (posterior sample = mean + std * eps with eps from the same generator
    state), i.e. the cache changes neither the distribution nor the RNG consumption."""
    from acestep_b200 import _lib
    from acestep_b200.backend import install
    from acestep_b200.vae import B200Vae

    vcfg = ovae.VaeConfig.tiny()
    vsd = make_vae_weights(vcfg, seed=3)
    vshape = VaeShape(encoder_hidden_size=vcfg.encoder_hidden_size, downsampling_ratios=vcfg.downsampling_ratios,
                      channel_multiples=vcfg.channel_multiples, decoder_channels=vcfg.decoder_channels)

    class Host:
        device, dtype = DEV, torch.bfloat16

        def _execute_service_generate_diffusion(self, *a, **k):
            raise AssertionError("not used")

        def tiled_decode(self, *a, **k):
            raise AssertionError("reference decode path must not run")

        def tiled_encode(self, audio, chunk_size=None, overlap=None, offload_latent_to_cpu=True):
            raise AssertionError("reference encode path must not run")

    h = install(Host())
    h.b200_vae, h.use_b200_vae = B200Vae(vsd, vshape, DEV), True
    lib = _lib.load()
    g = torch.Generator().manual_seed(5)
    ref_a = torch.rand(2, vcfg.hop * 40, generator=g) - 0.5   # CPU tensors, like loaded reference audio
    ref_b = torch.rand(2, vcfg.hop * 40, generator=g) - 0.5

    def call(audio):
        z = h.tiled_encode(audio, offload_latent_to_cpu=True)  # [64, T] for 2-D input
        assert z.is_cuda and z.shape == (64, 40)
        return z.to(h.device).to(h.dtype).transpose(0, 1)

    torch.manual_seed(123)
    n0 = lib.ace_launch_count()
    z1 = call(ref_a)
    n1 = lib.ace_launch_count()
    torch.manual_seed(123)
    z2 = call(ref_a)           # same clip, same seed: cache hit, identical sample
    n2 = lib.ace_launch_count()
    z3 = call(ref_a)           # same clip, generator moved on: a different posterior sample of the same moments
    zb = call(ref_b)           # different clip: miss
    torch.cuda.synchronize()
    assert (n1 - n0) > 5 and (n2 - n1) == 1, (n1 - n0, n2 - n1)  # encoder stack vs the posterior kernel alone
    assert h.b200_vae.moment_cache_hits == 2 and h.b200_vae.moment_cache_misses == 2
    assert torch.equal(z1, z2)
    assert not torch.equal(z1, z3) and not torch.equal(z1, zb)
    torch.manual_seed(123)
    fresh = h.b200_vae.encode(ref_a.unsqueeze(0), sample=True)[0].transpose(0, 1)  # uncached path, same seed
    assert torch.equal(z1, fresh)
    h.b200_vae.close()


def test_init_backends_and_cover_branch_of_prepare_condition_through_the_seam():
    """`_init_b200_backends` on a handler-shaped host whose model exposes decoder / encoder / tokenizer /
    detokenizer `state_dict()`s (the weight source the reference's MLX converters use too), then
    `_b200_prepare_condition` with a mixed is_covers batch: item 0's source latents are replaced by the LM hints
    (audio tokenizer -> FSQ -> detokenizer on the device, turbo :1630-1646), item 1 keeps its own; the context is
    [src | chunk_mask].  Checked against the oracle chain (oracle.cond + oracle.tokenizer) on bf16-rounded weights;
    bounds as in test_gpu_tokenizer.py / test_gpu_kernels.py (FSQ codes may flip next to a rounding boundary, so
    the hint rows are compared per 5-frame token and >= 80 % of the tokens must agree within 4e-2)."""
    import types

    from acestep_b200.backend import install
    from oracle import cond as ocond
    from oracle import tokenizer as otok
    from oracle.weights import bf16_round_ as r16

    tcfg = otok.TokConfig.tiny()
    ccfg = ocond.CondConfig.tiny()
    dcfg = DiTConfig.tiny()
    w_dit = r16(make_dit_weights(dcfg, seed=0))
    w_cond = r16(ocond.make_cond_weights(ccfg, seed=5))
    w_tok = r16(otok.make_tokenizer_weights(tcfg, seed=9))

    class _SD(torch.nn.Module):            # a module whose state_dict() is the oracle's weight dict
        def __init__(self, d):
            super().__init__()
            self._d = d

        def state_dict(self, *a, **k):
            return self._d

    sd = _SD
    strip = lambda p: {k[len(p):]: v for k, v in w_tok.items() if k.startswith(p)}
    cfg = types.SimpleNamespace(**{**vars(tcfg), **vars(ccfg), **vars(dcfg), "is_turbo": False})
    cfg.layer_types = dcfg.layer_types

    class Host:
        device, dtype = DEV, torch.bfloat16

        def _execute_service_generate_diffusion(self, *a, **k):
            raise AssertionError("not used")

        def tiled_decode(self, *a, **k):
            raise AssertionError("not used")

        def tiled_encode(self, *a, **k):
            raise AssertionError("not used")

    h = install(Host())
    h.model = types.SimpleNamespace(decoder=sd(w_dit), encoder=sd(w_cond), tokenizer=sd(strip("tokenizer.")),
                                    detokenizer=sd(strip("detokenizer.")), config=cfg,
                                    null_condition_emb=make_null_condition_emb(dcfg))
    h.config = cfg
    dit_status, vae_status = h._init_b200_backends(dit=True, vae=False, cond=True)
    assert "condition encoder" in dit_status and vae_status == "Disabled"
    assert h.use_b200_dit and h.use_b200_cond and h.b200_tok is not None

    g = torch.Generator().manual_seed(91)
    B, T = 2, 23
    text = torch.randn(B, 6, ccfg.text_hidden_dim, generator=g).to(torch.bfloat16)
    lyric = torch.randn(B, 12, ccfg.text_hidden_dim, generator=g).to(torch.bfloat16)
    refer = torch.randn(B, 10, ccfg.timbre_hidden_dim, generator=g).to(torch.bfloat16)
    tm, lm = torch.ones(B, 6, dtype=torch.long), torch.ones(B, 12, dtype=torch.long)
    order = torch.tensor([0, 1])
    hidden = torch.randn(B, T, 64, generator=g).to(torch.bfloat16)
    src = torch.randn(B, T, 64, generator=g).to(torch.bfloat16)
    silence = torch.randn(1, 40, 64, generator=g).to(torch.bfloat16)
    chunk = torch.ones(B, T, 64)
    is_covers = torch.tensor([1, 0])
    enc, enc_mask, ctx = h._b200_prepare_condition(
        text_hidden_states=text.to(DEV), text_attention_mask=tm.to(DEV), lyric_hidden_states=lyric.to(DEV),
        lyric_attention_mask=lm.to(DEV), refer_audio_acoustic_hidden_states_packed=refer.to(DEV),
        refer_audio_order_mask=order.to(DEV), hidden_states=hidden.to(DEV), attention_mask=torch.ones(B, T, device=DEV),
        silence_latent=silence.to(DEV), src_latents=src.to(DEV), chunk_masks=chunk.to(DEV), is_covers=is_covers.to(DEV))
    torch.cuda.synchronize()
    assert ctx.shape == (B, T, 128) and enc.dtype == torch.bfloat16
    want_enc, want_mask = ocond.condition_encoder(w_cond, ccfg, text.float(), tm, lyric.float(), lm, refer.float(), order)
    assert torch.equal(enc_mask.cpu(), want_mask)
    assert rel_l2(enc.cpu().float(), want_enc) <= 3e-2
    want_src = otok.lm_hints(w_tok, tcfg, hidden.float(), silence.float(), torch.ones(B, T), src.float(), is_covers)
    got_src = ctx[..., :64].cpu().float()
    assert torch.equal(ctx[..., 64:].cpu().float(), chunk)
    assert torch.equal(got_src[1], src[1].float())               # not a cover: untouched
    assert not torch.equal(got_src[0], src[0].float())           # cover: replaced by the hints
    tok_err = [rel_l2(got_src[0, 5 * k: 5 * k + 5], want_src[0, 5 * k: 5 * k + 5]) for k in range(4)]
    assert sum(e <= 4e-2 for e in tok_err) >= 3, tok_err
    for eng in (h.b200_dit, h.b200_cond, h.b200_tok):
        eng.close()


def test_async_pipeline_matches_blocking_calls_and_defers_the_latent_guard():
    """`generate_async` + `SongPipeline(depth=1)` (no host synchronisation inside a song; the next song is enqueued
    while the previous one runs) must hand out exactly what the blocking `generate` returns for the same seeds — the
    static I/O slots, the timestep cache and the two alternating pinned buffers are reused in stream order — and the
    reference's latent guard (generate_music_decode.py:66-77) must still fire, at `wait()`, before any audio is
    handed out."""
    from acestep_b200.output import NAN_MESSAGE
    from acestep_b200.pipeline import SongPipeline

    cfg, vcfg = DiTConfig.tiny(), ovae.VaeConfig.tiny()
    w = bf16_round_(make_dit_weights(cfg, seed=0))
    vsd = make_vae_weights(vcfg, seed=3)
    null = make_null_condition_emb(cfg).to(torch.bfloat16)
    vshape = VaeShape(encoder_hidden_size=vcfg.encoder_hidden_size, downsampling_ratios=vcfg.downsampling_ratios,
                      channel_multiples=vcfg.channel_multiples, decoder_channels=vcfg.decoder_channels)
    pipe = B200Pipeline(w, vsd, DiTShape.from_config(cfg), vshape, null, DEV, turbo=False)
    g = torch.Generator().manual_seed(21)
    T, E = 40, 12
    enc = torch.randn(1, E, cfg.hidden_size, generator=g).to(torch.bfloat16).pin_memory()
    src = torch.randn(1, T, 64, generator=g).to(torch.bfloat16)
    ctx = torch.cat([src, torch.ones(1, T, 64, dtype=torch.bfloat16)], -1).pin_memory()
    kw = dict(infer_steps=4, diffusion_guidance_sale=4.0, shift=3.0)
    want = [pipe.generate(enc, ctx, src, [s], **kw) for s in (5, 6, 7)]
    want = [{k: o[k].clone() for k in ("audio", "target_latents", "peak")} for o in want]
    q, got = SongPipeline(depth=1), []
    for s in (5, 6, 7):
        out = q.submit(pipe.generate_async(enc, ctx, src, [s], reuse_host_buffer=True, **kw))
        if out is not None:
            got.append({k: out[k].clone() for k in ("audio", "target_latents", "peak")})
    out = q.drain()
    got.append({k: out[k].clone() for k in ("audio", "target_latents", "peak")})
    assert len(got) == 3
    for a, b in zip(got, want):
        for k in a:
            assert torch.equal(a[k].cpu(), b[k].cpu()), k
    assert not torch.equal(got[0]["audio"], got[1]["audio"])  # different seeds: different songs
    for k in ("diffusion_time_cost", "diffusion_per_step_time_cost", "vae_decode_time_cost", "total_time_cost"):
        assert out["time_costs"][k] > 0.0

    bad = enc.clone()
    bad[0, 0, 0] = float("nan")
    pending = pipe.generate_async(bad, ctx, src, [5], **kw)   # enqueues fine
    with pytest.raises(RuntimeError, match=NAN_MESSAGE[:20]):
        pending.wait()
    assert torch.isfinite(pipe.generate(enc, ctx, src, [5], **kw)["audio"]).all()  # the engine is still usable
    pipe.close()


def test_two_stream_pipeline_matches_single_stream():
    """`overlap_codec=True` (loop on a high-priority stream, codec + D2H on a second one; off by default, see
    DESIGN.md) must return exactly what the single-stream pipeline returns, songs queued back to back."""
    from acestep_b200.pipeline import SongPipeline

    cfg, vcfg = DiTConfig.tiny(), ovae.VaeConfig.tiny()
    w = bf16_round_(make_dit_weights(cfg, seed=0))
    vsd = make_vae_weights(vcfg, seed=3)
    null = make_null_condition_emb(cfg).to(torch.bfloat16)
    vshape = VaeShape(encoder_hidden_size=vcfg.encoder_hidden_size, downsampling_ratios=vcfg.downsampling_ratios,
                      channel_multiples=vcfg.channel_multiples, decoder_channels=vcfg.decoder_channels)
    g = torch.Generator().manual_seed(31)
    T, E = 56, 9
    enc = torch.randn(1, E, cfg.hidden_size, generator=g).to(torch.bfloat16)
    src = torch.randn(1, T, 64, generator=g).to(torch.bfloat16)
    ctx = torch.cat([src, torch.ones(1, T, 64, dtype=torch.bfloat16)], -1)
    audio = torch.rand(1, 2, T * vcfg.hop, generator=g) - 0.5
    eps = torch.randn(1, T, 64, generator=g).to(torch.bfloat16)
    kw = dict(infer_steps=3, diffusion_guidance_sale=4.0, shift=3.0)
    results = []
    for overlap in (False, True):
        pipe = B200Pipeline(w, vsd, DiTShape.from_config(cfg), vshape, null, DEV, turbo=False, overlap_codec=overlap)
        q, outs = SongPipeline(depth=1), []
        for s in (1, 2, 3):
            o = q.submit(pipe.generate_async(enc, ctx, src, [s], **kw))
            if o is not None:
                outs.append(o["audio"].clone())
        outs.append(q.drain()["audio"].clone())
        o = pipe.repaint(enc, audio, 10, 30, src, [4], posterior_eps=eps, **kw)
        outs += [o["audio"].clone(), o["src_latents"].cpu().float()]
        results.append(outs)
        pipe.close()
    for a, b in zip(*results):
        assert torch.equal(a, b)


def test_wrapped_selection_sites_end_to_end_on_real_engines():
    """The drop-in boundary exercised the way the reference calls it, with REAL engines on the device (the CPU
    plumbing tests use stubs): `install()` on a handler-shaped host, `_init_b200_backends` from the modules'
    state dicts, then
      * `_execute_service_generate_diffusion(payload, generate_kwargs, seed_param, infer_method, shift,
        audio_cover_strength)` (handler/service_generate_execute.py:107-196): condition encoder -> context ->
        base sampler with CFG + APG -> (outputs, enc, enc_mask, ctx);
      * `tiled_decode(latents)` (handler/vae_decode.py:16-48) and `tiled_encode(audio)` (vae_encode.py:15-43).
    Checked against the oracle chain on bf16-rounded weights: condition encoder 3e-2, loop 3e-2 (started from the
    CUDA path's own conditioning so the comparison isolates the loop), waveform 2e-2 on the low-gain codec init (bounds of smoke())."""
    import contextlib
    import types

    from acestep_b200.backend import install
    from acestep_b200.sampler import prepare_noise
    from oracle import cond as ocond

    dcfg, ccfg, vcfg = DiTConfig.tiny(), ocond.CondConfig.tiny(), ovae.VaeConfig.tiny()
    w_dit = bf16_round_(make_dit_weights(dcfg, seed=0))
    w_cond = bf16_round_(ocond.make_cond_weights(ccfg, seed=5))
    vsd = make_vae_weights(vcfg, seed=3, gain=0.5)  # low-gain init: tight codec bound
    wf = folded_vae_state(vsd)
    null = make_null_condition_emb(dcfg).to(torch.bfloat16)

    class _SD(torch.nn.Module):
        def __init__(self, d, config=None):
            super().__init__()
            self._d, self.config = d, config

        def state_dict(self, *a, **k):
            return self._d

    cfg = types.SimpleNamespace(**{**vars(ccfg), **vars(dcfg), "is_turbo": False})
    cfg.layer_types = dcfg.layer_types

    class Host:
        device, dtype = DEV, torch.bfloat16

        def _execute_service_generate_diffusion(self, *a, **k):
            raise AssertionError("the reference path must not run while the B200 backend is active")

        def tiled_decode(self, *a, **k):
            raise AssertionError("the reference path must not run while the B200 backend is active")

        def tiled_encode(self, *a, **k):
            raise AssertionError("the reference path must not run while the B200 backend is active")

        @contextlib.contextmanager
        def _load_model_context(self, name):
            yield

    h = install(Host())
    h.model = types.SimpleNamespace(decoder=_SD(w_dit), encoder=_SD(w_cond), config=cfg, null_condition_emb=null,
                                    generate_audio=lambda **kw: None)  # plain base model: no `timesteps` parameter
    h.config = cfg
    h.vae = _SD(vsd, config=vcfg)
    dit_status, vae_status = h._init_b200_backends()
    assert dit_status.startswith("Active") and vae_status.startswith("Active")

    g = torch.Generator().manual_seed(77)
    B, T = 2, 36
    text = torch.randn(B, 5, ccfg.text_hidden_dim, generator=g).to(torch.bfloat16)
    lyric = torch.randn(B, 9, ccfg.text_hidden_dim, generator=g).to(torch.bfloat16)
    refer = torch.randn(B, 8, ccfg.timbre_hidden_dim, generator=g).to(torch.bfloat16)
    tm, lm = torch.ones(B, 5, dtype=torch.long), torch.ones(B, 9, dtype=torch.long)
    order = torch.tensor([0, 1])
    src = torch.randn(B, T, 64, generator=g).to(torch.bfloat16)
    chunk = torch.ones(B, T, 64, dtype=torch.bfloat16)
    h.silence_latent = torch.randn(1, 64, 64, generator=g).to(torch.bfloat16).to(DEV)
    dv = lambda x: x.to(DEV)
    payload = dict(src_latents=dv(src), text_hidden_states=dv(text), text_attention_mask=dv(tm),
                   lyric_hidden_states=dv(lyric), lyric_attention_mask=dv(lm),
                   refer_audio_acoustic_hidden_states_packed=dv(refer), refer_audio_order_mask=dv(order),
                   chunk_mask=dv(chunk), is_covers=torch.zeros(B, dtype=torch.long, device=DEV),
                   precomputed_lm_hints_25Hz=None, non_cover_text_hidden_states=None,
                   non_cover_text_attention_masks=None)
    gk = dict(infer_steps=4, diffusion_guidance_sale=5.0, cfg_interval_start=0.0, cfg_interval_end=1.0,
              use_adg=False, cover_noise_strength=0.0, timesteps=[1.0, 0.5, 0.0])  # ignored by the plain base model
    seeds = [11, 12]
    outputs, enc, enc_mask, ctx = h._execute_service_generate_diffusion(payload, gk, seeds, "ode", 3.0, 1.0)
    lat = outputs["target_latents"]
    assert lat.shape == (B, T, 64) and lat.dtype == torch.bfloat16 and lat.device.type == "cuda"
    assert "diffusion_time_cost" in outputs["time_costs"]
    want_enc, want_mask = ocond.condition_encoder(w_cond, ccfg, text.float(), tm, lyric.float(), lm, refer.float(), order)
    assert torch.equal(enc_mask.cpu(), want_mask)
    assert rel_l2(enc.cpu().float(), want_enc) <= 3e-2
    assert torch.equal(ctx.cpu(), torch.cat([src, chunk], -1))
    noise = prepare_noise((B, T, 64), seeds, DEV).cpu().float()  # the same RNG calls the seam makes
    vel = lambda xt, t, c, e, cache: dit_forward(w_dit, dcfg, xt, t, c, e, cache, bf16_time=True)
    want = osamp.sample_base(vel, enc.cpu().float(), ctx.cpu().float(), src.float(), None, null_emb=null.float(),
                             infer_steps=4, guidance_scale=5.0, shift=3.0, noise=noise, new_cache=CrossCache)
    err = rel_l2(lat.cpu().float(), want)
    assert err < 3e-2, err

    wav = h.tiled_decode(lat.transpose(1, 2).contiguous())  # [B, 64, T] like the reference's caller
    assert wav.shape == (B, 2, T * vcfg.hop) and wav.device.type == "cuda"
    want_wav = ovae.decode(wf, vcfg, lat.cpu().float().transpose(1, 2))
    peak = want_wav.abs().amax(dim=[1, 2], keepdim=True).clamp(min=1.0)
    assert rel_l2(wav.cpu().float(), want_wav / peak) < 2e-2
    z = h.tiled_encode(wav[0].float(), offload_latent_to_cpu=True)  # 2-D input -> 2-D result, stays on the device
    assert z.shape == (64, T) and z.device.type == "cuda"
    for eng in (h.b200_dit, h.b200_cond, h.b200_vae):
        eng.close()
