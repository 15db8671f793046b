"""GPU parity of the standalone pipeline's repaint path (BASELINE config 5: VAE encode of the source
audio -> src latents / chunk mask -> base sampler with CFG + APG -> VAE decode) against the fp32 oracle
chain on identical audio, posterior noise, starting noise and weights.

Tolerances: the source latents come out of a bf16 codec pass (bound: the measured spread of an
all-bf16 torch run of the oracle encoder, floor 2e-2, like tests/test_gpu_kernels.py); the denoised
latents and the waveform use the bounds of __graft_entry__.smoke() (3e-2 / 1e-1 rel-L2), with the
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
    vsd = make_vae_weights(vcfg, seed=3)
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
    assert werr < 1e-1, werr
    pipe.close()
