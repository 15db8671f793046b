#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REAL reference code (imported from /root/reference).

Runs only in the build container (the GPU box has no /root/reference); the vectors it writes are
committed so every test can check the oracle (and through it the CUDA path) without the reference.

What is executed from the reference, unmodified:
  * AceStepDiTModel (turbo modeling :1237-1504) with a tiny config and oracle.weights tensors,
  * AceStepConditionGenerationModel.generate_audio of the turbo AND base model files, driven
    through a stand-in `self` whose prepare_condition returns the synthetic conditioning (the
    same cut the handler's backend seam makes, handler/diffusion.py:18-33),
  * apg_forward / adg_forward / MomentumBuffer (base/apg_guidance.py),
  * the handler tiling mixins VaeDecodeChunksMixin / VaeEncodeMixin driven with the oracle VAE.
`vector_quantize_pytorch` (absent) is stubbed; it is never called on these paths.

Usage: python tools/make_golden.py   (writes tests/golden/)
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

stub = types.ModuleType("vector_quantize_pytorch")


class _ResidualFSQ(torch.nn.Module):
    def __init__(self, *a, **k):
        super().__init__()


stub.ResidualFSQ = _ResidualFSQ
sys.modules["vector_quantize_pytorch"] = stub

from acestep.models.turbo import modeling_acestep_v15_turbo as T  # noqa: E402
from acestep.models.base import modeling_acestep_v15_base as B  # noqa: E402
from acestep.models.base import apg_guidance as G  # noqa: E402
from acestep.models.turbo.configuration_acestep_v15 import AceStepConfig  # noqa: E402

from oracle.dit import DiTConfig  # noqa: E402
from oracle.weights import make_dit_weights, make_null_condition_emb, make_vae_weights  # noqa: E402
from oracle import vae as ovae  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
torch.set_grad_enabled(False)


def ref_decoder(cfg: DiTConfig, weights, mod=T):
    rc = AceStepConfig(hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                       num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
                       num_key_value_heads=cfg.num_key_value_heads, head_dim=cfg.head_dim,
                       sliding_window=cfg.sliding_window, rope_theta=cfg.rope_theta,
                       rms_norm_eps=cfg.rms_norm_eps, in_channels=cfg.in_channels,
                       audio_acoustic_hidden_dim=cfg.audio_acoustic_hidden_dim, patch_size=cfg.patch_size)
    rc._attn_implementation = "sdpa"
    m = mod.AceStepDiTModel(rc).float().eval()
    missing, unexpected = m.load_state_dict(weights, strict=False)
    assert not unexpected, unexpected
    assert all("rotary_emb" in k for k in missing), missing
    return m


def save(name, **arrs):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **{k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v))
                                 for k, v in arrs.items()})
    print("wrote", path, {k: tuple(np.asarray(v).shape) for k, v in arrs.items()})


def synth(cfg, B_, T_, E_, seed):
    g = torch.Generator().manual_seed(seed)
    enc = torch.randn(B_, E_, cfg.hidden_size, generator=g)
    src = torch.randn(B_, T_, 64, generator=g)
    ctx = torch.cat([src, torch.ones(B_, T_, 64)], dim=-1)
    xt = torch.randn(B_, T_, 64, generator=g)
    return enc, src, ctx, xt


class FakeSelf:
    """Stand-in for AceStepConditionGenerationModel: real loop code, injected conditioning."""

    def __init__(self, mod, decoder, null_emb, cond, cond_nc=None):
        cls = mod.AceStepConditionGenerationModel
        self.decoder = decoder
        self.null_condition_emb = null_emb
        self._cond = [cond, cond_nc]
        self._calls = 0
        self.prepare_noise = types.MethodType(cls.prepare_noise, self)
        self.get_x0_from_noise = types.MethodType(cls.get_x0_from_noise, self)
        self.renoise = types.MethodType(cls.renoise, self)
        self._gen = cls.generate_audio

    def prepare_condition(self, **kw):
        enc, ctx = self._cond[min(self._calls, 1)]
        self._calls += 1
        return enc, torch.ones(enc.shape[:2]), ctx

    def generate(self, src, **kw):
        z = torch.zeros(1)
        out = self._gen(self, text_hidden_states=z, text_attention_mask=z, lyric_hidden_states=z,
                        lyric_attention_mask=z, refer_audio_acoustic_hidden_states_packed=z,
                        refer_audio_order_mask=z, src_latents=src, chunk_masks=z,
                        is_covers=torch.zeros(src.shape[0]), silence_latent=torch.zeros(1, src.shape[1], 64),
                        **kw)
        return out["target_latents"]


def main():
    cfg = DiTConfig.tiny()
    w = make_dit_weights(cfg, seed=0)
    null_emb = make_null_condition_emb(cfg, seed=1)
    dec_t = ref_decoder(cfg, w, T)
    dec_b = ref_decoder(cfg, w, B)

    # ---- 1. single DiT forward (odd T exercises the patch pad/crop, S=19 > window 8) ----
    enc, src, ctx, xt = synth(cfg, 2, 37, 21, seed=10)
    t = torch.tensor([0.9, 0.35])
    vt = dec_t(hidden_states=xt, timestep=t, timestep_r=t, attention_mask=None,
               encoder_hidden_states=enc, encoder_attention_mask=None, context_latents=ctx,
               use_cache=False)[0]
    save("dit_forward_tiny", xt=xt, t=t, ctx=ctx, enc=enc, vt=vt)

    # ---- 2. turbo sampler: ODE shift 3, custom timesteps + SDE, cover-noise start ----
    enc, src, ctx, _ = synth(cfg, 2, 40, 17, seed=11)
    enc2, src2, ctx2, _ = synth(cfg, 2, 40, 17, seed=12)
    fs = FakeSelf(T, dec_t, null_emb, (enc, ctx))
    out = fs.generate(src, seed=[5, 6], shift=3.0, infer_method="ode")
    save("turbo_ode_shift3", enc=enc, ctx=ctx, src=src, seeds=[5, 6], out=out)
    fs = FakeSelf(T, dec_t, null_emb, (enc, ctx))
    torch.manual_seed(77)
    out = fs.generate(src, seed=7, shift=1.0, infer_method="sde",
                      timesteps=torch.tensor([0.97, 0.8, 0.52, 0.31, 0.0]))
    save("turbo_sde_custom", enc=enc, ctx=ctx, src=src, seed=7, timesteps=[0.97, 0.8, 0.52, 0.31, 0.0],
         rng_seed=77, out=out)
    fs = FakeSelf(T, dec_t, null_emb, (enc, ctx), (enc2, ctx2))
    out = fs.generate(src, seed=8, shift=2.0, infer_method="ode", cover_noise_strength=0.4,
                      audio_cover_strength=0.5, non_cover_text_hidden_states=torch.zeros(1),
                      non_cover_text_attention_mask=torch.zeros(1))
    save("turbo_cover", enc=enc, ctx=ctx, src=src, enc_nc=enc2, ctx_nc=ctx2, seed=8, out=out)

    # ---- 3. base sampler: CFG + APG, CFG interval, ADG (batch 1), no-CFG, SDE ----
    fs = FakeSelf(B, dec_b, null_emb, (enc, ctx))
    out = fs.generate(src, seed=[3, 4], infer_steps=6, diffusion_guidance_sale=7.0, shift=3.0,
                      use_progress_bar=False)
    save("base_apg_shift3", enc=enc, ctx=ctx, src=src, seeds=[3, 4], null_emb=null_emb, out=out)
    fs = FakeSelf(B, dec_b, null_emb, (enc, ctx))
    out = fs.generate(src, seed=9, infer_steps=5, diffusion_guidance_sale=4.0, shift=1.0,
                      cfg_interval_start=0.3, cfg_interval_end=0.85, use_progress_bar=False)
    save("base_apg_interval", enc=enc, ctx=ctx, src=src, seed=9, null_emb=null_emb, out=out)
    fs = FakeSelf(B, dec_b, null_emb, (enc[:1], ctx[:1]))
    out = fs.generate(src[:1], seed=2, infer_steps=4, diffusion_guidance_sale=5.0, shift=2.0,
                      use_adg=True, use_progress_bar=False)
    save("base_adg", enc=enc[:1], ctx=ctx[:1], src=src[:1], seed=2, null_emb=null_emb, out=out)
    fs = FakeSelf(B, dec_b, null_emb, (enc, ctx))
    torch.manual_seed(78)
    out = fs.generate(src, seed=1, infer_steps=4, diffusion_guidance_sale=1.0, shift=1.0,
                      infer_method="sde", use_progress_bar=False)
    save("base_nocfg_sde", enc=enc, ctx=ctx, src=src, seed=1, rng_seed=78, null_emb=null_emb, out=out)
    fs = FakeSelf(B, dec_b, null_emb, (enc, ctx), (enc2, ctx2))
    out = fs.generate(src, seed=11, infer_steps=6, diffusion_guidance_sale=3.0, shift=3.0,
                      cover_noise_strength=0.3, audio_cover_strength=0.5, use_progress_bar=False,
                      non_cover_text_hidden_states=torch.zeros(1), non_cover_text_attention_mask=torch.zeros(1))
    save("base_cover", enc=enc, ctx=ctx, src=src, enc_nc=enc2, ctx_nc=ctx2, seed=11, null_emb=null_emb, out=out)

    # ---- 4. guidance functions in isolation (two consecutive APG steps share the momentum) ----
    g = torch.Generator().manual_seed(20)
    pc, pu, pc2, pu2 = (torch.randn(2, 50, 64, generator=g) for _ in range(4))
    mb = G.MomentumBuffer()
    a1 = G.apg_forward(pc, pu, 7.0, mb, dims=[1])
    a2 = G.apg_forward(pc2, pu2, 7.0, mb, dims=[1])
    lat = torch.randn(1, 50, 64, generator=g)
    d1 = G.adg_forward(lat, pc[:1], pu[:1], torch.tensor(0.7), 5.0)
    save("guidance", pc=pc, pu=pu, pc2=pc2, pu2=pu2, apg1=a1, apg2=a2, lat=lat, adg=d1)

    # ---- 5. handler tiling glue, reference mixins over the oracle VAE ----
    from acestep.core.generation.handler.vae_decode_chunks import VaeDecodeChunksMixin
    from acestep.core.generation.handler.vae_encode import VaeEncodeMixin
    from acestep.core.generation.handler.vae_encode_chunks import VaeEncodeChunksMixin

    vcfg = ovae.VaeConfig.tiny()
    vw = make_vae_weights(vcfg, seed=3)

    class _Out:
        def __init__(self, s):
            self.sample = s

    class _Dist:
        def __init__(self, mean, scale, eps_fn):
            self.mean, self.scale, self.eps_fn = mean, scale, eps_fn

        def sample(self):
            return self.mean + (torch.nn.functional.softplus(self.scale) + 1e-4) * self.eps_fn(self.mean.shape)

    class _Enc:
        def __init__(self, d):
            self.latent_dist = d

    class OracleVae:
        dtype = torch.float32

        def decode(self, z):
            return _Out(ovae.decode(vw, vcfg, z))

        def encode(self, a):
            m, s = ovae.encode_moments(vw, vcfg, a)
            return _Enc(_Dist(m, s, lambda shp: torch.zeros(shp)))

    class Host(VaeDecodeChunksMixin, VaeEncodeMixin, VaeEncodeChunksMixin):
        vae = OracleVae()
        device = "cpu"
        disable_tqdm = True
        use_mlx_vae = False
        mlx_vae = None

        def _empty_cache(self):
            pass

    g = torch.Generator().manual_seed(30)
    z = torch.randn(2, 64, 90, generator=g)
    wav = Host()._tiled_decode_inner(z, 40, 8, False)
    save("vae_tiled_decode", z=z, chunk=40, overlap=8, wav=wav)
    audio = torch.rand(1, 2, 8 * 700, generator=g) - 0.5
    lat = Host().tiled_encode(audio, chunk_size=8 * 300, overlap=8 * 40, offload_latent_to_cpu=False)
    save("vae_tiled_encode", audio=audio, chunk=8 * 300, overlap=8 * 40, lat=lat)


if __name__ == "__main__":
    main()
