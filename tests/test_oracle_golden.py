"""Pins the oracle: every oracle restatement must reproduce the outputs of the REAL reference
modules captured in tests/golden/ by tools/make_golden.py (fp32, tiny config)."""
import functools

import pytest
import torch

from helpers import golden, max_abs, rel_l2
from oracle import guidance as og
from oracle import sampler as osamp
from oracle import vae as ovae
from oracle.dit import CrossCache, DiTConfig, dit_forward
from oracle.weights import make_dit_weights, make_vae_weights

TOL = 2e-5  # fp32 vs fp32, different op order (sdpa vs explicit softmax)


@pytest.fixture(scope="module")
def dit():
    cfg = DiTConfig.tiny()
    w = make_dit_weights(cfg, seed=0)
    vel = lambda xt, t, ctx, enc, cache: dit_forward(w, cfg, xt, t, ctx, enc, cache)
    return cfg, w, vel


def test_dit_forward_matches_reference(dit):
    cfg, w, _ = dit
    g = golden("dit_forward_tiny")
    vt = dit_forward(w, cfg, g["xt"], g["t"], g["ctx"], g["enc"])
    assert vt.shape == g["vt"].shape
    assert rel_l2(vt, g["vt"]) < TOL
    # the cross-KV cache must not change the result
    cache = CrossCache()
    dit_forward(w, cfg, g["xt"], g["t"], g["ctx"], g["enc"], cache)
    vt2 = dit_forward(w, cfg, g["xt"], g["t"], g["ctx"], g["enc"], cache)
    assert max_abs(vt2, vt) == 0.0


def test_turbo_ode(dit):
    _, _, vel = dit
    g = golden("turbo_ode_shift3")
    out = osamp.sample_turbo(vel, g["enc"], g["ctx"], g["src"], [int(s) for s in g["seeds"]], shift=3.0,
                             new_cache=CrossCache)
    assert rel_l2(out, g["out"]) < TOL


def test_turbo_sde_custom_timesteps(dit):
    _, _, vel = dit
    g = golden("turbo_sde_custom")
    torch.manual_seed(g["rng_seed"])
    out = osamp.sample_turbo(vel, g["enc"], g["ctx"], g["src"], g["seed"], shift=1.0, infer_method="sde",
                             timesteps=g["timesteps"].tolist(), new_cache=CrossCache)
    assert rel_l2(out, g["out"]) < TOL


def test_turbo_cover(dit):
    _, _, vel = dit
    g = golden("turbo_cover")
    out = osamp.sample_turbo(vel, g["enc"], g["ctx"], g["src"], g["seed"], shift=2.0,
                             cover_noise_strength=0.4, audio_cover_strength=0.5,
                             enc_non_cover=g["enc_nc"], ctx_non_cover=g["ctx_nc"], new_cache=CrossCache)
    assert rel_l2(out, g["out"]) < TOL


def test_base_apg(dit):
    _, _, vel = dit
    g = golden("base_apg_shift3")
    out = osamp.sample_base(vel, g["enc"], g["ctx"], g["src"], [int(s) for s in g["seeds"]],
                            null_emb=g["null_emb"], infer_steps=6, guidance_scale=7.0, shift=3.0,
                            new_cache=CrossCache)
    assert rel_l2(out, g["out"]) < TOL


def test_base_apg_interval(dit):
    _, _, vel = dit
    g = golden("base_apg_interval")
    out = osamp.sample_base(vel, g["enc"], g["ctx"], g["src"], g["seed"], null_emb=g["null_emb"],
                            infer_steps=5, guidance_scale=4.0, shift=1.0, cfg_interval_start=0.3,
                            cfg_interval_end=0.85, new_cache=CrossCache)
    assert rel_l2(out, g["out"]) < TOL


def test_base_adg(dit):
    _, _, vel = dit
    g = golden("base_adg")
    out = osamp.sample_base(vel, g["enc"], g["ctx"], g["src"], g["seed"], null_emb=g["null_emb"],
                            infer_steps=4, guidance_scale=5.0, shift=2.0, use_adg=True, new_cache=CrossCache)
    assert rel_l2(out, g["out"]) < 1e-4  # acos/sin chain amplifies fp32 noise


def test_base_nocfg_sde(dit):
    _, _, vel = dit
    g = golden("base_nocfg_sde")
    torch.manual_seed(g["rng_seed"])
    out = osamp.sample_base(vel, g["enc"], g["ctx"], g["src"], g["seed"], null_emb=g["null_emb"],
                            infer_steps=4, guidance_scale=1.0, shift=1.0, infer_method="sde",
                            new_cache=CrossCache)
    assert rel_l2(out, g["out"]) < TOL


def test_base_cover(dit):
    _, _, vel = dit
    g = golden("base_cover")
    out = osamp.sample_base(vel, g["enc"], g["ctx"], g["src"], g["seed"], null_emb=g["null_emb"],
                            infer_steps=6, guidance_scale=3.0, shift=3.0, cover_noise_strength=0.3,
                            audio_cover_strength=0.5, enc_non_cover=g["enc_nc"], ctx_non_cover=g["ctx_nc"],
                            new_cache=CrossCache)
    assert rel_l2(out, g["out"]) < TOL


def test_guidance_functions():
    g = golden("guidance")
    m = og.Momentum()
    a1 = og.apg(g["pc"], g["pu"], 7.0, m)
    a2 = og.apg(g["pc2"], g["pu2"], 7.0, m)
    assert rel_l2(a1, g["apg1"]) < 1e-6 and rel_l2(a2, g["apg2"]) < 1e-6
    d = og.adg(g["lat"], g["pc"][:1], g["pu"][:1], 0.7, 5.0)
    assert rel_l2(d, g["adg"]) < 1e-5


def test_schedules_are_dtype_rounded():
    t = osamp.base_schedule(27, 3.0, torch.bfloat16)
    assert t.dtype == torch.bfloat16 and t[0] == 1 and t[-1] == 0 and len(t) == 28
    assert osamp.turbo_schedule(2.6) == osamp.SHIFT_TIMESTEPS[3.0]
    assert osamp.turbo_schedule(3.0, [0.97, 0.8, 0.0, 0.0]) == [0.9545454545454546, 0.7692307692307693]


def test_vae_tiling_glue_matches_reference_mixins():
    cfg = ovae.VaeConfig.tiny()
    w = make_vae_weights(cfg, seed=3)
    g = golden("vae_tiled_decode")
    wav = ovae.tiled_decode(lambda z: ovae.decode(w, cfg, z), g["z"], g["chunk"], g["overlap"])
    assert wav.shape == g["wav"].shape and max_abs(wav, g["wav"]) < 1e-5
    g = golden("vae_tiled_encode")
    enc = lambda a, w0: ovae.encode_sample(w, cfg, a, torch.zeros(1))
    lat = ovae.tiled_encode(enc, g["audio"], g["chunk"], g["overlap"])
    assert lat.shape == g["lat"].shape and max_abs(lat, g["lat"]) < 1e-5


def test_vae_shapes_and_halo():
    cfg = ovae.VaeConfig.tiny()
    w = make_vae_weights(cfg, seed=3)
    z = torch.randn(1, 64, 120)
    full = ovae.decode(w, cfg, z)
    assert full.shape == (1, 2, 120 * cfg.hop)
    # overlap-discard tiling with a sufficient halo reproduces the untiled decode (the tiny
    # 2-stage codec has a wider receptive field in latent frames than the shipped 5-stage one)
    tiled = ovae.tiled_decode(lambda x: ovae.decode(w, cfg, x), z, 80, 24)
    assert max_abs(tiled, full) < 1e-3  # values are O(5); fp32 conv order differs
    audio = torch.rand(1, 2, 60 * cfg.hop) - 0.5
    m, s = ovae.encode_moments(w, cfg, audio)
    assert m.shape == (1, 64, 60) and s.shape == (1, 64, 60)


def test_condition_encoder_matches_reference():
    """oracle.cond vs the REAL AceStepConditionEncoder (tools/make_golden_cond.py): padded lyrics / text,
    packed timbre references, and rows whose whole sliding band is padding (uniform softmax quirk)."""
    from oracle.cond import CondConfig, condition_encoder, make_cond_weights

    g = golden("cond_encoder")
    cfg = CondConfig.tiny()
    w = make_cond_weights(cfg, seed=5)
    hs, m = condition_encoder(w, cfg, g["text"], g["text_mask"], g["lyric"], g["lyric_mask"], g["refer"], g["order"])
    assert torch.equal(m.long(), g["out_mask"])
    assert rel_l2(hs, g["out_hidden"]) <= 2e-5  # every row, padded ones included: the DiT attends to them


def test_condition_sequence_plumbing_matches_oracle():
    """acestep_b200.cond's device-side pack_sequences / unpack_timbre_embeddings (PyTorch host code,
    runs on CPU tensors too) against the oracle restatement of :135-166 / :1020-1071."""
    from acestep_b200.cond import pack_sequences, unpack_timbre_embeddings
    from oracle import cond as ocond

    gen = torch.Generator().manual_seed(3)
    h1, h2 = torch.randn(3, 7, 16, generator=gen), torch.randn(3, 4, 16, generator=gen)
    m1 = torch.tensor([[1] * 7, [1] * 3 + [0] * 4, [0] * 7])
    m2 = torch.tensor([[1, 1, 0, 0], [1, 0, 0, 0], [1, 1, 1, 1]])
    got, want = pack_sequences(h1, h2, m1, m2), ocond.pack_sequences(h1, h2, m1, m2)
    assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])
    embs = torch.randn(6, 8, generator=gen)
    order = torch.tensor([2, 0, 2, 1, 2, 0])
    got, want = unpack_timbre_embeddings(embs, order), ocond.unpack_timbre(embs, order)
    assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])


def test_output_path_restatements_match_reference_bit_exact():
    """oracle.output vs the real `normalize_audio` (compiled from the reference source by
    tools/make_golden_output.py) and the handler's peak-normalisation expressions: bit-exact."""
    from oracle import output as oout

    g = golden("output_normalize")
    wavs, peak = oout.peak_normalize(g["wav"])
    assert torch.equal(peak, g["peak"]) and torch.equal(wavs, g["stage1"])
    for db in (-1.0, 0.0, -6.0, -0.1):
        final, _ = oout.finalize(g["wav"], db)
        assert torch.equal(final, g[f"final_db{db}"]), db
    raw = torch.stack([oout.normalize_audio(g["raw"][i], -1.0) for i in range(g["raw"].shape[0])])
    assert torch.equal(raw, g["raw_db-1.0"])
    # silence and near-silence are returned unchanged (audio_utils.py:50-51)
    assert torch.equal(g["final_db-1.0"][5], g["stage1"][5]) and torch.equal(g["final_db-1.0"][6], g["stage1"][6])
    assert oout.latent_guard(torch.zeros(2, 3)) == (False, False)
    assert oout.latent_guard(torch.tensor([0.0, float("inf")])) == (True, True)


def test_cross_attention_probabilities_match_reference(dit):
    """oracle `cross_probs` vs `decoder(..., output_attentions=True, enable_early_exit=True)[2]` of the real
    reference (tools/make_golden_attn.py)."""
    cfg, w, _ = dit
    g = golden("dit_cross_attn_tiny")
    got = []
    vt = dit_forward(w, cfg, g["xt"], g["t"], g["ctx"], g["enc"], cross_probs=got)
    assert len(got) == cfg.num_hidden_layers == g["probs"].shape[0]
    assert rel_l2(torch.stack(got), g["probs"]) < TOL
    assert rel_l2(vt, g["vt"]) < TOL  # the eager path's velocity equals the sdpa one to fp32 round-off


def test_sft_explicit_timesteps(dit):
    """The SFT model's `timesteps` override (sft/modeling_acestep_v15_base.py:1864-1875: replaces the
    linspace/shift schedule and infer_steps) against the REAL sft module (tools/make_golden_sft.py)."""
    _, _, vel = dit
    g = golden("sft_timesteps")
    out = osamp.sample_base(vel, g["enc"], g["ctx"], g["src"], [int(s) for s in g["seeds"]], null_emb=g["null_emb"],
                            infer_steps=99, guidance_scale=6.0, shift=2.0, timesteps=g["timesteps"],
                            new_cache=CrossCache)
    assert rel_l2(out, g["out"]) < TOL
    g = golden("sft_timesteps_cover")
    out = osamp.sample_base(vel, g["enc"], g["ctx"], g["src"], g["seed"], null_emb=g["null_emb"], infer_steps=99,
                            guidance_scale=4.0, shift=3.0, timesteps=g["timesteps"], cover_noise_strength=0.25,
                            audio_cover_strength=0.6, enc_non_cover=g["enc_nc"], ctx_non_cover=g["ctx_nc"],
                            new_cache=CrossCache)
    assert rel_l2(out, g["out"]) < TOL


@pytest.mark.parametrize("tag,cfg", [
    ("tiny", ovae.VaeConfig.tiny()),
    ("odd", ovae.VaeConfig(encoder_hidden_size=32, downsampling_ratios=[2, 3, 5], channel_multiples=[1, 2, 4],
                           decoder_channels=32, decoder_input_channels=16, audio_channels=2)),
])
def test_vae_arithmetic_matches_reference_mlx_implementation(tag, cfg):
    """oracle.vae decode / encode vs the reference's OWN in-tree Oobleck implementation
    (acestep/models/mlx/vae_model.py + vae_convert.py, unmodified, run through tools/mlx_shim.py by
    tools/make_golden_vae_mlx.py): layer order, paddings (incl. ceil(s / 2) with odd strides, where every
    transposed conv yields L*s - 1 samples), dilations, Snake, weight-norm fusion, ConvTranspose weight axes,
    mean / scale split.  fp32 vs fp32 through 26+ layers with the fusion done in numpy there: 1e-4."""
    g = golden("vae_mlx_reference")
    sd = make_vae_weights(cfg, seed=int(g[f"{tag}_seed"]))
    dec = ovae.decode(sd, cfg, g[f"{tag}_z"].transpose(1, 2))
    want = g[f"{tag}_decoded"].transpose(1, 2)
    assert dec.shape == want.shape and rel_l2(dec, want) < 1e-4
    mean, scale = ovae.encode_moments(sd, cfg, g[f"{tag}_wav"].transpose(1, 2))
    moments = g[f"{tag}_moments"].transpose(1, 2)
    assert mean.shape[-1] == moments.shape[-1]
    assert rel_l2(torch.cat([mean, scale], dim=1), moments) < 1e-4
    assert rel_l2(mean, g[f"{tag}_mean"].transpose(1, 2)) < 1e-4


def test_audio_tokenizer_and_detokenizer_match_reference_modules():
    """oracle.tokenizer vs the REAL AceStepAudioTokenizer / AudioTokenDetokenizer and the model-level tokenize /
    detokenize / LM-hint substitution (turbo :1178-1218, :730-990, :1577-1600, :1646), run unmodified by
    tools/make_golden_tokenizer.py with the restated ResidualFSQ standing in for the absent third-party
    `vector_quantize_pytorch`: pins the silence padding, the 5 Hz mask pooling, the pooler's special token and
    position layout, the detokenizer's patch expansion, the crop and the is_covers substitution.  Indices are
    integer work: bit-exact."""
    from oracle.tokenizer import TokConfig, detokenize, lm_hints, make_tokenizer_weights, tokenize

    cfg = TokConfig.tiny()
    w = make_tokenizer_weights(cfg, seed=9)
    g = golden("tokenizer")
    q, idx, m5 = tokenize(w, cfg, g["x"], g["silence"], g["mask"])
    assert torch.equal(idx.long(), g["indices"]) and idx.shape == (2, 5, 1)
    assert len(set(idx.reshape(-1).tolist())) >= 5, "degenerate fixture: the codes must spread over the levels"
    assert torch.equal(m5, g["mask5"])
    assert rel_l2(q, g["quantized"]) < TOL
    assert rel_l2(detokenize(w, cfg, q), g["hints"]) < TOL
    new_src = lm_hints(w, cfg, g["x"], g["silence"], g["mask"], g["src"], g["is_covers"])
    assert rel_l2(new_src, g["new_src"]) < TOL and torch.equal(new_src[1], g["src"][1])


def test_fsq_restatement_properties():
    """The FSQ restatement (third-party library absent: UNPINNED against it) at least has the properties the
    published algorithm guarantees: codes lie on the per-dimension grid {-1, ..., 1} with `levels` points (even
    levels are offset by half a step), indices are a bijection of the code grid onto [0, prod(levels)), and
    quantisation is idempotent on its own codes."""
    from oracle.tokenizer import fsq_quantize

    levels = torch.tensor([8, 8, 8, 5, 5, 5])
    g = torch.Generator().manual_seed(2)
    z = torch.randn(4000, 6, generator=g) * 2.0
    codes, idx = fsq_quantize(z, levels)
    half = (levels // 2).float()
    q = codes * half
    assert torch.equal(q, q.round())
    for c, L in enumerate(levels.tolist()):
        lo, hi = (-(L // 2), L // 2 - 1) if L % 2 == 0 else (-(L // 2), L // 2)
        assert int(q[:, c].min()) >= lo and int(q[:, c].max()) <= hi
        assert len(set(q[:, c].tolist())) == L
    assert int(idx.min()) >= 0 and int(idx.max()) < int(levels.prod())
    basis = torch.cumprod(torch.cat([torch.ones(1, dtype=torch.long), levels[:-1]]), 0)
    back = torch.stack([(idx.long() // basis[c]) % levels[c] for c in range(6)], dim=-1).float() - half
    assert torch.equal(back, q)
