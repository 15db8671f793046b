"""GPU parity tests (run on the B200 box): CUDA path through the C ABI vs the CPU oracle.

Tolerances (floating point, bf16 storage / fp32 accumulate):
  * single GEMM / attention vs fp64 math on the same bf16 inputs: max-abs error <= 2 bf16 ulps of
    the output scale (2^-7 relative);
  * one DiT forward vs the fp32 oracle evaluated with the same bf16-rounded weights:
    rel-L2 <= 2e-2 — the reference's own bf16-vs-fp32 spread on this op is 1.67e-2;
  * VAE decode/encode vs the fp32 oracle with the same folded bf16 weights: the random-init codec
    amplifies bf16 rounding through 26 Snake/conv layers (a pure-torch bf16 run of the oracle — what
    the reference executes — sits 6e-2 from fp32), so the bound is measured in the test:
    rel-L2(cuda, fp32) <= max(1.1 * rel-L2(torch-bf16, fp32), 2e-2), i.e. no worse than the
    reference's own bf16 execution.
"""
import ctypes as C
import math

import pytest
import torch

from helpers import golden, max_abs, rel_l2

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("CUDA device required", allow_module_level=True)

from acestep_b200 import _lib  # noqa: E402
from acestep_b200.dit import B200DiT, DiTShape  # noqa: E402
from acestep_b200.pack import folded_vae_state  # noqa: E402
from acestep_b200.vae import B200Vae, VaeShape  # noqa: E402
from oracle import vae as ovae  # noqa: E402
from oracle.dit import DiTConfig, dit_forward  # noqa: E402
from oracle.weights import bf16_round_, make_dit_weights, make_vae_weights  # noqa: E402

DEV = torch.device("cuda:0")


@pytest.fixture(scope="module")
def lib():
    """The RELEASE library: one kernel per op, no switches."""
    l = _lib.load()
    _lib.check(l.ace_init(0))
    yield l


@pytest.fixture(scope="module")
def probe():
    """The probe build (-DACE_PROBE): same sources plus the A/B hooks (scalar reference GEMM, two-launch codec
    residual unit, P-through-shared-memory attention) that the two-path equivalence tests flip."""
    l = _lib.load_probe()
    _lib.check(l.ace_init(0), lib=l)
    yield l
    l.ace_debug_set_gemm_reference(0)
    l.ace_debug_set_vae_fused(-1)
    l.ace_debug_set_attention_p_in_tmem(-1)


def _pick(lib, probe, ref):
    """ref = 1: the scalar reference GEMM behind the same epilogues (probe build); ref = 0: the release library."""
    if ref:
        probe.ace_debug_set_gemm_reference(1)
        return probe
    return lib


def _stream():
    return torch.cuda.current_stream().cuda_stream


@pytest.mark.parametrize("m,n,k", [(128, 128, 64), (300, 256, 128), (1500, 2048, 2048), (77, 384, 6144)])
@pytest.mark.parametrize("ref", [1, 0])
def test_linear(lib, probe, m, n, k, ref):
    g = torch.Generator().manual_seed(m + n + k)
    a = torch.randn(m, k, generator=g).to(torch.bfloat16)
    b = (torch.randn(n, k, generator=g) * 0.05).to(torch.bfloat16)
    bias = torch.randn(n, generator=g).to(torch.bfloat16)
    want = a.double() @ b.double().T + bias.double()
    ad, bd, biasd = a.to(DEV), b.to(DEV), bias.to(DEV)
    out = torch.full((m, n), float("nan"), dtype=torch.bfloat16, device=DEV)
    use = _pick(lib, probe, ref)
    try:
        _lib.check(use.ace_linear(ad.data_ptr(), bd.data_ptr(), biasd.data_ptr(), out.data_ptr(), m, n, k,
                                  _stream()), lib=use)
        torch.cuda.synchronize()
    finally:
        probe.ace_debug_set_gemm_reference(0)
    got = out.cpu().double()
    assert torch.isfinite(got).all()
    assert max_abs(got, want) <= 2 ** -7 * float(want.abs().max()) + 1e-3


def _attn_ref(q, k, v, heads, kv_heads, window):
    B, Sq, _ = q.shape
    Skv = k.shape[1]
    qh = q.double().view(B, Sq, heads, 128).transpose(1, 2)
    kh = k.double().view(B, Skv, kv_heads, 128).transpose(1, 2).repeat_interleave(heads // kv_heads, 1)
    vh = v.double().view(B, Skv, kv_heads, 128).transpose(1, 2).repeat_interleave(heads // kv_heads, 1)
    s = qh @ kh.transpose(-1, -2) / math.sqrt(128)
    if window >= 0:
        i = torch.arange(Sq)[:, None]
        j = torch.arange(Skv)[None, :]
        s = s.masked_fill((i - j).abs() > window, float("-inf"))
    return (torch.softmax(s, -1) @ vh).transpose(1, 2).reshape(B, Sq, heads * 128)


@pytest.mark.parametrize("B,H,HK,Sq,Skv,win", [
    (1, 2, 1, 19, 19, -1), (2, 2, 1, 19, 19, 8), (1, 4, 2, 200, 200, 128), (2, 4, 2, 333, 333, -1),
    (1, 2, 2, 70, 257, -1), (1, 16, 8, 750, 750, 128), (1, 2, 1, 64, 1, -1), (1, 2, 1, 129, 129, 0),
])
def test_attention(lib, B, H, HK, Sq, Skv, win):
    g = torch.Generator().manual_seed(B * 1000 + Sq + Skv)
    q = torch.randn(B, Sq, H * 128, generator=g).to(torch.bfloat16)
    k = torch.randn(B, Skv, HK * 128, generator=g).to(torch.bfloat16)
    v = torch.randn(B, Skv, HK * 128, generator=g).to(torch.bfloat16)
    want = _attn_ref(q, k, v, H, HK, win)
    qd, kd, vd = q.to(DEV), k.to(DEV), v.to(DEV)
    out = torch.full((B, Sq, H * 128), float("nan"), dtype=torch.bfloat16, device=DEV)
    _lib.check(lib.ace_attention(qd.data_ptr(), kd.data_ptr(), vd.data_ptr(), out.data_ptr(), B, H, HK, Sq,
                                       Skv, win, _stream()))
    torch.cuda.synchronize()
    got = out.cpu().double()
    assert torch.isfinite(got).all()
    assert max_abs(got, want) <= 2e-2  # P is rounded to bf16 before P@V, outputs are O(1)
    assert rel_l2(got, want) <= 1e-2


def _tiny_dit():
    cfg = DiTConfig.tiny()
    w = bf16_round_(make_dit_weights(cfg, seed=0))
    return cfg, w


@pytest.mark.parametrize("ref", [1, 0])
def test_dit_forward_tiny_vs_oracle(lib, probe, ref):
    cfg, w = _tiny_dit()
    g = golden("dit_forward_tiny")
    xt, ctx, enc = (g[k].to(torch.bfloat16) for k in ("xt", "ctx", "enc"))
    t = g["t"].to(torch.bfloat16)
    want = dit_forward(w, cfg, xt.float(), t.float(), ctx.float(), enc.float(), bf16_time=True)
    use = _pick(lib, probe, ref)
    try:
        dit = B200DiT(w, DiTShape.from_config(cfg), DEV, lib=use)
        dit.bind(xt.shape[0], xt.shape[1], enc.shape[1])
        dit.set_condition(enc.to(DEV))
        vt = dit.step(xt.to(DEV), ctx.to(DEV), t.float().tolist())
        vt2 = dit.step(xt.to(DEV), ctx.to(DEV), t.float().tolist())  # second call: CUDA-graph replay
        torch.cuda.synchronize()
    finally:
        probe.ace_debug_set_gemm_reference(0)
    assert torch.isfinite(vt.float()).all()
    assert rel_l2(vt.cpu().float(), want) <= 2e-2
    assert max_abs(vt2, vt) == 0.0
    # and against the real reference's fp32 output (weights differ by bf16 rounding only)
    assert rel_l2(vt.cpu().float(), g["vt"]) <= 3e-2


def test_dit_forward_mid_size(lib):
    """A wider/deeper config with S > 128 so full, banded and cross attention all span several tiles."""
    cfg = DiTConfig(hidden_size=512, intermediate_size=1024, num_hidden_layers=4, num_attention_heads=4,
                    num_key_value_heads=2, sliding_window=128)
    w = bf16_round_(make_dit_weights(cfg, seed=5))
    g = torch.Generator().manual_seed(6)
    B, T, E = 2, 601, 130
    xt = torch.randn(B, T, 64, generator=g).to(torch.bfloat16)
    ctx = torch.cat([torch.randn(B, T, 64, generator=g), torch.ones(B, T, 64)], -1).to(torch.bfloat16)
    enc = torch.randn(B, E, cfg.hidden_size, generator=g).to(torch.bfloat16)
    t = torch.tensor([0.75, 0.25]).to(torch.bfloat16)
    want = dit_forward(w, cfg, xt.float(), t.float(), ctx.float(), enc.float(), bf16_time=True)
    dit = B200DiT(w, DiTShape.from_config(cfg), DEV)
    dit.bind(B, T, E)
    dit.set_condition(enc.to(DEV))
    vt = dit.step(xt.to(DEV), ctx.to(DEV), t.float().tolist())
    torch.cuda.synchronize()
    assert rel_l2(vt.cpu().float(), want) <= 2e-2


@pytest.mark.parametrize("hidden,T", [(512, 601), (2048, 1500), (512, 6403), (2048, 3000), (2048, 6001), (2048, 7001)])
def test_gemm_tail_path_matches_slab_path(lib, probe, hidden, T):
    """The residual GEMMs' tail path (residual tile by TMA into the freed operand ring, in-place update in shared
    memory, h and g out by TMA stores: gemm.cuh / EpiGatedResid::tail_box) rounds exactly like the slab path it
    replaces on a CTA's last tile, so a forward must be bit-identical with it switched off (probe build,
    ACE_NO_TMA_TAIL=1) — also against the release library.  The long cases are multi-wave problems (more pair
    tiles than CTA pairs: 192-wide tiles at M = 3000 / 6402, 256-wide at M = 6002 with a ragged last row tile), which
    run the every-tile variant (ALLTAIL: dedicated residual boxes, loader warp, g made in place).  T = 7001 (M = 7002)
    is the shape where down_proj runs 192-wide tiles in n-fastest order (A = 57 MB: GemmShape::raster_n), so a CTA pair
    meets the two-box last column tile BEFORE other tiles — the per-box barrier parity must not follow the tile
    counter (a bug this case caught)."""
    import os

    layers = 4 if hidden == 512 else 2
    cfg = DiTConfig(hidden_size=hidden, intermediate_size=2 * hidden, num_hidden_layers=layers,
                    num_attention_heads=hidden // 128, num_key_value_heads=max(1, hidden // 256), sliding_window=128)
    w = bf16_round_(make_dit_weights(cfg, seed=11))
    g = torch.Generator().manual_seed(12)
    B, E = 2, 70
    xt = torch.randn(B, T, 64, generator=g).to(torch.bfloat16).to(DEV)
    ctx = torch.randn(B, T, 128, generator=g).to(torch.bfloat16).to(DEV)
    enc = torch.randn(B, E, cfg.hidden_size, generator=g).to(torch.bfloat16).to(DEV)
    outs = []
    for use, env in ((probe, "1"), (probe, None), (lib, None)):
        if env is not None:
            os.environ["ACE_NO_TMA_TAIL"] = env
        try:
            dit = B200DiT(w, DiTShape.from_config(cfg), DEV, lib=use)
            dit.bind(B, T, E)
        finally:
            os.environ.pop("ACE_NO_TMA_TAIL", None)
        dit.set_condition(enc)
        outs.append(dit.step(xt, ctx, [0.5, 0.5]).clone())
        torch.cuda.synchronize()
        dit.close()
    assert torch.isfinite(outs[0].float()).all()
    assert torch.equal(outs[0], outs[1]), max_abs(outs[0], outs[1])
    assert torch.equal(outs[0], outs[2]), max_abs(outs[0], outs[2])


def _bf16_floor(fn_fp32, fn_bf16):
    """Spread between an all-bf16 torch run and the fp32 run of the same oracle op."""
    return rel_l2(fn_bf16().float(), fn_fp32())


def _tiny_vae():
    cfg = ovae.VaeConfig.tiny()
    sd = make_vae_weights(cfg, seed=3)
    shape = VaeShape(encoder_hidden_size=cfg.encoder_hidden_size, downsampling_ratios=cfg.downsampling_ratios,
                     channel_multiples=cfg.channel_multiples, decoder_channels=cfg.decoder_channels)
    return cfg, sd, shape


@pytest.mark.parametrize("ref", [1, 0])
def test_vae_decode_tiny(lib, probe, ref):
    cfg, sd, shape = _tiny_vae()
    wf = folded_vae_state(sd)
    g = torch.Generator().manual_seed(40)
    z = torch.randn(2, 64, 75, generator=g).to(torch.bfloat16)
    want = ovae.decode(wf, cfg, z.float())
    use = _pick(lib, probe, ref)
    try:
        vae = B200Vae(sd, shape, DEV, lib=use)
        got = vae.decode(z.to(DEV))
        torch.cuda.synchronize()
    finally:
        probe.ace_debug_set_gemm_reference(0)
    assert got.shape == want.shape and got.dtype == torch.float32
    wb = {k: v.to(torch.bfloat16) for k, v in wf.items()}
    floor = _bf16_floor(lambda: want, lambda: ovae.decode(wb, cfg, z))
    assert rel_l2(got.cpu(), want) <= max(1.1 * floor, 2e-2), floor


@pytest.mark.parametrize("ref", [1, 0])
def test_vae_encode_tiny(lib, probe, ref):
    cfg, sd, shape = _tiny_vae()
    wf = folded_vae_state(sd)
    g = torch.Generator().manual_seed(41)
    audio = (torch.rand(1, 2, 8 * 200, generator=g) - 0.5)
    eps = torch.randn(200, 64, generator=g).to(torch.bfloat16)
    mean, scale = ovae.encode_moments(wf, cfg, audio.to(torch.bfloat16).float())
    want_mean = mean[0].T
    want = (mean + (torch.nn.functional.softplus(scale) + 1e-4) * eps.float().T[None])[0].T
    use = _pick(lib, probe, ref)
    try:
        vae = B200Vae(sd, shape, DEV, lib=use)
        got_mean = vae.encode_samples(audio[0].to(DEV), None)
        got = vae.encode_samples(audio[0].to(DEV), eps.to(DEV))
        torch.cuda.synchronize()
    finally:
        probe.ace_debug_set_gemm_reference(0)
    wb = {k: v.to(torch.bfloat16) for k, v in wf.items()}
    floor = _bf16_floor(lambda: mean, lambda: ovae.encode_moments(wb, cfg, audio.to(torch.bfloat16))[0])
    tol = max(1.1 * floor, 2e-2)
    assert rel_l2(got_mean.cpu().float(), want_mean) <= tol, floor
    assert rel_l2(got.cpu().float(), want) <= tol, floor


@pytest.mark.parametrize("frames", [5, 75, 301, 13000])
def test_fused_res_unit_matches_two_launch_path(lib, probe, frames):
    """csrc/resunit.cuh (one kernel per residual unit at the 128-channel stages) rounds at the same
    points as the two-launch tap-shifted GEMM path, so decode and encode must agree bit for bit — the RELEASE
    library's fused kernels against the probe build's two-launch path (and the probe build's own fused path)."""
    cfg, sd, shape = _tiny_vae()
    g = torch.Generator().manual_seed(50 + frames)
    z = torch.randn(1, 64, frames, generator=g).to(torch.bfloat16)
    audio = torch.rand(2, cfg.hop * frames, generator=g) - 0.5
    eps = torch.randn(frames, 64, generator=g).to(torch.bfloat16)
    out = {}
    try:
        for mode in (0, 1, 2):  # 0: probe two-launch, 1: probe fused, 2: release
            if mode < 2:
                probe.ace_debug_set_vae_fused(mode)
            vae = B200Vae(sd, shape, DEV, lib=probe if mode < 2 else lib)
            dec = vae.decode(z.to(DEV))
            enc = vae.encode_samples(audio.to(DEV), eps.to(DEV))
            torch.cuda.synchronize()
            out[mode] = (dec.cpu(), enc.cpu())
            vae.close()
    finally:
        probe.ace_debug_set_vae_fused(-1)
    assert torch.isfinite(out[2][0]).all()
    for mode in (1, 2):
        assert torch.equal(out[0][0], out[mode][0]), (mode, max_abs(out[0][0], out[mode][0]))
        assert torch.equal(out[0][1], out[mode][1]), (mode, max_abs(out[0][1].float(), out[mode][1].float()))


def _cond_inputs(cfg, B, Ll, Lt, Lr, lyric_lens, text_lens, order, seed):
    g = torch.Generator().manual_seed(seed)
    lyric = torch.randn(B, Ll, cfg.text_hidden_dim, generator=g)
    text = torch.randn(B, Lt, cfg.text_hidden_dim, generator=g)
    refer = torch.randn(len(order), Lr, cfg.timbre_hidden_dim, generator=g)
    lm = (torch.arange(Ll)[None] < torch.tensor(lyric_lens)[:, None]).long()
    tm = (torch.arange(Lt)[None] < torch.tensor(text_lens)[:, None]).long()
    return text, tm, lyric, lm, refer, torch.tensor(order, dtype=torch.long)


@pytest.mark.parametrize("case", ["golden", "long"])
def test_condition_encoder_vs_oracle(lib, case):
    """csrc/cond.cu (lyric / timbre stacks with the key-padding mask) + acestep_b200.cond packing against
    oracle.cond on bf16-rounded weights and inputs; bound = measured spread of an all-bf16 torch run of the
    oracle (floor 2e-2), over EVERY row, padded ones included (the DiT cross-attends to them)."""
    from acestep_b200.cond import B200ConditionEncoder, CondShape
    from oracle.cond import CondConfig, condition_encoder, make_cond_weights

    cfg = CondConfig.tiny()
    w = bf16_round_(make_cond_weights(cfg, seed=5))
    if case == "golden":
        g = golden("cond_encoder")
        args = [g[k] for k in ("text", "text_mask", "lyric", "lyric_mask", "refer", "order")]
    else:  # several query tiles, padded tails longer than the window, ragged timbre packing
        args = list(_cond_inputs(cfg, 3, 300, 40, 150, [300, 131, 17], [40, 9, 1], [0, 1, 1, 2, 2, 2], seed=77))
    args = [a.to(torch.bfloat16).float() if a.is_floating_point() else a for a in args]
    want, want_mask = condition_encoder(w, cfg, *args)
    wb = {k: v.to(torch.bfloat16) for k, v in w.items()}
    a16 = [a.to(torch.bfloat16) if a.is_floating_point() else a for a in args]
    floor = rel_l2(condition_encoder(wb, cfg, *a16)[0].float(), want)
    enc = B200ConditionEncoder(w, CondShape.from_config(cfg), DEV)
    got, got_mask = enc(text_hidden_states=args[0], text_attention_mask=args[1], lyric_hidden_states=args[2],
                        lyric_attention_mask=args[3], refer_audio_acoustic_hidden_states_packed=args[4],
                        refer_audio_order_mask=args[5])
    torch.cuda.synchronize()
    assert got.shape == want.shape and got.dtype == torch.bfloat16
    assert torch.equal(got_mask.cpu(), want_mask)
    assert torch.isfinite(got).all()
    assert rel_l2(got.cpu().float(), want) <= max(1.5 * floor, 2e-2), floor
    enc.close()


def test_condition_encoder_rejects_unrepresentable_mask(lib):
    from acestep_b200.cond import B200ConditionEncoder, CondShape
    from oracle.cond import CondConfig, make_cond_weights

    cfg = CondConfig.tiny()
    enc = B200ConditionEncoder(make_cond_weights(cfg, seed=5), CondShape.from_config(cfg), DEV)
    text, tm, lyric, lm, refer, order = _cond_inputs(cfg, 1, 16, 4, 8, [16], [4], [0], seed=1)
    lm[0, 3] = 0  # a hole: not a right-padded mask
    with pytest.raises(ValueError):
        enc(text_hidden_states=text, text_attention_mask=tm, lyric_hidden_states=lyric, lyric_attention_mask=lm,
            refer_audio_acoustic_hidden_states_packed=refer, refer_audio_order_mask=order)
    enc.close()


@pytest.mark.parametrize("B,T,E", [(1, 1, 1), (3, 7, 2), (16, 5, 3)])
def test_dit_forward_edge_shapes(lib, B, T, E):
    """Smallest and most ragged shapes: one latent frame (one half-empty patch), one condition token, the
    maximum effective batch the handle accepts; every GEMM is a single partial tile (split-K path)."""
    cfg, w = _tiny_dit()
    g = torch.Generator().manual_seed(100 + B + T + E)
    xt = torch.randn(B, T, 64, generator=g).to(torch.bfloat16)
    ctx = torch.randn(B, T, 128, generator=g).to(torch.bfloat16)
    enc = torch.randn(B, E, cfg.hidden_size, generator=g).to(torch.bfloat16)
    t = torch.rand(B, generator=g).to(torch.bfloat16)
    want = dit_forward(w, cfg, xt.float(), t.float(), ctx.float(), enc.float(), bf16_time=True)
    dit = B200DiT(w, DiTShape.from_config(cfg), DEV)
    dit.bind(B, T, E)
    dit.set_condition(enc.to(DEV))
    vt = dit.step(xt.to(DEV), ctx.to(DEV), t.float().tolist())
    torch.cuda.synchronize()
    assert vt.shape == (B, T, 64) and torch.isfinite(vt.float()).all()
    assert rel_l2(vt.cpu().float(), want) <= 2e-2


def test_dit_forward_full_size_short_clip(lib):
    """The shipped architecture (2048 wide, 24 layers, 16/8 heads) at the 10 s shape of BASELINE config 0
    (T = 250 -> 125 tokens, no CFG): every GEMM runs its production tile / split-K plan.  Tolerance: the
    reference's own bf16-vs-fp32 spread at this size is 1.67e-2, bound 2e-2."""
    cfg = DiTConfig()
    w = bf16_round_(make_dit_weights(cfg, seed=0))
    g = torch.Generator().manual_seed(9)
    B, T, E = 1, 250, 64
    xt = torch.randn(B, T, 64, generator=g).to(torch.bfloat16)
    ctx = torch.cat([torch.randn(B, T, 64, generator=g), torch.ones(B, T, 64)], -1).to(torch.bfloat16)
    enc = torch.randn(B, E, cfg.hidden_size, generator=g).to(torch.bfloat16)
    t = torch.tensor([0.625]).to(torch.bfloat16)
    with torch.no_grad():
        want = dit_forward(w, cfg, xt.float(), t.float(), ctx.float(), enc.float(), bf16_time=True)
    dit = B200DiT(w, DiTShape.from_config(cfg), DEV)
    dit.bind(B, T, E)
    dit.set_condition(enc.to(DEV))
    vt = dit.step(xt.to(DEV), ctx.to(DEV), t.float().tolist())
    vt2 = dit.step(xt.to(DEV), ctx.to(DEV), t.float().tolist())
    torch.cuda.synchronize()
    assert torch.isfinite(vt.float()).all()
    assert max_abs(vt2, vt) == 0.0  # split-K reduces in a fixed order: replays are bit-identical
    assert rel_l2(vt.cpu().float(), want) <= 2e-2
    dit.close()


def test_c_abi_rejects_bad_arguments(lib):
    """Errors are int status codes + ace_last_error(), never exceptions or aborts (include/acestep_b200.h)."""
    cfg, w = _tiny_dit()
    dit = B200DiT(w, DiTShape.from_config(cfg), DEV)
    ws = torch.empty(1 << 20, dtype=torch.uint8, device=DEV)
    for bc, t, e in ((0, 8, 4), (17, 8, 4), (1, 0, 4), (1, 8, 0)):
        assert lib.ace_dit_bind(dit.handle, bc, t, e, ws.data_ptr(), ws.numel()) != 0
        assert lib.ace_last_error()
    assert lib.ace_dit_bind(dit.handle, 1, 64, 16, ws.data_ptr(), 1024) != 0  # workspace too small
    assert b"workspace" in lib.ace_last_error()
    with pytest.raises(_lib.B200Error):
        dit.step(torch.zeros(1, 8, 64, device=DEV, dtype=torch.bfloat16),
                 torch.zeros(1, 8, 128, device=DEV, dtype=torch.bfloat16), [0.5])  # not bound
    cfgv, sd, shape = _tiny_vae()
    vae = B200Vae(sd, shape, DEV)
    with pytest.raises(_lib.B200Error):
        vae.encode_samples(torch.zeros(2, cfgv.hop * 3 + 1, device=DEV), None)  # not a multiple of the hop
    with pytest.raises(ValueError):
        vae.encode(torch.zeros(1, 2, cfgv.hop - 1, device=DEV))  # shorter than one hop
    with pytest.raises(ValueError):
        vae.decode(torch.zeros(64, 4, device=DEV))  # wrong rank
    with pytest.raises(_lib.B200Error):
        B200DiT(w, DiTShape.from_config(cfg), "cpu")  # no CPU path


def test_vae_full_size_short_clip(lib):
    """The shipped Oobleck config (strides 2,4,4,6,10, channels 128..2048) on a 1 s clip: every stage's
    kernel variant runs (pair-tile convs, transposed / strided convs, fused residual units at 128 channels,
    SIMT end convs).  Bound like the tiny test: the spread of an all-bf16 torch run of the oracle."""
    cfg = ovae.VaeConfig()
    sd = make_vae_weights(cfg, seed=4)
    wf = folded_vae_state(sd)
    shape = VaeShape()
    g = torch.Generator().manual_seed(60)
    frames = 25
    z = torch.randn(1, 64, frames, generator=g).to(torch.bfloat16)
    audio = torch.rand(1, 2, frames * cfg.hop, generator=g) - 0.5
    vae = B200Vae(sd, shape, DEV)
    wb = {k: v.to(torch.bfloat16) for k, v in wf.items()}
    with torch.no_grad():
        want = ovae.decode(wf, cfg, z.float())
        floor = _bf16_floor(lambda: want, lambda: ovae.decode(wb, cfg, z))
        mean, _ = ovae.encode_moments(wf, cfg, audio.to(torch.bfloat16).float())
        efloor = _bf16_floor(lambda: mean, lambda: ovae.encode_moments(wb, cfg, audio.to(torch.bfloat16))[0])
    got = vae.decode(z.to(DEV))
    got_mean = vae.encode_samples(audio[0].to(DEV), None)
    torch.cuda.synchronize()
    assert got.shape == (1, 2, frames * 1920) and torch.isfinite(got).all()
    assert rel_l2(got.cpu(), want) <= max(1.1 * floor, 2e-2), floor
    assert got_mean.shape == (frames, 64)
    assert rel_l2(got_mean.cpu().float(), mean[0].T) <= max(1.1 * efloor, 2e-2), efloor
    vae.close()


def test_attention_p_in_tmem_stress(lib, probe):
    """P through tensor memory (default) against P through shared memory on long KV loops, many launches, both
    co-resident CTAs busy: the two paths round identically, so outputs must be bit-identical and finite.  (A
    missing PV_{j-2} -> S_j ordering passed every small test and produced NaNs only after ~10^3 launches.)"""
    g = torch.Generator().manual_seed(123)
    B, H, HK, S = 2, 16, 8, 1500
    q = (torch.randn(B, S, H * 128, generator=g) * 2).to(torch.bfloat16).to(DEV)
    k = (torch.randn(B, S, HK * 128, generator=g) * 2).to(torch.bfloat16).to(DEV)
    v = torch.randn(B, S, HK * 128, generator=g).to(torch.bfloat16).to(DEV)
    outs = {}
    try:
        for mode in (0, 1):  # 0: probe build, P through shared memory; 1: the release library (P in tensor memory)
            use = lib if mode == 1 else probe
            if mode == 0:
                probe.ace_debug_set_attention_p_in_tmem(0)
            o = torch.empty(B, S, H * 128, dtype=torch.bfloat16, device=DEV)
            ref = None
            for it in range(150 if mode == 1 else 2):
                for win in (-1, 128):
                    _lib.check(use.ace_attention(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), B, H, HK,
                                                 S, S, win, _stream()), lib=use)
                    if win == -1:
                        cur = o.clone()
                        if ref is None:
                            ref = cur
                        else:
                            assert torch.equal(cur, ref), f"launch {it}: result changed between launches"
            torch.cuda.synchronize()
            outs[mode] = ref
    finally:
        probe.ace_debug_set_attention_p_in_tmem(-1)
    assert torch.isfinite(outs[1].float()).all()
    assert torch.equal(outs[0], outs[1]), max_abs(outs[0].float(), outs[1].float())


def test_cross_attention_probabilities_tiny_vs_oracle_and_reference(lib):
    """ace_dit_cross_attentions (the lyric aligner's output_attentions call, handler/lyric_timestamp.py:78-91)
    vs the fp32 oracle with the same bf16-rounded weights, and vs the REAL reference decoder's fp32 output
    (tests/golden/dit_cross_attn_tiny.npz).  Tolerance: probabilities are bf16 (2^-9 relative) computed from
    scores the eager path rounds to bf16 twice (2^-8 relative on |score| <= ~6 -> a few 1e-2 relative on p):
    rel-L2 <= 3e-2 per layer, max-abs <= 2e-2, rows sum to 1 within 1e-2."""
    cfg, w = _tiny_dit()
    g = golden("dit_cross_attn_tiny")
    xt, ctx, enc = (g[k].to(torch.bfloat16) for k in ("xt", "ctx", "enc"))
    t = g["t"].to(torch.bfloat16)
    want = []
    dit_forward(w, cfg, xt.float(), t.float(), ctx.float(), enc.float(), bf16_time=True, cross_probs=want)
    L = cfg.num_hidden_layers
    assert len(want) == L
    dit = B200DiT(w, DiTShape.from_config(cfg), DEV)
    dit.bind(xt.shape[0], xt.shape[1], enc.shape[1])
    dit.set_condition(enc.to(DEV))
    for n_layers in (L, 1, 3):
        probs = dit.cross_attentions(xt.to(DEV), ctx.to(DEV), t.float().tolist(), n_layers)
        torch.cuda.synchronize()
        assert probs.shape == (n_layers, 2, cfg.num_attention_heads, 19, 21) and probs.dtype == torch.bfloat16
        p = probs.cpu().float()
        assert torch.isfinite(p).all() and p.min() >= 0 and p.max() <= 1
        assert (p.sum(-1) - 1).abs().max() <= 1e-2
        for l in range(n_layers):
            assert rel_l2(p[l], want[l]) <= 3e-2, (n_layers, l)
            assert max_abs(p[l], want[l]) <= 2e-2, (n_layers, l)
            assert rel_l2(p[l], g["probs"][l]) <= 4e-2, (n_layers, l)  # real reference, fp32 weights
    # the step after an attention extraction is unaffected (shared workspace, graph replay)
    vt = dit.step(xt.to(DEV), ctx.to(DEV), t.float().tolist())
    want_vt = dit_forward(w, cfg, xt.float(), t.float(), ctx.float(), enc.float(), bf16_time=True)
    assert rel_l2(vt.cpu().float(), want_vt) <= 2e-2
    with pytest.raises(ValueError):
        dit.cross_attentions(xt.to(DEV), ctx.to(DEV), t.float().tolist(), L + 1)


def test_cross_attention_probabilities_ragged_mid_size(lib):
    """S and E that are not multiples of the kernel's 16-row / 64-key tiles, GQA 4:2, odd E (2-byte rows)."""
    cfg = DiTConfig(hidden_size=512, intermediate_size=1024, num_hidden_layers=2, num_attention_heads=4,
                    num_key_value_heads=2, sliding_window=128)
    w = bf16_round_(make_dit_weights(cfg, seed=5))
    g = torch.Generator().manual_seed(8)
    B, T, E = 2, 333, 131
    xt = torch.randn(B, T, 64, generator=g).to(torch.bfloat16)
    ctx = torch.cat([torch.randn(B, T, 64, generator=g), torch.ones(B, T, 64)], -1).to(torch.bfloat16)
    enc = torch.randn(B, E, cfg.hidden_size, generator=g).to(torch.bfloat16)
    t = torch.tensor([0.125, 0.125]).to(torch.bfloat16)
    want = []
    dit_forward(w, cfg, xt.float(), t.float(), ctx.float(), enc.float(), bf16_time=True, cross_probs=want)
    dit = B200DiT(w, DiTShape.from_config(cfg), DEV)
    dit.bind(B, T, E)
    dit.set_condition(enc.to(DEV))
    p = dit.cross_attentions(xt.to(DEV), ctx.to(DEV), t.float().tolist(), 2).cpu().float()
    assert p.shape == (2, B, 4, 167, E)
    assert (p.sum(-1) - 1).abs().max() <= 1e-2
    for l in range(2):
        assert rel_l2(p[l], want[l]) <= 3e-2 and max_abs(p[l], want[l]) <= 2e-2, l


def test_timestep_cache_eviction_and_mixed_batches(lib):
    """The per-handle timestep cache (dit.cu: 96 entries, ring eviction, entries computed on first sight or by
    ace_dit_prepare_timesteps): more distinct timesteps than entries, batches whose items sit at DIFFERENT
    timesteps (one cached, one new — including the case where the new entry's slot range would evict the cached
    one), duplicates inside a batch, and a prepared schedule.  Every result must equal, bit for bit, what a fresh
    handle computes for the same inputs, and the fp32 oracle within the forward tolerance."""
    cfg, w = _tiny_dit()
    g = torch.Generator().manual_seed(77)
    B, T, E = 2, 33, 9
    xt = torch.randn(B, T, 64, generator=g).to(torch.bfloat16).to(DEV)
    ctx = torch.randn(B, T, 128, generator=g).to(torch.bfloat16).to(DEV)
    enc = torch.randn(B, E, cfg.hidden_size, generator=g).to(torch.bfloat16).to(DEV)

    def fresh(ts):
        d = B200DiT(w, DiTShape.from_config(cfg), DEV)
        d.bind(B, T, E)
        d.set_condition(enc)
        out = d.step(xt, ctx, ts).clone()
        torch.cuda.synchronize()
        d.close()
        return out

    dit = B200DiT(w, DiTShape.from_config(cfg), DEV)
    dit.bind(B, T, E)
    dit.set_condition(enc)
    bf = lambda x: float(torch.tensor(x, dtype=torch.bfloat16))
    ts = sorted({bf(i / 257.0) for i in range(1, 257)})[:130]  # 130 distinct bf16 timesteps > 96 entries
    assert len(ts) == 130
    first = dit.step(xt, ctx, [ts[0], ts[0]]).clone()
    mixed = {}
    for i in range(1, len(ts)):
        out = dit.step(xt, ctx, [ts[i], ts[i - 1]])  # item 1 cached by the previous step, item 0 new
        if i in (1, 95, 96, 97, 129):
            mixed[i] = out.clone()
    again = dit.step(xt, ctx, [ts[0], ts[0]])  # ts[0] was evicted long ago: recomputed
    dit.prepare_timesteps(ts[:64] + ts[:8])    # batched fill with duplicates and cached values
    prepared = dit.step(xt, ctx, [ts[5], ts[60]]).clone()
    torch.cuda.synchronize()
    assert torch.equal(first, again)
    assert torch.equal(first, fresh([ts[0], ts[0]]))
    for i, out in mixed.items():
        assert torch.equal(out, fresh([ts[i], ts[i - 1]])), i
    assert torch.equal(prepared, fresh([ts[5], ts[60]]))
    want = dit_forward(w, cfg, xt.cpu().float(), torch.tensor([ts[5], ts[60]]), ctx.cpu().float(), enc.cpu().float(),
                       bf16_time=True)
    assert rel_l2(prepared.cpu().float(), want) <= 2e-2
    dit.close()


def test_dit_forward_random_shapes_one_handle(lib):
    """Twelve seeded random (batch, frames, condition tokens) shapes through ONE handle, rebinding between them (the
    way a serving process sees requests): odd frame counts (half-empty last patch), sequence lengths either side of
    the 128-row tile and 64-key block boundaries, mixed timesteps per batch item.  Each forward against the fp32
    oracle, 2e-2; the last shape is replayed and must be bit-identical (graph re-capture per shape, timestep cache
    shared across shapes)."""
    cfg = DiTConfig(hidden_size=512, intermediate_size=1024, num_hidden_layers=3, num_attention_heads=4,
                    num_key_value_heads=2, sliding_window=16)
    w = bf16_round_(make_dit_weights(cfg, seed=21))
    dit = B200DiT(w, DiTShape.from_config(cfg), DEV)
    rng = torch.Generator().manual_seed(4242)
    pick = lambda lo, hi: int(torch.randint(lo, hi + 1, (1,), generator=rng))
    shapes = [(pick(1, 4), pick(1, 700), pick(1, 300)) for _ in range(9)] + [(2, 255, 64), (1, 257, 65), (3, 513, 129)]
    worst = 0.0
    for B, T, E in shapes:
        xt = torch.randn(B, T, 64, generator=rng).to(torch.bfloat16)
        ctx = torch.randn(B, T, 128, generator=rng).to(torch.bfloat16)
        enc = torch.randn(B, E, cfg.hidden_size, generator=rng).to(torch.bfloat16)
        t = torch.rand(B, generator=rng).to(torch.bfloat16)
        want = dit_forward(w, cfg, xt.float(), t.float(), ctx.float(), enc.float(), bf16_time=True)
        dit.bind(B, T, E)
        dit.set_condition(enc.to(DEV))
        vt = dit.step(xt.to(DEV), ctx.to(DEV), t.float().tolist())
        torch.cuda.synchronize()
        assert vt.shape == (B, T, 64) and torch.isfinite(vt.float()).all(), (B, T, E)
        err = rel_l2(vt.cpu().float(), want)
        assert err <= 2e-2, (B, T, E, err)
        worst = max(worst, err)
    again = dit.step(xt.to(DEV), ctx.to(DEV), t.float().tolist())
    torch.cuda.synchronize()
    assert torch.equal(again, vt)
    dit.close()


def test_vae_ragged_lengths_one_handle(lib):
    """Decode and encode at lengths either side of the 128-row tile edge (1, 2, 17, 127, 128, 129, 260 latent
    frames) through ONE codec handle, workspace regrown on demand; bound as in test_vae_decode_tiny."""
    cfg, sd, shape = _tiny_vae()
    wf = folded_vae_state(sd)
    wb = {k: v.to(torch.bfloat16) for k, v in wf.items()}
    vae = B200Vae(sd, shape, DEV)
    g = torch.Generator().manual_seed(60)
    for T in (129, 1, 260, 2, 128, 17, 127):
        z = torch.randn(1, 64, T, generator=g).to(torch.bfloat16)
        want = ovae.decode(wf, cfg, z.float())
        got = vae.decode(z.to(DEV))
        torch.cuda.synchronize()
        assert got.shape == (1, 2, T * cfg.hop) and torch.isfinite(got).all(), T
        floor = _bf16_floor(lambda: want, lambda: ovae.decode(wb, cfg, z))
        assert rel_l2(got.cpu().float(), want) <= max(1.1 * floor, 2e-2), (T, floor)
        audio = torch.rand(1, 2, T * cfg.hop, generator=g) - 0.5
        mean, _ = ovae.encode_moments(wf, cfg, audio.to(torch.bfloat16).float())
        got_mean = vae.encode_samples(audio[0].to(DEV), None)
        torch.cuda.synchronize()
        floor = _bf16_floor(lambda: mean, lambda: ovae.encode_moments(wb, cfg, audio.to(torch.bfloat16))[0])
        assert got_mean.shape == (T, 64)
        assert rel_l2(got_mean.cpu().float(), mean[0].T) <= max(1.1 * floor, 2e-2), (T, floor)
    vae.close()


def test_multiwave_residual_gemms_hand_back_protocol_under_a_forced_race(lib, probe):
    """M = 15000 (600 s, CFG): the residual GEMMs run 192-wide tiles in n-fastest order with the every-tile tail
    variant, so a CTA pair meets the two-box last column tile — a tile in which column group 1 has nothing to load —
    in the MIDDLE of its tile sequence.  A hand-back protocol that let the idle group arrive on `resid_empty` there
    deadlocks when group 1 gets a tile ahead of the loader warp (it completes two phases before the loader looks at
    the first): seen on some boxes only, caught by the 24-layer full-size test.  The probe build forces that schedule
    (ACE_RACE_DELAY=1: group 0 sleeps 30 us per tile — with it the old protocol deadlocks on the first launch): the
    forward must complete (the bounded barrier wait turns a deadlock into a launch failure) and equal the release
    library's result bit for bit; then a few undisturbed replays on the release library."""
    import os

    from acestep_b200.synthetic import random_dit_state

    shape = DiTShape(num_hidden_layers=2)
    state = random_dit_state(shape, 0, DEV)
    g = torch.Generator(device=DEV).manual_seed(15000)
    T, E = 15000, 70
    xt = torch.randn(2, T, 64, device=DEV, generator=g).bfloat16()
    ctx = torch.randn(2, T, 128, device=DEV, generator=g).bfloat16()
    enc = torch.randn(2, E, shape.hidden_size, device=DEV, generator=g).bfloat16()
    outs = []
    for use, env in ((probe, "1"), (lib, None)):
        if env is not None:
            os.environ["ACE_RACE_DELAY"] = env
        try:
            dit = B200DiT(state, shape, DEV, lib=use)
            dit.bind(2, T, E)  # the plans (and the probe switch) are made here
        finally:
            os.environ.pop("ACE_RACE_DELAY", None)
        dit.set_condition(enc)
        outs.append(dit.step(xt, ctx, [0.5, 0.25]).clone())
        torch.cuda.synchronize()
        if use is lib:
            for i in range(4):
                assert torch.equal(dit.step(xt, ctx, [0.5, 0.25]), outs[-1]), i
            torch.cuda.synchronize()
        dit.close()
    assert torch.isfinite(outs[0].float()).all()
    assert torch.equal(outs[0], outs[1])
