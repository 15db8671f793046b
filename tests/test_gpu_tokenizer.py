"""GPU parity of the audio tokenizer / residual FSQ / detokenizer (the LM-hint branch of
prepare_condition: turbo :1577-1600, :1630-1646) against oracle.tokenizer (itself pinned against the real reference
modules by tests/test_oracle_golden.py).

Tolerances:
  * pooler and detokenizer (floating point, 2-layer encoder stacks): rel-L2 <= max(1.5 x spread, 2e-2) vs the fp32
    oracle on bf16-rounded weights, spread = an all-bf16 torch run of the oracle, measured in the test;
  * FSQ indices (integer work): BIT-EXACT against the oracle executed with the reference's bf16 semantics (bf16
    project_in, fp32 FSQ) on every token whose bounded value is further than 0.05 from a rounding boundary on all
    channels; a token nearer than that may legitimately flip with the summation order of the 2048-long dot product
    (one bf16 ulp of y), so overall >= 97 % of tokens must match;
  * quantized vectors: exact-index tokens within 2 bf16 ulps of the oracle's.
"""
import pytest
import torch

from helpers import golden, max_abs, rel_l2

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("CUDA device required", allow_module_level=True)

from acestep_b200.tokenizer import B200AudioTokenizer, TokShape  # noqa: E402
from oracle import tokenizer as otok  # noqa: E402
from oracle.weights import bf16_round_  # noqa: E402

DEV = torch.device("cuda:0")


def _shape(cfg):
    return TokShape(hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                    num_attention_heads=cfg.num_attention_heads, num_key_value_heads=cfg.num_key_value_heads,
                    head_dim=cfg.head_dim, sliding_window=cfg.sliding_window, rope_theta=cfg.rope_theta,
                    rms_norm_eps=cfg.rms_norm_eps, fsq_dim=cfg.fsq_dim, fsq_input_levels=list(cfg.fsq_input_levels),
                    fsq_input_num_quantizers=cfg.fsq_input_num_quantizers, pool_window_size=cfg.pool_window_size,
                    num_attention_pooler_hidden_layers=cfg.num_attention_pooler_hidden_layers)


@pytest.fixture(scope="module", params=["tiny", "two_quantizers"])
def env(request):
    cfg = otok.TokConfig.tiny() if request.param == "tiny" else otok.TokConfig.tiny(
        fsq_input_levels=[8, 5, 5, 5], fsq_input_num_quantizers=2)
    w = bf16_round_(otok.make_tokenizer_weights(cfg, seed=9))
    tok = B200AudioTokenizer(w, _shape(cfg), DEV)
    yield cfg, w, tok
    tok.close()


def test_fsq_indices_bit_exact_and_quantized(env):
    cfg, w, tok = env
    g = torch.Generator().manual_seed(41)
    B, Tp = 3, 200
    h = (torch.randn(B, Tp, cfg.hidden_size, generator=g) * 1.5).to(torch.bfloat16)
    q, idx = tok.quantize(h.to(DEV))
    torch.cuda.synchronize()
    assert q.shape == (B, Tp, cfg.hidden_size) and idx.shape == (B, Tp, cfg.fsq_input_num_quantizers)
    assert idx.dtype == torch.int32
    # oracle in the reference's bf16 execution: bf16 Linear layers, fp32 FSQ arithmetic
    wb = {k: v.to(torch.bfloat16) for k, v in w.items()}
    want_q, want_idx = otok.residual_fsq(wb, "tokenizer.quantizer.", h, cfg.fsq_input_levels,
                                         cfg.fsq_input_num_quantizers)
    # distance of every bounded value of the FIRST round to its rounding boundary (later rounds inherit the risk)
    lv = torch.tensor(cfg.fsq_input_levels)
    y = torch.nn.functional.linear(h, wb["tokenizer.quantizer.project_in.weight"],
                                   wb["tokenizer.quantizer.project_in.bias"]).float()
    half_l = (lv - 1).float() * (1 + 1e-3) / 2
    offset = torch.where(lv % 2 == 0, 0.5, 0.0)
    bounded = (y + (offset / half_l).atanh()).tanh() * half_l - offset
    safe = ((bounded - bounded.floor() - 0.5).abs() > 0.05).all(dim=-1)
    match = (idx.cpu() == want_idx).all(dim=-1)
    assert int(idx.min()) >= 0 and int(idx.max()) < int(lv.prod())
    assert float(match.float().mean()) >= 0.97, float(match.float().mean())
    if cfg.fsq_input_num_quantizers == 1:
        assert bool(match[safe].all()), int((~match[safe]).sum())
        assert float(safe.float().mean()) > 0.5 and len(set(idx.reshape(-1).tolist())) > 50
    scale = float(want_q.float().abs().max())
    assert max_abs(q.cpu().float()[match], want_q.float()[match]) <= 2 ** -7 * scale


def test_pooler_and_detokenizer_vs_oracle(env):
    cfg, w, tok = env
    g = torch.Generator().manual_seed(42)
    B, Tp, P = 2, 61, cfg.pool_window_size
    x = torch.randn(B, Tp, P, 64, generator=g).to(torch.bfloat16)
    wb = {k: v.to(torch.bfloat16) for k, v in w.items()}
    proj = lambda ww, xx: torch.nn.functional.linear(xx, ww["tokenizer.audio_acoustic_proj.weight"],
                                                     ww["tokenizer.audio_acoustic_proj.bias"])
    want = otok.attention_pooler(w, cfg, "tokenizer.attention_pooler.", proj(w, x.float()))
    floor = rel_l2(otok.attention_pooler(wb, cfg, "tokenizer.attention_pooler.", proj(wb, x)).float(), want)
    got = tok.pool(x.to(DEV))
    torch.cuda.synchronize()
    assert got.shape == (B, Tp, cfg.hidden_size) and torch.isfinite(got.float()).all()
    assert rel_l2(got.cpu().float(), want) <= max(1.5 * floor, 2e-2), (rel_l2(got.cpu().float(), want), floor)

    qz = (torch.randn(B, Tp, cfg.hidden_size, generator=g) * 0.5).to(torch.bfloat16)
    want_d = otok.detokenize(w, cfg, qz.float())
    floor_d = rel_l2(otok.detokenize(wb, cfg, qz).float(), want_d)
    got_d = tok.detokenize(qz.to(DEV))
    torch.cuda.synchronize()
    assert got_d.shape == (B, Tp * P, 64)
    assert rel_l2(got_d.cpu().float(), want_d) <= max(1.5 * floor_d, 2e-2), (rel_l2(got_d.cpu().float(), want_d), floor_d)


def test_tokenize_detokenize_chain_matches_reference_golden():
    """The model-level chain on the fixture of the REAL reference modules (tests/golden/tokenizer.npz): silence padding
    of T = 23 to 25, 5 Hz mask, indices, detokenized hints, crop and is_covers substitution."""
    cfg = otok.TokConfig.tiny()
    w = bf16_round_(otok.make_tokenizer_weights(cfg, seed=9))
    tok = B200AudioTokenizer(w, _shape(cfg), DEV)
    g = golden("tokenizer")
    q, idx, m5 = tok.tokenize(g["x"].to(DEV), g["silence"].to(DEV), g["mask"].to(DEV))
    hints = tok.lm_hints(g["x"].to(DEV), g["silence"].to(DEV), g["mask"].to(DEV), g["src"].shape[1])
    torch.cuda.synchronize()
    assert q.shape == g["quantized"].shape and tuple(idx.shape) == tuple(g["indices"].shape)
    assert torch.equal(m5.cpu().float(), g["mask5"])
    same = (idx.cpu().long() == g["indices"]).all(dim=-1)
    assert float(same.float().mean()) >= 0.8, idx.cpu().reshape(-1).tolist()  # 10 tokens, fp32-weights golden
    assert hints.shape == (2, 23, 64)
    if bool(same.all()):
        assert rel_l2(hints.cpu().float(), g["hints"][:, :23]) <= 4e-2
    tok.close()
