"""Oracle: the audio tokenizer / residual FSQ / detokenizer of `prepare_condition`'s LM-hint branch — SURVEY §8f row 1.

TEST INFRASTRUCTURE ONLY (imported by tests/ and tools/make_golden_tokenizer.py).

Restates (modeling_acestep_v15_turbo.py):
  * AceStepConditionGenerationModel.tokenize / detokenize (:1577-1600) and the LM-hint branch of prepare_condition
    (:1630-1646): pad T to a multiple of pool_window_size with the silence latent, tokenize at 5 Hz, detokenize back to
    25 Hz, crop, and substitute the result for src_latents where is_covers > 0;
  * AceStepAudioTokenizer (:1178-1218): audio_acoustic_proj Linear(64 -> D) -> AttentionPooler -> ResidualFSQ;
  * AttentionPooler (:730-856): embed_tokens Linear(D -> D), prepend the special token, (B*T/P) sequences of P + 1
    tokens through num_attention_pooler_hidden_layers AceStepEncoderLayers (no padding mask), RMSNorm, token 0;
  * AudioTokenDetokenizer (:859-990): embed_tokens Linear(D -> D), each token repeated P times + special_tokens,
    the same encoder layers over sequences of P tokens, RMSNorm, proj_out Linear(D -> 64).

ResidualFSQ lives in the third-party `vector_quantize_pytorch` (unpinned in the reference's requirements, NOT
installed here and absent from /root/reference), so `ResidualFSQ` below restates its published algorithm
(lucidrains/vector-quantize-pytorch: residual_fsq.py + finite_scalar_quantization.py; FSQ = Mentzer et al. 2023):
    project_in Linear(dim -> len(levels)); per quantizer q (scale_q = (levels - 1)^-q):
        z = residual / scale;  bounded = tanh(z + shift) * half_l - offset,  half_l = (levels - 1)(1 + 1e-3) / 2,
        offset = 0.5 for even levels else 0,  shift = atanh(offset / half_l);
        codes = round(bounded) / (levels // 2);  index = sum((round(bounded) + levels // 2) * basis),
        basis = cumprod([1, levels[:-1]]);  residual -= codes * scale;  out += codes * scale;
    project_out Linear(len(levels) -> dim); the FSQ arithmetic itself runs in fp32 whatever the model dtype.
PARITY STATUS: the surrounding modules are PINNED against the real reference classes (tools/make_golden_tokenizer.py
runs AceStepAudioTokenizer / AudioTokenDetokenizer unmodified with this ResidualFSQ injected as the
`vector_quantize_pytorch` module); the FSQ arithmetic itself is UNPINNED against the absent library.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

from .cond import CondConfig, encoder_layer
from .dit import rms_norm, rope_tables


@dataclass
class TokConfig(CondConfig):
    """AceStepConfig fields of the tokenizer path (configuration_acestep_v15.py:148-260)."""

    audio_acoustic_hidden_dim: int = 64
    pool_window_size: int = 5
    fsq_dim: int = 2048
    fsq_input_levels: List[int] = field(default_factory=lambda: [8, 8, 8, 5, 5, 5])
    fsq_input_num_quantizers: int = 1
    num_attention_pooler_hidden_layers: int = 2

    @staticmethod
    def tiny(**kw) -> "TokConfig":
        base = dict(hidden_size=256, intermediate_size=512, num_attention_heads=2, num_key_value_heads=1,
                    head_dim=128, sliding_window=8, fsq_dim=256, num_attention_pooler_hidden_layers=2)
        base.update(kw)
        return TokConfig(**base)


# ------------------------------------------------------------------------------------------------------------
# FSQ (restated; see the module docstring)
# ------------------------------------------------------------------------------------------------------------
def fsq_quantize(z: torch.Tensor, levels: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """One FSQ layer in fp32: z [..., C] -> (codes in [-1, 1] [..., C], indices [...] int32)."""
    z = z.float()
    lv = levels.to(z.device)
    half_l = (lv - 1).float() * (1 + 1e-3) / 2
    offset = torch.where(lv % 2 == 0, 0.5, 0.0)
    shift = (offset / half_l).atanh()
    bounded = (z + shift).tanh() * half_l - offset
    q = bounded.round()
    half_width = (lv // 2).float()
    basis = torch.cumprod(torch.cat([torch.ones(1, dtype=torch.long), levels[:-1].long().cpu()]), 0).to(z.device)
    idx = ((q + half_width) * basis.float()).sum(dim=-1).to(torch.int32)
    return q / half_width, idx


def residual_fsq(w: Dict[str, torch.Tensor], prefix: str, x: torch.Tensor, levels: List[int],
                 num_quantizers: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """ResidualFSQ.forward: x [..., dim] -> (quantized [..., dim], indices [..., num_quantizers])."""
    lv = torch.tensor(levels)
    y = F.linear(x, w[prefix + "project_in.weight"], w[prefix + "project_in.bias"])
    out = torch.zeros_like(y, dtype=torch.float32)
    residual = y.float()
    all_idx = []
    for qi in range(num_quantizers):
        scale = (lv - 1).float().to(y.device) ** -qi
        codes, idx = fsq_quantize(residual / scale, lv)
        codes = codes.to(y.dtype).float() * scale  # FSQ hands its codes back in the caller's dtype
        residual = residual - codes
        out = out + codes
        all_idx.append(idx)
    q = F.linear(out.to(y.dtype), w[prefix + "project_out.weight"], w[prefix + "project_out.bias"])
    return q, torch.stack(all_idx, dim=-1)


class ResidualFSQ(torch.nn.Module):
    """nn.Module form with the library's constructor signature and parameter names (project_in / project_out), so
    the real AceStepAudioTokenizer (:1190-1194) can be built on it in tools/make_golden_tokenizer.py."""

    def __init__(self, *, dim, levels, num_quantizers, **_):
        super().__init__()
        self.levels, self.num_quantizers = list(levels), num_quantizers
        self.project_in = torch.nn.Linear(dim, len(levels))
        self.project_out = torch.nn.Linear(len(levels), dim)

    def forward(self, x):
        w = {"project_in.weight": self.project_in.weight, "project_in.bias": self.project_in.bias,
             "project_out.weight": self.project_out.weight, "project_out.bias": self.project_out.bias}
        return residual_fsq(w, "", x, self.levels, self.num_quantizers)


# ------------------------------------------------------------------------------------------------------------
def _layers(w, cfg: TokConfig, prefix: str, h: torch.Tensor) -> torch.Tensor:
    """num_attention_pooler_hidden_layers encoder layers over [N, S, D] sequences (no padding mask; S <= window, so
    the sliding and the full mask coincide) + the final RMSNorm."""
    S = h.shape[1]
    cos, sin = rope_tables(cfg.as_dit(), S, h.dtype, h.device)
    for i in range(cfg.num_attention_pooler_hidden_layers):
        h = encoder_layer(w, cfg, f"{prefix}layers.{i}.", h, cos, sin, None)
    return rms_norm(h, w[prefix + "norm.weight"], cfg.rms_norm_eps)


def attention_pooler(w, cfg: TokConfig, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """AttentionPooler.forward (:755-856): x [B, Tp, P, D] -> [B, Tp, D]."""
    B, Tp, P, D = x.shape
    assert P + 1 <= cfg.sliding_window + 1, "pool window wider than the sliding window: masks would differ"
    x = F.linear(x, w[prefix + "embed_tokens.weight"], w[prefix + "embed_tokens.bias"])
    x = torch.cat([w[prefix + "special_token"].to(x.dtype).expand(B, Tp, 1, D), x], dim=2).reshape(B * Tp, P + 1, D)
    return _layers(w, cfg, prefix, x)[:, 0, :].reshape(B, Tp, D)


def tokenizer_forward(w, cfg: TokConfig, x: torch.Tensor, prefix: str = "tokenizer."):
    """AceStepAudioTokenizer.forward (:1200-1213): x [B, Tp, P, 64] -> (quantized [B, Tp, D], indices [B, Tp, nq])."""
    h = F.linear(x, w[prefix + "audio_acoustic_proj.weight"], w[prefix + "audio_acoustic_proj.bias"])
    h = attention_pooler(w, cfg, prefix + "attention_pooler.", h)
    return residual_fsq(w, prefix + "quantizer.", h, cfg.fsq_input_levels, cfg.fsq_input_num_quantizers)


def tokenize(w, cfg: TokConfig, x: torch.Tensor, silence_latent: torch.Tensor, attention_mask: torch.Tensor):
    """AceStepConditionGenerationModel.tokenize (:1577-1588): x [B, T, 64] -> (quantized, indices, 5 Hz mask)."""
    P = cfg.pool_window_size
    if x.shape[1] % P != 0:
        pad = P - x.shape[1] % P
        x = torch.cat([x, silence_latent[:1, :pad].repeat(x.shape[0], 1, 1).to(x.dtype)], dim=1)
        attention_mask = F.pad(attention_mask, (0, pad), mode="constant", value=0)
    x = x.reshape(x.shape[0], x.shape[1] // P, P, x.shape[2])
    seq_len = x.shape[1]
    chunk = math.ceil(attention_mask.shape[1] / seq_len)
    m = F.max_pool1d(attention_mask.to(x.dtype).unsqueeze(1), kernel_size=chunk, stride=chunk, ceil_mode=True).squeeze(1)
    q, idx = tokenizer_forward(w, cfg, x)
    return q, idx, m


def detokenize(w, cfg: TokConfig, quantized: torch.Tensor, prefix: str = "detokenizer.") -> torch.Tensor:
    """AudioTokenDetokenizer.forward (:887-990): quantized [B, Tp, D] -> [B, Tp * P, 64]."""
    B, Tp, D = quantized.shape
    P = cfg.pool_window_size
    x = F.linear(quantized, w[prefix + "embed_tokens.weight"], w[prefix + "embed_tokens.bias"])
    x = x.unsqueeze(2).repeat(1, 1, P, 1) + w[prefix + "special_tokens"].to(x.dtype).expand(B, Tp, -1, -1)
    h = _layers(w, cfg, prefix, x.reshape(B * Tp, P, D))
    h = F.linear(h, w[prefix + "proj_out.weight"], w[prefix + "proj_out.bias"])
    return h.reshape(B, Tp * P, -1)


def lm_hints(w, cfg: TokConfig, hidden_states, silence_latent, attention_mask, src_latents, is_covers):
    """The LM-hint branch of prepare_condition (:1630-1646): returns the src_latents the context is built from."""
    q, _idx, _m = tokenize(w, cfg, hidden_states, silence_latent, attention_mask)
    hints = detokenize(w, cfg, q)[:, : src_latents.shape[1], :]
    return torch.where(is_covers.unsqueeze(-1).unsqueeze(-1) > 0, hints.to(src_latents.dtype), src_latents)


def make_tokenizer_weights(cfg: TokConfig, seed: int = 9, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Seeded random init with the reference model's state_dict key names (`tokenizer.*`, `detokenizer.*`).  The
    FSQ input projection gets a larger gain so the codes spread over the levels instead of sitting at 0."""
    g = torch.Generator().manual_seed(seed)
    D, I, hd = cfg.hidden_size, cfg.intermediate_size, cfg.head_dim
    nq, nkv = cfg.num_attention_heads * hd, cfg.num_key_value_heads * hd
    C = len(cfg.fsq_input_levels)
    lin = lambda o, i, std=0.02: torch.randn(o, i, generator=g) * std
    vec = lambda n, std, mean=0.0: mean + torch.randn(n, generator=g) * std
    w: Dict[str, torch.Tensor] = {}

    def stack(prefix):
        w[prefix + "embed_tokens.weight"] = lin(D, D, 0.05)
        w[prefix + "embed_tokens.bias"] = vec(D, 0.02)
        w[prefix + "norm.weight"] = vec(D, 0.1, 1.0)
        for i in range(cfg.num_attention_pooler_hidden_layers):
            p = f"{prefix}layers.{i}."
            w[p + "input_layernorm.weight"] = vec(D, 0.1, 1.0)
            w[p + "post_attention_layernorm.weight"] = vec(D, 0.1, 1.0)
            w[p + "self_attn.q_proj.weight"] = lin(nq, D)
            w[p + "self_attn.k_proj.weight"] = lin(nkv, D)
            w[p + "self_attn.v_proj.weight"] = lin(nkv, D)
            w[p + "self_attn.o_proj.weight"] = lin(D, nq)
            w[p + "self_attn.q_norm.weight"] = vec(hd, 0.1, 1.0)
            w[p + "self_attn.k_norm.weight"] = vec(hd, 0.1, 1.0)
            w[p + "mlp.gate_proj.weight"] = lin(I, D)
            w[p + "mlp.up_proj.weight"] = lin(I, D)
            w[p + "mlp.down_proj.weight"] = lin(D, I)

    w["tokenizer.audio_acoustic_proj.weight"] = lin(D, cfg.audio_acoustic_hidden_dim, 0.1)
    w["tokenizer.audio_acoustic_proj.bias"] = vec(D, 0.02)
    stack("tokenizer.attention_pooler.")
    w["tokenizer.attention_pooler.special_token"] = torch.randn(1, 1, D, generator=g) * 0.02
    w["tokenizer.quantizer.project_in.weight"] = lin(C, cfg.fsq_dim, 0.08)
    w["tokenizer.quantizer.project_in.bias"] = vec(C, 0.3)
    w["tokenizer.quantizer.project_out.weight"] = lin(cfg.fsq_dim, C, 0.3)
    w["tokenizer.quantizer.project_out.bias"] = vec(cfg.fsq_dim, 0.02)
    stack("detokenizer.")
    w["detokenizer.special_tokens"] = torch.randn(1, cfg.pool_window_size, D, generator=g) * 0.02
    w["detokenizer.proj_out.weight"] = lin(cfg.audio_acoustic_hidden_dim, D, 0.05)
    w["detokenizer.proj_out.bias"] = vec(cfg.audio_acoustic_hidden_dim, 0.02)
    return {k: v.to(dtype) for k, v in w.items()}
