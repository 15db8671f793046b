"""Oracle: the condition encoder (fp32, CPU, functional) — SURVEY §8f row 1.

TEST INFRASTRUCTURE ONLY (imported by tests/, tools/make_golden.py and bench.py's CPU legs).

Restates AceStepConditionEncoder.forward (/root/reference/acestep/models/turbo/
modeling_acestep_v15_turbo.py:1506-1552) and what it calls: text_projector (:1518), AceStepLyricEncoder
(:574-728), AceStepTimbreEncoder (:994-1175, incl. unpack_timbre_embeddings :1020-1071),
AceStepEncoderLayer (:371-437), pack_sequences (:135-166) and the padding-aware create_4d_mask
(:53-132).  Unlike the DiT (which nulls its masks, :1381), the lyric encoder DOES apply the key-padding
mask, so attention takes a per-sample valid length.  Weights are a flat dict with the reference's
`encoder.state_dict()` key names.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

from .dit import DiTConfig, _heads, apply_rope, attention, rms_norm, rope_tables


@dataclass
class CondConfig:
    """AceStepConfig fields the condition encoder reads (configuration_acestep_v15.py:148-260)."""

    hidden_size: int = 2048
    intermediate_size: int = 6144
    num_attention_heads: int = 16
    num_key_value_heads: int = 8
    head_dim: int = 128
    sliding_window: int = 128
    rope_theta: float = 1000000.0
    rms_norm_eps: float = 1e-6
    text_hidden_dim: int = 1024
    timbre_hidden_dim: int = 64
    num_lyric_encoder_hidden_layers: int = 8
    num_timbre_encoder_hidden_layers: int = 4

    def layer_type(self, i: int) -> str:
        # configuration_acestep_v15.py:251-254 (shared with the DiT): even layers slide, odd are full
        return "sliding_attention" if (i + 1) % 2 else "full_attention"

    def as_dit(self) -> DiTConfig:
        return DiTConfig(hidden_size=self.hidden_size, intermediate_size=self.intermediate_size,
                         num_attention_heads=self.num_attention_heads, num_key_value_heads=self.num_key_value_heads,
                         head_dim=self.head_dim, sliding_window=self.sliding_window, rope_theta=self.rope_theta,
                         rms_norm_eps=self.rms_norm_eps, num_hidden_layers=1)

    @staticmethod
    def tiny(**kw) -> "CondConfig":
        base = dict(hidden_size=256, intermediate_size=512, num_attention_heads=2, num_key_value_heads=1,
                    head_dim=128, sliding_window=8, text_hidden_dim=128, timbre_hidden_dim=64,
                    num_lyric_encoder_hidden_layers=3, num_timbre_encoder_hidden_layers=2)
        base.update(kw)
        return CondConfig(**base)


def padding_band_mask(seq_len: int, key_valid: Optional[torch.Tensor], window: Optional[int], dtype) -> Optional[torch.Tensor]:
    """create_4d_mask(is_causal=False) (:53-132): additive [B or 1, 1, S, S]; a key is visible when it is
    inside the +-window band (if any) AND not padding (key_valid [B, S] of 0/1)."""
    idx = torch.arange(seq_len)
    keep = torch.ones(seq_len, seq_len, dtype=torch.bool)
    if window is not None:
        keep = (idx[:, None] - idx[None, :]).abs() <= window
    keep = keep[None, None]
    if key_valid is not None:
        keep = keep & key_valid.bool()[:, None, None, :]
    m = torch.full(keep.shape, torch.finfo(dtype).min, dtype=dtype)
    return m.masked_fill(keep, 0.0)


def encoder_layer(w, cfg: CondConfig, p: str, h, cos, sin, mask) -> torch.Tensor:
    """AceStepEncoderLayer.forward (:398-437)."""
    eps, hd = cfg.rms_norm_eps, cfg.head_dim
    n_rep = cfg.num_attention_heads // cfg.num_key_value_heads
    x = rms_norm(h, w[p + "input_layernorm.weight"], eps)
    a = p + "self_attn."
    q = rms_norm(_heads(F.linear(x, w[a + "q_proj.weight"]), hd), w[a + "q_norm.weight"], eps)
    k = rms_norm(_heads(F.linear(x, w[a + "k_proj.weight"]), hd), w[a + "k_norm.weight"], eps)
    v = _heads(F.linear(x, w[a + "v_proj.weight"]), hd)
    q, k = apply_rope(q, cos, sin), apply_rope(k, cos, sin)
    o = attention(q, k, v, mask, n_rep, hd ** -0.5).transpose(1, 2).reshape(h.shape[0], h.shape[1], -1)
    h = h + F.linear(o, w[a + "o_proj.weight"])
    x = rms_norm(h, w[p + "post_attention_layernorm.weight"], eps)
    m = p + "mlp."
    return h + F.linear(F.silu(F.linear(x, w[m + "gate_proj.weight"])) * F.linear(x, w[m + "up_proj.weight"]),
                        w[m + "down_proj.weight"])


def encoder_stack(w, cfg: CondConfig, prefix: str, n_layers: int, x: torch.Tensor,
                  key_valid: Optional[torch.Tensor]) -> torch.Tensor:
    """embed_tokens -> n encoder layers (alternating band / full, both padding-aware) -> final norm."""
    h = F.linear(x, w[prefix + "embed_tokens.weight"], w[prefix + "embed_tokens.bias"])
    S = h.shape[1]
    cos, sin = rope_tables(cfg.as_dit(), S, h.dtype)
    masks = {"full_attention": padding_band_mask(S, key_valid, None, h.dtype) if key_valid is not None else None,
             "sliding_attention": padding_band_mask(S, key_valid, cfg.sliding_window, h.dtype)}
    for i in range(n_layers):
        h = encoder_layer(w, cfg, f"{prefix}layers.{i}.", h, cos, sin, masks[cfg.layer_type(i)])
    return rms_norm(h, w[prefix + "norm.weight"], cfg.rms_norm_eps)


def pack_sequences(h1, h2, m1, m2) -> Tuple[torch.Tensor, torch.Tensor]:
    """pack_sequences (:135-166): concatenate, move valid tokens first (stable), rebuild the mask."""
    hc, mc = torch.cat([h1, h2], dim=1), torch.cat([m1, m2], dim=1)
    B, L, D = hc.shape
    order = mc.argsort(dim=1, descending=True, stable=True)
    packed = torch.gather(hc, 1, order.unsqueeze(-1).expand(B, L, D))
    lengths = mc.sum(dim=1)
    return packed, torch.arange(L).unsqueeze(0) < lengths.unsqueeze(1)


def unpack_timbre(embs: torch.Tensor, order_mask: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """unpack_timbre_embeddings (:1020-1071): packed [N, d] + batch id per row -> [B, max_count, d], mask."""
    N, d = embs.shape
    B = int(order_mask.max().item()) + 1
    counts = torch.bincount(order_mask, minlength=B)
    mx = int(counts.max().item())
    out = torch.zeros(B, mx, d, dtype=embs.dtype)
    mask = torch.zeros(B, mx, dtype=torch.long)
    fill = [0] * B
    for n in range(N):  # rows keep their packed order within each sample
        b = int(order_mask[n])
        out[b, fill[b]] = embs[n]
        mask[b, fill[b]] = 1
        fill[b] += 1
    return out, mask


def condition_encoder(w: Dict[str, torch.Tensor], cfg: CondConfig, text_hs, text_mask, lyric_hs, lyric_mask,
                      refer_packed, refer_order_mask) -> Tuple[torch.Tensor, torch.Tensor]:
    """AceStepConditionEncoder.forward (:1524-1552) -> (encoder_hidden_states [B, Ll+Nt+Lt, D], mask)."""
    text = F.linear(text_hs, w["text_projector.weight"])
    lyric = encoder_stack(w, cfg, "lyric_encoder.", cfg.num_lyric_encoder_hidden_layers, lyric_hs, lyric_mask)
    timbre = encoder_stack(w, cfg, "timbre_encoder.", cfg.num_timbre_encoder_hidden_layers, refer_packed, None)
    t_unpack, t_mask = unpack_timbre(timbre[:, 0, :], refer_order_mask)
    h, m = pack_sequences(lyric, t_unpack, lyric_mask, t_mask)
    return pack_sequences(h, text, m, text_mask)


def make_cond_weights(cfg: CondConfig, seed: int = 5, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Seeded random init with `AceStepConditionEncoder.state_dict()` key names (Linear ~ N(0, 0.02),
    norms perturbed around 1 so a kernel that forgot them fails)."""
    g = torch.Generator().manual_seed(seed)
    D, I, hd = cfg.hidden_size, cfg.intermediate_size, cfg.head_dim
    nq, nkv = cfg.num_attention_heads * hd, cfg.num_key_value_heads * hd
    lin = lambda o, i, std=0.02: torch.randn(o, i, generator=g) * std
    vec = lambda n, std, mean=0.0: mean + torch.randn(n, generator=g) * std
    w: Dict[str, torch.Tensor] = {"text_projector.weight": lin(D, cfg.text_hidden_dim)}
    for name, n_layers, in_dim in (("lyric_encoder.", cfg.num_lyric_encoder_hidden_layers, cfg.text_hidden_dim),
                                   ("timbre_encoder.", cfg.num_timbre_encoder_hidden_layers, cfg.timbre_hidden_dim)):
        w[name + "embed_tokens.weight"] = lin(D, in_dim, 0.05)
        w[name + "embed_tokens.bias"] = vec(D, 0.02)
        w[name + "norm.weight"] = vec(D, 0.1, 1.0)
        for i in range(n_layers):
            p = f"{name}layers.{i}."
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
    w["timbre_encoder.special_token"] = torch.randn(1, 1, D, generator=g)  # present in the state dict, unused (:1084)
    return {k: v.to(dtype) for k, v in w.items()}
