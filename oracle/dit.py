"""Oracle: one DiT velocity prediction (fp32, CPU, functional).

Restates AceStepDiTModel.forward and the layer classes it calls
(/root/reference/acestep/models/turbo/modeling_acestep_v15_turbo.py:1300-1504 model,
:472-536 layer, :286-368 attention, :222-251 timestep embedding, :53-132 masks) together with
the transformers pieces they import (:30-39): Qwen3RMSNorm, Qwen3MLP, Qwen3RotaryEmbedding,
apply_rotary_pos_emb, sdpa attention with repeat_kv.  Weights are a flat dict using the
reference's `decoder.state_dict()` key names so the same tensors load into the real module
(tools/make_golden.py) and into the CUDA packer.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F


@dataclass
class DiTConfig:
    """Subset of AceStepConfig (configuration_acestep_v15.py:148-260) the decoder reads."""

    hidden_size: int = 2048
    intermediate_size: int = 6144
    num_hidden_layers: int = 24
    num_attention_heads: int = 16
    num_key_value_heads: int = 8
    head_dim: int = 128
    in_channels: int = 192
    audio_acoustic_hidden_dim: int = 64
    patch_size: int = 2
    sliding_window: int = 128
    rope_theta: float = 1000000.0
    rms_norm_eps: float = 1e-6
    layer_types: Optional[List[str]] = None

    def __post_init__(self):
        if self.layer_types is None:
            # configuration_acestep_v15.py:251-254 — even layers slide, odd layers are full
            self.layer_types = [
                "sliding_attention" if (i + 1) % 2 else "full_attention"
                for i in range(self.num_hidden_layers)
            ]

    @staticmethod
    def tiny(**kw) -> "DiTConfig":
        base = dict(hidden_size=256, intermediate_size=512, num_hidden_layers=4,
                    num_attention_heads=2, num_key_value_heads=1, head_dim=128, sliding_window=8)
        base.update(kw)
        return DiTConfig(**base)


def rms_norm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    """Qwen3RMSNorm.forward: fp32 variance, x * rsqrt(var + eps), then weight multiply."""
    xf = x.float()
    var = xf.pow(2).mean(-1, keepdim=True)
    return w * (xf * torch.rsqrt(var + eps)).to(x.dtype)


def rope_tables(cfg: DiTConfig, seq_len: int, dtype=torch.float32, device="cpu") -> Tuple[torch.Tensor, torch.Tensor]:
    """Qwen3RotaryEmbedding (default rope): returns cos, sin of shape [S, head_dim].  The tables are always
    built on the CPU (identical values whatever device the oracle runs on) and then moved."""
    d = cfg.head_dim
    inv_freq = 1.0 / (cfg.rope_theta ** (torch.arange(0, d, 2, dtype=torch.int64).float() / d))
    pos = torch.arange(seq_len, dtype=torch.float32)
    freqs = pos[:, None] * inv_freq[None, :]
    emb = torch.cat([freqs, freqs], dim=-1)
    return emb.cos().to(dtype).to(device), emb.sin().to(dtype).to(device)


def _rotate_half(x: torch.Tensor) -> torch.Tensor:
    h = x.shape[-1] // 2
    return torch.cat([-x[..., h:], x[..., :h]], dim=-1)


def apply_rope(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """apply_rotary_pos_emb for x of shape [B, H, S, D]; cos/sin [S, D]."""
    return x * cos[None, None] + _rotate_half(x) * sin[None, None]


def band_mask(seq_len: int, window: int, dtype=torch.float32, device="cpu") -> torch.Tensor:
    """create_4d_mask(..., is_sliding_window=True, is_causal=False) (:92-132): |i-j| <= window."""
    idx = torch.arange(seq_len, device=device)
    keep = (idx[:, None] - idx[None, :]).abs() <= window
    m = torch.full((seq_len, seq_len), torch.finfo(dtype).min, dtype=dtype, device=device)
    return m.masked_fill(keep, 0.0)[None, None]


# "eager" (matmul + fp32 softmax, eager_attention_forward :349-368) or "sdpa" (ALL_ATTENTION_FUNCTIONS["sdpa"]:
# repeat_kv + F.scaled_dot_product_attention with the dense additive mask — what the reference's GPU path runs,
# handler/init_service_loader.py:45-71).  bench.py's torch-gpu baseline arm switches this to "sdpa".
ATTN_IMPL = "eager"


def attention(q, k, v, mask, n_rep: int, scaling: float, probs_out: Optional[list] = None) -> torch.Tensor:
    """sdpa / eager attention with grouped KV heads (repeat_kv): q [B,H,S,D], k/v [B,Hkv,E,D].
    `probs_out` (a list) receives the probabilities, like output_attentions=True (:349-368)."""
    if n_rep > 1:
        k = k.repeat_interleave(n_rep, dim=1)
        v = v.repeat_interleave(n_rep, dim=1)
    if ATTN_IMPL == "sdpa" and probs_out is None:
        return F.scaled_dot_product_attention(q, k, v, attn_mask=mask, scale=scaling)
    s = torch.matmul(q, k.transpose(-1, -2)) * scaling
    if mask is not None:
        s = s + mask
    p = torch.softmax(s, dim=-1, dtype=torch.float32).to(q.dtype)
    if probs_out is not None:
        probs_out.append(p)  # [B, heads, Sq, Skv]: eager_attention_forward's second return value
    return torch.matmul(p, v)


def timestep_sinusoid(t: torch.Tensor, dim: int = 256, scale: float = 1000.0, max_period: float = 10000.0,
                      bf16_time: bool = False):
    """TimestepEmbedding.timestep_embedding (:222-243).  NOTE the `t * scale` happens in t's own
    dtype before the .float() (so a bf16 t gives the bf16-rounded 1000*t).  `bf16_time=True`
    reproduces that quirk of the reference's bf16 execution while the rest stays fp32."""
    t = (t.to(torch.bfloat16) * scale).float() if bf16_time else t * scale
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32) / half).to(t.device)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def timestep_embedding(w: Dict[str, torch.Tensor], prefix: str, t: torch.Tensor, bf16_time: bool = False):
    """TimestepEmbedding.forward (:245-251) -> (temb [B,D], proj [B,6,D])."""
    e = timestep_sinusoid(t, bf16_time=bf16_time).to(t.dtype)
    if bf16_time:
        e = e.to(torch.bfloat16).to(t.dtype)  # t_freq.to(t.dtype) with a bf16 t (:247)
    x = F.linear(e, w[prefix + "linear_1.weight"], w[prefix + "linear_1.bias"])
    x = F.silu(x)
    temb = F.linear(x, w[prefix + "linear_2.weight"], w[prefix + "linear_2.bias"])
    proj = F.linear(F.silu(temb), w[prefix + "time_proj.weight"], w[prefix + "time_proj.bias"])
    return temb, proj.unflatten(1, (6, -1))


def _heads(x: torch.Tensor, head_dim: int) -> torch.Tensor:
    b, s, _ = x.shape
    return x.view(b, s, -1, head_dim).transpose(1, 2)


def cross_kv(w, cfg: DiTConfig, layer: int, enc: torch.Tensor):
    """Cross-attention K/V for one layer (:317-318); `enc` is already condition_embedder'ed."""
    p = f"layers.{layer}.cross_attn."
    k = rms_norm(_heads(F.linear(enc, w[p + "k_proj.weight"]), cfg.head_dim), w[p + "k_norm.weight"], cfg.rms_norm_eps)
    v = _heads(F.linear(enc, w[p + "v_proj.weight"]), cfg.head_dim)
    return k, v


def dit_layer(w, cfg: DiTConfig, i: int, h, tproj, cos, sin, mask, enc_kv, cross_probs: Optional[list] = None) -> torch.Tensor:
    """AceStepDiTLayer.forward (:472-536)."""
    p = f"layers.{i}."
    eps = cfg.rms_norm_eps
    n_rep = cfg.num_attention_heads // cfg.num_key_value_heads
    scaling = cfg.head_dim ** -0.5
    mod = w[p + "scale_shift_table"] + tproj  # [B,6,D]
    shift, scale, gate, c_shift, c_scale, c_gate = mod.chunk(6, dim=1)

    # self attention with AdaLN
    x = rms_norm(h, w[p + "self_attn_norm.weight"], eps) * (1 + scale) + shift
    a = p + "self_attn."
    q = rms_norm(_heads(F.linear(x, w[a + "q_proj.weight"]), cfg.head_dim), w[a + "q_norm.weight"], eps)
    k = rms_norm(_heads(F.linear(x, w[a + "k_proj.weight"]), cfg.head_dim), w[a + "k_norm.weight"], eps)
    v = _heads(F.linear(x, w[a + "v_proj.weight"]), cfg.head_dim)
    q, k = apply_rope(q, cos, sin), apply_rope(k, cos, sin)
    o = attention(q, k, v, mask, n_rep, scaling).transpose(1, 2).reshape(h.shape[0], h.shape[1], -1)
    h = h + F.linear(o, w[a + "o_proj.weight"]) * gate

    # cross attention (no RoPE, no mask: the caller's masks are dropped at :1381-1382)
    x = rms_norm(h, w[p + "cross_attn_norm.weight"], eps)
    c = p + "cross_attn."
    q = rms_norm(_heads(F.linear(x, w[c + "q_proj.weight"]), cfg.head_dim), w[c + "q_norm.weight"], eps)
    ck, cv = enc_kv
    o = attention(q, ck, cv, None, n_rep, scaling, cross_probs).transpose(1, 2).reshape(h.shape[0], h.shape[1], -1)
    h = h + F.linear(o, w[c + "o_proj.weight"])

    # SwiGLU MLP with AdaLN
    x = rms_norm(h, w[p + "mlp_norm.weight"], eps) * (1 + c_scale) + c_shift
    m = p + "mlp."
    y = F.linear(F.silu(F.linear(x, w[m + "gate_proj.weight"])) * F.linear(x, w[m + "up_proj.weight"]),
                 w[m + "down_proj.weight"])
    return h + y * c_gate


@dataclass
class CrossCache:
    """Stands in for EncoderDecoderCache: per-layer cross K/V computed on first use (:307-330)."""

    kv: Dict[int, Tuple[torch.Tensor, torch.Tensor]] = field(default_factory=dict)


def dit_forward(w: Dict[str, torch.Tensor], cfg: DiTConfig, xt: torch.Tensor, t: torch.Tensor,
                ctx: torch.Tensor, enc: torch.Tensor, cache: Optional[CrossCache] = None,
                bf16_time: bool = False, cross_probs: Optional[list] = None) -> torch.Tensor:
    """AceStepDiTModel.forward (:1300-1504) with timestep_r == timestep (inference).

    xt [B,T,64], t [B], ctx [B,T,128], enc [B,E,hidden] -> vt [B,T,64].  `cross_probs` (a list) receives one
    [B, heads, S, E] tensor per layer: `all_cross_attentions` of output_attentions=True (:1448-1482).
    """
    B, T, _ = xt.shape
    temb_t, proj_t = timestep_embedding(w, "time_embed.", t, bf16_time)
    temb_r, proj_r = timestep_embedding(w, "time_embed_r.", t - t, bf16_time)
    temb, tproj = temb_t + temb_r, proj_t + proj_r

    x = torch.cat([ctx, xt], dim=-1)  # [ctx(128) | xt(64)] (:1344)
    ps = cfg.patch_size
    if T % ps:
        x = F.pad(x, (0, 0, 0, ps - T % ps))
    # proj_in: Conv1d(k=patch, stride=patch) over time
    h = F.conv1d(x.transpose(1, 2), w["proj_in.1.weight"], w["proj_in.1.bias"], stride=ps).transpose(1, 2)
    S = h.shape[1]
    enc_e = F.linear(enc, w["condition_embedder.weight"], w["condition_embedder.bias"])

    cos, sin = rope_tables(cfg, S, h.dtype, h.device)
    masks = {"full_attention": None, "sliding_attention": band_mask(S, cfg.sliding_window, h.dtype, h.device)}
    for i in range(cfg.num_hidden_layers):
        if cache is not None and i in cache.kv:
            kv = cache.kv[i]
        else:
            kv = cross_kv(w, cfg, i, enc_e)
            if cache is not None:
                cache.kv[i] = kv
        h = dit_layer(w, cfg, i, h, tproj, cos, sin, masks[cfg.layer_types[i]], kv, cross_probs)

    shift, scale = (w["scale_shift_table"] + temb.unsqueeze(1)).chunk(2, dim=1)
    h = rms_norm(h, w["norm_out.weight"], cfg.rms_norm_eps) * (1 + scale) + shift
    y = F.conv_transpose1d(h.transpose(1, 2), w["proj_out.1.weight"], w["proj_out.1.bias"], stride=ps).transpose(1, 2)
    return y[:, :T, :]
