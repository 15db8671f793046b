"""Random-init weights and synthetic conditioning of the reference's shapes (for benchmarks, smoke
tests and demos: there are no checkpoints and no network on the GPU box).

State-dict keys follow the reference modules (`AceStepDiTModel.state_dict()`,
`AutoencoderOobleck.state_dict()`), so the same packers load real checkpoints.  Init scales follow
AceStepPreTrainedModel._init_weights (modeling_acestep_v15_turbo.py:555-571): Linear ~ N(0, 0.02),
RMSNorm = 1, scale_shift_table ~ randn / sqrt(D) (:469, :1296).
"""
from __future__ import annotations

from typing import Dict

import torch

from .dit import DiTShape
from .vae import VaeShape


def random_dit_state(shape: DiTShape, seed: int = 0, device="cpu", dtype=torch.bfloat16) -> Dict[str, torch.Tensor]:
    g = torch.Generator(device=device).manual_seed(seed)
    D, I, hd = shape.hidden_size, shape.intermediate_size, shape.head_dim
    nq, nkv = shape.num_attention_heads * hd, shape.num_key_value_heads * hd

    def rn(*s, std=0.02):
        return (torch.randn(*s, generator=g, device=device, dtype=torch.float32) * std).to(dtype)

    def ones(n):
        return torch.ones(n, device=device, dtype=dtype)

    w: Dict[str, torch.Tensor] = {}
    for i in range(shape.num_hidden_layers):
        p = f"layers.{i}."
        for nm in ("self_attn_norm", "cross_attn_norm", "mlp_norm"):
            w[p + nm + ".weight"] = ones(D)
        for a in ("self_attn.", "cross_attn."):
            w[p + a + "q_proj.weight"] = rn(nq, D)
            w[p + a + "k_proj.weight"] = rn(nkv, D)
            w[p + a + "v_proj.weight"] = rn(nkv, D)
            w[p + a + "o_proj.weight"] = rn(D, nq)
            w[p + a + "q_norm.weight"] = ones(hd)
            w[p + a + "k_norm.weight"] = ones(hd)
        w[p + "mlp.gate_proj.weight"] = rn(I, D)
        w[p + "mlp.up_proj.weight"] = rn(I, D)
        w[p + "mlp.down_proj.weight"] = rn(D, I)
        w[p + "scale_shift_table"] = rn(1, 6, D, std=D ** -0.5)
    w["proj_in.1.weight"] = rn(D, 192, 2, std=0.05)
    w["proj_in.1.bias"] = rn(D)
    for te in ("time_embed.", "time_embed_r."):
        w[te + "linear_1.weight"] = rn(D, 256)
        w[te + "linear_1.bias"] = rn(D)
        w[te + "linear_2.weight"] = rn(D, D)
        w[te + "linear_2.bias"] = rn(D)
        w[te + "time_proj.weight"] = rn(6 * D, D)
        w[te + "time_proj.bias"] = rn(6 * D)
    w["condition_embedder.weight"] = rn(D, D)
    w["condition_embedder.bias"] = rn(D)
    w["norm_out.weight"] = ones(D)
    w["proj_out.1.weight"] = rn(D, 64, 2)
    w["proj_out.1.bias"] = rn(64)
    w["scale_shift_table"] = rn(1, 2, D, std=D ** -0.5)
    return w


def random_vae_state(shape: VaeShape, seed: int = 0, device="cpu") -> Dict[str, torch.Tensor]:
    g = torch.Generator(device=device).manual_seed(seed)
    w: Dict[str, torch.Tensor] = {}

    def conv(name, cout, cin, k, bias=True, transposed=False):
        s = (cin, cout, k) if transposed else (cout, cin, k)
        w[name + ".weight_v"] = torch.randn(*s, generator=g, device=device)
        w[name + ".weight_g"] = 0.6 + 0.2 * torch.rand(s[0], 1, 1, generator=g, device=device)
        if bias:
            w[name + ".bias"] = torch.randn(cout, generator=g, device=device) * 0.02

    def snk(name, c):
        w[name + ".alpha"] = torch.randn(1, c, 1, generator=g, device=device) * 0.2
        w[name + ".beta"] = torch.randn(1, c, 1, generator=g, device=device) * 0.2

    def ru(name, c):
        snk(name + ".snake1", c)
        conv(name + ".conv1", c, c, 7)
        snk(name + ".snake2", c)
        conv(name + ".conv2", c, c, 1)

    cm = [1] + list(shape.channel_multiples)
    H, C = shape.encoder_hidden_size, shape.decoder_channels
    ratios = shape.downsampling_ratios
    n = len(ratios)
    conv("encoder.conv1", H, shape.audio_channels, 7)
    for i, s in enumerate(ratios):
        cin, cout = H * cm[i], H * cm[i + 1]
        for j in (1, 2, 3):
            ru(f"encoder.block.{i}.res_unit{j}", cin)
        snk(f"encoder.block.{i}.snake1", cin)
        conv(f"encoder.block.{i}.conv1", cout, cin, 2 * s)
    snk("encoder.snake1", H * cm[-1])
    conv("encoder.conv2", H, H * cm[-1], 3)
    conv("decoder.conv1", C * cm[-1], shape.decoder_input_channels, 7)
    for i, s in enumerate(ratios[::-1]):
        cin, cout = C * cm[n - i], C * cm[n - i - 1]
        snk(f"decoder.block.{i}.snake1", cin)
        conv(f"decoder.block.{i}.conv_t1", cout, cin, 2 * s, transposed=True)
        for j in (1, 2, 3):
            ru(f"decoder.block.{i}.res_unit{j}", cout)
    snk("decoder.snake1", C)
    conv("decoder.conv2", shape.audio_channels, C, 7, bias=False)
    return w


def synthetic_conditioning(batch: int, frames: int, cond_tokens: int, hidden: int, seed: int = 1234,
                           device="cpu", dtype=torch.bfloat16, pin: bool = False):
    """Text2music-shaped inputs: encoder_hidden_states ~ N(0,1) [B,E,hidden];
    context_latents = [silence-like N(0,1) (64) | chunk mask ones (64)] [B,T,128]."""
    g = torch.Generator(device=device).manual_seed(seed)
    enc = torch.randn(batch, cond_tokens, hidden, generator=g, device=device).to(dtype)
    src = torch.randn(batch, frames, 64, generator=g, device=device).to(dtype)
    ctx = torch.cat([src, torch.ones(batch, frames, 64, device=device, dtype=dtype)], dim=-1)
    null = torch.randn(1, 1, hidden, generator=g, device=device).to(dtype)
    out = {"enc": enc, "ctx": ctx.contiguous(), "src": src, "null_emb": null}
    if pin and str(device) == "cpu":
        out = {k: v.pin_memory() for k, v in out.items()}
    return out
