"""Seeded random-init weight factories shared by the oracle, the golden generator and the tests.

There are no checkpoints in the container or on the GPU box (no network), so parity and
throughput work uses random weights of the reference architecture:
  * DiT: Linear ~ N(0, 0.02) like AceStepPreTrainedModel._init_weights
    (modeling_acestep_v15_turbo.py:555-571), scale_shift_table ~ randn/sqrt(D) (:469, :1296);
    RMSNorm weights and biases get a small random perturbation (instead of the 1 / 0 init) so a
    kernel that forgot them fails the test.
  * VAE: weight-norm (g, v) pairs, Snake alpha/beta ~ N(0, 0.3), small biases.
Keys follow the reference module state_dict names so the same dict loads into the real
modules (tools/make_golden.py) and into the CUDA weight packer.
"""
from __future__ import annotations

from typing import Dict

import torch

from .dit import DiTConfig
from .vae import VaeConfig


def make_dit_weights(cfg: DiTConfig, seed: int = 0, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    D, I, hd = cfg.hidden_size, cfg.intermediate_size, cfg.head_dim
    nq, nkv = cfg.num_attention_heads * hd, cfg.num_key_value_heads * hd

    def lin(o, i, std=0.02):
        return torch.randn(o, i, generator=g) * std

    def vec(n, std, mean=0.0):
        return mean + torch.randn(n, generator=g) * std

    w: Dict[str, torch.Tensor] = {}
    for i in range(cfg.num_hidden_layers):
        p = f"layers.{i}."
        w[p + "self_attn_norm.weight"] = vec(D, 0.1, 1.0)
        w[p + "cross_attn_norm.weight"] = vec(D, 0.1, 1.0)
        w[p + "mlp_norm.weight"] = vec(D, 0.1, 1.0)
        for a in ("self_attn.", "cross_attn."):
            w[p + a + "q_proj.weight"] = lin(nq, D)
            w[p + a + "k_proj.weight"] = lin(nkv, D)
            w[p + a + "v_proj.weight"] = lin(nkv, D)
            w[p + a + "o_proj.weight"] = lin(D, nq)
            w[p + a + "q_norm.weight"] = vec(hd, 0.1, 1.0)
            w[p + a + "k_norm.weight"] = vec(hd, 0.1, 1.0)
        w[p + "mlp.gate_proj.weight"] = lin(I, D)
        w[p + "mlp.up_proj.weight"] = lin(I, D)
        w[p + "mlp.down_proj.weight"] = lin(D, I)
        w[p + "scale_shift_table"] = torch.randn(1, 6, D, generator=g) / D ** 0.5
    w["proj_in.1.weight"] = torch.randn(D, cfg.in_channels, cfg.patch_size, generator=g) * 0.05
    w["proj_in.1.bias"] = vec(D, 0.02)
    for te in ("time_embed.", "time_embed_r."):
        w[te + "linear_1.weight"] = lin(D, 256)
        w[te + "linear_1.bias"] = vec(D, 0.02)
        w[te + "linear_2.weight"] = lin(D, D)
        w[te + "linear_2.bias"] = vec(D, 0.02)
        w[te + "time_proj.weight"] = lin(6 * D, D)
        w[te + "time_proj.bias"] = vec(6 * D, 0.02)
    w["condition_embedder.weight"] = lin(D, D)
    w["condition_embedder.bias"] = vec(D, 0.02)
    w["norm_out.weight"] = vec(D, 0.1, 1.0)
    w["proj_out.1.weight"] = torch.randn(D, cfg.audio_acoustic_hidden_dim, cfg.patch_size, generator=g) * 0.02
    w["proj_out.1.bias"] = vec(cfg.audio_acoustic_hidden_dim, 0.02)
    w["scale_shift_table"] = torch.randn(1, 2, D, generator=g) / D ** 0.5
    return {k: v.to(dtype) for k, v in w.items()}


def make_null_condition_emb(cfg: DiTConfig, seed: int = 1, dtype=torch.float32) -> torch.Tensor:
    """AceStepConditionGenerationModel.null_condition_emb ~ randn(1,1,D) (turbo modeling :1572)."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(1, 1, cfg.hidden_size, generator=g).to(dtype)


def make_vae_weights(cfg: VaeConfig, seed: int = 0, dtype=torch.float32, gain: float = 1.0) -> Dict[str, torch.Tensor]:
    """`gain` scales every weight-norm g (same RNG stream for every gain).  gain = 1 keeps activations O(1..10), where
    the Snake stack amplifies bf16 rounding noise to several percent; a low gain (0.35) keeps the stack near its
    linear regime, so a bf16 execution stays within ~5e-3 of fp32 and parity tests can use a tight bound."""
    g = torch.Generator().manual_seed(seed)
    w: Dict[str, torch.Tensor] = {}

    def conv(name, cout, cin, k, bias=True, transposed=False):
        shape = (cin, cout, k) if transposed else (cout, cin, k)
        v = torch.randn(*shape, generator=g)
        # gain chosen so activations stay O(1) through the stack: w rows have norm ~ g
        w[name + ".weight_v"] = v
        w[name + ".weight_g"] = (0.8 + 0.4 * torch.rand(shape[0], 1, 1, generator=g)) * gain
        if bias:
            w[name + ".bias"] = torch.randn(cout, generator=g) * 0.05

    def snk(name, c):
        w[name + ".alpha"] = torch.randn(1, c, 1, generator=g) * 0.3
        w[name + ".beta"] = torch.randn(1, c, 1, generator=g) * 0.3

    def res_unit(name, c):
        snk(name + ".snake1", c)
        conv(name + ".conv1", c, c, 7)
        snk(name + ".snake2", c)
        conv(name + ".conv2", c, c, 1)

    cm = [1] + list(cfg.channel_multiples)
    H = cfg.encoder_hidden_size
    conv("encoder.conv1", H, cfg.audio_channels, 7)
    for i, s in enumerate(cfg.downsampling_ratios):
        cin, cout = H * cm[i], H * cm[i + 1]
        for j in (1, 2, 3):
            res_unit(f"encoder.block.{i}.res_unit{j}", cin)
        snk(f"encoder.block.{i}.snake1", cin)
        conv(f"encoder.block.{i}.conv1", cout, cin, 2 * s)
    snk("encoder.snake1", H * cm[-1])
    conv("encoder.conv2", H, H * cm[-1], 3)

    C = cfg.decoder_channels
    ups = cfg.downsampling_ratios[::-1]
    n = len(ups)
    conv("decoder.conv1", C * cm[-1], cfg.decoder_input_channels, 7)
    for i, s in enumerate(ups):
        cin, cout = C * cm[n - i], C * cm[n - i - 1]
        snk(f"decoder.block.{i}.snake1", cin)
        conv(f"decoder.block.{i}.conv_t1", cout, cin, 2 * s, transposed=True)
        for j in (1, 2, 3):
            res_unit(f"decoder.block.{i}.res_unit{j}", cout)
    snk("decoder.snake1", C)
    conv("decoder.conv2", cfg.audio_channels, C, 7, bias=False)
    return {k: v.to(dtype) for k, v in w.items()}


def bf16_round_(w: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Round every tensor through bf16 (what the CUDA path stores) and return fp32 copies."""
    return {k: v.to(torch.bfloat16).to(torch.float32) for k, v in w.items()}
