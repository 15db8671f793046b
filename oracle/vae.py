"""Oracle: Oobleck VAE decode / encode and the handler's overlap-discard tiling (fp32, CPU).

The codec arithmetic is `diffusers.AutoencoderOobleck` (third-party, version unpinned in the
reference's pyproject.toml:27, NOT installed here and not vendored under /root/reference), so
this restates its published structure as mirrored in-tree by the MLX backend:
  * Snake1d                x + 1/(exp(beta)+1e-9) * sin(exp(alpha)*x)^2   (mlx/vae_model.py:24-56)
  * residual unit          snake -> conv k7 (dilation d, pad 3d) -> snake -> conv k1, + x (:62-87)
  * decoder block          snake -> ConvTranspose1d(k=2s, stride s, pad ceil(s/2)) -> 3 units d=1,3,9 (:119-142)
  * encoder block          3 units -> snake -> Conv1d(k=2s, stride s, pad ceil(s/2))            (:94-116)
  * decoder / encoder      (:149-230);  posterior sample mean + (softplus(scale)+1e-4)*eps     (:285-304)
  * weight norm            w = g * v / ||v||  (norm over all dims but 0)                      (mlx/vae_convert.py:18-34)
Tiling follows acestep/core/generation/handler/vae_decode_chunks.py:13-112 and
vae_encode.py:15-82 / vae_encode_chunks.py:10-41.

PINNED against the reference's in-tree implementation: tools/make_golden_vae_mlx.py executes the reference's own
`MLXAutoEncoderOobleck` + `convert_vae_weights`, unmodified, through a torch-backed stand-in for the `mlx`
primitives (tools/mlx_shim.py: conv1d / conv_transpose1d in NLC with MLX weight layouts, elementwise ops), and
tests/test_oracle_golden.py requires this file to reproduce its decode, encoder moments and mean to 1e-4 (measured
0.5-1.8e-5), including a config with odd strides.  What that pins: structure, paddings, strides, dilations, Snake,
weight-norm fusion, weight axis conventions, mean / scale split.  What it does not: `diffusers` itself (absent), so
a divergence between diffusers and the reference's own MLX mirror of it would go unseen.
Weights use the diffusers state_dict key names (encoder.block.{i}.res_unit{j}.conv1.weight_g ...).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List

import torch
import torch.nn.functional as F


@dataclass
class VaeConfig:
    """AutoencoderOobleck config.  Shipped ACE-Step values: hop 1920 = 2*4*4*6*10."""

    encoder_hidden_size: int = 128
    downsampling_ratios: List[int] = field(default_factory=lambda: [2, 4, 4, 6, 10])
    channel_multiples: List[int] = field(default_factory=lambda: [1, 2, 4, 8, 16])
    decoder_channels: int = 128
    decoder_input_channels: int = 64
    audio_channels: int = 2

    @property
    def hop(self) -> int:
        return math.prod(self.downsampling_ratios)

    @staticmethod
    def tiny() -> "VaeConfig":
        return VaeConfig(encoder_hidden_size=128, downsampling_ratios=[2, 4], channel_multiples=[1, 2],
                         decoder_channels=128, decoder_input_channels=64, audio_channels=2)


def fuse_weight_norm(g: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """torch weight_norm (dim=0): w = g * v / ||v||, norm over every dim except 0."""
    nrm = v.flatten(1).norm(dim=1).view(-1, *([1] * (v.dim() - 1)))
    return g * v / nrm


def _w(w: Dict[str, torch.Tensor], name: str) -> torch.Tensor:
    if name + ".weight" in w:
        return w[name + ".weight"]
    return fuse_weight_norm(w[name + ".weight_g"], w[name + ".weight_v"])


def _b(w, name):
    return w.get(name + ".bias")


def snake(x: torch.Tensor, alpha: torch.Tensor, beta: torch.Tensor) -> torch.Tensor:
    a = torch.exp(alpha).view(1, -1, 1)
    b = torch.exp(beta).view(1, -1, 1)
    return x + (b + 1e-9).reciprocal() * torch.sin(a * x).pow(2)


def _res_unit(w, p: str, x: torch.Tensor, dilation: int) -> torch.Tensor:
    y = snake(x, w[p + "snake1.alpha"], w[p + "snake1.beta"])
    y = F.conv1d(y, _w(w, p + "conv1"), _b(w, p + "conv1"), dilation=dilation, padding=3 * dilation)
    y = snake(y, w[p + "snake2.alpha"], w[p + "snake2.beta"])
    y = F.conv1d(y, _w(w, p + "conv2"), _b(w, p + "conv2"))
    return x + y


def decode(w: Dict[str, torch.Tensor], cfg: VaeConfig, z: torch.Tensor) -> torch.Tensor:
    """z [B, 64, L] -> audio [B, 2, L*hop]   (OobleckDecoder)."""
    p = "decoder."
    x = F.conv1d(z, _w(w, p + "conv1"), _b(w, p + "conv1"), padding=3)
    for i, s in enumerate(cfg.downsampling_ratios[::-1]):
        b = f"{p}block.{i}."
        x = snake(x, w[b + "snake1.alpha"], w[b + "snake1.beta"])
        x = F.conv_transpose1d(x, _w(w, b + "conv_t1"), _b(w, b + "conv_t1"), stride=s, padding=math.ceil(s / 2))
        for j, d in enumerate((1, 3, 9)):
            x = _res_unit(w, f"{b}res_unit{j + 1}.", x, d)
    x = snake(x, w[p + "snake1.alpha"], w[p + "snake1.beta"])
    return F.conv1d(x, _w(w, p + "conv2"), None, padding=3)


def encode_moments(w, cfg: VaeConfig, audio: torch.Tensor):
    """audio [B, 2, N] -> (mean, scale) each [B, 64, N/hop]   (OobleckEncoder + chunk)."""
    p = "encoder."
    x = F.conv1d(audio, _w(w, p + "conv1"), _b(w, p + "conv1"), padding=3)
    for i, s in enumerate(cfg.downsampling_ratios):
        b = f"{p}block.{i}."
        for j, d in enumerate((1, 3, 9)):
            x = _res_unit(w, f"{b}res_unit{j + 1}.", x, d)
        x = snake(x, w[b + "snake1.alpha"], w[b + "snake1.beta"])
        x = F.conv1d(x, _w(w, b + "conv1"), _b(w, b + "conv1"), stride=s, padding=math.ceil(s / 2))
    x = snake(x, w[p + "snake1.alpha"], w[p + "snake1.beta"])
    x = F.conv1d(x, _w(w, p + "conv2"), _b(w, p + "conv2"), padding=1)
    return x.chunk(2, dim=1)


def encode_sample(w, cfg: VaeConfig, audio: torch.Tensor, eps: torch.Tensor) -> torch.Tensor:
    """latent_dist.sample(): mean + (softplus(scale) + 1e-4) * eps."""
    mean, scale = encode_moments(w, cfg, audio)
    return mean + (F.softplus(scale) + 1e-4) * eps


# ------------------------------------------------------------------------------------------
# Overlap-discard tiling, restated from the handler mixins
# ------------------------------------------------------------------------------------------
def tiled_decode(decode_fn, latents: torch.Tensor, chunk_size: int = 512, overlap: int = 64) -> torch.Tensor:
    """_tiled_decode_inner / _tiled_decode_gpu (vae_decode_chunks.py:13-112); latents [B,C,T]."""
    B, _, T = latents.shape
    if B > 1:
        return torch.cat([tiled_decode(decode_fn, latents[b:b + 1], chunk_size, overlap) for b in range(B)], 0)
    while chunk_size - 2 * overlap <= 0 and overlap > 0:
        overlap //= 2
    if T <= chunk_size:
        return decode_fn(latents)
    stride = chunk_size - 2 * overlap
    parts, up = [], None
    for i in range(math.ceil(T / stride)):
        c0 = i * stride
        c1 = min(c0 + stride, T)
        w0, w1 = max(0, c0 - overlap), min(T, c1 + overlap)
        a = decode_fn(latents[:, :, w0:w1])
        if up is None:
            up = a.shape[-1] / (w1 - w0)
        t0 = int(round((c0 - w0) * up))
        t1 = int(round((w1 - c1) * up))
        parts.append(a[:, :, t0:a.shape[-1] - t1 if t1 > 0 else a.shape[-1]])
    return torch.cat(parts, dim=-1)


def tiled_encode(encode_fn, audio: torch.Tensor, chunk_size: int = 48000 * 30, overlap: int = 48000 * 2):
    """tiled_encode / _tiled_encode_gpu (vae_encode.py:45-82, vae_encode_chunks.py:10-41).
    encode_fn(audio_chunk, win_start) -> latents for that chunk."""
    N = audio.shape[-1]
    if N <= chunk_size:
        return encode_fn(audio, 0)
    stride = chunk_size - 2 * overlap
    parts, down = [], None
    for i in range(math.ceil(N / stride)):
        c0 = i * stride
        c1 = min(c0 + stride, N)
        w0, w1 = max(0, c0 - overlap), min(N, c1 + overlap)
        z = encode_fn(audio[:, :, w0:w1], w0)
        if down is None:
            down = (w1 - w0) / z.shape[-1]
        t0 = int(round((c0 - w0) / down))
        t1 = int(round((w1 - c1) / down))
        parts.append(z[:, :, t0:z.shape[-1] - t1 if t1 > 0 else z.shape[-1]])
    return torch.cat(parts, dim=-1)
