"""Host-side Oobleck VAE engine (decode / encode) and the reference-facing tiled_* mirrors.

Replaces `self.vae.decode(...)` / `self.vae.encode(...).latent_dist.sample()` behind
AceStepHandler.tiled_decode / tiled_encode (handler/vae_decode.py:16-48, vae_encode.py:15-43) the
same way the MLX backend does (handler/mlx_vae_decode_native.py:31-76).
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch

from . import _lib
from .pack import pack_vae


@dataclass
class VaeShape:
    """AutoencoderOobleck config (shipped ACE-Step values; hop 1920)."""

    encoder_hidden_size: int = 128
    downsampling_ratios: List[int] = field(default_factory=lambda: [2, 4, 4, 6, 10])
    channel_multiples: List[int] = field(default_factory=lambda: [1, 2, 4, 8, 16])
    decoder_channels: int = 128
    decoder_input_channels: int = 64
    audio_channels: int = 2

    @property
    def hop(self) -> int:
        return math.prod(self.downsampling_ratios)

    @classmethod
    def from_config(cls, cfg) -> "VaeShape":
        return cls(encoder_hidden_size=cfg.encoder_hidden_size,
                   downsampling_ratios=list(cfg.downsampling_ratios),
                   channel_multiples=list(cfg.channel_multiples), decoder_channels=cfg.decoder_channels,
                   decoder_input_channels=cfg.decoder_input_channels, audio_channels=cfg.audio_channels)


class B200Vae:
    """Whole-song (untiled) decode/encode; one launch sequence per sample."""

    # The codec's receptive field is < 10 latent frames each side, so decoding a
    # song in one pass equals the reference's 512/64 overlap-discard tiling in every kept core.
    MAX_FRAMES_PER_PASS = 8192  # ~5.5 min; longer inputs are tiled with HALO_FRAMES of overlap
    HALO_FRAMES = 16

    def __init__(self, state_dict: Dict[str, torch.Tensor], shape: VaeShape, device="cuda:0", lib=None):
        self.lib = lib or _lib.load()  # `lib`: tests / tools hand in the probe build (_lib.load_probe())
        self.device = torch.device(device)
        self.shape = shape
        if self.device.type != "cuda":
            raise _lib.B200Error("B200Vae needs a CUDA device (there is no CPU path)")
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ace_init(self.device.index or 0), "ace_init", self.lib)
            cfg = _lib.AceVaeConfig()
            cfg.num_stages = len(shape.downsampling_ratios)
            for i, (r, m) in enumerate(zip(shape.downsampling_ratios, shape.channel_multiples)):
                cfg.ratios[i], cfg.channel_multiples[i] = r, m
            cfg.encoder_hidden, cfg.decoder_channels = shape.encoder_hidden_size, shape.decoder_channels
            cfg.latent_channels, cfg.audio_channels = shape.decoder_input_channels, shape.audio_channels
            blob = pack_vae(state_dict, shape.downsampling_ratios, shape.channel_multiples,
                            shape.encoder_hidden_size, shape.decoder_channels)
            expect = self.lib.ace_vae_packed_bytes(C.byref(cfg))
            if blob.numel() != expect:
                raise _lib.B200Error(f"packed VAE blob has {blob.numel()} bytes, library expects {expect}")
            handle = C.c_void_p()
            _lib.check(self.lib.ace_vae_create(C.byref(handle), C.byref(cfg), blob.data_ptr(), blob.numel()), "ace_vae_create", self.lib)
            self.handle = handle
        self._ws: Optional[torch.Tensor] = None

    def close(self):
        if getattr(self, "handle", None):
            self.lib.ace_vae_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _workspace(self, nbytes: int) -> torch.Tensor:
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        return self._ws

    # ------------------------------------------------------------------
    def decode_frames(self, z_tc: torch.Tensor) -> torch.Tensor:
        """z_tc [T, 64] (time-major, the sampler's layout) -> wav [2, T*hop] fp32 on the device."""
        T = z_tc.shape[0]
        z = z_tc.to(device=self.device, dtype=torch.bfloat16).contiguous()
        wav = torch.empty(2, T * self.shape.hop, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            need = self.lib.ace_vae_decode_workspace_bytes(self.handle, T)
            ws = self._workspace(need)
            _lib.check(self.lib.ace_vae_decode(self.handle, z.data_ptr(), T, wav.data_ptr(), ws.data_ptr(),
                                               ws.numel(), _lib.stream_handle(self.device)), "ace_vae_decode", self.lib)
        return wav

    def decode(self, latents: torch.Tensor) -> torch.Tensor:
        """latents [B, 64, T] (the reference's vae.decode layout) -> [B, 2, T*hop] fp32."""
        if latents.dim() != 3:
            raise ValueError(f"decode expects [B, C, T], got {tuple(latents.shape)}")
        outs = []
        for b in range(latents.shape[0]):
            outs.append(self._decode_long(latents[b].transpose(0, 1)))
        return torch.stack(outs, dim=0)

    def decode_normalized(self, latents: torch.Tensor) -> torch.Tensor:
        """decode() followed by the caller's per-sample peak normalisation (generate_music_decode.py:191-195),
        in place on the device: samples whose max|x| exceeds 1 are divided by it."""
        from .output import peak_normalize_

        wav = self.decode(latents)
        peak_normalize_(wav)
        return wav

    def _decode_long(self, z_tc: torch.Tensor) -> torch.Tensor:
        T = z_tc.shape[0]
        if T <= self.MAX_FRAMES_PER_PASS:
            return self.decode_frames(z_tc)
        hop, halo = self.shape.hop, self.HALO_FRAMES
        stride = self.MAX_FRAMES_PER_PASS - 2 * halo
        parts = []
        for c0 in range(0, T, stride):
            c1 = min(c0 + stride, T)
            w0, w1 = max(0, c0 - halo), min(T, c1 + halo)
            a = self.decode_frames(z_tc[w0:w1])
            parts.append(a[:, (c0 - w0) * hop: a.shape[1] - (w1 - c1) * hop])
        return torch.cat(parts, dim=1)

    def encode_samples(self, wav: torch.Tensor, eps_tc: Optional[torch.Tensor]) -> torch.Tensor:
        """wav [2, N] fp32 (N % hop == 0), eps_tc [N/hop, 64] or None -> z [N/hop, 64] bf16."""
        N = wav.shape[1]
        wav = wav.to(device=self.device, dtype=torch.float32).contiguous()
        T = N // self.shape.hop
        z = torch.empty(T, 64, dtype=torch.bfloat16, device=self.device)
        if eps_tc is not None:
            eps_tc = eps_tc.to(device=self.device, dtype=torch.bfloat16).contiguous()
        with torch.cuda.device(self.device):
            need = self.lib.ace_vae_encode_workspace_bytes(self.handle, N)
            ws = self._workspace(need)
            _lib.check(self.lib.ace_vae_encode(self.handle, wav.data_ptr(), N, _lib.ptr(eps_tc), z.data_ptr(),
                                               ws.data_ptr(), ws.numel(), _lib.stream_handle(self.device)), "ace_vae_encode", self.lib)
        return z

    # ------------------------------------------------------------------
    # Posterior-moments path: the encoder output (mean | scale per frame) is a deterministic function of the audio,
    # only the noise of `latent_dist.sample()` is fresh per call.  Callers that encode the same reference / source
    # audio over and over (handler/conditioning_embed.py:18-69 re-encodes every reference of every request) keep the
    # moments and re-draw the sample: same distribution and same RNG consumption as encoding again.
    def encode_moments(self, wav: torch.Tensor) -> torch.Tensor:
        """wav [2, N] fp32 (N % hop == 0) -> moments [N/hop, 128] bf16 (mean | scale)."""
        N = wav.shape[1]
        wav = wav.to(device=self.device, dtype=torch.float32).contiguous()
        T = N // self.shape.hop
        mom = torch.empty(T, 2 * self.shape.decoder_input_channels, dtype=torch.bfloat16, device=self.device)
        with torch.cuda.device(self.device):
            need = self.lib.ace_vae_encode_workspace_bytes(self.handle, N)
            ws = self._workspace(need)
            _lib.check(self.lib.ace_vae_encode_moments(self.handle, wav.data_ptr(), N, mom.data_ptr(), ws.data_ptr(),
                                                       ws.numel(), _lib.stream_handle(self.device)),
                       "ace_vae_encode_moments", self.lib)
        return mom

    def posterior_sample(self, moments: torch.Tensor, eps_tc: Optional[torch.Tensor]) -> torch.Tensor:
        """moments [T, 128], eps_tc [T, 64] or None (the mean) -> z [T, 64] bf16; bit-identical to encode_samples."""
        T = moments.shape[0]
        z = torch.empty(T, self.shape.decoder_input_channels, dtype=torch.bfloat16, device=self.device)
        if eps_tc is not None:
            eps_tc = eps_tc.to(device=self.device, dtype=torch.bfloat16).contiguous()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ace_vae_posterior_sample(self.handle, moments.data_ptr(), _lib.ptr(eps_tc), z.data_ptr(),
                                                         T, _lib.stream_handle(self.device)),
                       "ace_vae_posterior_sample", self.lib)
        return z

    MOMENT_CACHE_ENTRIES = 16

    @staticmethod
    def _fingerprint(a: torch.Tensor):
        """128-bit content fingerprint computed on the device (two wrapping 64-bit sums over the raw fp32 words, the
        second position-weighted) + the shape.  A cache key, not a cryptographic hash: a collision needs two audio
        clips of identical length whose word sums and index-weighted word sums both agree modulo 2^64."""
        w = a.contiguous().view(torch.int32).reshape(-1).to(torch.int64)
        idx = torch.arange(1, w.numel() + 1, device=w.device, dtype=torch.int64)
        both = torch.stack([w.sum(), (w * idx).sum()]).tolist()
        return (tuple(a.shape), both[0], both[1])

    def encode_cached(self, audio: torch.Tensor, sample: bool = True, generator: Optional[torch.Generator] = None):
        """encode() with a device-resident cache of posterior moments keyed by audio content: a repeated clip skips
        the encoder (5-6 ms per 30 s reference on B200) and only re-draws the posterior sample.  Same returns as
        encode(): [B, 64, N // hop] bf16."""
        if audio.dim() == 2:
            audio = audio.unsqueeze(0)
        hop = self.shape.hop
        if not hasattr(self, "_moments"):
            self._moments, self.moment_cache_hits, self.moment_cache_misses = {}, 0, 0
        outs = []
        for b in range(audio.shape[0]):
            T = audio.shape[2] // hop
            if T < 1:
                raise ValueError(f"audio of {audio.shape[2]} samples is shorter than one hop ({hop})")
            a = audio[b, :, : T * hop].to(device=self.device, dtype=torch.float32)
            key = self._fingerprint(a)
            mom = self._moments.pop(key, None)
            if mom is None:
                self.moment_cache_misses += 1
                mom = self.encode_moments(a)
                while len(self._moments) >= self.MOMENT_CACHE_ENTRIES:
                    self._moments.pop(next(iter(self._moments)))  # least recently used
            else:
                self.moment_cache_hits += 1
            self._moments[key] = mom  # most recently used last
            eps = torch.randn(T, 64, device=self.device, dtype=torch.bfloat16, generator=generator) if sample else None
            outs.append(self.posterior_sample(mom, eps).transpose(0, 1))
        return torch.stack(outs, dim=0)

    def encode(self, audio: torch.Tensor, sample: bool = True, generator: Optional[torch.Generator] = None):
        """audio [B, 2, N] -> latents [B, 64, N // hop] (bf16); `sample` draws the posterior noise
        with torch's RNG on the device like latent_dist.sample()."""
        if audio.dim() == 2:
            audio = audio.unsqueeze(0)
        hop = self.shape.hop
        outs = []
        for b in range(audio.shape[0]):
            a = audio[b]
            # each strided conv floors its output length, so the reference yields N // hop frames;
            # samples past the last whole hop only touch the final frame's right edge — crop them.
            T = a.shape[1] // hop
            if T < 1:
                raise ValueError(f"audio of {a.shape[1]} samples is shorter than one hop ({hop})")
            a = a[:, : T * hop]
            eps = None
            if sample:
                eps = torch.randn(T, 64, device=self.device, dtype=torch.bfloat16, generator=generator)
            outs.append(self.encode_samples(a, eps).transpose(0, 1))
        return torch.stack(outs, dim=0)
