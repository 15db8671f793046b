"""Standalone hot path: conditioning tensors -> denoising loop -> VAE decode -> waveform.

This is the slice of `AceStepHandler.generate_music` between batch preparation and payload building
(handler/generate_music.py:22-190: service_generate -> _decode_generate_music_pred_latents), with
conditioning supplied as tensors exactly like the backend seam receives it.  It is the public call
bench.py times end to end (host buffers in, host waveform out).
"""
from __future__ import annotations

import time
from typing import Any, Dict, Optional

import torch

from .dit import B200DiT, DiTShape
from .sampler import B200Sampler
from .vae import B200Vae, VaeShape


class B200Pipeline:
    def __init__(self, dit_state: Dict[str, torch.Tensor], vae_state: Dict[str, torch.Tensor],
                 dit_shape: Optional[DiTShape] = None, vae_shape: Optional[VaeShape] = None,
                 null_condition_emb: Optional[torch.Tensor] = None, device="cuda:0", turbo: bool = False):
        self.device = torch.device(device)
        self.turbo = turbo
        self.dit = B200DiT(dit_state, dit_shape or DiTShape(), self.device)
        self.vae = B200Vae(vae_state, vae_shape or VaeShape(), self.device)
        self.sampler = B200Sampler(self.dit, null_condition_emb)
        self.sample_rate = 48000
        self._pinned_wav: Optional[torch.Tensor] = None

    def close(self):
        self.dit.close()
        self.vae.close()

    def _dev(self, x):
        if x is None:
            return None
        return x.to(self.device, torch.bfloat16, non_blocking=True)

    def generate(self, encoder_hidden_states, context_latents, src_latents=None, seed=None, *,
                 noise=None, to_host: bool = True, decode: bool = True, latent_shift: float = 0.0,
                 latent_rescale: float = 1.0, **sampler_kwargs) -> Dict[str, Any]:
        """Returns {"audio": fp32 [B,2,N] (peak-normalised like the reference when |x|max > 1),
        "target_latents": bf16 [B,T,64], "time_costs": {...}}."""
        t0 = time.time()
        enc, ctx = self._dev(encoder_hidden_states), self._dev(context_latents)
        src = self._dev(src_latents) if src_latents is not None else ctx[..., :64].contiguous()
        noise = self._dev(noise)
        fn = self.sampler.generate_turbo if self.turbo else self.sampler.generate_base
        out = fn(enc, ctx, src, seed, noise=noise, **sampler_kwargs)
        lat = out["target_latents"]
        # NaN / Inf / all-zero guard of _prepare_generate_music_decode_state (generate_music_decode.py:66-77)
        if torch.isnan(lat).any() or torch.isinf(lat).any():
            raise RuntimeError("Generation produced NaN or Inf latents.")
        if lat.numel() > 0 and lat.abs().sum() == 0:
            raise RuntimeError("Generation produced zero latents.")
        if latent_shift != 0.0 or latent_rescale != 1.0:
            lat = lat * latent_rescale + latent_shift
        res: Dict[str, Any] = {"target_latents": lat, "time_costs": out["time_costs"]}
        if decode:
            t1 = time.time()
            wav = torch.stack([self.vae.decode_frames(lat[b]) for b in range(lat.shape[0])], dim=0)
            # .float() + per-sample peak normalisation (generate_music_decode.py:191-195)
            peak = wav.abs().amax(dim=[1, 2], keepdim=True)
            if torch.any(peak > 1.0):
                wav = wav / peak.clamp(min=1.0)
            if to_host:
                if self._pinned_wav is None or self._pinned_wav.shape != wav.shape:
                    self._pinned_wav = torch.empty(wav.shape, dtype=torch.float32, pin_memory=True)
                self._pinned_wav.copy_(wav, non_blocking=True)
                torch.cuda.synchronize(self.device)
                wav = self._pinned_wav
            else:
                torch.cuda.synchronize(self.device)
            res["audio"] = wav
            res["time_costs"]["vae_decode_time_cost"] = time.time() - t1
        res["time_costs"]["total_time_cost"] = time.time() - t0
        return res
