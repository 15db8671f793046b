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
from .output import check_latents, peak_normalize_
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
                 noise=None, to_host: bool = True, reuse_host_buffer: bool = False, decode: bool = True,
                 latent_shift: float = 0.0,
                 latent_rescale: float = 1.0, normalization_db: Optional[float] = None,
                 **sampler_kwargs) -> Dict[str, Any]:
        """Returns {"audio": fp32 [B,2,N] (peak-normalised like the reference when |x|max > 1),
        "target_latents": bf16 [B,T,64], "peak": fp32 [B] raw peaks, "time_costs": {...}}.
        `normalization_db` (e.g. -1.0, GenerationParams.normalization_db) also applies the front-end's
        `normalize_audio` per song (inference.py:674-679) inside the same device pass, so the host copy that
        comes back is the final audio and the reference's three host passes over it are not needed.
        With `to_host` the waveform comes back in a pinned host tensor owned by the caller (a fresh one per call);
        `reuse_host_buffer=True` returns the pipeline's single reusable pinned buffer instead — valid only until
        the next call (serving loops / the benchmark, which consume each result before asking for the next)."""
        t0 = time.time()
        enc, ctx = self._dev(encoder_hidden_states), self._dev(context_latents)
        src = self._dev(src_latents) if src_latents is not None else ctx[..., :64].contiguous()
        noise = self._dev(noise)
        fn = self.sampler.generate_turbo if self.turbo else self.sampler.generate_base
        out = fn(enc, ctx, src, seed, noise=noise, **sampler_kwargs)
        lat = out["target_latents"]
        # NaN / Inf / all-zero guard of _prepare_generate_music_decode_state (generate_music_decode.py:66-77)
        check_latents(lat)
        if latent_shift != 0.0 or latent_rescale != 1.0:
            lat = lat * latent_rescale + latent_shift
        res: Dict[str, Any] = {"target_latents": lat, "time_costs": out["time_costs"]}
        if decode:
            t1 = time.time()
            wav = torch.stack([self.vae.decode_frames(lat[b]) for b in range(lat.shape[0])], dim=0)
            # per-sample peak normalisation (generate_music_decode.py:191-195), in place, no host decision
            res["peak"] = peak_normalize_(wav, normalization_db=normalization_db)
            if to_host:
                if not reuse_host_buffer:
                    host = torch.empty(wav.shape, dtype=torch.float32, pin_memory=True)
                else:
                    if self._pinned_wav is None or self._pinned_wav.shape != wav.shape:
                        self._pinned_wav = torch.empty(wav.shape, dtype=torch.float32, pin_memory=True)
                    host = self._pinned_wav
                host.copy_(wav, non_blocking=True)
                torch.cuda.synchronize(self.device)
                wav = host
            else:
                torch.cuda.synchronize(self.device)
            res["audio"] = wav
            res["time_costs"]["vae_decode_time_cost"] = time.time() - t1
        res["time_costs"]["total_time_cost"] = time.time() - t0
        return res

    # ------------------------------------------------------------------
    def repaint(self, encoder_hidden_states, src_audio, repaint_start_frame: int, repaint_end_frame: int,
                silence_latent, seed=None, *, posterior_eps=None, noise=None, to_host: bool = True,
                reuse_host_buffer: bool = False, **sampler_kwargs) -> Dict[str, Any]:
        """Repaint / edit (BASELINE config 5): reference audio -> VAE encode -> DiT loop -> VAE decode.

        Mirrors the slice of the reference between `_encode_audio_to_latents` (handler/batch_prep.py:63-76)
        and the decode: the source latents keep the encoded audio outside [start, end) and the silence
        latent inside, the chunk mask is 1 inside (handler/conditioning_masks.py:33-76), and
        context_latents = [src_latents | chunk_mask] (modeling_acestep_v15_base.py:1651).

        src_audio [B, 2, N] fp32 (host or device, N a multiple of the hop); silence_latent [1 or B, >=T, 64];
        posterior_eps [B, T, 64] fixes the posterior sample (else torch's device RNG, like
        latent_dist.sample()).  Returns generate()'s dict plus "src_latents"."""
        t0 = time.time()
        audio = src_audio.to(self.device, torch.float32, non_blocking=True)
        if audio.dim() == 2:
            audio = audio.unsqueeze(0)
        B = audio.shape[0]
        hop = self.vae.shape.hop
        T = audio.shape[-1] // hop
        lat = []
        for b in range(B):
            eps = None if posterior_eps is None else posterior_eps[b].to(self.device, torch.bfloat16)
            if eps is None:
                eps = torch.randn(T, 64, device=self.device, dtype=torch.bfloat16)
            lat.append(self.vae.encode_samples(audio[b, :, : T * hop], eps))
        target = torch.stack(lat, dim=0)  # [B, T, 64] bf16
        t_enc = time.time() - t0
        s0 = max(0, min(int(repaint_start_frame), T - 1))
        s1 = max(s0 + 1, min(int(repaint_end_frame), T))
        sil = self._dev(silence_latent)[:, :T, :].expand(B, -1, -1)
        src = target.clone()
        src[:, s0:s1] = sil[:, s0:s1]
        mask = torch.zeros(B, T, 64, device=self.device, dtype=torch.bfloat16)
        mask[:, s0:s1] = 1.0
        ctx = torch.cat([src, mask], dim=-1)
        out = self.generate(encoder_hidden_states, ctx, src, seed, noise=noise, to_host=to_host,
                            reuse_host_buffer=reuse_host_buffer, **sampler_kwargs)
        out["src_latents"] = target
        out["time_costs"]["vae_encode_time_cost"] = t_enc
        out["time_costs"]["total_time_cost"] = time.time() - t0
        return out
