"""Output path of the hot path: the latent sanity guard and the per-sample
peak normalisation the reference runs either side of the VAE decode, on the C ABI.

Reference: handler/generate_music_decode.py:66-77 (NaN / Inf / all-zero guard, three torch reductions and
three host syncs) and :191-195 (`peak = wav.abs().amax(dim=[1, 2]); if any(peak > 1): wav / peak.clamp(min=1)`,
an |x| temporary as large as the waveform plus a host sync for the `any`).  Here: one pass over the latents
with one 8-byte read-back, and one read pass over the waveform + a scale pass that leaves after 4 bytes when
the peak is <= 1; no host decision in between, so the waveform's D2H copy can be queued right behind it."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib

NAN_MESSAGE = "Generation produced NaN or Inf latents."
ZERO_MESSAGE = "Generation produced zero latents."


def peak_normalize_(wav: torch.Tensor, peak: Optional[torch.Tensor] = None,
                    normalization_db: Optional[float] = None) -> torch.Tensor:
    """In place on a CUDA fp32 tensor [B, C, N] (or [C, N] = one sample): every sample whose max|x| exceeds 1
    is divided by it.  With `normalization_db` (<= 0, the front-end's GenerationParams.normalization_db,
    inference.py:117-118, 674-679) the same kernel then applies `normalize_audio(sample, normalization_db)`
    (acestep/audio_utils.py:24-62) to each sample, bit-identically to the reference's host pass.
    Returns the per-sample peaks [B] (fp32, on the device, measured BEFORE scaling)."""
    if not wav.is_cuda:
        raise _lib.B200Error("peak_normalize_: the B200 output path needs a CUDA tensor (no CPU fallback)")
    if wav.dtype != torch.float32 or not wav.is_contiguous():
        raise ValueError(f"peak_normalize_ expects a contiguous fp32 tensor, got {wav.dtype}, "
                         f"contiguous={wav.is_contiguous()}")
    if wav.dim() not in (2, 3):
        raise ValueError(f"peak_normalize_ expects [B, C, N] or [C, N], got {tuple(wav.shape)}")
    batch = wav.shape[0] if wav.dim() == 3 else 1
    n = wav.numel() // batch if batch else 0
    if peak is None:
        peak = torch.empty(batch, dtype=torch.float32, device=wav.device)
    elif peak.numel() != batch or peak.dtype != torch.float32 or peak.device != wav.device:
        raise ValueError("peak buffer must be fp32 [B] on the waveform's device")
    lib = _lib.load()
    with torch.cuda.device(wav.device):
        if normalization_db is None:
            _lib.check(lib.ace_peak_normalize(wav.data_ptr(), batch, n, peak.data_ptr(),
                                              _lib.stream_handle(wav.device)), "ace_peak_normalize")
        else:
            if normalization_db > 0.0:  # the front-end skips normalisation for positive targets (inference.py:674)
                raise ValueError(f"normalization_db must be <= 0 dB, got {normalization_db}")
            target_amp = 10 ** (normalization_db / 20.0)  # audio_utils.py:54
            _lib.check(lib.ace_peak_normalize_db(wav.data_ptr(), batch, n, peak.data_ptr(), target_amp,
                                                 _lib.stream_handle(wav.device)), "ace_peak_normalize_db")
    return peak


def latent_flags_enqueue(lat: torch.Tensor) -> torch.Tensor:
    """Queue the guard kernel; returns its int32[2] device result (any NaN/Inf, any non-zero) without reading it."""
    if not lat.is_cuda:
        raise _lib.B200Error("latent_flags: the B200 output path needs a CUDA tensor (no CPU fallback)")
    if lat.dtype != torch.bfloat16:
        raise ValueError(f"latent_flags expects bf16 latents, got {lat.dtype}")
    lat = lat.contiguous()
    flags = torch.empty(2, dtype=torch.int32, device=lat.device)
    with torch.cuda.device(lat.device):
        _lib.check(_lib.load().ace_latent_guard(lat.data_ptr(), lat.numel(), flags.data_ptr(),
                                                _lib.stream_handle(lat.device)), "ace_latent_guard")
    return flags


def latent_flags(lat: torch.Tensor) -> Tuple[bool, bool]:
    """(any NaN/Inf, any non-zero) of a CUDA bf16 tensor — one kernel, one 8-byte device->host read."""
    bad, nonzero = latent_flags_enqueue(lat).tolist()
    return bool(bad), bool(nonzero)


def raise_on_flags(bad, nonzero, numel: int) -> None:
    """The reference's guard, same order and same leading sentence of each message: RuntimeError on NaN/Inf,
    then on all-zero latents (only when there is at least one element)."""
    if bad:
        raise RuntimeError(NAN_MESSAGE)
    if numel > 0 and not nonzero:
        raise RuntimeError(ZERO_MESSAGE)


def check_latents(lat: torch.Tensor) -> None:
    """Guard + host decision in one call (one host synchronisation)."""
    bad, nonzero = latent_flags(lat)
    raise_on_flags(bad, nonzero, lat.numel())
