"""Oracle restatements of the output path (TEST INFRASTRUCTURE): the reference's torch expressions either side
of the VAE decode and the front-end's peak normalisation.  Pinned bit-exactly against the real
`normalize_audio` (extracted from the reference source by tools/make_golden_output.py ->
tests/golden/output_normalize.npz)."""
import torch


def latent_guard(pred_latents: torch.Tensor):
    """(any NaN/Inf, any non-zero) — handler/generate_music_decode.py:66 and :73."""
    bad = bool(torch.isnan(pred_latents).any() or torch.isinf(pred_latents).any())
    nonzero = bool(pred_latents.numel() > 0 and pred_latents.abs().sum() != 0)
    return bad, nonzero


def peak_normalize(pred_wavs: torch.Tensor):
    """handler/generate_music_decode.py:191-195 on [B, C, N] fp32.  Returns (waveforms, per-sample peaks)."""
    if pred_wavs.dtype != torch.float32:
        pred_wavs = pred_wavs.float()
    peak = pred_wavs.abs().amax(dim=[1, 2], keepdim=True)
    if torch.any(peak > 1.0):
        pred_wavs = pred_wavs / peak.clamp(min=1.0)
    return pred_wavs, peak.flatten()


def normalize_audio(audio_data: torch.Tensor, target_db: float = -1.0) -> torch.Tensor:
    """acestep/audio_utils.py:24-62 (tensor branch): peak-normalise ONE song to `target_db` dBFS; silence
    (peak < 1e-6) is returned unchanged."""
    audio = audio_data.clone()
    peak = torch.max(torch.abs(audio))
    if peak < 1e-6:
        return audio_data
    target_amp = 10 ** (target_db / 20.0)
    gain = target_amp / peak
    return audio * gain


def finalize(pred_wavs: torch.Tensor, normalization_db=None):
    """The whole output chain for a batch: handler normalisation, then (front-end, per song,
    inference.py:674-679: only when normalization_db <= 0) normalize_audio."""
    wavs, peak = peak_normalize(pred_wavs)
    if normalization_db is not None and normalization_db <= 0.0:
        wavs = torch.stack([normalize_audio(wavs[i], normalization_db) for i in range(wavs.shape[0])])
    return wavs, peak
