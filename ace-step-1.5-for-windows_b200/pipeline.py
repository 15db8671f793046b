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
from .output import latent_flags_enqueue, peak_normalize_, raise_on_flags
from .sampler import B200Sampler
from .vae import B200Vae, VaeShape


class PendingSong:
    """A song in flight (B200Pipeline.generate_async).  `wait()` -> the result dict; raises the reference's latent
    guard errors.  Time costs are device times between CUDA events (the host only enqueued)."""

    def __init__(self, res, events, host_flags, numel, t0, decoded, steps):
        self._res, self._ev, self._flags, self._numel = res, events, host_flags, numel
        self._t0, self._decoded, self._steps = t0, decoded, max(1, int(steps))
        self._extra_time: Dict[str, float] = {}
        self.stream = None  # the CUDA stream the waveform is produced on

    def done(self) -> bool:
        return self._ev[2].query()

    @property
    def device_audio(self):
        """The waveform on the device ([B, 2, N] fp32), valid in stream order on `stream` (the caller's current stream
        unless the pipeline overlaps the codec) WITHOUT waiting — for callers that queue device work behind the song,
        e.g. the multi-GPU gather.  The latent guard has not been looked at yet: call wait() before handing it out."""
        return self._res.get("_device_audio", self._res.get("audio"))

    def wait(self) -> Dict[str, Any]:
        self._ev[2].synchronize()
        bad, nonzero = self._flags.tolist()
        raise_on_flags(bool(bad), bool(nonzero), self._numel)
        res = self._res
        res.pop("_device_audio", None)
        tc = res["time_costs"]
        diff = self._ev[0].elapsed_time(self._ev[1]) * 1e-3
        tc["diffusion_time_cost"], tc["diffusion_per_step_time_cost"] = diff, diff / self._steps
        if self._decoded:
            tc["vae_decode_time_cost"] = self._ev[1].elapsed_time(self._ev[2]) * 1e-3
        tc.update(self._extra_time)
        tc["total_time_cost"] = time.time() - self._t0
        return res


class SongPipeline:
    """Bounded queue of songs in flight for serving loops: `submit(pending)` returns the result of the OLDEST song once
    more than `depth` are queued (None before that), `drain()` waits for the rest and returns the last result.  With
    depth 1 the host prepares and enqueues song i + 1 while the GPU still runs song i, so the device never idles at a
    song boundary; results come back one submit later."""

    def __init__(self, depth: int = 1):
        self.depth, self._q = max(0, int(depth)), []

    def submit(self, pending: "PendingSong"):
        self._q.append(pending)
        if len(self._q) > self.depth:
            return self._q.pop(0).wait()
        return None

    def drain(self):
        res = None
        while self._q:
            res = self._q.pop(0).wait()
        return res


class B200Pipeline:
    def __init__(self, dit_state: Dict[str, torch.Tensor], vae_state: Dict[str, torch.Tensor],
                 dit_shape: Optional[DiTShape] = None, vae_shape: Optional[VaeShape] = None,
                 null_condition_emb: Optional[torch.Tensor] = None, device="cuda:0", turbo: bool = False,
                 overlap_codec: bool = False):
        """`overlap_codec=True`: the denoising loop runs on a high-priority stream and everything the codec does
        (decode, peak normalisation, the D2H copy; the encode of a repaint) on a second, normal-priority stream, so with
        songs queued back to back (SongPipeline) the decode of song i runs under song i + 1's steps instead of
        standing between two loops.  OFF by default: measured on B200 (bench.py, C2, 10 songs) it is slower —
        386.3 vs 393.7 audio-s/s — the codec's CTAs (a whole SM each, like the GEMMs') delay the step's dependent
        GEMM chain by more than the idle SM time they fill, and the device is power-capped either way.
        False: one stream (the caller's current one)."""
        self.device = torch.device(device)
        self.turbo = turbo
        self.overlap_codec = bool(overlap_codec)
        self._loop_stream = torch.cuda.Stream(self.device, priority=-1) if overlap_codec else None
        self._codec_stream = torch.cuda.Stream(self.device, priority=0) if overlap_codec else None
        self.dit = B200DiT(dit_state, dit_shape or DiTShape(), self.device)
        self.vae = B200Vae(vae_state, vae_shape or VaeShape(), self.device)
        self.sampler = B200Sampler(self.dit, null_condition_emb)
        self.sample_rate = 48000
        self._pinned_wav = [None, None]
        self._pinned_i = 0

    @property
    def codec_stream(self):
        """The stream decode / normalisation / D2H run on (None: the caller's current stream)."""
        return self._codec_stream

    def close(self):
        self.dit.close()
        self.vae.close()

    def _dev(self, x):
        if x is None:
            return None
        return x.to(self.device, torch.bfloat16, non_blocking=True)

    def generate(self, encoder_hidden_states, context_latents, src_latents=None, seed=None, **kwargs) -> Dict[str, Any]:
        """Returns {"audio": fp32 [B,2,N] (peak-normalised like the reference when |x|max > 1),
        "target_latents": bf16 [B,T,64], "peak": fp32 [B] raw peaks, "time_costs": {...}}.
        `normalization_db` (e.g. -1.0, GenerationParams.normalization_db) also applies the front-end's
        `normalize_audio` per song (inference.py:674-679) inside the same device pass, so the host copy that
        comes back is the final audio and the reference's three host passes over it are not needed.
        With `to_host` the waveform comes back in a pinned host tensor owned by the caller (a fresh one per call);
        `reuse_host_buffer=True` returns one of the pipeline's two reusable pinned buffers instead — valid only until
        the call after the next (serving loops / the benchmark, which consume each result before then).
        = generate_async(...).wait()."""
        return self.generate_async(encoder_hidden_states, context_latents, src_latents, seed, **kwargs).wait()

    def generate_async(self, encoder_hidden_states, context_latents, src_latents=None, seed=None, *,
                       noise=None, to_host: bool = True, reuse_host_buffer: bool = False, decode: bool = True,
                       latent_shift: float = 0.0, latent_rescale: float = 1.0,
                       normalization_db: Optional[float] = None, **sampler_kwargs) -> "PendingSong":
        """Queues the whole song — H2D copies, denoising loop, latent guard, decode, peak normalisation, D2H copy —
        without ONE host synchronisation and returns a PendingSong; `.wait()` blocks on its completion event, applies
        the reference's latent guard (generate_music_decode.py:66-77: RuntimeError on NaN / Inf / all-zero latents,
        raised before any audio is handed out) and returns generate()'s dict.  A serving loop that submits song i + 1
        before waiting for song i keeps the GPU busy across the song boundary (the host-side preparation of a song —
        conditioning copies, K/V projection launches, the graph launches of the first steps — is a few hundred
        microseconds during which the device would otherwise idle).  The loop runs on one stream (the handle's static
        I/O slots are reused in stream order) and the codec on another (see `overlap_codec`); `PendingSong.stream` is
        the stream the waveform was produced on, for callers that queue device work behind it without waiting."""
        t0 = time.time()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        cur = torch.cuda.current_stream(self.device)
        loop_s = self._loop_stream if self.overlap_codec else cur
        codec_s = self._codec_stream if self.overlap_codec else cur
        loop_s.wait_stream(cur)  # inputs the caller produced on its own stream
        with torch.cuda.stream(loop_s):
            ev[0].record()
            enc, ctx = self._dev(encoder_hidden_states), self._dev(context_latents)
            src = self._dev(src_latents) if src_latents is not None else ctx[..., :64].contiguous()
            noise = self._dev(noise)
            fn = self.sampler.generate_turbo if self.turbo else self.sampler.generate_base
            out = fn(enc, ctx, src, seed, noise=noise, sync=False, **sampler_kwargs)
            lat = out["target_latents"]
            # NaN / Inf / all-zero guard of _prepare_generate_music_decode_state (generate_music_decode.py:66-77):
            # the kernel runs here, the host looks at its two flags in wait()
            flags = latent_flags_enqueue(lat)
            numel = lat.numel()
            host_flags = torch.empty(2, dtype=torch.int32, pin_memory=True)
            host_flags.copy_(flags, non_blocking=True)
            if latent_shift != 0.0 or latent_rescale != 1.0:
                lat = lat * latent_rescale + latent_shift
            res: Dict[str, Any] = {"target_latents": lat, "time_costs": out["time_costs"]}
            ev[1].record()
        codec_s.wait_event(ev[1])
        with torch.cuda.stream(codec_s):
            if decode:
                lat.record_stream(codec_s)
                wav = torch.stack([self.vae.decode_frames(lat[b]) for b in range(lat.shape[0])], dim=0)
                # per-sample peak normalisation (generate_music_decode.py:191-195), in place, no host decision
                res["peak"] = peak_normalize_(wav, normalization_db=normalization_db)
                if to_host:
                    if not reuse_host_buffer:
                        host = torch.empty(wav.shape, dtype=torch.float32, pin_memory=True)
                    else:  # two buffers: the one handed out by the previous call stays valid while this song runs
                        self._pinned_i ^= 1
                        host = self._pinned_wav[self._pinned_i]
                        if host is None or host.shape != wav.shape:
                            host = torch.empty(wav.shape, dtype=torch.float32, pin_memory=True)
                            self._pinned_wav[self._pinned_i] = host
                    host.copy_(wav, non_blocking=True)
                    res["_device_audio"] = wav  # keeps the source alive until the copy has run
                    wav = host
                res["audio"] = wav
            ev[2].record()
        pending = PendingSong(res, ev, host_flags, numel, t0, decode, out.get("steps", 1))
        pending.stream = codec_s
        return pending

    # ------------------------------------------------------------------
    def repaint(self, *args, **kwargs) -> Dict[str, Any]:
        """= repaint_async(...).wait()."""
        return self.repaint_async(*args, **kwargs).wait()

    def repaint_async(self, encoder_hidden_states, src_audio, repaint_start_frame: int, repaint_end_frame: int,
                      silence_latent, seed=None, *, posterior_eps=None, noise=None, to_host: bool = True,
                      reuse_host_buffer: bool = False, **sampler_kwargs) -> PendingSong:
        """Repaint / edit (BASELINE config 5): reference audio -> VAE encode -> DiT loop -> VAE decode.

        Mirrors the slice of the reference between `_encode_audio_to_latents` (handler/batch_prep.py:63-76)
        and the decode: the source latents keep the encoded audio outside [start, end) and the silence
        latent inside, the chunk mask is 1 inside (handler/conditioning_masks.py:33-76), and
        context_latents = [src_latents | chunk_mask] (modeling_acestep_v15_base.py:1651).

        src_audio [B, 2, N] fp32 (host or device, N a multiple of the hop); silence_latent [1 or B, >=T, 64];
        posterior_eps [B, T, 64] fixes the posterior sample (else torch's device RNG, like
        latent_dist.sample()).  The PendingSong's result is generate()'s dict plus "src_latents"; like
        generate_async nothing here synchronises with the host."""
        t0 = time.time()
        cur = torch.cuda.current_stream(self.device)
        codec_s = self._codec_stream if self.overlap_codec else cur
        codec_s.wait_stream(cur)
        # the encode shares the codec handle's workspace with the previous song's decode: same stream
        with torch.cuda.stream(codec_s):
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
        cur.wait_stream(codec_s)  # generate_async makes its loop stream wait for the caller's stream
        for x in (ctx, src, target):
            x.record_stream(cur)
            if self.overlap_codec:
                x.record_stream(self._loop_stream)
        pending = self.generate_async(encoder_hidden_states, ctx, src, seed, noise=noise, to_host=to_host,
                                      reuse_host_buffer=reuse_host_buffer, **sampler_kwargs)
        pending._res["src_latents"] = target
        pending._extra_time["vae_encode_time_cost"] = t_enc  # enqueue time (the encode is not event-bracketed)
        pending._t0 = t0
        return pending
