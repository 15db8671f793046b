#!/usr/bin/env python
"""bench.py — headline benchmark of the ACE-Step hot path on B200 (contract: see DESIGN.md §Measurement).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference|torch-gpu] [--workload c1|c2|c3|c5]

A "step" is one pass of the hot path over one batch of synthetic input: ONE SONG per GPU —
denoising loop (base sampler, CFG + APG) on synthetic text-conditioning embeddings + VAE decode to
a 48 kHz stereo waveform.  Default workload = BASELINE.json configs[1]: text2music 60 s, 27 steps,
bf16, batch 1 (effective batch 2 with CFG), E = 512 condition tokens.

Metric: generated-audio seconds per wall second (whole job, all GPUs).
  value : inputs resident in HBM, timed with CUDA events, max over ranks.
  e2e   : same work through the public API (acestep_b200.pipeline.B200Pipeline.generate) with
          pinned HOST inputs and a HOST waveform result (H2D + D2H inside the timed region).
  roofline : tcgen05 GEMM launches of one song, event-timed per launch inside this run
             (ace_profile_start/stop), algorithmic FLOPs / summed duration vs the measured bf16 peak.
  cpu_baseline / --impl reference : the oracle port of the reference's PyTorch CPU path on this
             box's host cores, on a bounded sample of the same workload (stated in `sample`; the full C2 song
             is ~85 s of CPU time, so K + W songs would not fit the few-minute budget).
  gpu_baseline / --impl torch-gpu : the reference's own GPU path restated (oracle modules, bf16, SDPA + cuBLAS /
             cuDNN, tiled decode) on the same B200, same workload, no extrapolation.
  extra_workloads : c3 (240 s / 60 steps — north_star's target), c5 (repaint 120 s), c1 (10 s turbo) measured in
             the same run (value, e2e, DiT step alone vs both measured peaks).
  dit_step_tensor_util : the DiT step alone (song minus event-timed codec passes, / steps) vs the sustained and
             the burst bf16 peak of MEASURED_PEAKS.json, plus the whole-song figure.
Weights are random-init of the reference architecture, data synthetic (no checkpoints / network).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (seconds, frames T, steps, guidance, shift, cond tokens E, turbo)
    "c1": dict(seconds=10, T=250, steps=8, guidance=1.0, shift=3.0, E=512, turbo=True,
               desc="text2music 10 s, turbo 8 steps, batch 1 (no CFG)"),
    "c2": dict(seconds=60, T=1500, steps=27, guidance=7.0, shift=3.0, E=512, turbo=False,
               desc="text2music 60 s, base sampler 27 steps, CFG 7.0 + APG, batch 1 (effective 2)"),
    "c3": dict(seconds=240, T=6000, steps=60, guidance=7.0, shift=3.0, E=512, turbo=False,
               desc="text2music 240 s, base sampler 60 steps, CFG 7.0 + APG, batch 1 (effective 2)"),
    "c5": dict(seconds=120, T=3000, steps=27, guidance=7.0, shift=3.0, E=512, turbo=False, repaint=(750, 2250),
               desc="repaint 120 s: VAE encode of the source audio -> base sampler 27 steps, CFG 7.0 + APG, frames "
                    "[750, 2250) repainted -> VAE decode, batch 1 per GPU (effective 2)"),
}


def dit_flops(Bc, S, E, L=24, D=2048, I=6144, NQ=2048, NKV=1024, window=128):
    """Algorithmic FLOPs of one DiT forward."""
    lin = 2 * ((NQ + 2 * NKV) * D + NQ * D + NQ * D + NQ * D + 2 * I * D + D * I)  # per token per layer
    per_tok = L * lin + 2 * (384 * D + D * 128)
    attn = (L // 2) * 4 * S * S * NQ + (L - L // 2) * 4 * S * min(S, 2 * window + 1) * NQ + L * 4 * S * E * NQ
    return Bc * (S * per_tok + attn)


def vae_flops_per_frame():
    return 4.874e9


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            # nvidia-smi's start-up (NVML attach) can stall the GPU for tens of ms: let it deliver its first
            # sample BEFORE the timed region opens (seen once as +12 ms per song on the first region only)
            deadline = time.perf_counter() + 5.0
            while not self.lines and time.perf_counter() < deadline and self.proc.poll() is None:
                time.sleep(0.02)
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def window(self, t0, t1):
        """Median SM clock of the samples taken in [t0, t1] (perf_counter seconds)."""
        v = []
        for ts, ln in self.lines:
            if t0 <= ts <= t1:
                try:
                    v.append(float(ln.split(",")[0]))
                except ValueError:
                    pass
        return statistics.median(v) if v else None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for _ts, ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
def cpu_sample(wl, threads=None):
    """Bounded CPU sample of the workload with the oracle port (fp32, all host threads):
    one DiT forward at the workload's (Bc, T, E) + one 64-frame VAE decode window, extrapolated
    linearly to the whole song.  Returns (audio_s_per_s, sample_seconds, description)."""
    import torch

    from oracle import vae as ovae
    from oracle.dit import CrossCache, DiTConfig, dit_forward
    from oracle.weights import make_dit_weights, make_vae_weights

    import oracle.dit as odit

    odit.ATTN_IMPL = "sdpa"  # the reference's default attention implementation (handler/init_service_loader.py:43)
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    cfg = DiTConfig()
    w = make_dit_weights(cfg, seed=0)
    vcfg = ovae.VaeConfig()
    vw = make_vae_weights(vcfg, seed=0)
    vw = {k: v for k, v in vw.items()}
    Bc = 2 if wl["guidance"] > 1.0 else 1
    g = torch.Generator().manual_seed(1)
    xt = torch.randn(Bc, wl["T"], 64, generator=g)
    ctx = torch.randn(Bc, wl["T"], 128, generator=g)
    enc = torch.randn(Bc, wl["E"], cfg.hidden_size, generator=g)
    t = torch.full((Bc,), 0.5)
    z = torch.randn(1, 64, 64, generator=g)
    state = {"cache": CrossCache()}

    def one_sample():
        t0 = time.perf_counter()
        with torch.no_grad():
            dit_forward(w, cfg, xt, t, ctx, enc, state["cache"])
        t1 = time.perf_counter()
        with torch.no_grad():
            ovae.decode(vw, vcfg, z)
        t2 = time.perf_counter()
        return t1 - t0, t2 - t1

    return one_sample, threads


def run_reference(args, wl):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the reference
    itself is pure Python on torch CPU ops, and /root/reference does not exist on the GPU box)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    one_sample, threads = cpu_sample(wl)
    for _ in range(max(args.warmup, 1)):
        one_sample()
    fw, dec = [], []
    for _ in range(args.steps):
        a, b = one_sample()
        fw.append(a)
        dec.append(b)
    t_fwd, t_win = statistics.mean(fw), statistics.mean(dec)
    # repaint also encodes the source audio: the encoder mirrors the decoder (same FLOPs per frame), so it is
    # charged as a second pass of the timed decode window
    song = wl["steps"] * t_fwd + (wl["T"] / 64.0) * t_win * (2 if wl.get("repaint") else 1)
    value = wl["seconds"] / song
    sample = (f"per step: 1 DiT forward at effective batch {2 if wl['guidance'] > 1 else 1}, T={wl['T']}, "
              f"E={wl['E']} (cross-KV cached) + one 64-frame VAE decode window; extrapolated to "
              f"{wl['steps']} forwards + {wl['T']}/64 windows per song")
    line = {
        "impl": "reference", "metric": "generated-audio-sec/wall-sec", "value": value, "unit": "audio-s/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": (t_fwd + t_win) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["desc"], "extrapolated_song_seconds": song},
        "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
def torch_gpu_arm(wl, dev, dshape, vshape, dit_state, vae_state, cond, null_emb):
    """The reference's GPU path restated: the oracle modules (same arithmetic as the reference's PyTorch modules,
    pinned by tests/test_oracle_golden.py) in bf16 on the cuda device with SDPA attention + cuBLAS / cuDNN — what
    `initialize_service(device="cuda")` runs (handler/init_service_loader.py:45-71: bf16, attn_implementation="sdpa";
    dense additive masks, per-step Python loop, overlap-discard tiled decode with chunk 512 / overlap 64).
    Returns song(seed) -> fp32 waveform on the device.  Baseline only: nothing of the product path runs here."""
    import torch

    import oracle.dit as odit
    from acestep_b200.pack import folded_vae_state
    from oracle import sampler as osamp
    from oracle import vae as ovae

    odit.ATTN_IMPL = "sdpa"
    cfg = odit.DiTConfig(hidden_size=dshape.hidden_size, intermediate_size=dshape.intermediate_size,
                         num_hidden_layers=dshape.num_hidden_layers, num_attention_heads=dshape.num_attention_heads,
                         num_key_value_heads=dshape.num_key_value_heads, sliding_window=dshape.sliding_window)
    vcfg = ovae.VaeConfig()
    wb = {k: v.to(dev, torch.bfloat16) for k, v in dit_state.items()}
    vb = {k: v.to(dev, torch.bfloat16) for k, v in folded_vae_state(vae_state).items()}
    vel = lambda xt, t, c, e, cache: odit.dit_forward(wb, cfg, xt, t, c, e, cache)
    enc, ctx, src = cond["enc"], cond["ctx"], cond["src"]

    def song(seed):
        with torch.no_grad():
            noise = osamp.prepare_noise((1, wl["T"], 64), [seed], torch.bfloat16, dev)
            if wl["turbo"]:
                lat = osamp.sample_turbo(vel, enc, ctx, src, None, shift=wl["shift"], noise=noise,
                                         new_cache=odit.CrossCache)
            else:
                lat = osamp.sample_base(vel, enc, ctx, src, None, null_emb=null_emb, infer_steps=wl["steps"],
                                        guidance_scale=wl["guidance"], shift=wl["shift"], noise=noise,
                                        new_cache=odit.CrossCache)
            wav = ovae.tiled_decode(lambda z: ovae.decode(vb, vcfg, z), lat.transpose(1, 2).contiguous(), 512, 64).float()
            peak = wav.abs().amax(dim=[1, 2], keepdim=True)
            return torch.where(peak > 1.0, wav / peak, wav)

    return song


def time_torch_gpu(wl, dev, dshape, vshape, dit_state, vae_state, cond, null_emb, songs=2, warmup=1):
    import torch

    song = torch_gpu_arm(wl, dev, dshape, vshape, dit_state, vae_state, cond, null_emb)
    for i in range(warmup):
        song(i)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(songs):
        out = song(100 + i)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / songs
    finite = bool(torch.isfinite(out).all())
    del song, out
    torch.cuda.empty_cache()
    return {"value": wl["seconds"] / (ms * 1e-3), "unit": "audio-s/s", "ms_per_song": ms, "songs": songs,
            "kind": "oracle modules (= the reference's PyTorch modules restated), bf16, cuda, SDPA + cuBLAS/cuDNN, "
                    "dense masks, tiled decode 512/64; same workload, same box, same run",
            "outputs_finite": finite}


def cublas_gemm_ab(shapes, dev, reps=20):
    """cuBLAS (torch.matmul, bf16, fp32 accumulate) on the DiT's GEMM problem shapes, timed like bench.py's own
    per-launch profile: one CUDA event pair per launch, eager, mean of `reps` — a PLAIN GEMM (no fused epilogue, bf16
    output), so it bounds what the library path would cost before its separate norm / RoPE / SwiGLU / residual kernels."""
    import torch

    out = []
    for sh in shapes:
        m, n, k = sh["m"], sh["n"], sh["k"]
        if m > 100000:  # codec shapes: not a library GEMM (tap-shifted convolution)
            continue
        a = torch.randn(m, k, device=dev, dtype=torch.bfloat16)
        b = torch.randn(n, k, device=dev, dtype=torch.bfloat16)
        for _ in range(3):
            torch.matmul(a, b.t())
        torch.cuda.synchronize(dev)
        tot = 0.0
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b.t())
            e1.record()
            e1.synchronize()
            tot += e0.elapsed_time(e1)
        us = tot / reps * 1e3
        out.append({"m": m, "n": n, "k": k, "cublas_us": round(us, 2), "cublas_tflops": 2.0 * m * n * k / (us * 1e-6) / 1e12,
                    "this_repo_us": round(sh["ms_total"] / sh["launches"] * 1e3, 2) if sh["launches"] else None,
                    "this_repo_tflops": sh["tflops"]})
    return out


def run_torch_gpu(args, wl):
    """--impl torch-gpu: the reference's GPU path (see torch_gpu_arm) as a stand-alone arm, rank 0 only."""
    import torch

    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    from acestep_b200.dit import DiTShape
    from acestep_b200.synthetic import random_dit_state, random_vae_state, synthetic_conditioning
    from acestep_b200.vae import VaeShape

    dev = torch.device("cuda:0")
    dshape, vshape = DiTShape(), VaeShape()
    cond = synthetic_conditioning(1, wl["T"], wl["E"], dshape.hidden_size, seed=1234, device=dev)
    r = time_torch_gpu(wl, dev, dshape, vshape, random_dit_state(dshape, 0, dev), random_vae_state(vshape, 0, dev),
                       cond, cond["null_emb"], songs=max(args.steps, 1), warmup=max(args.warmup, 1))
    line = {"impl": "torch-gpu", "metric": "generated-audio-sec/wall-sec", "value": r["value"], "unit": "audio-s/s",
            "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_song"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": wl["desc"], "kind": r["kind"]}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
class Runner:
    """One workload on this rank's GPU: the device-resident and the host-buffer song closures."""

    def __init__(self, pipe, wl, dshape, vshape, dev, rank):
        import torch

        from acestep_b200.synthetic import synthetic_conditioning

        self.pipe, self.wl, self.dev = pipe, wl, dev
        T, E = wl["T"], wl["E"]
        self.T, self.E = T, E
        self.n_samples = T * vshape.hop
        self.host = synthetic_conditioning(1, T, E, dshape.hidden_size, seed=1234 + rank, device="cpu", pin=True)
        self.dev_in = {k: v.to(dev) for k, v in self.host.items()}
        pipe.sampler.null_condition_emb = self.dev_in["null_emb"]
        pipe.turbo = wl["turbo"]
        if wl["turbo"]:
            self.skw = dict(shift=wl["shift"])
        else:
            self.skw = dict(infer_steps=wl["steps"], diffusion_guidance_sale=wl["guidance"], shift=wl["shift"])
        self.rp = wl.get("repaint")
        self.seed0 = 1000 * rank
        if self.rp:  # config 5: the source audio is encoded inside the timed region
            ga = torch.Generator().manual_seed(99 + rank)
            self.audio_h = (torch.rand(1, 2, self.n_samples, generator=ga) - 0.5).pin_memory()
            self.eps_h = torch.randn(1, T, 64, generator=ga).to(torch.bfloat16).pin_memory()
            self.sil_h = torch.randn(1, T, 64, generator=ga).to(torch.bfloat16).pin_memory()
            self.audio_d, self.eps_d, self.sil_d = self.audio_h.to(dev), self.eps_h.to(dev), self.sil_h.to(dev)

    # the starting noise is drawn INSIDE the song (seeded torch generator on the device), like the reference's
    # generate_audio -> prepare_noise (turbo :1917, base :1899)
    # Both closures only ENQUEUE the song (B200Pipeline.generate_async / repaint_async -> PendingSong); the loops
    # below keep one song in flight behind the one being waited for (SongPipeline(depth=1)), which is how a serving
    # loop calls the product API.
    def song_device(self, i):
        if self.rp:
            return self.pipe.repaint_async(self.dev_in["enc"], self.audio_d, self.rp[0], self.rp[1], self.sil_d,
                                           [self.seed0 + i], posterior_eps=self.eps_d, to_host=False, **self.skw)
        return self.pipe.generate_async(self.dev_in["enc"], self.dev_in["ctx"], self.dev_in["src"], [self.seed0 + i],
                                        to_host=False, **self.skw)

    def song_host(self, i):
        if self.rp:
            return self.pipe.repaint_async(self.host["enc"], self.audio_h, self.rp[0], self.rp[1], self.sil_h,
                                           [self.seed0 + i], posterior_eps=self.eps_h, to_host=True,
                                           reuse_host_buffer=True, **self.skw)
        return self.pipe.generate_async(self.host["enc"], self.host["ctx"], self.host["src"], [self.seed0 + i],
                                        to_host=True, reuse_host_buffer=True, **self.skw)

    def h2d_bytes(self):
        if self.rp:
            return (self.host["enc"].numel() + self.eps_h.numel() + self.sil_h.numel()) * 2 + self.audio_h.numel() * 4
        return sum(self.host[k].numel() * self.host[k].element_size() for k in ("enc", "ctx", "src"))

    def codec_ms(self, reps=3):
        """Event-timed VAE decode (+ encode for repaint) of this workload's length, to split the song time."""
        import torch

        z = self.dev_in["src"][0]
        self.pipe.vae.decode_frames(z)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(self.dev)
        e0.record()
        for _ in range(reps):
            self.pipe.vae.decode_frames(z)
            if self.rp:
                self.pipe.vae.encode_samples(self.audio_d[0], self.eps_d[0])
        e1.record()
        torch.cuda.synchronize(self.dev)
        return e0.elapsed_time(e1) / reps

    def flops(self):
        Bc = 2 if self.wl["guidance"] > 1.0 else 1
        S = (self.T + 1) // 2
        step = dit_flops(Bc, S, self.E)
        return step, self.wl["steps"] * step + vae_flops_per_frame() * self.T * (2 if self.rp else 1)


def util_block(r, ms_song, codec_ms, peaks):
    """DiT step alone (song time minus the event-timed codec passes, divided by the steps — i.e. including
    guidance, Euler update and every host gap of the loop) against both measured peaks, and the whole song."""
    step_fl, song_fl = r.flops()
    step_ms = max(ms_song - codec_ms, 1e-6) / r.wl["steps"]
    tf = step_fl / (step_ms * 1e-3) / 1e12
    sus, burst = float(peaks.get("bf16_tflops_sustained", 1400.0)), float(peaks.get("bf16_tflops", 1590.0))
    return {"dit_step_ms": step_ms, "dit_step_algorithmic_tflop": step_fl / 1e12, "dit_step_tflops": tf,
            "dit_step_frac_of_sustained_peak": tf / sus, "dit_step_frac_of_burst_peak": tf / burst,
            "codec_ms_per_song": codec_ms,
            "whole_song_algorithmic_tflop": song_fl / 1e12,
            "whole_song_tflops": song_fl / (ms_song * 1e-3) / 1e12,
            "whole_song_frac_of_sustained_peak": song_fl / (ms_song * 1e-3) / 1e12 / sus,
            "peaks_tflops": {"sustained": sus, "burst": burst}}


def run_b200(args, wl):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)

    from acestep_b200 import _lib
    from acestep_b200.dit import DiTShape
    from acestep_b200.multi_gpu import GatherPipeline, generate_sharded
    from acestep_b200.pipeline import B200Pipeline, SongPipeline
    from acestep_b200.synthetic import random_dit_state, random_vae_state
    from acestep_b200.vae import VaeShape

    lib = _lib.load()
    dshape, vshape = DiTShape(), VaeShape()
    dit_state, vae_state = random_dit_state(dshape, 0, dev), random_vae_state(vshape, 0, dev)
    pipe = B200Pipeline(dit_state, vae_state, dshape, vshape, device=dev, turbo=wl["turbo"])
    torch.cuda.empty_cache()
    peaks, peak_kind = measured_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def measure(r, steps, warmup, clocks=None, verify=False):
        """value: EXACTLY `steps` songs per GPU, device-resident inputs, CUDA events, barrier + synchronize on both
        sides; N > 1 goes through the product's multi-GPU API (acestep_b200.multi_gpu.generate_sharded: song i on
        rank i, asynchronous waveform gather to rank 0 that overlaps the next song, drained inside the region).
        e2e: the same through the public API with pinned HOST inputs and a HOST waveform (wall clock)."""
        lengths = [r.n_samples] * world
        pipe_g, pipe_s, keep = GatherPipeline(depth=1), SongPipeline(depth=1), {}
        counter = [0]

        def retire(out):
            if out is not None:
                keep["out"] = out

        def step_device():
            i = counter[0]
            counter[0] += 1
            if world == 1:
                retire(pipe_s.submit(r.song_device(i)))  # song i enqueued; song i - 1 waited for and guard-checked
                return
            def one(_song_index):
                pend = r.song_device(i)
                wav = pend.device_audio[0]  # device tensor, stream-ordered: the gather is queued behind the song
                retire(pipe_s.submit(pend))
                return wav
            # at most one gather in flight: the previous song's gather (finished long ago) is retired here, which
            # hands its buffers back to the allocator before this song's successor needs memory
            pg = generate_sharded(one, world, dst=0, device=dev, lengths=lengths, async_op=True,
                                  gather_stream=pipe.codec_stream)
            with torch.cuda.stream(pipe.codec_stream):  # retiring the previous gather must not stall the loop stream
                pipe_g.submit(pg)

        def drain():
            retire(pipe_s.drain())
            with torch.cuda.stream(pipe.codec_stream):
                return pipe_g.drain()

        # Warm-up keeps the previous song's result alive while the next one runs, exactly like the timed loop;
        # the clock sampler attaches BEFORE the warm-up and the warm-up runs back to back into the timed region
        # (after an idle gap a power-capped B200 runs one song fast and then over-corrects for about one song).
        barrier()
        if verify and world > 1:
            # once, through the RAGGED mode of the product API (length exchange + padded gather) and checked
            def one(_i):
                keep["out"] = r.song_device(0).wait()
                return keep["out"]["audio"][0]
            songs = generate_sharded(one, world, dst=0, device=dev)  # (the song was waited for: any stream will do)
            if rank == 0:
                assert len(songs) == world and all(s.shape == (2, r.n_samples) for s in songs)
                assert torch.equal(songs[0], keep["out"]["audio"][0]) and all(bool(torch.isfinite(s).all()) for s in songs)
        for _ in range(warmup):
            step_device()
        drain()
        l0 = lib.ace_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        w0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            step_device()
        gathered = drain()  # all gathers complete inside the timed region
        e1.record()
        barrier()
        w1 = time.perf_counter()
        ms = e0.elapsed_time(e1)
        launches = lib.ace_launch_count() - l0
        finite = bool(torch.isfinite(keep["out"]["audio"]).all())
        if gathered is not None:
            finite = finite and all(bool(torch.isfinite(s).all()) for s in gathered)
        # ---- e2e
        r.song_host(0).wait()
        barrier()
        barrier()
        t0 = time.perf_counter()
        pipe_h, oh = SongPipeline(depth=1), None
        for i in range(steps):
            oh = pipe_h.submit(r.song_host(10 + i)) or oh  # host waveform of song i - 1 (pinned, two buffers alternate)
        oh = pipe_h.drain() or oh
        barrier()
        t1 = time.perf_counter()
        res = {"ms": ms, "launches": int(launches), "finite": finite, "e2e_s": t1 - t0,
               "d2h": oh["audio"].numel() * 4, "h2d": r.h2d_bytes(), "win_value": (w0, w1), "win_e2e": (t0, t1)}
        keep.clear()
        return res

    clocks = ClockSampler(local)
    clocks.start()
    main = Runner(pipe, wl, dshape, vshape, dev, rank)
    m = measure(main, args.steps, args.warmup, clocks, verify=True)
    my_clk = clocks.window(*m["win_value"])
    ms_rank, e2e_rank = m["ms"], m["e2e_s"]
    per_rank = [{"rank": rank, "ms_per_step": ms_rank / args.steps, "sm_mhz": my_clk}]
    ms, e2e_s = ms_rank, e2e_rank
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, per_rank[0])
        per_rank = gathered
        tt = torch.tensor([ms, e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, e2e_s = float(tt[0]), float(tt[1])

    line = None
    if rank == 0:
        # per-launch event profile of one song (rank 0 only)
        import ctypes as C

        lib.ace_profile_start()
        main.song_device(0).wait()  # rank-local: no collective here, the other ranks are already past the timed region
        pms, pfl, pby, pln = (C.c_float * 4)(), (C.c_double * 4)(), (C.c_double * 4)(), (C.c_int * 4)()
        _lib.check(lib.ace_profile_stop(pms, pfl, pby, pln))
        peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1400.0)))
        gemm_tf = (pfl[0] / (pms[0] * 1e-3)) / 1e12 if pms[0] > 0 else 0.0
        # dominant kernel = the GEMM problem shape with the largest total time in this song
        NS = 16
        sm_, sn_, sk_, sl_, st_ = ((C.c_int * NS)(), (C.c_int * NS)(), (C.c_int * NS)(), (C.c_int * NS)(),
                                   (C.c_float * NS)())
        ns = lib.ace_profile_gemm_shapes(NS, sm_, sn_, sk_, sl_, st_)
        shapes = [{"m": sm_[i], "n": sn_[i], "k": sk_[i], "launches": sl_[i], "ms_total": round(float(st_[i]), 3),
                   "tflops": (2.0 * sm_[i] * sn_[i] * sk_[i] * sl_[i] / (st_[i] * 1e-3) / 1e12) if st_[i] > 0 else 0.0}
                  for i in range(ns)]
        dom = shapes[0] if shapes else {"m": 0, "n": 0, "k": 0, "launches": 0, "ms_total": 0.0, "tflops": 0.0}
        traffic = None
        for name in ("r2_gemm_traffic.json", "r1_gemm_traffic.json"):
            tpath = os.path.join(ROOT, "profiles", name)
            if os.path.exists(tpath):  # dram__bytes_read+write per launch from the committed ncu --set full capture
                with open(tpath) as f:
                    traffic = json.load(f).get(f"{dom['m']}x{dom['n']}x{dom['k']}")
                if traffic is not None:
                    break
        audio_s = wl["seconds"] * world * args.steps
        value = audio_s / (ms * 1e-3)
        clk = clocks.stop()
        clk["sm_mhz"] = my_clk or clk["sm_mhz"]
        clk["sm_mhz_e2e_region"] = clocks.window(*m["win_e2e"])
        line = {
            "metric": "generated-audio-sec/wall-sec", "value": value, "unit": "audio-s/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic (random-init weights, N(0,1) text-conditioning embeddings, seeded noise drawn inside "
                    "the timed region like the reference's prepare_noise)",
            "config": {"workload": wl["desc"], "songs_per_gpu_per_step": 1, "latent_frames": main.T,
                       "cond_tokens": main.E,
                       "parallelism": (f"songs sharded 1/GPU x{world} through acestep_b200.multi_gpu.generate_sharded "
                                       "(async NCCL waveform gather to rank 0, drained inside the timed region)"
                                       if world > 1 else "1 GPU"),
                       "l2_policy": "working set per step (3.2 GB weights) exceeds the 126 MB L2",
                       "outputs_finite": m["finite"]},
            "e2e": {"value": audio_s / e2e_s, "unit": "audio-s/s", "h2d_bytes_per_step": m["h2d"],
                    "d2h_bytes_per_step": m["d2h"]},
            "gpu_launches": m["launches"],
            "clocks": clk,
            "per_rank": per_rank,
            "roofline": {"bound": "tensor",
                         "kernel": (f"gemm_tc2_kernel M={dom['m']} N={dom['n']} K={dom['k']} (largest share of the "
                                    f"song: {dom['launches']} launches, {dom['ms_total']} ms)"),
                         "achieved": dom["tflops"], "peak": peak, "unit": "TFLOP/s", "frac": dom["tflops"] / peak,
                         "peak_source": f"{peak_kind} bf16_tflops_sustained (kernel timed inside a long step)",
                         "traffic": traffic,
                         "algorithmic_flops_per_launch": 2.0 * dom["m"] * dom["n"] * dom["k"],
                         "avg_launch_us": (dom["ms_total"] / dom["launches"] * 1e3) if dom["launches"] else None,
                         "all_gemm_launches": {"achieved": gemm_tf, "frac": gemm_tf / peak, "launches": int(pln[0]),
                                               "ms_per_song": float(pms[0])},
                         "by_shape": shapes[:8]},
            "breakdown_ms_per_song": {"gemm": float(pms[0]), "attention": float(pms[1]),
                                      "elementwise": float(pms[2]), "simt_conv": float(pms[3]),
                                      "launches": [int(pln[i]) for i in range(4)],
                                      "attention_tflops": (pfl[1] / (pms[1] * 1e-3)) / 1e12 if pms[1] > 0 else 0.0,
                                      "elementwise_gbs": (pby[2] / (pms[2] * 1e-3)) / 1e9 if pms[2] > 0 else 0.0},
            "dit_step_tensor_util": util_block(main, ms_rank / args.steps, main.codec_ms(), peaks),
        }
    else:
        clocks.stop()

    # ---- the other single-GPU workloads of BASELINE.json in the same run (N = 1 only): 240 s / 60 steps is
    # north_star's own target, c5 the repaint chain, c1 the 10 s turbo clip
    if world == 1 and not args.no_extra:
        extra = {}
        for name in ("c3", "c5", "c1", "c2"):
            if WORKLOADS[name] is wl:
                continue
            w2 = WORKLOADS[name]
            try:
                ck = ClockSampler(local)
                ck.start()
                r2 = Runner(pipe, w2, dshape, vshape, dev, rank)
                k = max(2, min(args.steps, 3))
                m2 = measure(r2, k, 1)
                c2 = ck.stop()
                entry = {"workload": w2["desc"], "steps": k, "warmup": 1,
                         "value": w2["seconds"] * k / (m2["ms"] * 1e-3), "ms_per_step": m2["ms"] / k,
                         "e2e": {"value": w2["seconds"] * k / m2["e2e_s"], "h2d_bytes_per_step": m2["h2d"],
                                 "d2h_bytes_per_step": m2["d2h"]},
                         "gpu_launches": m2["launches"], "outputs_finite": m2["finite"],
                         "clocks": {"sm_mhz": ck.window(*m2["win_value"]) or c2["sm_mhz"], "reasons": c2["reasons"]}}
                entry.update(util_block(r2, m2["ms"] / k, r2.codec_ms(), peaks))
                extra[name] = entry
                del r2
                torch.cuda.empty_cache()
            except Exception as exc:  # an extra workload never takes the headline line down with it
                extra[name] = {"error": f"{type(exc).__name__}: {exc}"}
        line["extra_workloads"] = extra
    pipe.close()
    main.pipe = None
    del pipe
    torch.cuda.empty_cache()

    if rank == 0 and world == 1 and not args.no_gpu_baseline and not main.rp:
        # the reference's own GPU path on this B200, same workload, same run (SURVEY §8c last row)
        try:
            gb = time_torch_gpu(wl, dev, dshape, vshape, dit_state, vae_state, main.dev_in, main.dev_in["null_emb"])
            gb["speedup_of_this_repo"] = line["value"] / gb["value"]
            gb["cublas_vs_this_repo_by_gemm_shape"] = cublas_gemm_ab(line["roofline"]["by_shape"], dev)
            line["gpu_baseline"] = gb
        except Exception as exc:
            line["gpu_baseline"] = {"value": None, "error": f"{type(exc).__name__}: {exc}"}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        T, E, rp = main.T, main.E, main.rp
        try:
            one_sample, threads = cpu_sample(wl)
            one_sample()
            a, b = one_sample()
            song = wl["steps"] * a + (T / 64.0) * b * (2 if rp else 1)
            line["cpu_baseline"] = {
                "value": wl["seconds"] / song, "unit": "audio-s/s", "cores": threads, "kind": "port",
                "sample": (f"1 DiT forward (effective batch {2 if wl['guidance'] > 1 else 1}, T={T}, E={E}) = {a:.2f} s and one "
                           f"64-frame VAE decode window = {b:.2f} s on the host CPU (fp32 oracle port), "
                           f"extrapolated to {wl['steps']} forwards + {T}/64 windows (the full song would take "
                           f"{song:.0f} s of CPU time per step)")}
        except Exception as exc:  # the baseline is reported, never required
            line["cpu_baseline"] = {"value": None, "unit": "audio-s/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"failed: {exc}"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "torch-gpu"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra_workloads block (c3 / c5 / c1)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, wl)
    if args.impl == "torch-gpu":
        return run_torch_gpu(args, wl)
    return run_b200(args, wl)


if __name__ == "__main__":
    sys.exit(main())
