#!/usr/bin/env python
"""Output-path kernels between cudaProfilerStart/Stop for `ncu --profile-from-start off`: peak normalisation of a
batch of two 60 s stereo waveforms (one loud: scaled; one quiet: the scale kernel leaves after 4 bytes), the
fused -1 dB variant, the latent guard, and the lyric-alignment attention extraction at the C2 shape
(Bc = 2, T = 1500, E = 512, layers 0..6).  Also prints CUDA-event timings (unprofiled runs only)."""
import os
import sys

# (ncu profiles the kernel nodes of a CUDA-graph launch individually: the release library is profiled as shipped)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from acestep_b200.dit import B200DiT, DiTShape
from acestep_b200.output import latent_flags, peak_normalize_
from acestep_b200.synthetic import random_dit_state

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
N = 2_880_000
base = torch.randn(2, 2, N, device=dev, generator=g) * 0.25
base[0] *= 3.0
lat = torch.randn(2, 1500, 64, device=dev, generator=g).bfloat16()
T, E, Bc = 1500, 512, 2
dit = B200DiT(random_dit_state(DiTShape(), 0, dev), DiTShape(), dev)
xt = torch.randn(Bc, T, 64, device=dev, generator=g).bfloat16()
ctx = torch.randn(Bc, T, 128, device=dev, generator=g).bfloat16()
enc = torch.randn(Bc, E, 2048, device=dev, generator=g).bfloat16()
dit.bind(Bc, T, E)
dit.set_condition(enc)


def timed(fn, reps=20):
    fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    big = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    tot = 0.0
    for _ in range(reps):
        big.zero_()  # flush L2 between repetitions
        ev[0].record()
        fn()
        ev[1].record()
        torch.cuda.synchronize()
        tot += ev[0].elapsed_time(ev[1])
    return tot / reps * 1e3  # us


wavs = [base.clone() for _ in range(64)]
it = iter(wavs)
us_plain = timed(lambda: peak_normalize_(next(it)), reps=20)
it = iter(wavs[32:])
us_db = timed(lambda: peak_normalize_(next(it), normalization_db=-1.0), reps=20)
us_guard = timed(lambda: latent_flags(lat), reps=20)
us_attn = timed(lambda: dit.cross_attentions(xt, ctx, [0.125] * Bc, 7), reps=5)
mb = base.numel() * 4 / 1e6
# mb MB at 1 us = mb TB/s = 1000 mb GB/s
print(f"peak_normalize (2 x 60 s, one scaled): {us_plain:.1f} us  ({mb * 2.0 / us_plain * 1e3:.0f} GB/s algorithmic: "
      f"read all, then read + write the loud half; memset + 2 launches included)")
print(f"peak_normalize -1 dB (both scaled):    {us_db:.1f} us  ({mb * 3 / us_db * 1e3:.0f} GB/s algorithmic: read, read + write)")
print(f"latent guard (2 x 1500 x 64 bf16, incl. the 8-byte read-back): {us_guard:.1f} us")
print(f"cross attentions, 7 layers at C2 (incl. the partial forward):   {us_attn:.1f} us")

w = base.clone()
torch.cuda.synchronize()
torch.cuda.profiler.start()
peak_normalize_(w)
peak_normalize_(base.clone(), normalization_db=-1.0)
latent_flags(lat)
dit.cross_attentions(xt, ctx, [0.125] * Bc, 2)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled the output-path kernels")
