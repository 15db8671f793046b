#!/usr/bin/env python
"""One DiT step (C2 shape: effective batch 2, T=1500, E=512) + one 1500-frame VAE decode between
cudaProfilerStart/Stop, for `ncu --profile-from-start off` (see profiles/README.md)."""
import os
import sys

# (ncu profiles the kernel nodes of a CUDA-graph launch individually: the release library is profiled as shipped)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from acestep_b200.dit import B200DiT, DiTShape
from acestep_b200.synthetic import random_dit_state, random_vae_state
from acestep_b200.vae import B200Vae, VaeShape

dev = torch.device("cuda:0")
T = int(os.environ.get("PROF_T", "1500"))
E, Bc = 512, 2
dit = B200DiT(random_dit_state(DiTShape(), 0, dev), DiTShape(), dev)
vae = B200Vae(random_vae_state(VaeShape(), 0, dev), VaeShape(), dev)
g = torch.Generator(device=dev).manual_seed(0)
xt = torch.randn(Bc, T, 64, device=dev, generator=g).bfloat16()
ctx = torch.randn(Bc, T, 128, device=dev, generator=g).bfloat16()
enc = torch.randn(Bc, E, 2048, device=dev, generator=g).bfloat16()
dit.bind(Bc, T, E)
dit.set_condition(enc)
for _ in range(2):
    dit.step(xt, ctx, [0.5] * Bc)
vae.decode_frames(xt[0])
torch.cuda.synchronize()
torch.cuda.profiler.start()
dit.step(xt, ctx, [0.5] * Bc)
if os.environ.get("PROF_VAE", "1") == "1":
    vae.decode_frames(xt[0])
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one DiT step + one VAE decode")
