#!/usr/bin/env python
"""Real (CUDA-graph + PDL) time of one DiT step and one VAE decode at a bench shape; run several
times with ACE_SKIP=attn|norm|gemm to read each kernel class's share off by difference."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if any(k.startswith("ACE_") and k != "ACE_B200_LIB" for k in os.environ):
    # the A/B switches only exist in the probe build (-DACE_PROBE); the release library ignores the environment
    os.environ.setdefault("ACE_B200_LIB", os.path.join(ROOT, "ace-step-1.5-for-windows_b200", "libacestep_b200_probe.so"))
import torch

from acestep_b200.dit import B200DiT, DiTShape
from acestep_b200.synthetic import random_dit_state, random_vae_state
from acestep_b200.vae import B200Vae, VaeShape

dev = torch.device("cuda:0")
T = int(os.environ.get("PROF_T", "1500"))
E, Bc = int(os.environ.get("PROF_E", "512")), int(os.environ.get("PROF_BC", "2"))  # PROF_E: condition tokens (SURVEY §8d side table: 64 / 512 / 2305)
dit = B200DiT(random_dit_state(DiTShape(), 0, dev), DiTShape(), dev)
g = torch.Generator(device=dev).manual_seed(0)
xt = torch.randn(Bc, T, 64, device=dev, generator=g).bfloat16()
ctx = torch.randn(Bc, T, 128, device=dev, generator=g).bfloat16()
enc = torch.randn(Bc, E, 2048, device=dev, generator=g).bfloat16()
dit.bind(Bc, T, E)
dit.set_condition(enc)
out = torch.empty_like(xt)
for _ in range(3):
    dit.step(xt, ctx, [0.5] * Bc, out=out)
torch.cuda.synchronize()
n = 30
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n):
    dit.step(xt, ctx, [0.5] * Bc, out=out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
msg = f"skip={os.environ.get('ACE_SKIP', '-'):10s} T={T} Bc={Bc} E={E}: DiT step {ms:.3f} ms"
if os.environ.get("PROF_VAE", "1") == "1" and not os.environ.get("ACE_SKIP"):
    vae = B200Vae(random_vae_state(VaeShape(), 0, dev), VaeShape(), dev)
    z = xt[0]
    for _ in range(2):
        vae.decode_frames(z)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        vae.decode_frames(z)
    e1.record()
    torch.cuda.synchronize()
    msg += f"; VAE decode {e0.elapsed_time(e1) / 5:.3f} ms"
print(msg)
