#!/usr/bin/env python
"""Where one song's wall time goes (C2 by default): sampler loop, VAE decode, post-processing, host
copies.  Wall clock with synchronisation around each piece + CUDA events; run on the GPU box."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from acestep_b200.dit import DiTShape
from acestep_b200.pipeline import B200Pipeline
from acestep_b200.synthetic import random_dit_state, random_vae_state, synthetic_conditioning
from acestep_b200.vae import VaeShape

dev = torch.device("cuda:0")
T = int(os.environ.get("PROF_T", "1500"))
STEPS = int(os.environ.get("PROF_STEPS", "27"))
E = 512
dshape, vshape = DiTShape(), VaeShape()
pipe = B200Pipeline(random_dit_state(dshape, 0, dev), random_vae_state(vshape, 0, dev), dshape, vshape, device=dev)
host = synthetic_conditioning(1, T, E, dshape.hidden_size, seed=1234, device="cpu", pin=True)
noise_h = torch.randn(1, T, 64, generator=torch.Generator().manual_seed(0)).to(torch.bfloat16).pin_memory()
dev_in = {k: v.to(dev) for k, v in host.items()}
noise_d = noise_h.to(dev)
pipe.sampler.null_condition_emb = dev_in["null_emb"]
skw = dict(infer_steps=STEPS, diffusion_guidance_sale=7.0, shift=3.0)


def sync():
    torch.cuda.synchronize(dev)


def wall(fn, n=3):
    fn()
    sync()
    t0 = time.perf_counter()
    for _ in range(n):
        r = fn()
    sync()
    return (time.perf_counter() - t0) / n * 1e3, r


ms_full_dev, _ = wall(lambda: pipe.generate(dev_in["enc"], dev_in["ctx"], dev_in["src"], None, noise=noise_d, to_host=False, **skw))
ms_full_host, _ = wall(lambda: pipe.generate(host["enc"], host["ctx"], host["src"], None, noise=noise_h, to_host=True, **skw))
ms_nodec, out = wall(lambda: pipe.generate(dev_in["enc"], dev_in["ctx"], dev_in["src"], None, noise=noise_d, to_host=False, decode=False, **skw))
ms_samp, out = wall(lambda: pipe.sampler.generate_base(dev_in["enc"], dev_in["ctx"], dev_in["src"], None, noise=noise_d, **skw))
lat = out["target_latents"]
ms_vae, wav = wall(lambda: pipe.vae.decode_frames(lat[0]))
wav3 = wav[None]


def post():
    peak = wav3.abs().amax(dim=[1, 2], keepdim=True)
    w = wav3
    if torch.any(peak > 1.0):
        w = wav3 / peak.clamp(min=1.0)
    return w


ms_post, w = wall(post)
pinned = torch.empty(w.shape, dtype=torch.float32, pin_memory=True)
ms_d2h, _ = wall(lambda: pinned.copy_(w, non_blocking=True))
ms_h2d, _ = wall(lambda: [host[k].to(dev, non_blocking=True) for k in ("enc", "ctx", "src")])
# the bare DiT step in a tight loop (graph replay)
Bc = 2
x2 = torch.randn(Bc, T, 64, device=dev).bfloat16()
ctx2 = torch.cat([dev_in["ctx"], dev_in["ctx"]], 0)
vt = torch.empty_like(x2)
ms_step, _ = wall(lambda: pipe.dit.step(x2, ctx2, [0.5] * Bc, out=vt), n=30)
print(f"T={T} steps={STEPS}: generate(dev)={ms_full_dev:.2f} ms  generate(host,to_host)={ms_full_host:.2f} ms  "
      f"no-decode={ms_nodec:.2f}  sampler={ms_samp:.2f} (bare step x{STEPS} = {ms_step * STEPS:.2f}, step {ms_step:.3f})  "
      f"vae={ms_vae:.2f}  post={ms_post:.2f}  d2h={ms_d2h:.2f} ({w.numel() * 4 / ms_d2h / 1e6:.1f} GB/s)  h2d={ms_h2d:.3f}")
