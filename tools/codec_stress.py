#!/usr/bin/env python
"""Replay stress for the codec kernels (fused residual unit with the row-stepped xs box + weight ring, tap-shifted
GEMMs): the same decode / encode many times at several lengths, interleaved with an L2-thrashing fill; any difference
between replays (or a launch failure from the bounded barrier waits) is a race.  Dev tool."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from acestep_b200.synthetic import random_vae_state
from acestep_b200.vae import B200Vae, VaeShape

dev = torch.device("cuda:0")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
vae = B200Vae(random_vae_state(VaeShape(), 0, dev), VaeShape(), dev)
junk = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
bad = 0
for T in (37, 300, 1500, 3000):
    g = torch.Generator(device=dev).manual_seed(T)
    z = torch.randn(T, 64, device=dev, generator=g).bfloat16()
    ref = vae.decode_frames(z).clone()
    audio = ref[:, : T * vae.shape.hop].contiguous()
    ref_m = vae.encode_samples(audio, None).clone()
    diff = 0
    for i in range(reps):
        if i % 3 == 0:
            junk.fill_(i & 255)
        if not torch.equal(vae.decode_frames(z), ref):
            diff += 1
        if not torch.equal(vae.encode_samples(audio, None), ref_m):
            diff += 1
    torch.cuda.synchronize()
    print(f"T={T}: {diff} of {2 * reps} codec replays differ; finite={bool(torch.isfinite(ref).all())}")
    bad += diff
sys.exit(1 if bad else 0)
