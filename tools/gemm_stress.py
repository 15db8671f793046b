#!/usr/bin/env python
"""Replay stress for the multi-wave residual GEMMs (every-tile tail variant): the same DiT step, many times, at
shapes with more pair tiles than CTA pairs, interleaved with an L2-thrashing copy so that TMA latencies move around;
any difference between replays is a race.  Dev tool.  usage: python tools/gemm_stress.py [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from acestep_b200.dit import B200DiT, DiTShape
from acestep_b200.synthetic import random_dit_state

dev = torch.device("cuda:0")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
shape = DiTShape(num_hidden_layers=2)
dit = B200DiT(random_dit_state(shape, 0, dev), shape, dev)
junk = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
bad = 0
for T in (3000, 6001, 7001, 12000):
    g = torch.Generator(device=dev).manual_seed(T)
    xt = torch.randn(2, T, 64, device=dev, generator=g).bfloat16()
    ctx = torch.randn(2, T, 128, device=dev, generator=g).bfloat16()
    enc = torch.randn(2, 70, shape.hidden_size, device=dev, generator=g).bfloat16()
    dit.bind(2, T, 70)
    dit.set_condition(enc)
    ref = dit.step(xt, ctx, [0.5, 0.25]).clone()
    diff = 0
    for i in range(reps):
        if i % 3 == 0:
            junk.fill_(i & 255)
        out = dit.step(xt, ctx, [0.5, 0.25])
        if not torch.equal(out, ref):
            diff += 1
    torch.cuda.synchronize()
    print(f"T={T}: {diff} of {reps} replays differ; finite={bool(torch.isfinite(ref.float()).all())}")
    bad += diff
sys.exit(1 if bad else 0)
