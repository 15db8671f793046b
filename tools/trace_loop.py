#!/usr/bin/env python
"""Kernel timeline of the sampler loop (CUPTI through torch.profiler): where the time between two DiT-step graphs
goes.  Dev tool.  usage: python tools/trace_loop.py [T] [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import ProfilerActivity, profile

from acestep_b200.dit import B200DiT, DiTShape
from acestep_b200.sampler import B200Sampler
from acestep_b200.synthetic import random_dit_state

dev = torch.device("cuda:0")
T = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 27
dit = B200DiT(random_dit_state(DiTShape(), 0, dev), DiTShape(), dev)
g = torch.Generator(device=dev).manual_seed(0)
enc = torch.randn(1, 512, 2048, device=dev, generator=g).bfloat16()
ctx = torch.randn(1, T, 128, device=dev, generator=g).bfloat16()
src = torch.randn(1, T, 64, device=dev, generator=g).bfloat16()
null = torch.randn(1, 1, 2048, device=dev, generator=g).bfloat16()
smp = B200Sampler(dit, null_condition_emb=null)
kw = dict(infer_steps=steps, diffusion_guidance_sale=7.0, shift=3.0, seed=[1])
for _ in range(2):
    smp.generate_base(enc, ctx, src, **kw)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
smp.generate_base(enc, ctx, src, **kw)
e1.record()
torch.cuda.synchronize()
print(f"loop {e0.elapsed_time(e1):.2f} ms for {steps} steps = {e0.elapsed_time(e1) / steps:.3f} ms per step")
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    smp.generate_base(enc, ctx, src, **kw)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
print(len(ev), "device activities")
# find the boundaries of the graphs: the set_slots kernel precedes each graph launch
idx = [i for i, e in enumerate(ev) if "set_slots" in e.name]
print("steps seen:", len(idx))
if len(idx) > 12:
    a, b = idx[10], idx[11]
    step_us = ev[b].time_range.start - ev[a].time_range.start
    print(f"step 10: {step_us:.1f} us between consecutive set_slots kernels; {b - a} activities")
    busy = sum(e.time_range.end - e.time_range.start for e in ev[a:b])
    print(f"  sum of kernel durations {busy:.1f} us; idle {step_us - busy:.1f} us")
    # biggest gaps
    gaps = []
    for i in range(a, b):
        gap = ev[i + 1].time_range.start - ev[i].time_range.end
        gaps.append((gap, ev[i].name[:60], ev[i + 1].name[:60]))
    gaps.sort(reverse=True)
    for gp in gaps[:8]:
        print(f"  gap {gp[0]:7.1f} us after {gp[1]} before {gp[2]}")
    # the non-graph part: from the last graph kernel to the next set_slots
    for e in ev[b - 8:b + 2]:
        print(f"  {e.time_range.start - ev[a].time_range.start:9.1f} +{e.time_range.end - e.time_range.start:7.1f} us  {e.name[:90]}")
