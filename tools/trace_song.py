#!/usr/bin/env python
"""Kernel timeline of whole songs through B200Pipeline (CUPTI via torch.profiler): where a song's time goes outside
the DiT-step graphs and the decode.  Dev tool.  usage: python tools/trace_song.py [T] [steps]"""
import os
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import ProfilerActivity, profile

from acestep_b200.dit import DiTShape
from acestep_b200.pipeline import B200Pipeline, SongPipeline
from acestep_b200.synthetic import random_dit_state, random_vae_state, synthetic_conditioning
from acestep_b200.vae import VaeShape

dev = torch.device("cuda:0")
T = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 27
ds, vs = DiTShape(), VaeShape()
pipe = B200Pipeline(random_dit_state(ds, 0, dev), random_vae_state(vs, 0, dev), ds, vs, device=dev)
c = synthetic_conditioning(1, T, 512, ds.hidden_size, seed=1, device=dev)
pipe.sampler.null_condition_emb = c["null_emb"]
kw = dict(infer_steps=steps, diffusion_guidance_sale=7.0, shift=3.0, to_host=False)
q = SongPipeline(1)
for i in range(3):
    q.submit(pipe.generate_async(c["enc"], c["ctx"], c["src"], [i], **kw))
q.drain()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(3):
        q.submit(pipe.generate_async(c["enc"], c["ctx"], c["src"], [i], **kw))
    q.drain()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
slots = [i for i, e in enumerate(ev) if "set_slots" in e.name]
# song 2 = activities between the last set_slots of song 1 ... : take the middle song: slots[steps] .. slots[2*steps]
a, b = slots[steps], slots[2 * steps]
song_us = ev[b].time_range.start - ev[a].time_range.start
print(f"song period (first set_slots of song 2 -> first set_slots of song 3): {song_us / 1e3:.2f} ms")
last_step_start = ev[slots[2 * steps - 1]].time_range.start
print(f"  27 steps (first set_slots -> 27th set_slots) {(last_step_start - ev[a].time_range.start) / 1e3:.2f} ms "
      f"= {(last_step_start - ev[a].time_range.start) / (steps - 1):.1f} us per step")
# everything after the last step's euler kernel up to the next song's first set_slots
tail = [e for e in ev[slots[2 * steps - 1]:b]]
eul = max(i for i, e in enumerate(tail) if "euler" in e.name)
t_end_loop = tail[eul].time_range.end
print(f"  last step ends at +{(t_end_loop - ev[a].time_range.start) / 1e3:.2f} ms; between-songs section "
      f"{(ev[b].time_range.start - t_end_loop) / 1e3:.2f} ms:")
agg = defaultdict(lambda: [0, 0.0])
prev_end = t_end_loop
idle = 0.0
for e in tail[eul + 1:]:
    k = e.name.split("(")[0][-70:]
    agg[k][0] += 1
    agg[k][1] += e.time_range.end - e.time_range.start
    if e.time_range.start > prev_end:
        idle += e.time_range.start - prev_end
    prev_end = max(prev_end, e.time_range.end)
idle += max(0.0, ev[b].time_range.start - prev_end)
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"    {us:9.1f} us  x{n:4d}  {k}")
print(f"    device idle inside the section: {idle:.1f} us")
