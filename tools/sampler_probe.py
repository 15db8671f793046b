#!/usr/bin/env python
"""Where the base sampler's time goes besides the DiT steps: CUDA-event time of set_condition, of the loop,
and of the whole generate_base call (C2 shape)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from acestep_b200.dit import B200DiT, DiTShape
from acestep_b200.sampler import B200Sampler
from acestep_b200.synthetic import random_dit_state, synthetic_conditioning
dev = torch.device("cuda:0")
T, E = 1500, 512
dit = B200DiT(random_dit_state(DiTShape(), 0, dev), DiTShape(), dev)
c = synthetic_conditioning(1, T, E, 2048, seed=1, device=dev)
s = B200Sampler(dit, c["null_emb"])
noise = torch.randn(1, T, 64, device=dev).bfloat16()
kw = dict(infer_steps=27, diffusion_guidance_sale=7.0, shift=3.0, noise=noise)
for _ in range(2):
    s.generate_base(c["enc"], c["ctx"], c["src"], None, **kw)
def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e
res = []
for rep in range(4):
    torch.cuda.synchronize()
    a = ev(); t0 = time.perf_counter()
    out = s.generate_base(c["enc"], c["ctx"], c["src"], None, **kw)
    t1 = time.perf_counter(); b = ev(); torch.cuda.synchronize()
    enc2 = torch.cat([c["enc"], c["null_emb"].expand_as(c["enc"])], 0)
    c0 = ev(); dit.set_condition(enc2); c1 = ev()
    xin, ctxin, vt = dit.io_views()
    l0 = ev()
    for i in range(27):
        dit.step(xin, ctxin, [0.5, 0.5], out=vt)
    l1 = ev(); torch.cuda.synchronize()
    res.append((a.elapsed_time(b), (t1 - t0) * 1e3, c0.elapsed_time(c1), l0.elapsed_time(l1)))
for r in res:
    print(f"generate_base: GPU {r[0]:.2f} ms (wall {r[1]:.2f}) | set_condition {r[2]:.3f} ms | 27 bare steps {r[3]:.2f} ms | rest {r[0] - r[2] - r[3]:.2f} ms")
