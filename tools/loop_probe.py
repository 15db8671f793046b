import sys, time
sys.path.insert(0, "/root/repo")
import torch
from acestep_b200 import _lib
from acestep_b200.dit import B200DiT, DiTShape
from acestep_b200.synthetic import random_dit_state
dev = torch.device("cuda:0")
T, E, Bc, B = 1500, 512, 2, 1
dit = B200DiT(random_dit_state(DiTShape(), 0, dev), DiTShape(), dev)
lib = dit.lib
g = torch.Generator(device=dev).manual_seed(0)
dit.bind(Bc, T, E)
enc = torch.randn(Bc, E, 2048, device=dev, generator=g).bfloat16()
dit.set_condition(enc)
xin, ctxin, vt = dit.io_views()
xin.copy_(torch.randn(Bc, T, 64, device=dev, generator=g).bfloat16()); ctxin.copy_(torch.randn(Bc, T, 128, device=dev, generator=g).bfloat16())
vg = torch.empty(B, T, 64, device=dev, dtype=torch.bfloat16); mom = torch.zeros_like(vg)
st = _lib.stream_handle(dev)
def run(n, with_apg, with_euler, with_sync_free=True):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for i in range(n):
        dit.step(xin, ctxin, [0.5] * Bc, out=vt)
        if with_apg:
            lib.ace_apg(vt[:B].data_ptr(), vt[B:].data_ptr(), mom.data_ptr(), 0, -0.75, 2.5, 7.0, vg.data_ptr(), B, T, st)
        if with_euler:
            lib.ace_euler_step_dup(xin[:B].data_ptr(), vg.data_ptr(), 1e-6, B * T * 64, xin[B:].data_ptr(), st)
    t1 = time.perf_counter(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (t1 - t0) / n * 1e3
for _ in range(2): run(5, True, True)
# alternate the variants so that clock drift cancels
acc = {}
for rep in range(6):
    for name, a, e in (("step only", False, False), ("step+apg", True, False), ("step+apg+euler", True, True)):
        gpu, cpu = run(27, a, e)
        acc.setdefault(name, []).append(gpu)
for name, v in acc.items():
    print(f"{name:16s}: GPU {sum(v) / len(v):.3f} ms/iter  (runs: {' '.join(f'{x:.3f}' for x in v)})")
