#!/usr/bin/env python
"""Times the waveform gather of acestep_b200.multi_gpu on its own (torchrun, NCCL): one 60 s stereo fp32 waveform
(23 MB) per rank to rank 0, synchronous and asynchronous, to attribute multi-GPU step time."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

from acestep_b200.multi_gpu import gather_waveforms

local = int(os.environ.get("LOCAL_RANK", "0"))
dev = torch.device(f"cuda:{local}")
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
n = 2880000
wav = torch.randn(2, n, device=dev)
for mode in ("ragged", "lengths", "async"):
    for it in range(6):
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        if mode == "ragged":
            out = gather_waveforms([wav], world, dst=0, device=dev)
        elif mode == "lengths":
            out = gather_waveforms([wav], world, dst=0, device=dev, lengths=[n] * world)
        else:
            p = gather_waveforms([wav], world, dst=0, device=dev, lengths=[n] * world, async_op=True)
            t_issue = time.perf_counter() - t0
            out = p.wait()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if rank == 0 and it >= 2:
            extra = f" (issue {t_issue * 1e3:.2f} ms)" if mode == "async" else ""
            print(f"{mode:8s} it{it}: {dt * 1e3:.2f} ms{extra}", flush=True)
# raw collective, preallocated outputs
outs = [torch.empty(1, 2, n, device=dev) for _ in range(world)] if rank == 0 else None
for it in range(5):
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    dist.gather(wav.unsqueeze(0), outs, dst=0)
    torch.cuda.synchronize()
    if rank == 0 and it >= 2:
        print(f"raw gather it{it}: {(time.perf_counter() - t0) * 1e3:.2f} ms", flush=True)
dist.destroy_process_group()
