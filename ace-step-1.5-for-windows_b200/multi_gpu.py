"""Whole-song sharding across the GPUs of one box: one process per GPU, zero collectives during
denoise/decode, ONE gather of the finished waveforms.

The reference batches up to 8 songs on a single GPU (gpu_config.py:294-299) and decodes them
sequentially (handler/vae_decode_chunks.py:18-29); each song's loop is independent (per-sample
seeds: prepare_noise list branch, turbo modeling :1749-1761), so song i simply goes to rank
i % world with replicated weights and the result is independent of placement.

Works with any initialised torch.distributed backend: NCCL over NVLink on the B200 box,
gloo in the CPU tests.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import torch
import torch.distributed as dist


def shard_indices(n_items: int, rank: int, world: int) -> List[int]:
    """Songs handled by `rank`: i ≡ rank (mod world)."""
    return list(range(rank, n_items, world))


def gather_waveforms(local: Sequence[torch.Tensor], n_items: int, dst: int = 0, group=None,
                     device: Optional[torch.device] = None) -> Optional[List[torch.Tensor]]:
    """Gather per-song waveforms ([C, N_i], ragged N_i allowed) produced under `shard_indices`
    onto rank `dst`, returned in global song order (None on other ranks).

    Collectives: one all_gather of the lengths (a few bytes) + one gather of the padded waveforms.
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    mine = shard_indices(n_items, rank, world)
    assert len(local) == len(mine), f"rank {rank}: expected {len(mine)} waveforms, got {len(local)}"
    per_rank = (n_items + world - 1) // world
    if device is None:
        device = local[0].device if len(local) else torch.device("cpu")
    chans = local[0].shape[0] if len(local) else 2
    dtype = local[0].dtype if len(local) else torch.float32
    lens = torch.zeros(per_rank, dtype=torch.int64, device=device)
    for j, w in enumerate(local):
        lens[j] = w.shape[-1]
    all_lens = [torch.zeros_like(lens) for _ in range(world)]
    dist.all_gather(all_lens, lens, group=group)
    max_len = int(torch.stack(all_lens).max().item())
    buf = torch.zeros(per_rank, chans, max_len, dtype=dtype, device=device)
    for j, w in enumerate(local):
        buf[j, :, : w.shape[-1]] = w
    out = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, out, dst=dst, group=group)
    if rank != dst:
        return None
    songs: List[Optional[torch.Tensor]] = [None] * n_items
    for r in range(world):
        for j, i in enumerate(shard_indices(n_items, r, world)):
            songs[i] = out[r][j, :, : int(all_lens[r][j])]
    return songs  # type: ignore[return-value]


def generate_sharded(generate_one: Callable[[int], torch.Tensor], n_items: int, dst: int = 0, group=None,
                     device: Optional[torch.device] = None) -> Optional[List[torch.Tensor]]:
    """Run `generate_one(i) -> waveform [C, N]` for this rank's songs, then gather on `dst`."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    local = [generate_one(i) for i in shard_indices(n_items, rank, world)]
    return gather_waveforms(local, n_items, dst=dst, group=group, device=device)
