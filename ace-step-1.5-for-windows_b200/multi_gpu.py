"""Whole-song sharding across the GPUs of one box: one process per GPU, zero collectives during
denoise/decode, ONE gather of the finished waveforms.

The reference batches up to 8 songs on a single GPU (gpu_config.py:294-299) and decodes them
sequentially (handler/vae_decode_chunks.py:18-29); each song's loop is independent (per-sample
seeds: prepare_noise list branch, turbo modeling :1749-1761), so song i simply goes to rank
i % world with replicated weights and the result is independent of placement.

Works with any initialised torch.distributed backend: NCCL over NVLink on the B200 box,
gloo in the CPU tests.

Two gather modes:
  * ragged (default): lengths are all-gathered first (one tiny collective + a host read), then one padded
    gather.  Needed when ranks cannot know each other's durations.
  * `lengths=` given (every rank knows every song's sample count — it is `duration * 48000`, request
    metadata): no length exchange, no host synchronisation; with `async_op=True` the gather is only ENQUEUED
    (on NCCL's stream) and a `PendingGather` is returned, so rank r's next song overlaps the transfer.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import torch
import torch.distributed as dist


def shard_indices(n_items: int, rank: int, world: int) -> List[int]:
    """Songs handled by `rank`: i ≡ rank (mod world)."""
    return list(range(rank, n_items, world))


class PendingGather:
    """A waveform gather in flight.  `wait()` -> songs in global order on `dst`, None elsewhere."""

    def __init__(self, work, out, lens, n_items: int, world: int, is_dst: bool, keepalive):
        self._work, self._out, self._lens = work, out, lens
        self._n, self._world, self._is_dst = n_items, world, is_dst
        self._keepalive = keepalive  # the send buffer must outlive the collective

    def wait(self) -> Optional[List[torch.Tensor]]:
        if self._work is not None:
            self._work.wait()
            self._work = None
        self._keepalive = None
        if not self._is_dst:
            return None
        songs: List[Optional[torch.Tensor]] = [None] * self._n
        for r in range(self._world):
            for j, i in enumerate(shard_indices(self._n, r, self._world)):
                songs[i] = self._out[r][j, :, : int(self._lens[i])]
        return songs  # type: ignore[return-value]


class GatherPipeline:
    """Bounded queue of in-flight gathers for serving loops: `submit()` waits for (and returns the songs of) the
    oldest gather once more than `depth` are pending.  A gather issued one song ago has long finished, so that wait
    is free — what it buys is that its send / receive buffers go back to the caching allocator before the next song
    needs memory (torch keeps the tensors of an un-waited NCCL op alive; a loop that only waits at the very end
    makes the allocator cudaMalloc fresh blocks every song, ~20 ms each on a B200 box)."""

    def __init__(self, depth: int = 1):
        self.depth, self._q = max(0, int(depth)), []

    def submit(self, pending: "PendingGather"):
        self._q.append(pending)
        if len(self._q) > self.depth:
            return self._q.pop(0).wait()
        return None

    def drain(self):
        """Wait for everything still in flight; returns the songs of the LAST gather (None off `dst`)."""
        res = None
        while self._q:
            res = self._q.pop(0).wait()
        return res


def gather_waveforms(local: Sequence[torch.Tensor], n_items: int, dst: int = 0, group=None,
                     device: Optional[torch.device] = None, lengths: Optional[Sequence[int]] = None,
                     async_op: bool = False):
    """Gather per-song waveforms ([C, N_i], ragged N_i allowed) produced under `shard_indices`
    onto rank `dst`, returned in global song order (None on other ranks).

    Collectives: (ragged mode) one all_gather of the lengths + one gather of the padded waveforms;
    (`lengths` given: global per-song sample counts) the one gather only.  `async_op=True` returns a
    PendingGather instead of waiting.
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
    if lengths is None:
        lens = torch.zeros(per_rank, dtype=torch.int64, device=device)
        for j, w in enumerate(local):
            lens[j] = w.shape[-1]
        all_lens = [torch.zeros_like(lens) for _ in range(world)]
        dist.all_gather(all_lens, lens, group=group)
        table = torch.stack(all_lens).cpu()  # the one host read of the ragged mode
        glob = [0] * n_items
        for r in range(world):
            for j, i in enumerate(shard_indices(n_items, r, world)):
                glob[i] = int(table[r][j])
    else:
        glob = [int(x) for x in lengths]
        assert len(glob) == n_items, f"lengths has {len(glob)} entries for {n_items} songs"
        for j, i in enumerate(mine):
            assert local[j].shape[-1] == glob[i], f"song {i}: {local[j].shape[-1]} samples, lengths says {glob[i]}"
    max_len = max(glob) if glob else 0
    if per_rank == 1 and len(local) == 1 and local[0].shape[-1] == max_len and local[0].is_contiguous():
        buf = local[0].unsqueeze(0)  # the common case (one equal-length song per rank): no staging copy
    else:
        buf = torch.zeros(per_rank, chans, max_len, dtype=dtype, device=device)
        for j, w in enumerate(local):
            buf[j, :, : w.shape[-1]] = w
    out = [torch.empty(per_rank, chans, max_len, dtype=dtype, device=device) for _ in range(world)] if rank == dst else None
    work = dist.gather(buf, out, dst=dst, group=group, async_op=True)
    pending = PendingGather(work, out, glob, n_items, world, rank == dst, buf)
    return pending if async_op else pending.wait()


def generate_sharded(generate_one: Callable[[int], torch.Tensor], n_items: int, dst: int = 0, group=None,
                     device: Optional[torch.device] = None, lengths: Optional[Sequence[int]] = None,
                     async_op: bool = False, gather_stream=None):
    """Run `generate_one(i) -> waveform [C, N]` for this rank's songs, then gather on `dst`
    (see gather_waveforms for `lengths` / `async_op`).  `gather_stream`: the CUDA stream the waveforms are produced
    on when that is not the current one (B200Pipeline.codec_stream): the collective is queued behind it, so the
    caller's stream — and the next song's denoising loop — does not wait for this song's decode."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    local = [generate_one(i) for i in shard_indices(n_items, rank, world)]
    if gather_stream is None:
        return gather_waveforms(local, n_items, dst=dst, group=group, device=device, lengths=lengths,
                                async_op=async_op)
    with torch.cuda.stream(gather_stream):
        return gather_waveforms(local, n_items, dst=dst, group=group, device=device, lengths=lengths,
                                async_op=async_op)
