"""world_size-2 gloo test (CPU) of the N>1 path: song sharding + ragged waveform gather."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from acestep_b200.multi_gpu import GatherPipeline, gather_waveforms, generate_sharded, shard_indices


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _song(i):
    n = 1000 + 37 * i  # ragged lengths
    return torch.arange(2 * n, dtype=torch.float32).view(2, n) + 1000.0 * i


def _worker(rank, world, port, n_items, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        out = generate_sharded(_song, n_items, dst=0)  # ragged mode: lengths exchanged first
        # known-lengths mode, asynchronous: no length exchange, the gather is waited for later
        lens = [_song(i).shape[-1] for i in range(n_items)]
        pend = generate_sharded(_song, n_items, dst=0, lengths=lens, async_op=True)
        # one equal-length song per rank (the benchmark's shape): zero-copy send buffer
        pend1 = generate_sharded(lambda i: _song(0) + i, world, dst=0, lengths=[_song(0).shape[-1]] * world, async_op=True)
        gp = GatherPipeline(depth=1)
        assert gp.submit(pend) is None          # one in flight: nothing retired yet
        out2 = gp.submit(pend1)                  # second submit retires (waits for) the first
        out3 = gp.drain()
        if rank == 0:
            ok = len(out) == n_items and all(torch.equal(out[i], _song(i)) for i in range(n_items))
            ok = ok and len(out2) == n_items and all(torch.equal(out2[i], _song(i)) for i in range(n_items))
            ok = ok and len(out3) == world and all(torch.equal(out3[i], _song(0) + i) for i in range(world))
            q.put(bool(ok))
        else:
            assert out is None and out2 is None and out3 is None
    finally:
        dist.destroy_process_group()


def test_shard_indices_partition():
    for n, w in [(8, 2), (5, 2), (3, 4), (0, 2)]:
        got = sorted(i for r in range(w) for i in shard_indices(n, r, w))
        assert got == list(range(n))


def test_sharded_generate_and_gather_world2():
    for n_items in (5, 2):
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, q)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
            assert p.exitcode == 0
        assert q.get(timeout=5) is True
