"""Executable model of the barrier protocol of the fused codec residual unit (csrc/resunit.cuh): TMA producer warp,
MMA issuer, epilogue 1 (D1 -> snake -> hs tile) and epilogue 2 (D2 + x -> x', snake(x')), with the shared-memory
resources of the shipped kernel: ONE xs box (fetched right after the first weight blocks of its tile, once the
previous tile's conv7 has retired), a 4-stage weight ring that carries the 14 W1 blocks of a tile AND the 2 W2 blocks
of the previous tile's conv1 (slotted in after block RU_MMA2_AT), one hs tile, TWO residual tiles (epilogue 2 is what
bounds the kernel: the next tile's residual rows are fetched while it still works on the previous ones), and the
double-buffered D1 / D2 accumulators.  Same idea as tests/test_tail_protocol_model.py: coroutines that block on
mbarrier parity waits like the device code, a seeded random scheduler, and checks that (a) nothing deadlocks for any
number of tiles, (b) every consumer sees the tile it expects in the buffer it reads, (c) no buffer is overwritten
while in use.  The positions at which the producer fetches the next xs box / the residual rows and the MMA warp issues
conv1 are the kernel's constants; moving conv1's slot on ONE side only must break the model (the ring is consumed in
program order)."""
import random

import pytest

STAGES, KB, MMA2_AT, LOADX_AT, LOADA_AT = 4, 14, 11, 8, 2


class Barrier:
    def __init__(self):
        self.phase = 0

    def arrive(self):
        self.phase += 1

    def ready(self, parity):
        return (self.phase & 1) != parity


class ProtocolError(AssertionError):
    pass


def simulate(n_tiles, rng, mma2_at_mma=MMA2_AT, loada_at=LOADA_AT, max_steps=400000):
    full = [Barrier() for _ in range(STAGES)]
    empty = [Barrier() for _ in range(STAGES)]
    af, ae = Barrier(), Barrier()
    d1f, d1e = [Barrier(), Barrier()], [Barrier(), Barrier()]
    d2f, d2e = [Barrier(), Barrier()], [Barrier(), Barrier()]
    hsf, hse = Barrier(), Barrier()
    xf, xe = [Barrier(), Barrier()], [Barrier(), Barrier()]
    ring = [None] * STAGES            # ("w1", tile, kb) / ("w2", tile, half)
    xs_box = [None]
    xs_busy = [False]
    hs_tile, x_tile = [None], [None, None]
    d1, d2 = [None, None], [None, None]

    def producer():
        st = {"stage": 0, "phase": 0}

        def load_w(tag):
            yield ("wait", empty[st["stage"]], st["phase"] ^ 1)
            ring[st["stage"]] = tag
            full[st["stage"]].arrive()
            st["stage"] += 1
            if st["stage"] == STAGES:
                st["stage"], st["phase"] = 0, st["phase"] ^ 1

        def load_a(j):
            yield ("wait", ae, (j & 1) ^ 1)
            if xs_busy[0]:
                raise ProtocolError(f"xs box of tile {xs_box[0]} overwritten by tile {j} while conv7 reads it")
            xs_box[0], xs_busy[0] = j, True
            af.arrive()

        def load_x(j):
            yield ("wait", xe[j & 1], ((j >> 1) & 1) ^ 1)
            x_tile[j & 1] = j
            xf[j & 1].arrive()

        for it in range(n_tiles):
            for kb in range(KB):
                yield from load_w(("w1", it, kb))
                if kb == loada_at:
                    yield from load_a(it)
                if kb == LOADX_AT and it > 0:
                    yield from load_x(it - 1)
                if kb == MMA2_AT and it > 0:
                    yield from load_w(("w2", it - 1, 0))
                    yield from load_w(("w2", it - 1, 1))
                yield ("step",)
        if n_tiles > 0:
            yield from load_w(("w2", n_tiles - 1, 0))
            yield from load_w(("w2", n_tiles - 1, 1))
            yield from load_x(n_tiles - 1)

    def mma():
        st = {"stage": 0, "phase": 0}

        def take(expect):
            yield ("wait", full[st["stage"]], st["phase"])
            if ring[st["stage"]] != expect:
                raise ProtocolError(f"MMA expects {expect}, ring holds {ring[st['stage']]}")
            empty[st["stage"]].arrive()   # tcgen05.commit (instantaneous in the model)
            st["stage"] += 1
            if st["stage"] == STAGES:
                st["stage"], st["phase"] = 0, st["phase"] ^ 1

        def mma2(j):
            b = j & 1
            yield ("wait", hsf, j & 1)
            yield ("wait", d2e[b], ((j >> 1) & 1) ^ 1)
            if hs_tile[0] != j:
                raise ProtocolError(f"conv1 of tile {j} reads hs of tile {hs_tile[0]}")
            for half in range(2):
                yield from take(("w2", j, half))
            d2[b] = j
            hse.arrive()
            d2f[b].arrive()

        for it in range(n_tiles):
            b = it & 1
            yield ("wait", d1e[b], ((it >> 1) & 1) ^ 1)
            yield ("wait", af, it & 1)
            if xs_box[0] != it:
                raise ProtocolError(f"conv7 of tile {it} reads the xs box of tile {xs_box[0]}")
            for kb in range(KB):
                yield from take(("w1", it, kb))
                if kb == KB - 1:
                    d1[b] = it
                    xs_busy[0] = False
                    d1f[b].arrive()
                    ae.arrive()
                if kb == mma2_at_mma and it > 0:
                    yield from mma2(it - 1)
                yield ("step",)
        if n_tiles > 0:
            yield from mma2(n_tiles - 1)

    def epilogue1():
        for it in range(n_tiles):
            b = it & 1
            yield ("wait", d1f[b], (it >> 1) & 1)
            yield ("wait", hse, (it & 1) ^ 1)
            if d1[b] != it:
                raise ProtocolError(f"epilogue 1 of tile {it} reads D1 of tile {d1[b]}")
            yield ("step",)
            hs_tile[0] = it
            d1e[b].arrive()
            hsf.arrive()

    def epilogue2():
        for it in range(n_tiles):
            b = it & 1
            yield ("wait", d2f[b], (it >> 1) & 1)
            yield ("wait", xf[b], (it >> 1) & 1)
            if d2[b] != it or x_tile[b] != it:
                raise ProtocolError(f"epilogue 2 of tile {it}: D2 of {d2[b]}, residual rows of {x_tile[b]}")
            yield ("step",)
            d2e[b].arrive()
            yield ("step",)   # the two TMA stores read the slab (cp.async.bulk.wait_group.read) before it is handed back
            if x_tile[b] != it:
                raise ProtocolError(f"residual tile of {it} overwritten by tile {x_tile[b]} during the output copies")
            xe[b].arrive()

    agents = {"producer": producer(), "mma": mma(), "epi1": epilogue1(), "epi2": epilogue2()}
    pending = {}
    for k in list(agents):
        try:
            pending[k] = next(agents[k])
        except StopIteration:
            del agents[k]
    for _ in range(max_steps):
        if not agents:
            return
        runnable = [k for k, r in pending.items() if r[0] == "step" or r[1].ready(r[2])]
        if not runnable:
            raise ProtocolError(f"deadlock with {n_tiles} tiles: {sorted(pending)}")
        k = rng.choice(runnable)
        try:
            pending[k] = next(agents[k])
        except StopIteration:
            del agents[k], pending[k]
    raise ProtocolError("no progress bound hit")


def test_residual_unit_protocol_survives_random_schedules():
    rng = random.Random(11)
    for n_tiles in list(range(0, 9)) + [17, 40, 153]:
        for _ in range(12 if n_tiles < 20 else 3):
            simulate(n_tiles, rng)


def test_the_xs_box_must_be_requested_within_the_ring_depth():
    """The one xs box of tile it is requested by the producer after weight block RU_LOADA_AT of that tile; the MMA
    warp cannot consume any of those blocks before the box has landed, so a position at or beyond the ring depth
    deadlocks (the kernel has a static_assert for it)."""
    rng = random.Random(2)
    with pytest.raises(ProtocolError):
        simulate(3, rng, loada_at=STAGES)
    for _ in range(10):
        simulate(3, rng, loada_at=STAGES - 1)


def test_ring_is_consumed_in_program_order():
    """conv1's weight blocks sit in the ring right after W1 block RU_MMA2_AT: issuing conv1 elsewhere on the MMA side
    only must be caught (wrong block, or a deadlock), otherwise the model would not be checking the order at all."""
    rng = random.Random(5)
    with pytest.raises(ProtocolError):
        for _ in range(20):
            simulate(5, rng, mma2_at_mma=MMA2_AT - 2)
