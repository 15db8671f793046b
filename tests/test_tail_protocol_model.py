"""Executable model of the barrier protocol between the residual-loader warp and the two epilogue column groups of
the every-tile tail variant of the CTA-pair GEMM (csrc/gemm.cuh, ALLTAIL).

The kernel's warps are modelled as coroutines that block on mbarrier parity waits exactly like the device code
(`try_wait.parity(x)` succeeds iff the barrier's current phase has the other parity — which is also how a waiter that
is two phases behind gets stuck), and a seeded random scheduler explores interleavings: the tensor-core warp running
ahead, column group 1 skipping the tiles in which it has no box (the two-box last column tile of N = 2048 at 192-wide
tiles), the loader lagging.  The model is run on the SHIPPED rules and on the two earlier rule sets that were wrong:

  * `parity_from_tile_counter`: the epilogue waits for residual box i with parity `it & 1` — wrong as soon as a box
    is absent from some tile in the middle of a CTA's sequence (it reads a box before its TMA load has landed);
  * `arrive_every_tile`: a group hands its boxes back (arrives on resid_empty) in every tile, also in tiles where it
    had none, and the loader waits once per tile — the idle group can complete two phases before the loader looks at
    the first, and the kernel deadlocks (seen on B200 at M = 15000, n-fastest tile order).

No GPU: this is the host-side guard for a class of bug that only shows up under particular schedules on the device
(the device-side guard is tests/test_gpu_kernels.py::test_multiwave_residual_gemms_hand_back_protocol_under_a_forced_race).
"""
import random

import pytest


class Barrier:
    """count-1 mbarrier: every arrive completes a phase."""

    def __init__(self):
        self.phase = 0

    def arrive(self):
        self.phase += 1

    def ready(self, parity):  # try_wait.parity(parity)
        return (self.phase & 1) != parity


class ProtocolError(AssertionError):
    pass


def simulate(tiles, rules, rng, max_steps=200000):
    """tiles: list of box-existence tuples per tile, e.g. (True, True, True, False) for boxes 0..3 (box 2g belongs to
    group g together with box 2g + 1).  Returns None, raises ProtocolError on a deadlock or a wrong box."""
    nbox = len(tiles[0])
    resid_full = [Barrier() for _ in range(nbox)]
    resid_empty = [Barrier(), Barrier()]
    box_tile = [None] * nbox      # which tile's residual a box currently holds
    box_busy = [False] * nbox     # loaded and not yet handed back
    acc_done = [0]                # tiles whose accumulator the tensor-core warp has finished (runs ahead freely:
    released = [0, 0]             # at most 2 beyond what BOTH groups have released, as the double-buffered TMEM allows)

    def group_has(t, g):
        return 2 * g < nbox and tiles[t][2 * g]

    def loader():
        in_use, parity = [False, False], [0, 0]
        for t, boxes in enumerate(tiles):
            for g in (0, 1):
                if rules == "arrive_every_tile":
                    yield ("wait", resid_empty[g], (t & 1) ^ 1)
                else:
                    if not group_has(t, g):
                        continue
                    if in_use[g]:
                        yield ("wait", resid_empty[g], parity[g])
                        parity[g] ^= 1
                    in_use[g] = True
                for b in (2 * g, 2 * g + 1):
                    if b < nbox and boxes[b]:
                        if box_busy[b]:
                            raise ProtocolError(f"loader overwrites box {b} (tile {box_tile[b]}) with tile {t}")
                        box_tile[b], box_busy[b] = t, True
                        resid_full[b].arrive()   # (the TMA completion; instantaneous in the model)
                yield ("step",)

    def epilogue(g):
        box_parity = [0] * nbox
        for t, boxes in enumerate(tiles):
            yield ("acc", t)                      # tfull of tile t
            for b in (2 * g, 2 * g + 1):
                if b < nbox and boxes[b]:
                    if rules == "parity_from_tile_counter":
                        yield ("wait", resid_full[b], t & 1)
                    else:
                        yield ("wait", resid_full[b], box_parity[b])
                        box_parity[b] ^= 1
                    if box_tile[b] != t:
                        raise ProtocolError(f"group {g} reads box {b} holding tile {box_tile[b]} as tile {t}")
                    yield ("step",)               # h pass, stores, g pass ...
            released[g] = t + 1                   # tempty arrive
            had = group_has(t, g)
            if had:
                for b in (2 * g, 2 * g + 1):
                    if b < nbox and boxes[b]:
                        box_busy[b] = False
            if had or rules == "arrive_every_tile":
                resid_empty[g].arrive()
            yield ("step",)

    def tensor_core():
        for t in range(len(tiles)):
            yield ("tmem", t)
            acc_done[0] = t + 1
            yield ("step",)

    agents = {"loader": loader(), "epi0": epilogue(0), "epi1": epilogue(1), "mma": tensor_core()}
    pending = {k: next(v) for k, v in agents.items()}
    for _ in range(max_steps):
        if not agents:
            return
        runnable = []
        for k, req in pending.items():
            if req[0] == "step":
                runnable.append(k)
            elif req[0] == "wait" and req[1].ready(req[2]):
                runnable.append(k)
            elif req[0] == "acc" and acc_done[0] > req[1]:
                runnable.append(k)
            elif req[0] == "tmem" and req[1] < min(released) + 2:
                runnable.append(k)
        if not runnable:
            raise ProtocolError(f"deadlock: {sorted((k, r[0]) for k, r in pending.items())}")
        k = rng.choice(runnable)
        try:
            pending[k] = next(agents[k])
        except StopIteration:
            del agents[k], pending[k]
    raise ProtocolError("no progress bound hit")


def tile_sequences(rng, n_cases):
    """CTA-pair tile sequences as the kernel sees them: 192-wide tiles (3 boxes) or 256-wide (4), N = 2048 or ragged,
    m-fastest (the partial column tile only at the end) or n-fastest (anywhere)."""
    for _ in range(n_cases):
        nbox = rng.choice((3, 4))
        n_tiles = rng.randint(1, 12)
        full = (True,) * nbox
        partials = [tuple(i < k for i in range(nbox)) for k in range(1, nbox)]
        seq = []
        for _t in range(n_tiles):
            seq.append(rng.choice(partials) if rng.random() < 0.3 else full)
        yield seq


def test_shipped_rules_survive_random_schedules():
    rng = random.Random(20260117)
    for seq in tile_sequences(rng, 300):
        for _ in range(6):
            simulate(seq, "shipped", rng)


@pytest.mark.parametrize("rules", ["parity_from_tile_counter", "arrive_every_tile"])
def test_the_two_earlier_rule_sets_fail_in_the_model(rules):
    """The model is only worth something if it sees the bugs the device showed."""
    rng = random.Random(7)
    failures = 0
    for seq in tile_sequences(rng, 300):
        for _ in range(6):
            try:
                simulate(seq, rules, rng)
            except ProtocolError:
                failures += 1
    assert failures > 0


def test_the_device_case_in_the_model():
    """N = 2048 at 192-wide tiles, n-fastest order: a CTA pair's sequence with the two-box column tile in the middle."""
    full, last = (True, True, True), (True, True, False)
    seq = [full, full, last, full, full, full, last, full]
    rng = random.Random(3)
    for _ in range(200):
        simulate(seq, "shipped", rng)
    with pytest.raises(ProtocolError):
        for _ in range(2000):
            simulate(seq, "arrive_every_tile", rng)


# ---------------------------------------------------------------------------------------------------------------------
# Single-wave tail path: the residual boxes of a CTA pair's LAST tile ride the operand ring as extra "k-blocks" that
# the tensor-core warp never consumes (gemm.cuh, the `Epi::kTmaTail && !ALLTAIL` branch).
# ---------------------------------------------------------------------------------------------------------------------
def simulate_ring_ride(my_tiles, total_kb, nbox, stages, rng, max_steps=200000):
    full = [Barrier() for _ in range(stages)]
    empty = [Barrier() for _ in range(stages)]
    resid = [Barrier() for _ in range(nbox)]
    slot = [None] * stages            # ("op", tile, kb) / ("box", i)
    last_mma_done = [False]

    def producer():
        stage, phase = 0, 0
        for t in range(my_tiles):
            for kb in range(total_kb):
                yield ("wait", empty[stage], phase ^ 1)
                slot[stage] = ("op", t, kb)
                full[stage].arrive()
                stage += 1
                if stage == stages:
                    stage, phase = 0, phase ^ 1
                yield ("step",)
        for i in range(nbox):
            yield ("wait", empty[stage], phase ^ 1)
            if slot[stage] is not None and slot[stage][0] == "box":
                raise ProtocolError("a residual box overwrites another one")
            slot[stage] = ("box", i)
            resid[i].arrive()
            stage += 1
            if stage == stages:
                stage, phase = 0, phase ^ 1
            yield ("step",)

    def mma():
        stage, phase = 0, 0
        for t in range(my_tiles):
            for kb in range(total_kb):
                yield ("wait", full[stage], phase)
                if slot[stage] != ("op", t, kb):
                    raise ProtocolError(f"MMA reads {slot[stage]} as operand block ({t}, {kb})")
                yield ("step",)                 # the MMAs run ...
                empty[stage].arrive()           # ... and their commit frees the slot
                stage += 1
                if stage == stages:
                    stage, phase = 0, phase ^ 1
        last_mma_done[0] = True

    def epilogue():
        ring0 = (my_tiles * total_kb) % stages
        while not last_mma_done[0]:
            yield ("step",)
        for i in range(nbox):
            yield ("wait", resid[i], 0)
            if slot[(ring0 + i) % stages] != ("box", i):
                raise ProtocolError(f"epilogue finds {slot[(ring0 + i) % stages]} where box {i} should be")
            yield ("step",)

    agents = {"producer": producer(), "mma": mma(), "epi": epilogue()}
    pending = {k: next(v) for k, v in agents.items()}
    for _ in range(max_steps):
        if not agents:
            return
        runnable = [k for k, r in pending.items() if r[0] == "step" or r[1].ready(r[2])]
        if not runnable:
            raise ProtocolError(f"deadlock: {sorted(pending)}")
        k = rng.choice(runnable)
        try:
            pending[k] = next(agents[k])
        except StopIteration:
            del agents[k], pending[k]
    raise ProtocolError("no progress bound hit")


def test_residual_boxes_riding_the_operand_ring():
    """Every (tiles per CTA pair, k-blocks per tile, boxes) combination the DiT produces and then some: the boxes land
    in the slots the epilogue computes from `ring0`, behind MMAs that have retired, for any schedule."""
    rng = random.Random(99)
    for my_tiles in (1, 2, 3):
        for total_kb in (1, 2, 5, 6, 7, 32, 96):
            for nbox in (1, 2, 3, 4):
                for _ in range(4):
                    simulate_ring_ride(my_tiles, total_kb, nbox, 6, rng)
