"""CPU-only: discrete simulation of the barrier protocol of the EXPERIMENTAL persistent GEMM (csrc/kernels/gemm_tc_persist.cu), which has
not run on hardware yet.  It restates the kernel's phase arithmetic -- ring slots from a running k-block counter, two accumulators with
acc_full (one tcgen05.commit arrival) / acc_empty (four epilogue-warp arrivals) -- and checks under random interleavings of the TMA
producer, the MMA issuer and the four epilogue warps that nothing deadlocks, no ring slot or accumulator is overwritten while in use, and
every tile is drained by all four warps.  It guards the parity reasoning, not the CUDA code itself."""
import itertools
import random

import pytest


class MBar:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        if self.pending == 0:
            self.phase ^= 1
            self.pending = self.count

    def test(self, parity):          # mbarrier.try_wait.parity: true once the phase with this parity has completed
        return self.phase != parity


def simulate(tiles, nkb, stages, seed):
    rnd = random.Random(seed)
    full = [MBar(1) for _ in range(stages)]
    empty = [MBar(1) for _ in range(stages)]
    acc_full, acc_empty = [MBar(1), MBar(1)], [MBar(4), MBar(4)]
    smem = [None] * stages
    tmem = [None, None]
    drained = {}
    loads = [{}, {}]

    def producer():
        it = 0
        for t in range(tiles):
            for kb in range(nkb):
                s, ph = it % stages, (it // stages) & 1
                while not empty[s].test(ph ^ 1):
                    yield
                assert smem[s] is None, "ring slot overwritten while in use"
                smem[s] = (t, kb)
                full[s].arrive()
                it += 1
                yield

    def mma():
        it = 0
        for j, t in enumerate(range(tiles)):
            buf, aph = j & 1, (j >> 1) & 1
            while not acc_empty[buf].test(aph ^ 1):
                yield
            assert tmem[buf] is None, "accumulator overwritten before the epilogue drained it"
            tmem[buf] = dict(tile=t, kbs=[])
            for kb in range(nkb):
                s, ph = it % stages, (it // stages) & 1
                while not full[s].test(ph):
                    yield
                assert smem[s] == (t, kb)
                tmem[buf]["kbs"].append(kb)
                smem[s] = None
                empty[s].arrive()
                if kb == nkb - 1:
                    acc_full[buf].arrive()
                it += 1
                yield

    def epilogue(w):
        for j, t in enumerate(range(tiles)):
            buf, aph = j & 1, (j >> 1) & 1
            while not acc_full[buf].test(aph):
                yield
            assert tmem[buf] is not None and tmem[buf]["tile"] == t and tmem[buf]["kbs"] == list(range(nkb))
            drained.setdefault(t, set()).add(w)
            yield
            loads[buf][j] = loads[buf].get(j, 0) + 1
            if loads[buf][j] == 4:
                tmem[buf] = None
            acc_empty[buf].arrive()
            yield

    procs = [producer(), mma()] + [epilogue(w) for w in range(4)]
    alive, steps = set(range(len(procs))), 0
    while alive:
        i = rnd.choice(sorted(alive))
        try:
            next(procs[i])
        except StopIteration:
            alive.discard(i)
        steps += 1
        assert steps < 2_000_000, "deadlock"
    assert all(drained.get(t) == {0, 1, 2, 3} for t in range(tiles))


@pytest.mark.parametrize("stages", [4, 6, 8])
def test_persistent_gemm_barrier_protocol(stages):
    for tiles, nkb in itertools.product((1, 2, 3, 5, 8), (1, 2, 4, 7, 20)):
        for seed in range(3):
            simulate(tiles, nkb, stages, seed)
