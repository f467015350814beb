"""Executable models of three device-side protocols, run under random interleavings (no GPU):

1. the persistent tcgen05 GEMM pipeline (gemm_tcgen05_sm100a.cu / gemm_tcgen05_2cta_sm100a.cu): TMA producer, MMA
   issuer and requant epilogue with RUNNING stage / phase counters across tiles, mbarrier parity semantics,
   asynchronous TMA completion and tcgen05.commit arrival.  Checked: no deadlock, every shared-memory stage is written
   only when free and read only when full, a TMEM accumulator stage is never overwritten before the epilogue has read
   it, every tile's slices are consumed in order.
2. the two-level arrival / re-arming protocol of the BatchNorm statistics kernels (bn_act_sm100.cu::arrive_and_reduce):
   exactly one CTA per launch finalises, it sees every group partial, all counters return to zero for the next launch.
3. the cross-GPU epoch-flag protocol of the fused all-reduce (allreduce_fused_sm100a.cu): flags that are never cleared,
   parity-indexed buffers, no trailing barrier, variants interleaving from launch to launch, ranks running at their own
   pace - no read ever sees stale or already-overwritten data (and with a single buffer it does: the model is not vacuous).

The models restate the kernels' index arithmetic (it % NUM_STAGES, (it / NUM_STAGES) & 1, sl_run % NUM_ACC, ...), so an
off-by-one in a phase bit or a wait on the wrong parity shows up here as a deadlock or an invariant violation."""
import random

import pytest


class MBar:
    """mbarrier: `count` arrivals complete the current phase; wait(parity) passes once the phase with that parity is complete."""
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, "more arrivals than the barrier expects in one phase"
        if self.pending == 0:
            self.pending, self.phase = self.count, self.phase ^ 1

    def passed(self, parity):
        return self.phase != parity


def run_gemm_pipeline(tiles, kblocks, kb_per_slice, stages, accs, epi_warps, rng):
    full, empty = [MBar(1) for _ in range(stages)], [MBar(1) for _ in range(stages)]
    tfull, tempty = [MBar(1) for _ in range(accs)], [MBar(epi_warps) for _ in range(accs)]
    smem = ["free"] * stages                 # free -> loading -> full -> (mma reads) -> free
    tmem = ["free"] * accs                   # free -> acc -> full -> free
    tmem_tag = [None] * accs
    inflight_tma, inflight_commit = [], []   # asynchronous completions, delivered later in issue order
    slices = (kblocks + kb_per_slice - 1) // kb_per_slice
    consumed = []

    def producer():
        it = 0
        for t in range(tiles):
            for kb in range(kblocks):
                s, ph = it % stages, (it // stages) & 1
                while not empty[s].passed(ph ^ 1):
                    yield
                assert smem[s] == "free", f"TMA into stage {s} that is {smem[s]}"
                smem[s] = "loading"
                inflight_tma.append(s)
                it += 1
                yield

    def mma():
        it = sl_run = 0
        for t in range(tiles):
            kb = 0
            for sl in range(slices):
                a, aph = sl_run % accs, (sl_run // accs) & 1
                while not tempty[a].passed(aph ^ 1):
                    yield
                assert tmem[a] == "free", f"MMA into TMEM stage {a} that is {tmem[a]}"
                tmem[a], tmem_tag[a] = "acc", (t, sl)
                kb_end = min(kblocks, kb + kb_per_slice)
                while kb < kb_end:
                    s, ph = it % stages, (it // stages) & 1
                    while not full[s].passed(ph):
                        yield
                    assert smem[s] == "full", f"MMA reads stage {s} that is {smem[s]}"
                    inflight_commit.append(("smem", s))          # tcgen05.commit -> empty[s] once the MMAs retire
                    kb += 1
                    it += 1
                    yield
                inflight_commit.append(("tmem", a))
                sl_run += 1
                yield

    def epilogue(w):
        sl_run = 0
        for t in range(tiles):
            for sl in range(slices):
                a, aph = sl_run % accs, (sl_run // accs) & 1
                while not tfull[a].passed(aph):
                    yield
                assert tmem[a] == "full" and tmem_tag[a] == (t, sl), f"warp {w} reads TMEM {a}: {tmem[a]} {tmem_tag[a]} != {(t, sl)}"
                if w == 0:
                    consumed.append((t, sl))
                yield                                            # tcgen05.ld ... then hand the stage back
                tempty[a].arrive()
                if tempty[a].pending == tempty[a].count:         # last warp's arrival completed the phase
                    tmem[a] = "free"
                sl_run += 1
                yield

    actors = [producer(), mma()] + [epilogue(w) for w in range(epi_warps)]
    alive = list(range(len(actors)))
    for _ in range(200000):
        choices = list(alive)
        if inflight_tma:
            choices.append("tma")
        if inflight_commit:
            choices.append("commit")
        if not choices:
            break
        c = rng.choice(choices)
        if c == "tma":
            s = inflight_tma.pop(0)
            smem[s] = "full"
            full[s].arrive()
        elif c == "commit":
            kind, i = inflight_commit.pop(0)
            if kind == "smem":
                smem[i] = "free"
                empty[i].arrive()
            else:
                tmem[i] = "full"
                tfull[i].arrive()
        else:
            try:
                next(actors[c])
            except StopIteration:
                alive.remove(c)
    assert not alive and not inflight_tma and not inflight_commit, f"deadlock: actors {alive} still waiting"
    assert consumed == [(t, sl) for t in range(tiles) for sl in range(slices)]


@pytest.mark.parametrize("stages,accs", [(4, 2), (5, 2), (2, 2), (3, 1)])
def test_persistent_gemm_pipeline_never_deadlocks_or_overwrites(stages, accs):
    rng = random.Random(1234 + stages * 10 + accs)
    for tiles in (1, 2, 3, 5):
        for kblocks in (1, 2, 3, 9):
            for kb_per_slice in (1, 2, 4, 9):
                for _ in range(3):
                    run_gemm_pipeline(tiles, kblocks, kb_per_slice, stages, accs, epi_warps=4, rng=rng)


def test_gemm_pipeline_model_detects_a_wrong_parity():
    """The model is not vacuous: waiting on the wrong parity of the empty barrier deadlocks or overwrites a stage."""
    class Bad(MBar):
        def passed(self, parity):
            return self.phase == parity                          # inverted
    g = globals()
    good = g["MBar"]
    g["MBar"] = Bad
    try:
        with pytest.raises(AssertionError):
            run_gemm_pipeline(2, 3, 1, 4, 2, 4, random.Random(0))
    finally:
        g["MBar"] = good


def run_bn_arrival(slabs, group, rng, counters):
    """One launch of arrive_and_reduce for one channel tile; returns (finalisers, groups seen by the finaliser)."""
    groups = (slabs + group - 1) // group
    gpart_written = [False] * groups
    finalised = []

    def cta(y):
        grp = y // group
        gcount = min(group, slabs - grp * group)
        yield                                                    # part[y] written, fence
        done = counters[1 + grp]; counters[1 + grp] += 1         # atomicAdd (one step: atomic)
        last = done == gcount - 1
        if not last:
            return
        counters[1 + grp] = 0
        yield                                                    # reduce the group's partials
        if groups > 1:
            gpart_written[grp] = True
            yield                                                # fence
            done = counters[0]; counters[0] += 1
            if done != groups - 1:
                return
            counters[0] = 0
            yield
            finalised.append((y, list(gpart_written)))
        else:
            finalised.append((y, [True]))

    actors = [cta(y) for y in range(slabs)]
    alive = list(range(slabs))
    while alive:
        c = rng.choice(alive)
        try:
            next(actors[c])
        except StopIteration:
            alive.remove(c)
    return finalised


@pytest.mark.parametrize("slabs", [1, 2, 15, 16, 17, 37, 592])
def test_bn_two_level_arrival_has_one_finaliser_and_rearms(slabs):
    rng = random.Random(slabs)
    counters = [0] * (2 + (slabs + 15) // 16)
    for launch in range(3):                                      # the same counters, launch after launch (CUDA-graph replays)
        fin = run_bn_arrival(slabs, 16, rng, counters)
        assert len(fin) == 1, fin
        assert all(fin[0][1]), "the finaliser ran before every group partial was written"
        assert not any(counters), counters


def run_allreduce_epochs(W, variants, rng, double_buffered=True):
    """The cross-GPU flag protocol of allreduce_fused_sm100a.cu ("Why the protocol is safe without a trailing
    barrier or host resets"): monotonic epoch flags that are never cleared, buffers indexed by epoch parity, no barrier at
    the end of a launch.  Every rank runs the launches back to back at its own pace; every read asserts that it sees the
    value its producer wrote IN THIS LAUNCH (not a stale one, not one already overwritten by a rank that ran ahead)."""
    par = (lambda e: e & 1) if double_buffered else (lambda e: 0)
    flags = [[[0] * W for _ in range(W)] for _ in range(3)]           # flags[phase][dst][src] = epoch of src's last signal
    mailbox = [[None, None] for _ in range(W)]                        # mailbox[rank][parity]   (pulled by the peers)
    data = [[None, None] for _ in range(W)]                           # wire buffer of a rank   (pulled by the peers)
    gather = [[[None] * W for _ in range(2)] for _ in range(W)]       # gather[dst][parity][src] (pushed by the owners)

    def rank(me):
        peers = [p for p in range(W) if p != me]
        for e, v in enumerate(variants, start=1):
            p = par(e)
            phases = ([] if v == "raw" else [0]) + [1]
            for ph in phases:
                buf = mailbox if ph == 0 else data
                buf[me][p] = (me, e)
                yield
                for q in peers:                                      # st.release.sys flag[peer][phase][me] = epoch
                    flags[ph][q][me] = e
                    yield
                for q in peers:                                      # ld.acquire.sys until the peer's epoch arrived
                    while flags[ph][me][q] - e < 0:
                        yield
                for q in peers:
                    assert buf[q][p] == (q, e), f"rank {me} launch {e} phase {ph}: read {buf[q][p]} from rank {q}"
                    yield
            if v == "twoshot":
                for q in peers:                                      # push my reduced chunk into every peer's gather buffer
                    gather[q][p][me] = (me, e)
                    yield
                for q in peers:
                    flags[2][q][me] = e
                    yield
                for q in peers:
                    while flags[2][me][q] - e < 0:
                        yield
                for q in peers:
                    assert gather[me][p][q] == (q, e), f"rank {me} launch {e} gather: read {gather[me][p][q]} from rank {q}"
                    yield

    actors = [rank(r) for r in range(W)]
    alive = list(range(W))
    for _ in range(1000000):
        if not alive:
            return
        c = rng.choice(alive)
        try:
            next(actors[c])
        except StopIteration:
            alive.remove(c)
    raise AssertionError(f"deadlock: ranks {alive} never finished")


@pytest.mark.parametrize("W", [2, 3, 4, 8])
def test_allreduce_epoch_flags_never_show_stale_or_overwritten_data(W):
    rng = random.Random(W)
    for trial in range(60 if W <= 4 else 10):
        variants = [rng.choice(["raw", "oneshot", "twoshot"]) for _ in range(7)]       # variants interleave freely
        run_allreduce_epochs(W, variants, rng)


def test_allreduce_model_needs_the_double_buffering():
    """Not vacuous: with ONE buffer per rank a rank that runs ahead overwrites data a slow peer has not read yet."""
    rng = random.Random(7)
    with pytest.raises(AssertionError):
        for trial in range(200):
            run_allreduce_epochs(3, ["raw"] * 6, rng, double_buffered=False)
