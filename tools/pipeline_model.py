"""Discrete-event model of the mbarrier protocol of the CTA-pair conv kernel (csrc/conv_tc2.cu).

Not a kernel and not on any product path.  The kernel was written without GPU time (round 1) and ran correctly on its
first execution (round 2, ten kernel-level parity tests); the model is kept as the executable description of its
protocol.  A protocol mistake (a wrong
parity, a barrier count, a ring that is refilled while the tensor core still reads it) shows up on hardware as a hang
- which costs a GPU strike - or as silent corruption.  This model replays the role loops of the kernels (producer, MMA
issuer, epilogue warps; both CTAs of a pair) line by line against mbarrier semantics, with asynchronous TMA loads and
MMA commits completing after random delays and a randomised scheduler, and checks

  * liveness: every role terminates (no deadlock) under many random schedules;
  * safety: a shared-memory stage / TMEM accumulator is never overwritten while an earlier consumer still reads it,
    and every consumer sees exactly the (work item, chunk, tap) it expects.

tests/test_pipeline_model.py runs it on CPU.  mbarrier semantics used: a barrier holds a phase bit, a pending-arrival
count and a transaction-byte count; the current phase completes when both reach zero; wait(parity) returns once the
phase with that parity has completed (a fresh barrier passes wait(1) immediately).
"""
import random


class MBar:
    def __init__(self, count, name=""):
        self.count, self.pending, self.tx, self.phase, self.name = count, count, 0, 0, name

    def _check(self):
        if self.pending == 0 and self.tx == 0:
            self.phase ^= 1
            self.pending = self.count
        assert self.pending >= 0, f"{self.name}: more arrivals than the barrier expects in one phase"

    def arrive(self):
        self.pending -= 1
        self._check()

    def arrive_expect_tx(self, nbytes):
        self.tx += nbytes
        self.pending -= 1
        self._check()

    def complete_tx(self, nbytes):
        self.tx -= nbytes
        self._check()

    def done(self, parity):
        return self.phase != parity


class Sim:
    """Cooperative scheduler: roles are generators that yield a zero-argument predicate to wait on (or None to yield
    the processor); asynchronous completions are (time, callback) events."""

    def __init__(self, seed):
        self.rng = random.Random(seed)
        self.roles, self.events, self.now, self.seq = [], [], 0, 0

    def at(self, t, prio, fn):
        self.seq += 1
        self.events.append((t, prio, self.seq, fn))

    def spawn(self, name, gen):
        self.roles.append([name, gen, None])

    def after(self, lo, hi, fn):
        self.at(self.now + self.rng.randint(lo, hi), self.rng.random(), fn)

    def run(self, max_steps=2_000_000):
        for _ in range(max_steps):
            self.now += 1
            due = [e for e in self.events if e[0] <= self.now]
            if due:
                self.events = [e for e in self.events if e[0] > self.now]
                for e in sorted(due, key=lambda e: e[:3]):
                    e[3]()
            ready = [r for r in self.roles if r[2] is None or r[2]()]
            if not ready:
                if not self.roles:
                    return
                if not self.events:
                    raise AssertionError("deadlock: " + ", ".join(r[0] for r in self.roles))
                continue
            r = self.rng.choice(ready)
            try:
                r[2] = next(r[1])
            except StopIteration:
                self.roles.remove(r)
        raise AssertionError("did not finish")


class InOrderPipe:
    """The tensor pipe: MMAs complete in issue order; a commit fires its barrier arrivals when everything issued before it
    has completed."""

    def __init__(self, sim):
        self.sim, self.t_free = sim, 0

    def mma(self, reads, check):
        """reads: list of buffers (dicts with 'readers'); check(): called at execution time to validate contents."""
        start = max(self.sim.now, self.t_free)
        dur = self.sim.rng.randint(1, 4)
        self.t_free = start + dur
        for b in reads:
            b["readers"] += 1

        def fin():
            check()
            for b in reads:
                b["readers"] -= 1
        self.sim.at(self.t_free, 1.0, fin)                           # MMAs retire in issue order (seq breaks ties)

    def commit(self, fns):
        t = max(self.sim.now, self.t_free)
        self.sim.at(t, 2.0, lambda: [f() for f in fns])              # after the MMA completions of the same tick


def _buf():
    return {"tag": None, "readers": 0}


def _tma_write(sim, buf, tag, bar, nbytes):
    def land():
        assert buf["readers"] == 0, f"TMA overwrote a stage the tensor core still reads (new {tag}, old {buf['tag']})"
        buf["tag"] = tag
        bar.complete_tx(nbytes)
    sim.after(3, 40, land)


# ----------------------------------------------------------------------------------------------------------------------
# csrc/conv_tc2.cu: CTA pair.  Both CTAs run a producer (stage ring of STAGES) whose loads complete on the LEADER's full
# barrier (count 2: leader's arrive.expect_tx for both CTAs' bytes + the peer's plain arrive); the leader's commits are
# multicast to both CTAs' empty / acc_full barriers; all 8 epilogue warps arrive on the leader's acc_empty.
# ----------------------------------------------------------------------------------------------------------------------
def run_conv_pair(tiles, iters_per_tile, seed, stages=4):
    sim = Sim(seed)
    pipe = InOrderPipe(sim)
    full = [MBar(2, f"full{s}") for s in range(stages)]                               # leader's
    empty = [[MBar(1, f"empty{r}.{s}") for s in range(stages)] for r in range(2)]     # per CTA
    acc_full = [[MBar(1, f"acc_full{r}.{s}") for s in range(2)] for r in range(2)]
    acc_empty = [MBar(8, f"acc_empty{s}") for s in range(2)]                          # leader's
    bufs = [[_buf() for _ in range(stages)] for _ in range(2)]
    acc = [[{"tag": None, "readers": 0, "writes": 0} for _ in range(2)] for _ in range(2)]
    BYTES = 500
    drained = []

    def producer(rank):
        stage, phase = 0, 0
        for tile in range(tiles):
            for it in range(iters_per_tile):
                e = empty[rank][stage]
                while not e.done(phase ^ 1):
                    yield lambda e=e, phase=phase: e.done(phase ^ 1)
                if rank == 0:
                    full[stage].arrive_expect_tx(2 * BYTES)
                _tma_write(sim, bufs[rank][stage], (tile, it), full[stage], BYTES)
                if rank == 1:
                    full[stage].arrive()
                stage += 1
                if stage == stages:
                    stage, phase = 0, phase ^ 1
                yield None

    def mma():
        stage, phase, as_, aphase = 0, 0, 0, 0
        for tile in range(tiles):
            while not acc_empty[as_].done(aphase ^ 1):
                yield lambda as_=as_, aphase=aphase: acc_empty[as_].done(aphase ^ 1)
            for it in range(iters_per_tile):
                while not full[stage].done(phase):
                    yield lambda stage=stage, phase=phase: full[stage].done(phase)
                b0, b1, want = bufs[0][stage], bufs[1][stage], (tile, it)
                a0, a1 = acc[0][as_], acc[1][as_]

                def check(b0=b0, b1=b1, want=want, a0=a0, a1=a1, tile=tile):
                    assert b0["tag"] == want and b1["tag"] == want, f"pair MMA read {b0['tag']} / {b1['tag']}, wanted {want}"
                    for a in (a0, a1):
                        assert a["readers"] == 0, "pair MMA wrote an accumulator an epilogue still reads"
                        a["tag"] = tile
                        a["writes"] += 1
                pipe.mma([b0, b1], check)
                pipe.commit([empty[0][stage].arrive, empty[1][stage].arrive])
                stage += 1
                if stage == stages:
                    stage, phase = 0, phase ^ 1
                yield None
            pipe.commit([acc_full[0][as_].arrive, acc_full[1][as_].arrive])
            as_ += 1
            if as_ == 2:
                as_, aphase = 0, aphase ^ 1

    def epilogue(rank, q):
        as_, aphase = 0, 0
        for tile in range(tiles):
            f = acc_full[rank][as_]
            while not f.done(aphase):
                yield lambda f=f, aphase=aphase: f.done(aphase)
            a = acc[rank][as_]
            a["readers"] += 1
            assert a["tag"] == tile and a["writes"] == iters_per_tile * (tile // 2 + 1)
            for _ in range(sim.rng.randint(0, 6)):
                yield None
            a["readers"] -= 1
            drained.append((rank, q, tile))
            acc_empty[as_].arrive()
            as_ += 1
            if as_ == 2:
                as_, aphase = 0, aphase ^ 1

    for r in range(2):
        sim.spawn(f"producer{r}", producer(r))
        for q in range(4):
            sim.spawn(f"epilogue{r}.{q}", epilogue(r, q))
    sim.spawn("mma", mma())
    sim.run()
    assert len(drained) == 8 * tiles
    return True


if __name__ == "__main__":
    for seed in range(20):
        run_conv_pair(5, 18, seed)
    print("ok")
