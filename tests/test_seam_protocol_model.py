"""CPU model check of the warp-seam protocol of csrc/multi_reg3.cuh (protocols 3 and 4).

`compute-sanitizer --tool racecheck` is closed on this pool, so the protocol the register-resident LangmuirMulti
kernel synchronises its warps with is restated here as a small executable model and run under thousands of random
warp interleavings (the GPU counterpart is tests/test_gpu_race_stress.py).  The model follows the kernel statement
by statement where synchronisation is concerned:

  * input i of global step g lives in ring slot (g & 1) * 4 + i and is use g >> 1 of that slot's mbarrier (count 1);
    a producer's arrival completes one phase; a consumer waits with `try_wait.parity`, which can only tell the
    current phase from the one before it - so the model asserts, at EVERY wait, that the barrier has completed
    exactly k or k + 1 phases when phase k is awaited (k + 2 would hang the kernel, k - 1 would let it read early);
  * protocol 3 keeps one CTA barrier per step; protocol 4 replaces it by the split step barrier (every warp arrives
    at the top of a step, count = warps, and waits for that phase before the step's last post) and polls the redo
    flag with a CTA barrier every `poll` steps;
  * a run may be abandoned at a poll (the column failed the no-row-exchange bound) and is then repeated by the
    guarded run; the barriers are never re-initialised: the global step counter continues and `settle` tops every
    barrier up to the phases of all steps below the next even step number.

Checked: no deadlock, no wait outside its two-phase window, every consumer reads exactly the value its upstream warp
posted for that (step, input), and no ring slot is overwritten before it was read.
"""
import random

import pytest


class Hang(Exception):
    pass


class Model:
    def __init__(self, W, k_in, sync, poll, rng, settle=True, step_wait=True):
        self.W, self.k_in, self.sync, self.poll, self.rng, self.settle = W, k_in, sync, poll, rng, settle
        self.step_wait = step_wait
        self.full = [[0] * W for _ in range(8)]      # completed phases of full_bar(slot, seam)
        self.step_arrivals = 0                        # arrivals on the split step barrier (count = W)
        self.ring = [[None] * W for _ in range(8)]   # (tag, consumed) per (slot, seam)
        self.cta_count, self.cta_gen, self.cta_or = 0, 0, False
        self.gbase = 0

    # ---- primitives; generators yield while blocked -------------------------------------------------
    def wait_phase(self, completed_fn, k, what):
        while True:
            c = completed_fn()
            assert c in (k, k + 1), f"{what}: waiting for phase {k} with {c} completed (parity wait is ambiguous)"
            if (c & 1) != (k & 1):  # try_wait.parity(k & 1): the phase of that parity is the one BEFORE the current
                return
            yield "blocked"

    def cta_barrier(self, flag=False):
        gen = self.cta_gen
        self.cta_or |= flag
        self.cta_count += 1
        if self.cta_count == self.W:
            self.cta_count = 0
            self.cta_result = self.cta_or
            self.cta_or = False
            self.cta_gen += 1
        while self.cta_gen == gen:
            yield "blocked"
        return self.cta_result

    # ---- the kernel's seam operations ----------------------------------------------------------------
    def post(self, w, slot, g, i, run_id):
        if w + 1 < self.W:
            old = self.ring[slot][w]
            assert old is None or old[1], f"warp {w} overwrites slot {slot} holding unread {old[0]}"
            self.ring[slot][w] = ((run_id, g, i), False)
            self.full[slot][w] += 1

    def consume(self, w, slot, g, i, run_id):
        if w >= 1:
            yield from self.wait_phase(lambda: self.full[slot][w - 1], g >> 1, f"warp {w} full[{slot}]")
            tag, _ = self.ring[slot][w - 1]
            assert tag == (run_id, g, i), f"warp {w} expected {(run_id, g, i)} in slot {slot}, found {tag}"
            self.ring[slot][w - 1] = (tag, True)

    def run(self, w, steps, raise_at, guard, run_id):
        """One integration of a column by warp w; returns True when the column has to be integrated again."""
        W, k_in, sync = self.W, self.k_in, self.sync
        yield from self.cta_barrier()
        yield from self.cta_barrier()
        G0 = self.gbase
        assert G0 % 2 == 0
        self.post(w, 0, G0, 0, run_id)
        sus, abandon, done = False, False, steps
        for step in range(steps):
            g = G0 + step
            spar = g & 1
            if sync == 3:
                res = yield from self.cta_barrier(sus and not guard)
                if not guard and res:
                    abandon, done = True, step
                    break
            else:
                if not guard and step % self.poll == self.poll - 1:
                    res = yield from self.cta_barrier(sus)
                    if res:
                        abandon, done = True, step
                        break
                self.step_arrivals += 1  # lane 0 of the warp arrives for phase g
            s0, s1 = spar * 4, (spar ^ 1) * 4
            for i in range(k_in):
                last = i == k_in - 1
                slot, slot_next, g_next, i_next = s0 + i, (s1 if last else s0 + i + 1), (g + 1 if last else g), (0 if last else i + 1)
                yield "cell"  # cell(1): a scheduling point
                if sync == 4 and last and self.step_wait:
                    yield from self.wait_phase(lambda: self.step_arrivals // W, g, f"warp {w} step barrier")
                self.post(w, slot_next, g_next, i_next, run_id)
                yield from self.consume(w, slot, g, i, run_id)
                yield "cell"  # cell(0)
            if raise_at is not None and step == raise_at and w == raise_at % W:
                sus = True  # one warp's cell failed the bound in this step
        again = abandon
        if not abandon:
            res = yield from self.cta_barrier(sus and not guard)
            again = (not guard) and res
        # settle (after a CTA barrier: nobody waits on a seam barrier)
        gend = G0 + done
        G1 = (gend + 2) & ~1
        if w + 1 < W and self.settle:
            for g2 in range(gend, G1):
                for i in range(1 if g2 == gend else 0, k_in):
                    self.full[(g2 & 1) * 4 + i][w] += 1
        if sync == 4 and self.settle:
            self.step_arrivals += G1 - gend
        if w == 0:
            self._next_gbase = G1
        return again

    def warp(self, w, columns):
        for ci, (steps, raise_at) in enumerate(columns):
            again = yield from self.run(w, steps, raise_at, False, (ci, 0))
            yield from self.cta_barrier()  # (the column loop's __syncthreads)
            if w == 0:
                self.gbase = self._next_gbase
                for slot in range(8):  # values nobody will read any more (the last post of a run)
                    for s in range(self.W):
                        if self.ring[slot][s] is not None:
                            self.ring[slot][s] = (self.ring[slot][s][0], True)
            yield from self.cta_barrier()
            assert again == (raise_at is not None), (ci, again, raise_at)
            if again:
                a2 = yield from self.run(w, steps, None, True, (ci, 1))
                assert not a2
                yield from self.cta_barrier()
                if w == 0:
                    self.gbase = self._next_gbase
                    for slot in range(8):
                        for s in range(self.W):
                            if self.ring[slot][s] is not None:
                                self.ring[slot][s] = (self.ring[slot][s][0], True)
                yield from self.cta_barrier()

    def execute(self, columns, max_ticks=2_000_000):
        warps = [self.warp(w, columns) for w in range(self.W)]
        alive = list(range(self.W))
        blocked_streak = 0
        w, burst = None, 0
        for _ in range(max_ticks):
            if not alive:
                return
            if burst == 0 or w not in alive:  # bursts let one warp race as far ahead as the protocol allows
                w = self.rng.choice(alive)
                burst = self.rng.choice([1, 1, 2, 5, 40, 400])
            burst -= 1
            try:
                state = next(warps[w])
            except StopIteration:
                alive.remove(w)
                blocked_streak = 0
                continue
            if state == "blocked":
                burst = 0
                blocked_streak += 1
            else:
                blocked_streak = 0
            if blocked_streak > 200 * self.W:  # every warp has been polled many times without progress
                raise Hang(f"no warp can proceed: full={self.full} step_arrivals={self.step_arrivals}")
        raise Hang("tick limit")


@pytest.mark.parametrize("sync", [3, 4])
@pytest.mark.parametrize("k_in", [4, 1], ids=["rk4", "euler"])
def test_seam_protocol_under_random_interleavings(sync, k_in):
    for seed in range(500):
        rng = random.Random(1000 * sync + 10 * k_in + seed)
        W = rng.choice([1, 2, 3, 8])
        poll = rng.choice([2, 3, 8])
        columns = []
        for _ in range(rng.randint(1, 4)):  # columns of one persistent CTA, some of them abandoned and repeated
            steps = rng.randint(1, 21)
            raise_at = rng.randrange(steps) if rng.random() < 0.5 else None
            columns.append((steps, raise_at))
        Model(W, k_in, sync, poll, rng).execute(columns)


def test_the_model_catches_a_missing_settle():
    """Sanity of the checker itself: without the top-up at the end of a run the next run's parity waits are off
    (a wait outside its two-phase window, or a warp that never wakes up)."""
    caught = 0
    for seed in range(20):
        try:
            Model(4, 4, 4, 3, random.Random(seed), settle=False).execute([(5, None), (4, None), (3, None)])
        except (AssertionError, Hang):
            caught += 1
    assert caught == 20


def test_the_model_catches_an_unbounded_skew():
    """Protocol 4 without the wait on the step barrier lets a warp run ahead of its downstream neighbour by more than
    the ring holds: some interleaving must overwrite an unread slot or push a barrier two phases ahead."""
    caught = 0
    for seed in range(60):
        try:
            Model(4, 4, 4, 64, random.Random(seed), step_wait=False).execute([(12, None)])
        except (AssertionError, Hang):
            caught += 1
    assert caught > 0
