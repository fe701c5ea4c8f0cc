"""Executable model of the cross-rank protocol of the fused USP kernels (``csrc/usp_comm.cuh`` +
``parallel/fused_engine.py``): ready-to-receive handshake, arrival counters, owner completion counters, all of them
monotonic "epochs" that are never reset; three kernels per training step (forward, backward dQ pass with the pushes,
backward dK/dV pass without a comm role) on every rank of a U x R mesh, several steps back to back.

Each rank runs its kernels in stream order; inside a kernel a push thread (the comm CTAs) and a compute thread (the
attention CTAs) run concurrently, and the ranks drift apart at random.  Staging buffers carry the epoch of their
contents: a reader must find exactly the current call's data (never a stale or a too-new version), i.e. the model checks
that no rank can overwrite a peer's staging while that peer may still read the previous contents, and that nothing
deadlocks.  Destination masks are modelled too (round 2): under a causal / window mask a source sends a destination the
arrival SIGNAL but no data -- and does not wait for its ready-to-receive flag -- while the consumer never visits those
segments; the set of skipped pairs changes from call to call (layers with different masks), so the monotonic counters
must stay aligned.  (The NVLS-style broadcast push of round 1 is gone from the kernels and from the model.)"""
import random
import threading
import time

import pytest

KV, Q, RTR, ODONE, DKV, QA = 0, 16, 32, 48, 49, 64


class Sig:
    """Signal pad of one rank: monotonic counters, waiters block until a slot reaches a target."""

    def __init__(self):
        self.v = [0] * 96
        self.cv = threading.Condition()

    def add(self, slot, n=1):
        with self.cv:
            self.v[slot] += n
            self.cv.notify_all()

    def store_max(self, slot, val):
        with self.cv:
            self.v[slot] = max(self.v[slot], val)
            self.cv.notify_all()

    TIMEOUT = 20.0

    def wait_ge(self, slot, target):
        with self.cv:
            assert self.cv.wait_for(lambda: self.v[slot] >= target, Sig.TIMEOUT), f"deadlock: slot {slot} < {target}"


class Mesh:
    def __init__(self, U, R, n_comm, seed, use_rtr=True, slow_reader=None, masked=False, signal_skipped=True):
        self.slow_reader = slow_reader              # (rank, seconds): that rank dwells on every staging read
        self.U, self.R, self.P, self.n_comm = U, R, U * R, n_comm
        self.use_rtr = use_rtr
        self.masked = masked                        # destination masks that differ from call to call
        self.signal_skipped = signal_skipped        # False = the bug the kernels must not have (negative control)
        self.rng = random.Random(seed)
        self.sig = [Sig() for _ in range(self.P)]
        self.lock = threading.Lock()
        # staging[dst][(cls, src)] = epoch of the contents;  readers[dst][(cls, src)] = epoch currently being read
        self.staging = [dict() for _ in range(self.P)]
        self.reading = [dict() for _ in range(self.P)]
        self.errors = []

    def jitter(self, scale=3e-4):
        time.sleep(self.rng.random() * scale)

    def check(self, cond, msg):
        if not cond:
            self.errors.append(msg)
            raise AssertionError(msg)

    def needs_data(self, epoch, src, dst) -> bool:
        """Does `dst` ever visit rows of `src` in call `epoch`?  Ranks of the same ring block always do (stationary
        operands); otherwise a call-dependent pseudo-random subset (every rank evaluates the same function, like the
        engines evaluate the same mask)."""
        if not self.masked or src // self.U == dst // self.U:
            return True
        return (epoch * 7 + src * 3 + dst) % 3 != 0

    # ------------------------------------------------------------------ one kernel on one rank
    def push(self, me, epoch, classes_all, classes_ring):
        """comm CTAs: classes_all go to every rank, classes_ring (Ulysses Q) to the ranks of my ring index."""
        u, r = me % self.U, me // self.U
        for t in range(self.P):
            self.sig[t].store_max(RTR + me, epoch)                    # my staging is free for this call
        order = [(me + i) % self.P for i in range(self.P)]
        for d in order:
            if self.needs_data(epoch, me, d):
                if self.use_rtr:
                    self.sig[me].wait_ge(RTR + d, epoch)
                self.jitter()
                with self.lock:
                    for c in classes_all:
                        self._write(d, (c, me), epoch)
                    if d // self.U == r:
                        for c in classes_ring:
                            self._write(d, (c, me), epoch)
            elif not self.signal_skipped:
                continue
            self._signal(d, me, u, d // self.U == r)

    def _write(self, dst, key, epoch):
        rd = self.reading[dst].get(key)
        self.check(rd is None, f"rank {key[1]} overwrites staging {key} of rank {dst} (epoch {epoch}) while it reads epoch {rd}")
        self.staging[dst][key] = epoch

    def _signal(self, d, me, u, same_ring):
        self.sig[d].add(KV + me, self.n_comm)
        if same_ring:
            self.sig[d].add(Q + u, self.n_comm)
        self.sig[d].add(QA + me, self.n_comm)

    def consume(self, me, epoch, needs):
        """compute CTAs: needs = [(flag slot, (cls, src))...] in visiting order."""
        for slot, key in needs:
            self.sig[me].wait_ge(slot, epoch * self.n_comm)
            with self.lock:
                got = self.staging[me].get(key)
                self.check(got == epoch, f"rank {me} reads {key}: epoch {got}, wants {epoch}")
                self.reading[me][key] = epoch
            self.jitter()
            if self.slow_reader is not None and self.slow_reader[0] == me:
                time.sleep(self.slow_reader[1])
            with self.lock:
                self.reading[me][key] = None

    def kernel(self, me, epoch, classes_all, classes_ring, needs, owners, done_slot, target):
        """One fused launch: push thread || compute thread; the launch ends when both are done and, if this rank owns
        outputs produced elsewhere, when its completion counter reached the host-computed target."""
        res = {}

        def guard(fn, *a):
            try:
                fn(*a)
            except Exception as e:  # noqa: BLE001
                res[fn.__name__] = e

        th = []
        if classes_all or classes_ring:
            th.append(threading.Thread(target=guard, args=(self.push, me, epoch, classes_all, classes_ring), daemon=True))
        th.append(threading.Thread(target=guard, args=(self.consume, me, epoch, needs), daemon=True))
        for t in th:
            t.start()
        for t in th:
            t.join(60)
            assert not t.is_alive(), f"deadlock in kernel of rank {me}, epoch {epoch}: {res}"
        assert not res, res
        for o in owners:                                              # tiles scattered to the token owners
            self.sig[o].add(done_slot, 1)
        if target is not None:
            self.sig[me].wait_ge(done_slot, target)

    # ------------------------------------------------------------------ a rank's stream
    def rank(self, me, steps):
        U, R, P = self.U, self.R, self.P
        u, r = me % U, me // U
        ring_block = [r * U + x for x in range(U)]                    # the ranks whose queries I compute (my ring index)
        epoch, o_total, dkv_total = 0, 0, 0
        for _ in range(steps):
            self.jitter(2e-3)                                         # host-side drift between ranks
            # forward: K/V to everyone, Q to my Ulysses peers; I read Q of my ring block and every K/V
            epoch += 1
            needs = ([(Q + x, ("q", r * U + x)) for x in range(U)] if U > 1 else []) + \
                    [(KV + s, ("kv", s)) for s in [(me + i) % P for i in range(P)] if self.needs_data(epoch, s, me)]
            if U > 1:
                o_total += U
            self.kernel(me, epoch, ["kv"], ["q"] if U > 1 else [], needs, ring_block if U > 1 else [], ODONE,
                        o_total if U > 1 else None)
            # backward, dQ pass: q/dO/stats and K/V to EVERY rank; dQ tiles go to the owners in my ring block
            epoch += 1
            needs = [(QA + s, ("qa", s)) for s in ring_block] + \
                    [(KV + s, ("kv", s)) for s in range(P) if self.needs_data(epoch, s, me)]
            o_total += U
            self.kernel(me, epoch, ["kv", "qa"], [], needs, ring_block, ODONE, o_total)
            # backward, dK/dV pass: no comm role; reads what pass 1 delivered; dK/dV tiles go to the owners
            needs = [(KV + s, ("kv", s)) for s in ring_block] + \
                    [(QA + s, ("qa", s)) for s in range(P) if self.needs_data(epoch, s, me)]
            dkv_total += U
            self.kernel(me, epoch, [], [], needs, ring_block, DKV, dkv_total)    # target = the host's symm_wait


def _run_mesh(mesh, steps, join=120):
    res = {}

    def guard(me):
        try:
            mesh.rank(me, steps=steps)
        except Exception as e:  # noqa: BLE001
            res[me] = e

    threads = [threading.Thread(target=guard, args=(me,), daemon=True) for me in range(mesh.P)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(join)
        assert not t.is_alive(), f"deadlock: {res}"
    return res


@pytest.mark.parametrize("U,R", [(1, 2), (2, 1), (2, 2), (1, 8), (4, 2)])
@pytest.mark.parametrize("masked", [False, True])
def test_fused_cross_rank_protocol(U, R, masked):
    mesh = Mesh(U, R, n_comm=2, seed=U * 10 + R, masked=masked)
    res = _run_mesh(mesh, steps=3)
    assert not res, res
    assert not mesh.errors, mesh.errors


def test_model_detects_a_skipped_destination_that_gets_no_signal():
    """Negative control for the destination masks: if a source skipped the arrival signal together with the data, its
    counter would lag one call behind for ever and the next call that does need its rows would wait for ever."""
    Sig.TIMEOUT = 2.0
    try:
        mesh = Mesh(1, 4, n_comm=2, seed=5, masked=True, signal_skipped=False)
        res = {}

        def guard(me):
            try:
                mesh.rank(me, steps=3)
            except Exception as e:  # noqa: BLE001
                res[me] = e

        threads = [threading.Thread(target=guard, args=(me,), daemon=True) for me in range(4)]
        for t in threads:
            t.start()
        for t in threads:
            t.join(60)
        assert res, "a lagging arrival counter went unnoticed"
    finally:
        Sig.TIMEOUT = 20.0


def test_model_detects_a_missing_ready_to_receive_handshake():
    """Negative control: without the RTR wait a fast rank overwrites a slow peer's staging (or the peer reads data of
    the wrong epoch) -- the model must notice."""
    Sig.TIMEOUT = 2.0
    try:
        for attempt in range(3):
            mesh = Mesh(1, 4, n_comm=2, seed=3 + attempt, use_rtr=False, slow_reader=(0, 0.3))

            def guard(me):
                try:
                    mesh.rank(me, steps=4)
                except Exception:  # noqa: BLE001
                    pass

            threads = [threading.Thread(target=guard, args=(me,), daemon=True) for me in range(4)]
            for t in threads:
                t.start()
            for t in threads:
                t.join(60)
            if mesh.errors:
                return
        pytest.fail("the model did not detect the missing handshake in three runs")
    finally:
        Sig.TIMEOUT = 20.0
