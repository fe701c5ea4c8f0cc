"""Executable model of the synchronisation protocol of ``csrc/fmha_fwd_bn64_sm100.cu`` (64-row K/V tiles, two score
stages per Q tile, QK issued two tiles ahead, ``pv_done`` guard for the lazy O rescale).

Four role threads (TMA producer, MMA issuer, two softmax warpgroups) run the same control flow as the kernel against
modelled mbarriers (phase/parity semantics, arrival counts) and an in-order asynchronous "tensor pipe" that executes
the queued MMAs and ``tcgen05.commit`` arrivals with random delays.  Every resource carries a state machine, so a
protocol error shows up as an assertion (wrong tile in a K/V slot, QK overwriting a score stage whose P was not
consumed, softmax reading a stage that does not hold its tile, O rescaled while a PV of that tile is in flight) or as
a deadlock (join timeout).  The model is a transcription of the kernel's barrier logic, kept next to it on purpose."""
import queue
import random
import threading
import time

import pytest

STAGES = 10


class MBar:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0
        self.cv = threading.Condition()

    def arrive(self):
        with self.cv:
            self.pending -= 1
            assert self.pending >= 0
            if self.pending == 0:
                self.pending = self.count
                self.phase += 1
                self.cv.notify_all()

    TIMEOUT = 20.0

    def wait(self, parity, timeout=None):
        """mbarrier.try_wait.parity: returns once the phase with this parity has completed."""
        with self.cv:
            ok = self.cv.wait_for(lambda: (self.phase & 1) != parity, timeout or MBar.TIMEOUT)
            assert ok, "deadlock: barrier wait timed out"


class Model:
    def __init__(self, work_items, seed):
        self.work = work_items                      # list of (ntile, n_kv_tiles)
        self.rng = random.Random(seed)
        self.q_full = [MBar(1), MBar(1)]
        self.q_empty = [MBar(1), MBar(1)]
        self.o_full = [MBar(1), MBar(1)]
        self.pv_done = [MBar(1), MBar(1)]
        self.s_full = [[MBar(1), MBar(1)], [MBar(1), MBar(1)]]
        self.p_full = [[MBar(4), MBar(4)], [MBar(4), MBar(4)]]
        self.kv_full = [MBar(1) for _ in range(STAGES)]
        self.kv_empty = [MBar(1) for _ in range(STAGES)]
        self.pipe = queue.Queue()
        self.errors = []
        # resource state
        self.kv_slot = [None] * STAGES              # ("K"|"V", work, tile) or None
        self.s_stage = [[("empty",)] * 2 for _ in range(2)]      # per (t, stage)
        self.pv_inflight = [0, 0]
        self.q_tile = [None, None]
        self.lock = threading.Lock()

    def jitter(self, scale=1e-4):
        time.sleep(self.rng.random() * scale)

    def check(self, cond, msg):
        if not cond:
            self.errors.append(msg)
            raise AssertionError(msg)

    # ------------------------------------------------------------------ tensor pipe (in-order, asynchronous)
    def pipe_thread(self):
        while True:
            op = self.pipe.get()
            if op is None:
                return
            self.jitter(3e-4)
            kind = op[0]
            with self.lock:
                if kind == "QK":
                    _, w, t, s, tile, kslot = op
                    self.check(self.kv_slot[kslot] == ("K", w, tile), f"QK reads slot {kslot}: {self.kv_slot[kslot]} != K{tile}")
                    self.check(self.s_stage[t][s][0] in ("empty", "consumed"), f"QK overwrites live stage {self.s_stage[t][s]}")
                    self.check(self.q_tile[t] == w, "QK without its Q tile")
                    self.s_stage[t][s] = ("S", w, tile)
                elif kind == "PV":
                    _, w, t, s, tile, vslot = op
                    self.check(self.kv_slot[vslot] == ("V", w, tile), f"PV reads slot {vslot}: {self.kv_slot[vslot]} != V{tile}")
                    self.check(self.s_stage[t][s] == ("P", w, tile), f"PV reads stage {self.s_stage[t][s]} != P{tile}")
                    self.s_stage[t][s] = ("consumed",)
                    self.pv_inflight[t] -= 1
                elif kind == "COMMIT":
                    op[1].arrive()
                elif kind == "FREE":                # commit on kv_empty: the slot's readers have completed
                    self.kv_slot[op[2]] = None
                    op[1].arrive()

    # ------------------------------------------------------------------ roles
    def producer(self):
        qc, kvc = [0, 0], 0
        for w, (ntile, nkv) in enumerate(self.work):
            for t in range(ntile):
                self.q_empty[t].wait((qc[t] & 1) ^ 1)
                with self.lock:
                    self.q_tile[t] = w
                self.q_full[t].arrive()
                qc[t] += 1
            for tile in range(nkv):
                for kind in ("K", "V"):
                    slot, par = kvc % STAGES, (kvc // STAGES) & 1
                    self.kv_empty[slot].wait(par ^ 1)
                    self.jitter()
                    with self.lock:
                        self.check(self.kv_slot[slot] is None, f"TMA overwrites live slot {slot}")
                        self.kv_slot[slot] = (kind, w, tile)
                    self.kv_full[slot].arrive()
                    kvc += 1

    def mma(self):
        qc, pcbits, kvc = [0, 0], 0, 0
        for w, (nt, nkv) in enumerate(self.work):
            tiles = iter(range(nkv))
            more = next(tiles, None) is not None
            for t in range(nt):
                self.q_full[t].wait(qc[t] & 1)
                qc[t] += 1
            if not more:
                for t in range(nt):                 # empty item: Q tiles are handed back through o_full
                    self.pipe.put(("COMMIT", self.o_full[t]))
                continue
            base, n_qk = kvc, 0
            while more and n_qk < 2:
                idx = base + 2 * n_qk
                slot = idx % STAGES
                self.kv_full[slot].wait((idx // STAGES) & 1)
                for t in range(nt):
                    self.pipe.put(("QK", w, t, n_qk, n_qk, slot))
                    self.pipe.put(("COMMIT", self.s_full[t][n_qk]))
                self.pipe.put(("FREE", self.kv_empty[slot], slot))
                n_qk += 1
                more = next(tiles, None) is not None
            j = 0
            while j < n_qk:
                s = j & 1
                vidx = base + 2 * j + 1
                vslot = vidx % STAGES
                issue_next = more
                kidx = base + 2 * n_qk
                kslot = kidx % STAGES
                self.kv_full[vslot].wait((vidx // STAGES) & 1)
                for t in range(nt):
                    self.p_full[t][s].wait((pcbits >> (2 * t + s)) & 1)
                    pcbits ^= 1 << (2 * t + s)
                    with self.lock:
                        self.pv_inflight[t] += 1
                    self.pipe.put(("PV", w, t, s, j, vslot))
                    self.pipe.put(("COMMIT", self.pv_done[t]))
                    if t == nt - 1:
                        self.pipe.put(("FREE", self.kv_empty[vslot], vslot))
                    if issue_next:
                        if t == 0:
                            self.kv_full[kslot].wait((kidx // STAGES) & 1)
                        self.pipe.put(("QK", w, t, s, n_qk, kslot))
                        self.pipe.put(("COMMIT", self.s_full[t][s]))
                        if t == nt - 1:
                            self.pipe.put(("FREE", self.kv_empty[kslot], kslot))
                    elif j == n_qk - 1:
                        self.pipe.put(("COMMIT", self.o_full[t]))
                if issue_next:
                    n_qk += 1
                    more = next(tiles, None) is not None
                j += 1
            kvc = base + 2 * n_qk

    def softmax(self, t):
        scbits, oc, pvc = 0, 0, 0
        for w, (ntile, nkv) in enumerate(self.work):
            if t >= ntile:
                continue
            j = 0
            for tile in range(nkv):
                s = j & 1
                self.s_full[t][s].wait((scbits >> s) & 1)
                scbits ^= 1 << s
                with self.lock:
                    self.check(self.s_stage[t][s] == ("S", w, tile), f"softmax reads stage {self.s_stage[t][s]} != S{tile}")
                self.jitter(4e-4)
                if j > 0 and self.rng.random() < 0.3:          # a lazy rescale happens now and then
                    self.pv_done[t].wait((pvc + j - 1) & 1)
                    with self.lock:
                        self.check(self.pv_inflight[t] == 0, f"O_{t} rescaled while {self.pv_inflight[t]} PV in flight")
                with self.lock:
                    self.s_stage[t][s] = ("P", w, tile)
                for _ in range(4):                             # one arrive per softmax warp
                    self.p_full[t][s].arrive()
                j += 1
            pvc += j
            self.o_full[t].wait(oc & 1)
            oc += 1
            with self.lock:
                self.check(self.pv_inflight[t] == 0, "epilogue before the last PV")
                self.check(self.q_tile[t] == w, "epilogue reuses the Q buffer of another work item")
            with self.lock:
                self.q_tile[t] = None
            self.q_empty[t].arrive()

    def run(self):
        threads = [threading.Thread(target=f, args=a, daemon=True) for f, a in
                   [(self.pipe_thread, ()), (self.producer, ()), (self.mma, ()), (self.softmax, (0,)), (self.softmax, (1,))]]
        results = {}

        def guard(fn, args, name):
            try:
                fn(*args)
            except Exception as e:  # noqa: BLE001
                results[name] = e

        threads = [threading.Thread(target=guard, args=(f, a, n), daemon=True) for f, a, n in
                   [(self.pipe_thread, (), "pipe"), (self.producer, (), "tma"), (self.mma, (), "mma"),
                    (self.softmax, (0,), "wg0"), (self.softmax, (1,), "wg1")]]
        for th in threads:
            th.start()
        for th in threads[1:]:
            th.join(60)
            assert not th.is_alive(), "deadlock: a role did not finish"
        self.pipe.put(None)
        threads[0].join(10)
        assert not results, results
        assert not self.errors, self.errors
        assert all(x is None for x in self.kv_slot)


@pytest.mark.parametrize("seed", range(6))
def test_bn64_pipeline_protocol(seed):
    rng = random.Random(100 + seed)
    # work items: (Q tiles in the pair, K/V tiles); include empty, single-tile, odd and long items
    work = [(2, 1), (1, 2), (2, 0), (2, 3), (1, 1), (2, 7)] + [(rng.choice([1, 2]), rng.randint(0, 9)) for _ in range(6)]
    Model(work, seed).run()
