"""Executable model of the barrier protocol of ``csrc/fmha_fwd_sm100.cu`` (one 128-column score buffer per Q tile,
tensor-pipe order ``QK0(j+1) | PV1(j) | QK1(j+1) | PV0(j+1)``, 5-slot K/V ring); empty work items are handed back
through ``o_full`` (``qf=True``, what every instantiation does since round 2; ``qf=False`` is the round-1 rule, kept as the
negative control).  Same construction as the backward model."""
import queue
import random
import threading
import time

import pytest

from pipeline_model import MBar

STAGES = 5


class FwdModel:
    def __init__(self, work, seed, qf=True, slow_mma=0.0):
        self.work, self.qf, self.slow_mma = work, qf, slow_mma      # work: list of (ntile, n_kv_tiles)
        self.rng = random.Random(seed)
        self.q_full, self.q_empty = [MBar(1), MBar(1)], [MBar(1), MBar(1)]
        self.s_full, self.p_full = [MBar(1), MBar(1)], [MBar(4), MBar(4)]
        self.o_full = [MBar(1), MBar(1)]
        self.kv_full = [MBar(1) for _ in range(STAGES)]
        self.kv_empty = [MBar(1) for _ in range(STAGES)]
        self.pipe = queue.Queue()
        self.lock = threading.Lock()
        self.errors = []
        self.kv_slot = [None] * STAGES
        self.s_buf = [("empty",), ("empty",)]
        self.pv_inflight = [0, 0]
        self.q_tile = [None, None]

    def jitter(self, scale=1e-4):
        time.sleep(self.rng.random() * scale)

    def check(self, cond, msg):
        if not cond:
            self.errors.append(msg)
            raise AssertionError(msg)

    def pipe_thread(self):
        while True:
            op = self.pipe.get()
            if op is None:
                return
            self.jitter(3e-4)
            with self.lock:
                if op[0] == "QK":
                    _, w, t, tile, slot = op
                    self.check(self.kv_slot[slot] == ("K", w, tile), f"QK reads slot {slot}: {self.kv_slot[slot]}")
                    self.check(self.s_buf[t][0] in ("empty", "consumed"), f"QK overwrites live scores {self.s_buf[t]}")
                    self.check(self.q_tile[t] == w, "QK without its Q tile")
                    self.s_buf[t] = ("S", w, tile)
                elif op[0] == "PV":
                    _, w, t, tile, slot = op
                    self.check(self.kv_slot[slot] == ("V", w, tile), f"PV reads slot {slot}: {self.kv_slot[slot]}")
                    self.check(self.s_buf[t] == ("P", w, tile), f"PV reads {self.s_buf[t]} != P{tile}")
                    self.s_buf[t] = ("consumed",)
                    self.pv_inflight[t] -= 1
                elif op[0] == "COMMIT":
                    op[1].arrive()
                elif op[0] == "FREE":
                    self.kv_slot[op[2]] = None
                    op[1].arrive()

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
        qc, pc, kvc = [0, 0], [0, 0], 0
        for w, (nt, nkv) in enumerate(self.work):
            time.sleep(self.slow_mma)
            tiles = iter(range(nkv))
            cur = next(tiles, None)
            for t in range(nt):
                self.q_full[t].wait(qc[t] & 1)
                qc[t] += 1
            if cur is None:
                if self.qf:
                    for t in range(nt):
                        self.pipe.put(("COMMIT", self.o_full[t]))
                continue
            kslot = kvc % STAGES
            self.kv_full[kslot].wait((kvc // STAGES) & 1)
            kvc += 1
            for t in range(nt):
                self.pipe.put(("QK", w, t, cur, kslot))
                self.pipe.put(("COMMIT", self.s_full[t]))
            self.pipe.put(("FREE", self.kv_empty[kslot], kslot))
            j = 0
            while True:
                vslot, vpar = kvc % STAGES, (kvc // STAGES) & 1
                kvc += 1
                nxt = next(tiles, None)
                if nxt is not None:
                    kslot, kpar = kvc % STAGES, (kvc // STAGES) & 1
                    kvc += 1
                self.kv_full[vslot].wait(vpar)
                for t in range(nt):
                    self.p_full[t].wait(pc[t] & 1)
                    pc[t] += 1
                    with self.lock:
                        self.pv_inflight[t] += 1
                    self.pipe.put(("PV", w, t, cur, vslot))
                    if t == nt - 1:
                        self.pipe.put(("FREE", self.kv_empty[vslot], vslot))
                    if nxt is not None:
                        if t == 0:
                            self.kv_full[kslot].wait(kpar)
                        self.pipe.put(("QK", w, t, nxt, kslot))
                        self.pipe.put(("COMMIT", self.s_full[t]))
                        if t == nt - 1:
                            self.pipe.put(("FREE", self.kv_empty[kslot], kslot))
                    else:
                        self.pipe.put(("COMMIT", self.o_full[t]))
                if nxt is None:
                    break
                cur = nxt
                j += 1

    def softmax(self, t):
        sc = oc = qc = 0
        for w, (ntile, nkv) in enumerate(self.work):
            if t >= ntile:
                continue
            j = 0
            for tile in range(nkv):
                self.s_full[t].wait(sc & 1)
                sc += 1
                with self.lock:
                    self.check(self.s_buf[t] == ("S", w, tile), f"softmax reads {self.s_buf[t]} != S{tile}")
                    if j > 0:
                        self.check(self.pv_inflight[t] == 0, "O touched while a PV of this tile is in flight")
                self.jitter(4e-4)
                with self.lock:
                    self.s_buf[t] = ("P", w, tile)
                for _ in range(4):
                    self.p_full[t].arrive()
                j += 1
            if j > 0 or self.qf:
                self.o_full[t].wait(oc & 1)
                oc += 1
            else:
                self.q_full[t].wait(qc & 1)
            qc += 1
            with self.lock:
                self.check(self.pv_inflight[t] == 0, "epilogue before the last PV")
                self.check(self.q_tile[t] == w, "epilogue on another item's Q buffer")
                self.q_tile[t] = None
            self.q_empty[t].arrive()

    def run(self):
        results = {}

        def guard(fn, args, name):
            try:
                fn(*args)
            except Exception as e:  # noqa: BLE001
                results[name] = e

        roles = [(self.pipe_thread, (), "pipe"), (self.producer, (), "tma"), (self.mma, (), "mma"),
                 (self.softmax, (0,), "wg0"), (self.softmax, (1,), "wg1")]
        threads = [threading.Thread(target=guard, args=r, daemon=True) for r in roles]
        for th in threads:
            th.start()
        for th in threads[1:]:
            th.join(90)
            assert not th.is_alive(), f"deadlock: a role did not finish ({results})"
        self.pipe.put(None)
        threads[0].join(10)
        assert not results, results
        assert not self.errors, self.errors


@pytest.mark.parametrize("seed", range(4))
def test_forward_pipeline_protocol(seed):
    rng = random.Random(31 * seed + 5)
    work = [(2, 1), (1, 2), (2, 0), (2, 3), (1, 1), (2, 7), (2, 0), (1, 0)] + \
           [(rng.choice([1, 2]), rng.randint(0, 9)) for _ in range(6)]
    FwdModel(work, seed).run()


def test_forward_empty_items_without_o_full_handback_need_a_prompt_mma_warp():
    """What ``kQf`` closes: with the default rule (the warpgroup of an EMPTY work item waits ``q_full`` itself and
    releases ``q_empty``) an MMA warp that lagged a whole work item would see ``q_full`` complete two phases under its
    one-bit parity wait.  The model forces that lag; on hardware the MMA warp reaches the wait microseconds before the Q
    tile can be reloaded, which is why the validated default has never shown it."""
    MBar.TIMEOUT = 3.0
    try:
        for attempt in range(3):
            try:
                FwdModel([(1, 1), (1, 0), (1, 0), (1, 1)], attempt, qf=False, slow_mma=0.4).run()
            except AssertionError:
                break
        else:
            pytest.fail("the default rule survived three forced-lag runs")
        FwdModel([(1, 1), (1, 0), (1, 0), (1, 1)], 0, qf=True, slow_mma=0.4).run()
    finally:
        MBar.TIMEOUT = 20.0
