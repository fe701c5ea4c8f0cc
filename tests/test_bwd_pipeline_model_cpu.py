"""Executable model of the synchronisation protocol of ``csrc/fmha_bwd_sm100.cu`` for both passes (dQ / dK,dV) and
both element-wise modes: the hardware-validated default (warpgroup ``wg`` owns the streamed tiles with ``j % 2 == wg``)
and the opt-in ``kSplit`` variant (both warpgroups work on every tile, 32 columns each: ``p_full`` and ``st_empty`` take
8 arrivals instead of 4, ``t_full`` is tracked per stage).

Role threads (producer warp, MMA issuer, two element-wise warpgroups) run the kernel's control flow against modelled
mbarriers and an in-order asynchronous tensor pipe; every buffer (X tile, Y ring, statistics ring, the two T/dS TMEM
stages, the accumulators) carries a state machine, so a protocol error is an assertion and a missed phase a timeout.
Modelled mbarriers: ``pipeline_model.py``."""
import queue
import random
import threading
import time

import pytest

from pipeline_model import MBar

STAGES = 4


class BwdModel:
    def __init__(self, work, split, is_dkv, seed, xfix=True, slow_epilogue=0.0):
        self.work, self.split, self.is_dkv = work, split, is_dkv       # work: list of n_streamed_tiles per item
        self.xfix = xfix and not is_dkv      # kXfix instantiation: element-wise warps also arrive on x_empty (count 9)
        self.slow_epilogue = slow_epilogue
        self.rng = random.Random(seed)
        n = 8 if split else 4
        self.x_full, self.x_empty = MBar(1), MBar(9 if (xfix and not is_dkv) else 1)
        self.acc_full, self.acc_empty = MBar(1), MBar(8)
        self.t_full = [MBar(1), MBar(1)]
        self.p_full = [MBar(n), MBar(n)]
        self.y_full = [MBar(1) for _ in range(STAGES)]
        self.y_empty = [MBar(1) for _ in range(STAGES)]
        self.st_full = [MBar(32) for _ in range(STAGES)]
        self.st_empty = [MBar(n) for _ in range(STAGES)]
        self.pipe = queue.Queue()
        self.lock = threading.Lock()
        self.errors = []
        self.x_tile = None
        self.y_slot = [None] * STAGES
        self.st_slot = [None] * STAGES
        self.t_stage = [("empty",), ("empty",)]
        self.t_done = [0, 0]                 # element-wise halves finished on the current contents of a stage
        self.acc_owner = None                # work item whose partial sums live in the accumulators
        self.acc_readers = 0

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
                if op[0] == "T":
                    _, w, tile, st, s = op
                    self.check(self.x_tile == w, f"T GEMM reads X of {self.x_tile}, wants {w}")
                    self.check(self.y_slot[st] == (w, tile), f"T GEMM reads Y slot {st}: {self.y_slot[st]} != {(w, tile)}")
                    self.check(self.t_stage[s][0] in ("empty", "consumed"), f"T GEMM overwrites live stage {self.t_stage[s]}")
                    self.t_stage[s] = ("T", w, tile)
                    self.t_done[s] = 0
                elif op[0] == "ACC":
                    _, w, tile, st, s, first = op
                    self.check(self.y_slot[st] == (w, tile), f"accumulate reads Y slot {st}: {self.y_slot[st]}")
                    self.check(self.t_stage[s] == ("P", w, tile), f"accumulate reads stage {self.t_stage[s]} != P{(w, tile)}")
                    self.check(self.acc_readers == 0, "accumulate while the epilogue still reads the accumulators")
                    self.check(first or self.acc_owner == w, "accumulate into another work item's sums")
                    self.acc_owner = w
                    self.t_stage[s] = ("consumed",)
                elif op[0] == "COMMIT":
                    op[1].arrive()
                elif op[0] == "FREE_Y":
                    self.y_slot[op[2]] = None
                    op[1].arrive()
                elif op[0] == "FREE_X":          # every T GEMM of the item has completed: X is dead for the tensor pipe
                    self.x_tile = None
                    op[1].arrive()

    def producer(self):
        xc = yc = 0
        for w, ntiles in enumerate(self.work):
            self.x_empty.wait((xc & 1) ^ 1)
            with self.lock:
                self.check(self.x_tile is None, "TMA overwrites a live X tile")
                self.x_tile = w
            self.x_full.arrive()
            xc += 1
            for tile in range(ntiles):
                st, par = yc % STAGES, (yc // STAGES) & 1
                self.y_empty[st].wait(par ^ 1)
                self.jitter()
                with self.lock:
                    self.check(self.y_slot[st] is None, f"TMA overwrites live Y slot {st}")
                    self.y_slot[st] = (w, tile)
                self.y_full[st].arrive()
                if self.is_dkv:
                    self.st_empty[st].wait(par ^ 1)
                    with self.lock:
                        self.st_slot[st] = (w, tile)
                    for _ in range(32):
                        self.st_full[st].arrive()
                yc += 1

    def mma(self):
        xc = yc = ac = 0
        pc = [0, 0]
        for w, ntiles in enumerate(self.work):
            tiles = iter(range(ntiles))
            self.x_full.wait(xc & 1)
            xc += 1
            cur = next(tiles, None)
            if cur is None:
                self.pipe.put(("FREE_X", self.x_empty))
                continue
            stage_q, tile_q = [0, 0], [0, 0]
            n_issued = 0
            for s in range(2):
                if cur is None:
                    break
                st = yc % STAGES
                self.y_full[st].wait((yc // STAGES) & 1)
                yc += 1
                self.pipe.put(("T", w, cur, st, s))
                self.pipe.put(("COMMIT", self.t_full[s]))
                stage_q[s], tile_q[s] = st, cur
                n_issued += 1
                cur = next(tiles, None)
            if cur is None:
                self.pipe.put(("FREE_X", self.x_empty))
            j = 0
            while j < n_issued:
                s = j & 1
                self.p_full[s].wait(pc[s] & 1)
                pc[s] += 1
                if j == 0:
                    self.acc_empty.wait((ac & 1) ^ 1)
                    ac += 1
                self.pipe.put(("ACC", w, tile_q[s], stage_q[s], s, j == 0))
                self.pipe.put(("FREE_Y", self.y_empty[stage_q[s]], stage_q[s]))
                if cur is not None:
                    st = yc % STAGES
                    self.y_full[st].wait((yc // STAGES) & 1)
                    yc += 1
                    self.pipe.put(("T", w, cur, st, s))
                    self.pipe.put(("COMMIT", self.t_full[s]))
                    stage_q[s], tile_q[s] = st, cur
                    n_issued += 1
                    cur = next(tiles, None)
                    if cur is None:
                        self.pipe.put(("FREE_X", self.x_empty))
                j += 1
            self.pipe.put(("COMMIT", self.acc_full))

    def elementwise(self, wg):
        tc = yc = afc = xcw = 0
        tcs = [0, 0]
        for w, ntiles in enumerate(self.work):
            if not self.is_dkv:
                self.x_full.wait(xcw & 1)
                xcw += 1
                if self.xfix:
                    for _ in range(4):               # one arrive per warp of this warpgroup
                        self.x_empty.arrive()
            j = 0
            for tile in range(ntiles):
                st, ypar = yc % STAGES, (yc // STAGES) & 1
                yc += 1
                if not self.split and (j & 1) != wg:
                    j += 1
                    continue
                sT = (j & 1) if self.split else wg
                if self.split:
                    self.t_full[sT].wait(tcs[sT] & 1)
                    tcs[sT] += 1
                else:
                    self.t_full[wg].wait(tc & 1)
                    tc += 1
                if self.is_dkv:
                    self.st_full[st].wait(ypar)
                with self.lock:
                    self.check(self.t_stage[sT][:3] in (("T", w, tile), ("P", w, tile)), f"wg{wg} reads stage {self.t_stage[sT]} != {(w, tile)}")
                    if self.is_dkv:
                        self.check(self.st_slot[st] == (w, tile), f"wg{wg} reads statistics of {self.st_slot[st]}")
                self.jitter(4e-4)
                with self.lock:
                    self.t_done[sT] += 1
                    if self.t_done[sT] == (2 if self.split else 1):
                        self.t_stage[sT] = ("P", w, tile)
                for _ in range(4):
                    self.p_full[sT].arrive()
                    if self.is_dkv:
                        self.st_empty[st].arrive()
                j += 1
            if j > 0:
                self.acc_full.wait(afc & 1)
                afc += 1
                with self.lock:
                    self.check(self.acc_owner == w, f"epilogue of item {w} reads sums of {self.acc_owner}")
                    self.acc_readers += 1
                self.jitter(3e-4)
                time.sleep(self.slow_epilogue)
                with self.lock:
                    self.acc_readers -= 1
                for _ in range(4):
                    self.acc_empty.arrive()

    def run(self):
        results = {}

        def guard(fn, args, name):
            try:
                fn(*args)
            except Exception as e:  # noqa: BLE001
                results[name] = e

        roles = [(self.pipe_thread, (), "pipe"), (self.producer, (), "producer"), (self.mma, (), "mma"),
                 (self.elementwise, (0,), "wg0"), (self.elementwise, (1,), "wg1")]
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


@pytest.mark.parametrize("split", [False, True])
@pytest.mark.parametrize("is_dkv", [False, True])
@pytest.mark.parametrize("seed", range(3))
def test_backward_pipeline_protocol(split, is_dkv, seed):
    rng = random.Random(7 * seed + 1)
    work = [1, 2, 0, 3, 1, 7, 0, 0, 2] + [rng.randint(0, 9) for _ in range(6)]
    BwdModel(work, split, is_dkv, seed).run()


@pytest.mark.parametrize("work", [[2, 0, 1, 1], [3, 2, 2, 2], [3, 1, 1, 1]])
def test_dq_pass_without_xfix_can_miss_a_phase(work):
    """The hazard kXfix removes, made deterministic: a work item whose T GEMMs are all issued in the MMA warp's
    prologue (no visible streamed tile, or at most two) right behind another item, while the element-wise warpgroups
    are still in the previous epilogue.  ``x_empty`` is released by the MMA warp alone, the producer reloads X,
    ``x_full`` completes two phases before the warpgroup's one-bit parity wait -> it blocks forever.  With kXfix every
    element-wise warp also arrives on ``x_empty``, so the same schedule completes."""
    MBar.TIMEOUT = 3.0
    try:
        for attempt in range(3):            # the forced lag is generous, but thread scheduling is not ours: retry
            try:
                BwdModel(work, False, False, attempt, xfix=False, slow_epilogue=0.3).run()
            except AssertionError:
                break
        else:
            pytest.fail("the pre-fix protocol survived three forced-lag runs")
        BwdModel(work, False, False, 0, xfix=True, slow_epilogue=0.3).run()
    finally:
        MBar.TIMEOUT = 20.0
