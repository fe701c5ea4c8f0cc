"""Modelled mbarrier (phase/parity semantics, arrival counts) shared by the executable protocol models of the kernels'
barrier logic (``test_fwd_pipeline_model_cpu.py``, ``test_bwd_pipeline_model_cpu.py``)."""
import threading


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
