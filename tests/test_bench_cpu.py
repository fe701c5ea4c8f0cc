"""Control-flow dry runs of the headline benchmark (``bench.py --device cpu``: gloo, PyTorch engine, host timers).
The numbers are meaningless; what is checked is that every phase of the script (warm-up, device-timed loop, pipelined
end-to-end loop with H2D of q/k/v/dO and a D2H result per step, the compute-only communication probe, the max-over-
ranks reductions) runs to completion and that the ONE JSON line carries the keys of the driver's contract."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches"}


def _run(n, port, *flags):
    cmd = [sys.executable]
    if n > 1:
        cmd += ["-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
                "--master-port", str(port)]
    cmd += [os.path.join(ROOT, "bench.py"), "--gpus", str(n), "--steps", "2", "--warmup", "1", "--device", "cpu",
            "--seq", "256", "--heads", "4", "--head-dim", "16", *flags]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    return json.loads(lines[0])


def test_bench_single_process_contract():
    d = _run(1, 0)
    assert KEYS <= set(d)
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] >= 3 and d["higher_is_better"] is True
    assert d["metric"] == "attention_tflops_fwd_bwd" and d["config"]["seq_len"] == 256
    e = d["e2e"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(e)
    assert e["h2d_bytes_per_step"] == 4 * 256 * 4 * 16 * 2          # q, k, v and dO shards in bf16
    assert e["d2h_bytes_per_step"] == 4 and d["ms_per_step"] > 0 and e["ms_per_step"] > 0


@pytest.mark.parametrize("flags", [
    (),
    ("--mode", "fwd", "--ring-impl", "strip"),
    ("--ulysses", "2", "--window", "64", "--kv-heads", "2"),
    ("--ulysses", "2", "--qkvpacked", "--ring-impl", "basic"),
])
def test_bench_two_ranks_all_phases(flags):
    d = _run(2, 29871 + len(flags), *flags)
    assert KEYS <= set(d) and d["n_gpus"] == 2
    assert d["e2e"]["h2d_bytes_per_step"] > 0
    comm = d["comm"]
    assert comm is not None and "error" not in comm, comm
    assert comm["compute_only_ms"] > 0 and "exposed_comm_ms" in comm


def test_bench_multi_config_in_one_process_group():
    """``--configs 2,3 --modes fwd,fwdbwd``: one JSON line per config x mode from ONE launch (the presets' shapes are
    overridden by explicit flags, so the dry run stays tiny)."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29741", os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "1", "--warmup", "1",
           "--device", "cpu", "--seq", "128", "--heads", "4", "--head-dim", "16", "--configs", "2,3", "--modes", "fwd,fwdbwd",
           "--no-comm-probe"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
    assert [(d["config_id"], d["config"]["mode"]) for d in lines] == [(2, "fwd"), (2, "fwdbwd"), (3, "fwd"), (3, "fwdbwd")]
    assert lines[0]["config"]["parallelism"] == "ulysses2xring1" and lines[2]["config"]["parallelism"] == "ulysses1xring2"
