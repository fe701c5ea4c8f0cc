"""Run a function on N local processes over gloo (CPU) or nccl (GPU) and collect failures."""
import os
import socket
import traceback

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, backend, fn, args, errq):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["RANK"] = str(rank)
    os.environ["WORLD_SIZE"] = str(world)
    os.environ["LOCAL_RANK"] = str(rank)
    try:
        if backend == "nccl":
            torch.cuda.set_device(rank)
        dist.init_process_group(backend, rank=rank, world_size=world)
        fn(rank, world, *args)
        dist.barrier()
    except Exception:  # noqa: BLE001
        errq.put((rank, traceback.format_exc()))
    finally:
        if dist.is_initialized():
            try:
                dist.destroy_process_group()
            except Exception:  # noqa: BLE001
                pass


def run_distributed(fn, world, *args, backend="gloo", timeout=300):
    ctx = mp.get_context("spawn")
    errq = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, backend, fn, args, errq)) for r in range(world)]
    import time
    for p in procs:
        p.start()
    errs = []
    t0 = time.time()
    while any(p.is_alive() for p in procs) and time.time() - t0 < timeout:
        while not errq.empty():
            errs.append(errq.get())
        if errs:                       # fail fast: peers of a dead rank would spin on its flags forever
            time.sleep(2.0)
            break
        time.sleep(0.2)
    hung = [p for p in procs if p.is_alive()]
    for p in hung:
        p.kill()
    while not errq.empty():
        errs.append(errq.get())
    if errs:
        hung = []
    assert not hung, f"{len(hung)} rank(s) hung"
    assert not errs, "\n".join(f"[rank {r}]\n{tb}" for r, tb in errs)
    bad = [p.exitcode for p in procs if p.exitcode != 0]
    assert not bad, f"non-zero exit codes: {bad}"
