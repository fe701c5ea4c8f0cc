"""In-tree build of the sm_100a extension ``lca_b200/ops/_C*.so``.

* ``.cu`` files are compiled by nvcc for ``compute_100a/sm_100a`` only (``-lineinfo`` so ncu's
  source page maps to our code); they include no torch headers, so a kernel edit rebuilds in
  seconds.
* ``.cpp`` files (pybind11/torch bindings, symmetric-memory host code) are compiled by g++.
* objects are cached under ``lca_b200/ops/build/`` keyed by source+header mtimes.

The resulting ``.so`` lives in the package directory (git-ignored, but shipped to the GPU box
by gpurun) -- there is no JIT cache under ``~/.cache`` to miss on a fresh box.

Usage: ``python -m lca_b200.ops.build [--force] [--verbose]``
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
import sysconfig
from pathlib import Path
from typing import List

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
BUILD = HERE / "build"
EXT_NAME = "_C"

NVCC_ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _cuda_home() -> str:
    for c in (os.environ.get("CUDA_HOME"), os.environ.get("CUDA_PATH"), "/usr/local/cuda"):
        if c and os.path.exists(os.path.join(c, "bin", "nvcc")):
            return c
    raise RuntimeError("nvcc not found (set CUDA_HOME)")


def so_path() -> Path:
    return HERE / f"{EXT_NAME}{sysconfig.get_config_var('EXT_SUFFIX')}"


def _run(cmd: List[str], verbose: bool) -> None:
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError(f"build step failed: {' '.join(cmd[:3])} ...")
    if verbose and r.stdout.strip():
        print(r.stdout)


def _stamp(paths) -> str:
    h = hashlib.sha1()
    for p in sorted(paths):
        st = os.stat(p)
        h.update(f"{p}:{st.st_mtime_ns}:{st.st_size};".encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    import torch  # noqa: F401  (headers + libs)
    from torch.utils import cpp_extension as ce

    cuda = _cuda_home()
    nvcc = os.path.join(cuda, "bin", "nvcc")
    BUILD.mkdir(exist_ok=True)
    headers = sorted(str(p) for p in CSRC.glob("*.h")) + sorted(str(p) for p in CSRC.glob("*.cuh"))
    cus = sorted(CSRC.glob("*.cu"))
    cpps = sorted(CSRC.glob("*.cpp"))
    tinc = [f"-I{p}" for p in ce.include_paths()]
    pyinc = f"-I{sysconfig.get_paths()['include']}"
    cxx11 = int(bool(torch._C._GLIBCXX_USE_CXX11_ABI))

    # every translation unit is an independent job; nvcc/g++ run concurrently (a cold build is dominated by the two
    # attention kernels' template instantiations)
    jobs = []      # (cmd, stamp_file, stamp)
    objs: List[str] = []
    for src in cus + cpps:
        obj = BUILD / (src.stem + ".o")
        stamp_file = BUILD / (src.stem + ".stamp")
        stamp = _stamp([str(src)] + headers)
        objs.append(str(obj))
        if not (force or not obj.exists() or not stamp_file.exists() or stamp_file.read_text() != stamp):
            continue
        if src.suffix == ".cu":
            cmd = [nvcc, *NVCC_ARCH, "-O3", "-std=c++17", "-lineinfo", "--use_fast_math", "-Xcompiler", "-fPIC",
                   "-Xptxas", "-v" if verbose else "-O3", f"-I{CSRC}", "-c", str(src), "-o", str(obj)]
        else:
            cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-Wno-deprecated-declarations",
                   f"-D_GLIBCXX_USE_CXX11_ABI={cxx11}", f"-DTORCH_EXTENSION_NAME={EXT_NAME}",
                   "-DTORCH_API_INCLUDE_EXTENSION_H", f"-I{CSRC}", f"-I{cuda}/include", *tinc, pyinc, "-c", str(src),
                   "-o", str(obj)]
        jobs.append((cmd, stamp_file, stamp))
    rebuilt = bool(jobs)
    if jobs:
        from concurrent.futures import ThreadPoolExecutor
        workers = max(1, min(len(jobs), int(os.environ.get("LCA_B200_BUILD_JOBS", os.cpu_count() or 1))))
        with ThreadPoolExecutor(max_workers=workers) as pool:
            for fut in [pool.submit(_run, cmd, verbose) for cmd, _, _ in jobs]:
                fut.result()                      # re-raises the first failing step
        for _, stamp_file, stamp in jobs:
            stamp_file.write_text(stamp)

    out = so_path()
    if rebuilt or force or not out.exists():
        tlib = ce.library_paths()[0]
        _run(["g++", "-shared", "-o", str(out), *objs, f"-L{tlib}", f"-L{cuda}/lib64", "-lc10", "-lc10_cuda",
              "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python", "-lcudart",
              f"-Wl,-rpath,{tlib}", f"-Wl,-rpath,{cuda}/lib64"], verbose)
    return out


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv or "-v" in sys.argv)
    print(p)
