"""Guard for the hardware-validated kernels: the default (non-experimental) instantiations must still compile to the
exact instruction streams that ran on the B200s in the last GPU session (``profiles/sass_hashes_r2.json``, produced
by ``tools/sass_identity.py hash``).  Opt-in variants are template-gated, so adding one must not perturb a single
SASS instruction of the defaults.  After an intentional, re-validated change of a default kernel regenerate the file:

    python tools/sass_identity.py hash lca_b200/ops/build/*.o > profiles/sass_hashes_rN.json
"""
import json
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("cuobjdump") is None and not os.path.exists("/usr/local/cuda/bin/cuobjdump"),
                    reason="CUDA toolkit (cuobjdump) not available")
def test_validated_kernel_instantiations_are_bit_identical():
    env = dict(os.environ, PATH=os.environ.get("PATH", "") + ":/usr/local/cuda/bin")
    subprocess.run([sys.executable, "-m", "lca_b200.ops.build"], check=True, cwd=ROOT, env=env, capture_output=True,
                   timeout=1500)
    objs = [os.path.join(ROOT, "lca_b200", "ops", "build", f + ".o")
            for f in ("fmha_fwd_sm100", "fmha_bwd_sm100", "util_kernels", "fmha_fwd_fp8_sm100")]
    assert all(os.path.exists(o) for o in objs)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_identity.py"), "checkhash",
                        os.path.join(ROOT, "profiles", "sass_hashes_r2.json"), *objs], capture_output=True, text=True,
                       env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    n = len(json.load(open(os.path.join(ROOT, "profiles", "sass_hashes_r2.json"))))
    assert f"{n}/{n} validated instantiations identical" in r.stdout
