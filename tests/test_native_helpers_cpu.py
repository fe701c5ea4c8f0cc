"""Host-side helpers of the native path that do not need a GPU."""
import os

import pytest
import torch

from lca_b200.ops import native
from lca_b200.ops.attention import AttnParams, block_is_visible
from lca_b200.parallel.layout import Seg, ring_positions, varlen_positions


def test_window_bounds_fold_causal_into_right_bound():
    q = torch.zeros(1, 4, 1, 8)
    assert native.window_bounds(AttnParams.make(q, None, True)) == (-1, 0)
    assert native.window_bounds(AttnParams.make(q, None, True, (7, -1))) == (7, 0)
    assert native.window_bounds(AttnParams.make(q, None, False, (3, 5))) == (3, 5)
    assert native.window_bounds(AttnParams.make(q, None, False)) == (-1, -1)


def test_rows_and_strides():
    spec = ring_positions("zigzag", 1, 4, 8)
    assert native._rows(spec) == [(0, 4, 4, 0), (4, 4, 24, 0)]
    assert native._common_stride(ring_positions("stripe", 2, 4, 6)) == 4
    assert native._common_stride(spec) == 1


def test_chunk_by_group_splits_many_sequences():
    n = native.MAX_SEG + 8
    cu = list(range(0, (n + 1) * 16, 16))               # more sequences (= groups) than one launch takes segments
    spec = varlen_positions("basic", 0, 1, cu)
    rows = native._rows(spec)
    chunks = list(native._chunk_by_group(rows, rows))
    assert len(chunks) == 2 and all(len(q) <= native.MAX_SEG and len(k) <= native.MAX_SEG for q, k in chunks)
    assert sorted(r for q, _ in chunks for r in q) == sorted(rows)
    for q, k in chunks:                                  # a chunk holds whole groups on both sides
        assert {r[3] for r in q} == {r[3] for r in k}


def test_block_visibility_matches_reference_step_rule():
    """basic causal ring: step visible iff source rank <= my rank (ring_flash_attn.py:35); zigzag: always."""
    q = torch.zeros(1, 8, 1, 8)
    p = AttnParams.make(q, None, True)
    R, L = 4, 8
    for r in range(R):
        for src in range(R):
            vis = block_is_visible(ring_positions("basic", r, R, L), ring_positions("basic", src, R, L), p)
            assert vis == (src <= r)
            assert block_is_visible(ring_positions("zigzag", r, R, L), ring_positions("zigzag", src, R, L), p)
    pw = AttnParams.make(q, None, True, (3, 0))          # window 3: only the adjacent earlier block matters
    assert not block_is_visible((Seg(24, 8),), (Seg(0, 8),), pw)
    assert block_is_visible((Seg(24, 8),), (Seg(16, 8),), pw)


def test_padded_dim_and_support_messages():
    assert native._padded_dim(32) == 64 and native._padded_dim(96) == 128 and native._padded_dim(128) == 128
    assert "not on CUDA" in native.why_not(torch.zeros(1, 1, 1, 64))


def test_dropout_hash_cxx_recipe_matches_python_spec(tmp_path):
    """The kernels' integer dropout recipe (``csrc/sm100_ptx.cuh: mix32 / dropout_row_key / dropout_word /
    dropout_keep``, declared ``__host__ __device__``) compiled for the HOST and compared bit for bit with
    ``lca_b200/ops/dropout.py`` on a grid of coordinates."""
    import shutil
    import subprocess
    import torch
    from lca_b200.ops import dropout as d
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    csrc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lca_b200", "ops", "csrc")
    src = tmp_path / "h.cu"
    src.write_text('''
#include <cstdio>
#include <cstdint>
#include "sm100_ptx.cuh"
int main() {
  const uint32_t seed = 20240921u, p8 = 77u;
  for (uint32_t b = 0; b < 2; ++b) for (uint32_t h = 0; h < 3; ++h)
    for (uint32_t q = 0; q < 5; ++q) {
      const uint32_t qpos = 1000003u * q + 17u;
      const uint32_t rk = lca::ptx::dropout_row_key(qpos, seed, b, h + 5u);
      for (uint32_t k = 0; k < 37; ++k) {
        const uint32_t kpos = 262139u * (k / 9) + k;
        std::printf("%d", lca::ptx::dropout_keep(lca::ptx::dropout_word(rk, kpos), kpos, p8) ? 1 : 0);
      }
      std::printf("\\n");
    }
  return 0;
}
''')
    exe = tmp_path / "h"
    subprocess.run([nvcc, "-std=c++17", f"-I{csrc}", "-o", str(exe), str(src)], check=True, capture_output=True)
    lines = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    qpos = torch.tensor([1000003 * q + 17 for q in range(5)])
    kpos = torch.tensor([262139 * (k // 9) + k for k in range(37)])
    m = d.keep_mask(20240921, 2, 3, qpos, kpos, 77 / 256.0, head_offset=5)
    want = ["".join(str(int(x)) for x in m[b, h, q].tolist()) for b in range(2) for h in range(3) for q in range(5)]
    assert lines == want


def test_exp2_polynomial_accuracy_claim():
    """``ex2_poly`` (FMA-pipe exp2 of the forward softmax, ``csrc/sm100_ptx.cuh``): the coefficients are read from the
    header and the recipe is replayed in float32 -- max relative error must stay at the documented 1.0e-4, far below
    the bf16 rounding (3.9e-3) applied to P right after."""
    import re
    import numpy as np
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lca_b200", "ops", "csrc",
                            "sm100_ptx.cuh")).read()
    body = src[src.index("float ex2_poly(float x)"):]
    c3, c2, c1 = (np.float32(v) for v in re.search(
        r"fmaf\(fmaf\(fmaf\(([0-9.]+)f, f, ([0-9.]+)f\), f, ([0-9.]+)f\), f, 1\.0f\)", body).groups())
    magic = np.float32(12582912.0)
    x = np.linspace(-40, 0, 1_000_001, dtype=np.float32)
    t = (np.maximum(x, np.float32(-126)) + magic).astype(np.float32)
    f = (x - (t - magic).astype(np.float32)).astype(np.float32)
    p = ((((c3 * f + c2).astype(np.float32) * f + c1).astype(np.float32)) * f + np.float32(1)).astype(np.float32)
    y = (p.view(np.int32) + (t.view(np.int32) << 23)).view(np.float32).astype(np.float64)
    ref = np.exp2(x.astype(np.float64))
    assert float(np.max(np.abs(y - ref) / ref)) < 1.1e-4


def test_extension_entry_points_accept_the_argument_lists_python_passes():
    """pybind arity / type check without a GPU: with CPU tensors every kernel entry point must get past argument
    conversion (a mismatch is a TypeError) and fail later in the CUDA guard (RuntimeError)."""
    import torch
    if not native.extension_loaded():
        pytest.skip("extension not built")
    C = native.ext()._mod
    q = torch.zeros(1, 128, 1, 128, dtype=torch.bfloat16)
    lse, out = torch.zeros(1, 1, 128), torch.zeros_like(q)
    qs, ks = [[0, 128, 0, -1, 0, 0, 0, 0]], [[0, 128, 0, -1, 0]]
    xq, yk = [[0, 128, 0, 0, 0]], [[0, 128, 0, -1, 0]]
    calls = {
        "fmha_fwd": lambda: C.fmha_fwd(q, q, q, qs, ks, 1, 1, out, 0, lse, 0.1, -1, 0, 0.0, None, 0, 0, 0),
        "fmha_fwd_drop": lambda: C.fmha_fwd_drop(q, q, q, qs, ks, 1, 1, out, 0, lse, 0.1, -1, 0, 0.0, None, 0, 0, 0, [26, 1, 0]),
        "fmha_bwd_pass": lambda: C.fmha_bwd_pass(False, q, q, q, q, xq, yk, 1, 1, lse, lse, out, None, False, 0.1, -1, 0, 0.0,
                                                 None, 0),
        "fmha_bwd_pass_drop": lambda: C.fmha_bwd_pass_drop(False, q, q, q, q, xq, yk, 1, 1, lse, lse, out, None, False, 0.1, -1,
                                                           0, 0.0, None, 0, [26, 1, 0]),
    }
    for name, fn in calls.items():
        with pytest.raises(RuntimeError):
            fn()
    C.set_next_dropout([])
