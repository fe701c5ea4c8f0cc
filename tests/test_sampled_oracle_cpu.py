"""The long-sequence sampled oracle (``ops/sampled_oracle.py``) against the dense fp32 oracle on a small problem."""
import pytest
import torch

from lca_b200.kernels.attention import pytorch_attn_func
from lca_b200.ops.sampled_oracle import head_oracle, rel_err


@pytest.mark.parametrize("causal,window", [(True, (-1, -1)), (True, (37, 0)), (False, (-1, -1)), (False, (20, 11))])
def test_head_oracle_matches_dense_reference(causal, window):
    g = torch.Generator().manual_seed(3)
    S, H, Hkv, D = 192, 4, 2, 32
    q, k, v, do = (torch.randn(1, S, h, D, generator=g, dtype=torch.float64) for h in (H, Hkv, Hkv, H))
    q1, k1, v1 = (t.clone().float().requires_grad_() for t in (q, k, v))
    ref = pytorch_attn_func(q1, k1, v1, causal=causal, window_size=window)
    ref.backward(do.float())
    rows = torch.tensor([0, 1, 17, 100, 191])
    cols = torch.tensor([0, 5, 64, 150, 191])
    G = H // Hkv
    for hk in range(Hkv):
        qs = q[0, :, hk * G:(hk + 1) * G].float()
        dos = do[0, :, hk * G:(hk + 1) * G].float()
        r = head_oracle(qs, k[0, :, hk].float(), v[0, :, hk].float(), dos, rows, cols, causal=causal, window=window, chunk=50)
        assert rel_err(r["out"], ref[0, rows, hk * G:(hk + 1) * G].detach()) < 1e-4
        assert rel_err(r["dq"], q1.grad[0, rows, hk * G:(hk + 1) * G]) < 1e-4
        assert rel_err(r["dk"], k1.grad[0, cols, hk]) < 1e-4
        assert rel_err(r["dv"], v1.grad[0, cols, hk]) < 1e-4
