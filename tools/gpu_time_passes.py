"""Time fwd / dQ pass / dKV pass separately (CUDA events) -- used standalone and under ncu."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lca_b200.ops import native
from lca_b200.ops.attention import AttnParams
from lca_b200.parallel.layout import Seg

S = int(os.environ.get("S", 8192)); H = int(os.environ.get("H", 8)); D = int(os.environ.get("D", 128))
n = int(os.environ.get("N", 5)); causal = os.environ.get("CAUSAL", "1") == "1"
torch.manual_seed(0)
q, k, v = (torch.randn(1, S, H, D, device="cuda", dtype=torch.bfloat16) for _ in range(3))
do = torch.randn_like(q)
p = AttnParams.make(q, None, causal)
qp = kp = (Seg(0, S, 1),)
out, lse = native.fmha_fwd(q, k, v, qp, kp, p)
delta, lse2 = native.attn_delta(out, do, lse)
C = native.ext()
dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
wl, wr = native.window_bounds(p)
xq = [[0, S, 0, 0, 0]]; yk = [[0, S, 0, -1, 0]]


def t(fn, name):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    fl = 4.0 * H * S * S * D * (0.5 if causal else 1.0)
    print(json.dumps(dict(name=name, ms=ms, fwd_equiv_tflops=fl / ms / 1e9)), flush=True)


t(lambda: native.fmha_fwd(q, k, v, qp, kp, p, out=out, lse=lse), "fwd")
t(lambda: C.fmha_bwd_pass(False, q, do, k, v, xq, yk, 1, 1, lse2, delta, dq, None, False, p.softmax_scale, wl, wr, 0.0, None, 0), "dq_pass(3 gemms = 1.5x fwd flops)")
t(lambda: C.fmha_bwd_pass(True, k, v, q, do, xq, yk, 1, 1, lse2, delta, dk, dv, False, p.softmax_scale, wr, wl, 0.0, None, 0), "dkv_pass(4 gemms = 2x fwd flops)")
