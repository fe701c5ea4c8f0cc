"""fp8 (e4m3, block-scaled) forward vs bf16 forward on one GPU: kernel-only time (pre-quantised operands), time with the
quantisation pass included, and the error of both against the fp32 sampled oracle."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("LCA_B200_EXPERIMENTAL_FP8", "1")
from lca_b200.ops import native
from lca_b200.ops.attention import AttnParams
from lca_b200.ops.fp8 import quantize_blockwise
from lca_b200.ops.sampled_oracle import head_oracle
from lca_b200.parallel.layout import Seg

S = int(os.environ.get("S", 32768)); H = int(os.environ.get("H", 8)); D = 128; n = int(os.environ.get("N", 5))
torch.manual_seed(0)
q, k, v = (torch.randn(1, S, H, D, device="cuda", dtype=torch.bfloat16) for _ in range(3))
p = AttnParams.make(q, None, True)
pos = (Seg(0, S, 1),)
C = native.ext()
wl, wr = native.window_bounds(p)
out = torch.empty_like(q); lse = torch.empty(1, H, S, dtype=torch.float32, device="cuda")
qsegs = [[0, S, 0, -1, 0, 0, 0, 0]]; ksegs = [[0, S, 0, -1, 0]]


def quant():
    q8, sq = quantize_blockwise(q); k8, sk = quantize_blockwise(k); v8, sv = quantize_blockwise(v, per_head=True)
    return q8.view(torch.uint8), k8.view(torch.uint8), v8.view(torch.uint8), sq, sk, sv


qq = quant()


def fp8_kernel():
    C.fmha_fwd_fp8(qq[0], qq[1], qq[2], qq[3], qq[4], qq[5], qsegs, ksegs, 1, 1, out, lse, float(p.softmax_scale), wl, wr, 0.0, None)


def t(fn, name):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    fl = 4.0 * H * S * S * D * 0.5
    print(json.dumps(dict(name=name, ms=round(ms, 4), tflops=round(fl / ms / 1e9, 1))), flush=True)
    return ms


o16, l16 = native.fmha_fwd(q, k, v, pos, pos, p)
t(lambda: native.fmha_fwd(q, k, v, pos, pos, p, out=o16, lse=l16), "bf16 fwd")
t(fp8_kernel, "fp8 fwd (kernel only)")
t(quant, "quantise q,k,v (e4m3 + scales)")
fp8_kernel()
rows = torch.randint(0, S, (256,)); rows[0] = 0; rows[1] = S - 1
ref = head_oracle(q[0, :, :1], k[0, :, 0], v[0, :, 0], None, rows, causal=True)["out"][:, 0]
rows = rows.cuda()
print(json.dumps(dict(err_bf16=float((o16[0, rows, 0].float() - ref).abs().max()),
                      err_fp8=float((out[0, rows, 0].float() - ref).abs().max()),
                      mean_err_fp8=float((out[0, rows, 0].float() - ref).abs().mean()))))
