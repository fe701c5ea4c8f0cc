"""Single-GPU peers of our kernels on the same shapes: cuDNN SDPA (tcgen05 fused attention shipped with cuDNN 9, through
torch.nn.functional.scaled_dot_product_attention with the CUDNN_ATTENTION backend) and flash-attn 2.8 (mma.sync build).
CUDA-event timings, forward and forward+backward, same FLOP convention as bench.py (4*B*H*S^2*D, halved when causal,
backward = 2.5x forward).  Library code -- for the comparison row in profiles/README.md only."""
import json, os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

S = int(os.environ.get("S", 32768)); H = int(os.environ.get("H", 8)); D = int(os.environ.get("D", 128))
n = int(os.environ.get("N", 5)); causal = os.environ.get("CAUSAL", "1") == "1"
torch.manual_seed(0)
fl = 4.0 * H * S * S * D * (0.5 if causal else 1.0)


def t(fn, name, mult):
    try:
        for _ in range(2): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(json.dumps(dict(name=name, S=S, H=H, D=D, causal=causal, ms=round(ms, 4), tflops=round(mult * fl / ms / 1e9, 1))), flush=True)
    except Exception as e:  # noqa: BLE001 - a peer that is not available on this box is reported, not fatal
        print(json.dumps(dict(name=name, error=str(e)[:200])), flush=True)


def fb(f, *ts):
    def run():
        for x in ts: x.grad = None
        o = f()
        o.backward(do_)
    return run


# ---- ours (public single-device entry point)
from lca_b200.kernels.attention import flash_attn_func as attention
q, k, v = (torch.randn(1, S, H, D, device="cuda", dtype=torch.bfloat16, requires_grad=True) for _ in range(3))
do_ = torch.randn(1, S, H, D, device="cuda", dtype=torch.bfloat16)
with torch.no_grad():
    t(lambda: attention(q, k, v, causal=causal), "ours fwd", 1.0)
t(fb(lambda: attention(q, k, v, causal=causal), q, k, v), "ours fwd+bwd", 3.5)

# ---- cuDNN SDPA
from torch.nn.attention import SDPBackend, sdpa_kernel
qh, kh, vh = (torch.randn(1, H, S, D, device="cuda", dtype=torch.bfloat16, requires_grad=True) for _ in range(3))
do_h = torch.randn(1, H, S, D, device="cuda", dtype=torch.bfloat16)


def cudnn():
    with sdpa_kernel(SDPBackend.CUDNN_ATTENTION):
        return F.scaled_dot_product_attention(qh, kh, vh, is_causal=causal)


def cudnn_fb():
    for x in (qh, kh, vh): x.grad = None
    cudnn().backward(do_h)


with torch.no_grad():
    t(cudnn, "cudnn sdpa fwd", 1.0)
t(cudnn_fb, "cudnn sdpa fwd+bwd", 3.5)

# ---- flash-attn 2.8 (the reference's kernel)
try:
    import importlib
    flash_attn_func = importlib.import_module('flash_attn').flash_attn_func
    with torch.no_grad():
        t(lambda: flash_attn_func(q, k, v, causal=causal), "flash-attn2 fwd", 1.0)
    t(fb(lambda: flash_attn_func(q, k, v, causal=causal), q, k, v), "flash-attn2 fwd+bwd", 3.5)
except Exception as e:  # noqa: BLE001
    print(json.dumps(dict(name="flash-attn2", error=str(e)[:200])))
