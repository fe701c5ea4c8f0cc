"""Single-GPU harness of the fused kernels' push engine: a 1x1 "mesh" whose only destination is this GPU's own slab.
The push CTAs copy k/v (and q) into the staging buffers, bump the arrival counters, and the compute CTAs consume the
staging exactly as on a multi-GPU mesh -- so the copy loops, the mbarrier/bulk-group pipeline and the flag protocol can
be debugged (and run under compute-sanitizer) without a second GPU.

    python tools/debug_push_1gpu.py                 # bulk engine (default)
    LCA_B200_PUSH=scalar python tools/debug_push_1gpu.py
    compute-sanitizer --tool memcheck python tools/debug_push_1gpu.py
"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lca_b200.ops import native
from lca_b200.ops.attention import AttnParams
from lca_b200.parallel.layout import Seg

B = int(os.environ.get("B", 1)); S = int(os.environ.get("S", 2048)); H = int(os.environ.get("H", 4))
Hkv = int(os.environ.get("HKV", 2)); D = int(os.environ.get("D", 128)); n_comm = int(os.environ.get("NCOMM", 4))
iters = int(os.environ.get("ITERS", 3))
dev = torch.device("cuda", 0)
torch.manual_seed(0)
C = native.ext()
q, k, v = (torch.randn(B, S, h, D, device=dev, dtype=torch.bfloat16) for h in (H, Hkv, Hkv))
p = AttnParams.make(q, None, True)
ref, ref_lse = native.fmha_fwd(q, k, v, (Seg(0, S, 1),), (Seg(0, S, 1),), p)
esz = 2
skv = B * S * Hkv * D * esz
off_k, off_v = 0, (skv + 1023) // 1024 * 1024
slab = torch.zeros(off_v + skv + 4096, dtype=torch.uint8, device=dev)
sig = torch.zeros(1024, dtype=torch.int32, device=dev)
kst = slab[off_k:off_k + skv].view(torch.bfloat16).view(B, S, Hkv, D)
vst = slab[off_v:off_v + skv].view(torch.bfloat16).view(B, S, Hkv, D)
wl, wr = native.window_bounds(p)
SIG_KV = 0
for it in range(1, iters + 1):
    slab.zero_()
    out = torch.empty_like(q)
    lse = torch.empty(B, H, S, dtype=torch.float32, device=dev)
    qsegs = [[0, S, 0, -1, 0, 0, 0, 0]]
    ksegs = [[0, S, 0, SIG_KV + 0, 0]]
    C.usp_fwd(q, kst, vst, q, k, v, qsegs, ksegs, 1, 1, out, 0, lse, float(p.softmax_scale), wl, wr, 0.0, None,
              [1, 1, 1, 0, 0, S, n_comm], [0, off_k, off_v, S, S], [slab.data_ptr()], [sig.data_ptr()], sig.data_ptr(), it, 0)
    torch.cuda.synchronize()
    ek = (kst.float() - k.float()).abs().max().item()
    ev = (vst.float() - v.float()).abs().max().item()
    eo = (out.float() - ref.float()).abs().max().item()
    print(f"iter {it}: staged k err {ek} v err {ev} out err {eo:.5f} flags {sig[:4].tolist()} rtr {sig[32:34].tolist()}", flush=True)
    assert ek == 0 and ev == 0 and eo < 1e-2
print("push harness ok")
