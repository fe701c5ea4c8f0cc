"""torchrun --nproc-per-node 2 tools/debug_fused.py : step-by-step fused fwd/bwd with progress markers."""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
import lca_b200
from lca_b200 import EXTRACT_FUNC_DICT, LongContextAttention, set_seq_parallel_pg
from lca_b200.kernels.attention import pytorch_attn_func
U = int(os.environ.get("U", 2)); R = world // U
variant = os.environ.get("VARIANT", "zigzag")
def mark(s):
    print(f"[r{rank} {time.time()%1000:.2f}] {s}", flush=True)
g = torch.Generator().manual_seed(11)
B, S, H, Hkv, D = 1, 1024, 4, 4, 128
q, k, v, do = (torch.randn(B, S, h, D, generator=g).to("cuda", torch.bfloat16) for h in (H, Hkv, Hkv, H))
q1, k1, v1 = (t.clone().requires_grad_() for t in (q, k, v))
ref = pytorch_attn_func(q1, k1, v1, causal=True); ref.backward(do)
set_seq_parallel_pg(U, R, rank, world)
key = {"basic": "basic", "zigzag": "zigzag", "stripe": "strip"}[variant]
sh = lambda t: EXTRACT_FUNC_DICT[key](t, rank, world, rd=R, ud=U).detach().clone()
lq, lk, lv = (sh(t).requires_grad_() for t in (q, k, v))
attn = LongContextAttention(ring_impl_type=key, backend="fused")
for i in range(2):
    mark(f"fwd {i} launch")
    out = attn(lq, lk, lv, causal=True)
    torch.cuda.synchronize(); mark(f"fwd {i} done err={(out.float()-sh(ref.detach()).float()).abs().max().item():.4f}")
eng = attn._fused
mark("bwd launch")
from lca_b200.ops import native
C = native.ext()
orig = C._mod.usp_bwd_pass
def traced(*a, **kw):
    mark(f"usp_bwd_pass is_dkv={a[0]} launch"); r = orig(*a, **kw); torch.cuda.synchronize(); mark(f"usp_bwd_pass is_dkv={a[0]} done"); return r
C._mod = type("M", (), {"__getattr__": lambda self, n: traced if n == "usp_bwd_pass" else getattr(sys.modules["lca_b200.ops._C"], n)})()
out.backward(sh(do))
torch.cuda.synchronize(); mark("bwd done")
for a, b, n in ((lq.grad, q1.grad, "dq"), (lk.grad, k1.grad, "dk"), (lv.grad, v1.grad, "dv")):
    mark(f"{n} err {(a.float()-sh(b).float()).abs().max().item():.4f} ref {sh(b).float().abs().max().item():.3f}")
dist.destroy_process_group()
