"""Standalone numerics/perf probe of the sm_100a forward kernel vs the fp32 torch oracle.
Run on a B200: python tools/gpu_check_fwd.py [--quick] ; appends JSON lines to gpurun_out/fwd_check.jsonl"""
import json, math, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lca_b200.ops import native
from lca_b200.ops.attention import AttnParams
from lca_b200.ops.ref_attention import attn_block_fwd_ref
from lca_b200.parallel.layout import Seg, pos_tensor

os.makedirs("gpurun_out", exist_ok=True)
LOG = open("gpurun_out/fwd_check.jsonl", "a")

def emit(d):
    print(json.dumps(d), flush=True)
    LOG.write(json.dumps(d) + "\n"); LOG.flush()

def run(name, B, Sq, Sk, H, Hkv, D, causal=False, window=(-1, -1), softcap=0.0, alibi=False, dtype=torch.bfloat16,
        q_pos=None, k_pos=None, time_it=False):
    torch.manual_seed(0)
    dev = "cuda"
    q = torch.randn(B, Sq, H, D, device=dev, dtype=dtype)
    k = torch.randn(B, Sk, Hkv, D, device=dev, dtype=dtype)
    v = torch.randn(B, Sk, Hkv, D, device=dev, dtype=dtype)
    q_pos = q_pos or (Seg(Sk - Sq if causal and Sk >= Sq else 0, Sq, 1),)
    k_pos = k_pos or (Seg(0, Sk, 1),)
    slopes = (torch.rand(H, device=dev) * 0.5) if alibi else None
    p = AttnParams.make(q, None, causal, window, softcap, slopes)
    out, lse = native.fmha_fwd(q, k, v, q_pos, k_pos, p)
    torch.cuda.synchronize()
    ro, rl = attn_block_fwd_ref(q, k, v, pos_tensor(q_pos, dev), pos_tensor(k_pos, dev), p.softmax_scale, causal,
                                window, softcap, slopes)
    of, rf = out.float(), ro.float()
    err = (of - rf).abs()
    fin = torch.isfinite(rl)
    lerr = (lse[fin] - rl[fin]).abs().max().item() if fin.any() else 0.0
    inf_match = bool((torch.isinf(lse) == torch.isinf(rl)).all())
    d = dict(name=name, shape=[B, Sq, Sk, H, Hkv, D], causal=causal, window=list(window), softcap=softcap, alibi=alibi,
             dtype=str(dtype), max_err=err.max().item(), mean_err=err.mean().item(), lse_err=lerr, inf_match=inf_match,
             nan=bool(torch.isnan(of).any()))
    if err.max().item() > 0.05:
        # localise: per 128-row block x 64-col block error
        blocks = {}
        for r0 in range(0, min(Sq, 512), 128):
            for c0 in range(0, D, 64):
                blocks[f"r{r0}c{c0}"] = round(err[:, r0:r0 + 128, :, c0:c0 + 64].max().item(), 4)
        d["blocks"] = blocks
    if time_it:
        for _ in range(3):
            native.fmha_fwd(q, k, v, q_pos, k_pos, p, out=out, lse=lse)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 10
        e0.record()
        for _ in range(n):
            native.fmha_fwd(q, k, v, q_pos, k_pos, p, out=out, lse=lse)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        fl = 4.0 * B * H * Sq * Sk * D * (0.5 if causal else 1.0)
        d["ms"] = ms; d["tflops"] = fl / ms / 1e9
        try:
            from flash_attn import flash_attn_func
            for _ in range(3): flash_attn_func(q, k, v, causal=causal)
            torch.cuda.synchronize(); e0.record()
            for _ in range(n): flash_attn_func(q, k, v, causal=causal)
            e1.record(); torch.cuda.synchronize()
            ms2 = e0.elapsed_time(e1) / n
            d["fa2_ms"] = ms2; d["fa2_tflops"] = fl / ms2 / 1e9
        except Exception as e:
            d["fa2_err"] = str(e)[:100]
    emit(d)

if __name__ == "__main__":
    quick = "--quick" in sys.argv
    emit(dict(device=torch.cuda.get_device_name(0), native=native.available(), why=native.why_not(torch.empty(1,1,1,128,device="cuda",dtype=torch.bfloat16))))
    run("1tile", 1, 128, 128, 1, 1, 128)
    run("2q", 1, 256, 128, 1, 1, 128)
    run("2k", 1, 128, 256, 1, 1, 128)
    run("2q4k", 1, 256, 512, 1, 1, 128)
    run("d64", 1, 256, 512, 1, 1, 64)
    run("fp16", 1, 256, 512, 2, 2, 128, dtype=torch.float16)
    run("causal", 1, 512, 512, 2, 2, 128, causal=True)
    run("causal_ragged", 2, 333, 333, 3, 3, 128, causal=True)
    run("ragged_nc", 2, 200, 777, 4, 2, 64)
    run("gqa", 2, 1024, 1024, 8, 2, 128, causal=True)
    run("window", 1, 1024, 1024, 2, 2, 128, causal=True, window=(300, 0))
    run("window_nc", 1, 1024, 1024, 2, 2, 128, window=(100, 200))
    run("softcap", 1, 512, 512, 2, 2, 128, causal=True, softcap=15.0)
    run("alibi", 1, 512, 512, 4, 4, 128, causal=True, alibi=True)
    run("zigzag_q", 1, 512, 1024, 2, 2, 128, causal=True, q_pos=(Seg(256, 256, 1), Seg(1536, 256, 1)),
        k_pos=(Seg(0, 512, 1), Seg(1024, 512, 1)))
    run("stripe", 1, 512, 512, 2, 2, 128, causal=True, q_pos=(Seg(1, 512, 4),), k_pos=(Seg(2, 512, 4),))
    run("multiwork", 2, 4096, 4096, 16, 16, 128, causal=True)
    if not quick:
        run("perf_nc_8k", 1, 8192, 8192, 16, 16, 128, time_it=True)
        run("perf_c_8k", 1, 8192, 8192, 16, 16, 128, causal=True, time_it=True)
        run("perf_c_32k", 1, 32768, 32768, 8, 8, 128, causal=True, time_it=True)
        run("perf_c_8k_d64", 1, 8192, 8192, 16, 16, 64, causal=True, time_it=True)
    emit(dict(done=True))
