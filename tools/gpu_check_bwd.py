"""Numerics/perf probe of the sm_100a backward passes vs the fp32 torch oracle (+ FA2 timing)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lca_b200.ops import native
from lca_b200.ops.attention import AttnParams
from lca_b200.ops.ref_attention import attn_block_bwd_ref
from lca_b200.parallel.layout import Seg, pos_tensor

os.makedirs("gpurun_out", exist_ok=True)
LOG = open("gpurun_out/bwd_check.jsonl", "a")


def emit(d):
    print(json.dumps(d), flush=True)
    LOG.write(json.dumps(d) + "\n"); LOG.flush()


def run(name, B, Sq, Sk, H, Hkv, D, causal=False, window=(-1, -1), softcap=0.0, alibi=False, dtype=torch.bfloat16,
        time_it=False, check=True):
    torch.manual_seed(0)
    dev = "cuda"
    q = torch.randn(B, Sq, H, D, device=dev, dtype=dtype)
    k = torch.randn(B, Sk, Hkv, D, device=dev, dtype=dtype)
    v = torch.randn(B, Sk, Hkv, D, device=dev, dtype=dtype)
    do = torch.randn_like(q)
    qp, kp = (Seg(Sk - Sq if causal and Sk >= Sq else 0, Sq, 1),), (Seg(0, Sk, 1),)
    slopes = (torch.rand(H, device=dev) * 0.5) if alibi else None
    p = AttnParams.make(q, None, causal, window, softcap, slopes)
    out, lse = native.fmha_fwd(q, k, v, qp, kp, p)
    dq, dk, dv = native.fmha_bwd(do, q, k, v, out, lse, qp, kp, p)
    torch.cuda.synchronize()
    d = dict(name=name, shape=[B, Sq, Sk, H, Hkv, D], causal=causal, window=list(window), softcap=softcap, alibi=alibi,
             dtype=str(dtype))
    if check:
        rq, rk, rv = attn_block_bwd_ref(do, q, k, v, out, lse, pos_tensor(qp, dev), pos_tensor(kp, dev), p.softmax_scale,
                                        causal, window, softcap, slopes)
        for nm, a, b in (("dq", dq, rq), ("dk", dk, rk), ("dv", dv, rv)):
            e = (a.float() - b).abs()
            d[nm + "_err"] = e.max().item(); d[nm + "_ref_max"] = b.abs().max().item(); d[nm + "_nan"] = bool(torch.isnan(a).any())
            if e.max().item() > 0.05 * (b.abs().max().item() + 1e-6):
                rows = a.shape[1]
                d[nm + "_blocks"] = {f"r{r0}c{c0}": round(e[:, r0:r0 + 64, :, c0:c0 + 64].max().item(), 4)
                                     for r0 in range(0, min(rows, 256), 64) for c0 in range(0, D, 64)}
    if time_it:
        delta, lse2 = native.attn_delta(out, do, lse)
        for _ in range(3):
            native.fmha_bwd(do, q, k, v, out, lse, qp, kp, p, delta=delta, lse2=lse2, dq=dq, dk=dk, dv=dv)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 5
        e0.record()
        for _ in range(n):
            native.fmha_bwd(do, q, k, v, out, lse, qp, kp, p, delta=delta, lse2=lse2, dq=dq, dk=dk, dv=dv)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        fl = 2.5 * 4.0 * B * H * Sq * Sk * D * (0.5 if causal else 1.0)
        d["ms"] = ms; d["tflops"] = fl / ms / 1e9
        try:
            from flash_attn import flash_attn_func
            q2, k2, v2 = (t.clone().requires_grad_() for t in (q, k, v))
            o2 = flash_attn_func(q2, k2, v2, causal=causal)
            for _ in range(2): o2.backward(do, retain_graph=True)
            torch.cuda.synchronize(); e0.record()
            for _ in range(n): o2.backward(do, retain_graph=True)
            e1.record(); torch.cuda.synchronize()
            ms2 = e0.elapsed_time(e1) / n
            d["fa2_ms"] = ms2; d["fa2_tflops"] = fl / ms2 / 1e9
        except Exception as e:
            d["fa2_err"] = str(e)[:100]
    emit(d)


if __name__ == "__main__":
    emit(dict(device=torch.cuda.get_device_name(0), native=native.available(), has_bwd=native.has_bwd()))
    run("1tile", 1, 128, 64, 1, 1, 128)
    run("1x2", 1, 128, 128, 1, 1, 128)
    run("2x3", 1, 256, 192, 1, 1, 128)
    run("4x8", 1, 512, 512, 2, 2, 128)
    run("d64", 1, 256, 512, 1, 1, 64)
    run("fp16", 1, 256, 512, 2, 2, 128, dtype=torch.float16)
    run("causal", 1, 512, 512, 2, 2, 128, causal=True)
    run("causal_ragged", 2, 333, 333, 3, 3, 128, causal=True)
    run("ragged_nc_gqa", 2, 200, 777, 4, 2, 64)
    run("gqa", 2, 1024, 1024, 8, 2, 128, causal=True)
    run("window", 1, 1024, 1024, 2, 2, 128, causal=True, window=(300, 0))
    run("window_nc", 1, 1024, 1024, 2, 2, 128, window=(100, 200))
    run("softcap", 1, 512, 512, 2, 2, 128, causal=True, softcap=15.0)
    run("alibi", 1, 512, 512, 4, 4, 128, causal=True, alibi=True)
    run("multiwork", 2, 4096, 4096, 16, 16, 128, causal=True)
    if "--quick" not in sys.argv:
        run("perf_c_8k", 1, 8192, 8192, 16, 16, 128, causal=True, time_it=True, check=False)
        run("perf_c_32k", 1, 32768, 32768, 8, 8, 128, causal=True, time_it=True, check=False)
        run("perf_nc_8k", 1, 8192, 8192, 16, 16, 128, time_it=True, check=False)
        run("perf_c_8k_d64", 1, 8192, 8192, 16, 16, 64, causal=True, time_it=True, check=False)
    emit(dict(done=True))
