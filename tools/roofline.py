"""Achieved fraction of roofline for the benchmark JSON lines in profiles/ (and any given as arguments).

Roofline time of a fused compute+collective path = max(FLOPs / measured GEMM peak, NVLink bytes / link bandwidth)
per GPU (B200_PROFILING.md): peaks from MEASURED_PEAKS.json (fallback 1.59 PFLOP/s), link = measured 770 GB/s/direction.
Bytes that must cross NVLink per rank (inbound) on a U x R mesh: fwd: K,V head slices of the P-1 peers, Q of the U-1
Ulysses peers, my share of the O tiles computed elsewhere; fwd+bwd adds q, dO, k, v of the P-1 peers once more
(owner-computes backward) and the scattered dq / dk / dv -- 16-bit elements throughout.  The FLOP count ignores sliding
windows (like bench.py), so windowed configs read above 1.
"""
import glob
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
try:
    PK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    BURST, SUST, SRC = PK["bf16_tflops"], PK["bf16_tflops_sustained"], "measured"
except Exception:  # noqa: BLE001
    BURST, SUST, SRC = 1590.0, 1400.0, "fallback"
LINK = 770e9


def analyse(d):
    cfg = d["config"]
    N, S, H, D = d["n_gpus"], cfg["seq_len"], cfg["heads"], cfg["head_dim"]
    Hkv = cfg.get("kv_heads", H)
    fb = cfg.get("mode", "fwd") == "fwdbwd"
    flops = 4.0 * cfg["global_batch"] * H * S * S * D * (0.5 if cfg.get("causal", True) else 1.0) * (3.5 if fb else 1.0)
    # inbound NVLink bytes per rank on the fused path, U x R mesh (parallelism = "ulysses<U>xring<R>"): every peer's K/V
    # head slice once (forward and again in the backward), my ring block's Q from the U-1 Ulysses peers + my share of
    # the O tiles the others computed (forward), every peer's q / dO head slice + the scattered dq / dk / dv (backward)
    import re
    m = re.match(r"ulysses(\d+)xring(\d+)", cfg.get("parallelism", ""))
    U = int(m.group(1)) if m else 1
    rows = cfg["global_batch"] * (S // N)
    q_slice, kv_slice = rows * (H // U) * D * 2, rows * max(Hkv // U, 1) * D * 2
    o_in = rows * H * D * 2 * (U - 1) // U
    nv = (N - 1) * 2 * kv_slice + (U - 1) * q_slice + o_in
    if fb:
        nv += (N - 1) * (2 * q_slice + 2 * kv_slice) + o_in + 2 * rows * Hkv * D * 2 * (U - 1) // U
    t = d["ms_per_step"] * 1e-3
    t_c_s, t_c_b = flops / N / (SUST * 1e12), flops / N / (BURST * 1e12)
    t_l = nv / LINK
    return dict(impl=d.get("impl", "ours"), n=N, cfg=d.get("config_id", 3), par=cfg.get("parallelism", ""), mode=cfg.get("mode"), ms=d["ms_per_step"], tflops=d["value"],
                compute_ms_sustained=t_c_s * 1e3, nvlink_ms=t_l * 1e3, bound="compute" if t_c_s >= t_l else "nvlink",
                frac_sustained=max(t_c_s, t_l) / t, frac_burst=max(t_c_b, t_l) / t, nvlink_MB=nv / 1e6)


def main():
    files = sys.argv[1:] or sorted(glob.glob(os.path.join(ROOT, "profiles", "*.json")))
    rows = []
    for f in files:
        for line in open(f):
            line = line.strip()
            if line.startswith("{") and '"metric"' in line:
                try:
                    rows.append((os.path.basename(f), analyse(json.loads(line))))
                except Exception:  # noqa: BLE001
                    pass
    print(f"# Roofline fractions ({SRC} peaks: {BURST} TFLOPS burst / {SUST} sustained bf16; NVLink 770 GB/s per direction)\n")
    print("| file | impl | N | config | mode | ms/step | TFLOPS | compute ms @sustained | NVLink ms (MB in) | bound | of roofline (sustained) | (burst) |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|")
    for f, r in rows:
        print(f"| {f} | {r['impl']} | {r['n']} | {r['cfg']} ({r['par']}) | {r['mode']} | {r['ms']:.1f} | {r['tflops']:.0f} | {r['compute_ms_sustained']:.1f} | "
              f"{r['nvlink_ms']:.2f} ({r['nvlink_MB']:.0f}) | {r['bound']} | {r['frac_sustained']:.2f} | {r['frac_burst']:.2f} |")


if __name__ == "__main__":
    main()
