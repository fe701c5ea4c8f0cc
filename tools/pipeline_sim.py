"""Timing model of the kernels' pipelines: in-order tensor pipe + element-wise warpgroups + the barrier dependencies of
each design, to turn a measured tensor-pipe utilisation into the implied element-wise time W and to predict what the
opt-in variants do to the utilisation.  Pure arithmetic on clock counts (B200: 8192 dense bf16 FLOP/clk/SM), no GPU.

    python tools/pipeline_sim.py            # prints the markdown table kept in profiles/pipeline_sim_r1.md

Rules of the model: the MMA warp issues in program order and blocks on a barrier wait before issuing what follows;
the tensor pipe executes in issue order, an op starts when the pipe is free and its operands' barriers have completed;
`tcgen05.commit` completes when the op it follows completes; a warpgroup serves one tile at a time.  LAT is the barrier
round trip (commit -> try_wait wake-up -> first instruction), applied to every hand-over.
"""
import sys

LAT = 60          # clk per barrier hand-over
N_TILES = 400     # tiles per work item in the simulation (steady state dominates)


class Pipe:
    def __init__(self):
        self.free = 0.0      # tensor pipe free at
        self.issue = 0.0     # MMA warp's program counter time
        self.busy = 0.0

    def op(self, dur, deps=()):
        """Issue one MMA (group) after waiting for `deps` (times); returns its completion time."""
        self.issue = max([self.issue] + [d + LAT for d in deps])
        start = max(self.free, self.issue)
        self.free = start + dur
        self.busy += dur
        return self.free


def fwd_default(W, qk=512, pv=512):
    """fmha_fwd_sm100.cu: one 128-column score buffer per Q tile; order QK0 QK1 | PV0 QK0' | PV1 QK1' ..."""
    p = Pipe()
    wg_free = [0.0, 0.0]
    s_full = [p.op(qk), p.op(qk)]
    for j in range(N_TILES):
        for t in (0, 1):
            start = max(wg_free[t], s_full[t] + LAT)
            p_full = start + W
            wg_free[t] = p_full
            p.op(pv, [p_full])
            if j + 1 < N_TILES:
                s_full[t] = p.op(qk)
    return p.busy / p.free


def fwd_bn64(W64, qk=256, pv=256):
    """fmha_fwd_bn64_sm100.cu: two 64-column score stages per Q tile, QK two tiles ahead of PV."""
    p = Pipe()
    wg_free = [0.0, 0.0]
    s_full = {}
    for j in (0, 1):
        for t in (0, 1):
            s_full[(t, j)] = p.op(qk)
    for j in range(N_TILES):
        for t in (0, 1):
            start = max(wg_free[t], s_full[(t, j)] + LAT)
            p_full = start + W64
            wg_free[t] = p_full
            p.op(pv, [p_full])
            if j + 2 < N_TILES:
                s_full[(t, j + 2)] = p.op(qk)
    return p.busy / p.free


def bwd(W, t_dur, a_dur, split):
    """fmha_bwd_sm100.cu: T GEMMs of tile j+2 reuse the TMEM stage of tile j after its accumulate GEMMs.
    default: warpgroup j%2 owns tile j (time W); split: both warpgroups work on every tile (time W/2 + LAT)."""
    p = Pipe()
    wg_free = [0.0, 0.0]
    t_full = {0: p.op(t_dur), 1: p.op(t_dur)}
    for j in range(N_TILES):
        if split:
            start = max(max(wg_free), t_full[j] + LAT)
            p_full = start + W / 2 + LAT
            wg_free = [p_full, p_full]
        else:
            g = j & 1
            start = max(wg_free[g], t_full[j] + LAT)
            p_full = start + W
            wg_free[g] = p_full
        p.op(a_dur, [p_full])
        if j + 2 < N_TILES:
            t_full[j + 2] = p.op(t_dur)
    return p.busy / p.free


def solve(fn, target, lo=100.0, hi=6000.0):
    """element-wise time that reproduces a measured utilisation (utilisation falls monotonically with W)."""
    for _ in range(60):
        mid = 0.5 * (lo + hi)
        if fn(mid) > target:
            lo = mid
        else:
            hi = mid
    return 0.5 * (lo + hi)


def main():
    out = []
    w = out.append
    w("# Pipeline timing model (tools/pipeline_sim.py): measured utilisation -> implied element-wise time -> predictions\n")
    w("Measured inputs (profiles/README.md, burst peak 1729.8 TFLOPS): forward 1097 TFLOPS = 0.63; backward dQ pass 1147")
    w("TFLOPS of MMA work = 0.66; dK/dV pass 1214 = 0.70.  Clock counts per MMA group at D=128: QK / PV 512 (128 columns)")
    w(f"or 256 (64 columns); T GEMM pair 512; accumulate GEMMs 256 (dQ) / 512 (dK/dV).  Barrier hand-over {LAT} clk.\n")
    wf = solve(fwd_default, 0.63)
    wq = solve(lambda x: bwd(x, 512, 256, False), 0.66)
    wk = solve(lambda x: bwd(x, 512, 512, False), 0.70)
    w("| kernel | measured | implied element-wise time per tile |")
    w("|---|---|---|")
    w(f"| forward (128 columns, one warp per sub-partition) | 0.63 | W_f = {wf:.0f} clk |")
    w(f"| backward dQ pass (64 columns) | 0.66 | W = {wq:.0f} clk |")
    w(f"| backward dK/dV pass (64 columns) | 0.70 | W = {wk:.0f} clk |\n")
    w("## Forward: what each opt-in variant buys (tensor-pipe utilisation predicted by the model)\n")
    w("| softmax time per 128 columns | default pipeline | BN64 pipeline (two score stages, W/2 + 100 clk per 64 columns) |")
    w("|---|---|---|")
    for scale, label in [(1.0, "as measured"), (0.8, "-20 % (packed arithmetic, static count -31 % of the hot block)"),
                         (0.7, "-30 % (packed + poly 1/3)"), (0.6, "-40 %")]:
        W = wf * scale
        w(f"| {W:.0f} clk ({label}) | {fwd_default(W):.2f} | {fwd_bn64(W / 2 + 100):.2f} |")
    w("\n## Forward at head_dim 64 (QK / PV take half the clocks, the softmax does not shrink)\n")
    w("Measured: 490 TFLOPS = 0.28 of burst peak at D=64.  The implied softmax time from the D=128 fit predicts")
    w(f"{fwd_default(wf, 256, 256):.2f} for the default pipeline (measured 0.28: the D=64 kernel is even more softmax-bound than the")
    w("latency model says, i.e. throughput-bound on MUFU/issue), and the table shows that only a faster softmax helps there:\n")
    w("| softmax time per 128 columns | default pipeline, D=64 | BN64 pipeline, D=64 |")
    w("|---|---|---|")
    for scale in (1.0, 0.7, 0.5):
        W = wf * scale
        w(f"| {W:.0f} clk | {fwd_default(W, 256, 256):.2f} | {fwd_bn64(W / 2 + 100, 128, 128):.2f} |")
    w("\n## Backward: alternating warpgroups (default) vs both warpgroups on every tile (`kSplit`)\n")
    w("| element-wise time per 64 columns | dQ default | dQ split | dK/dV default | dK/dV split |")
    w("|---|---|---|---|---|")
    for scale, label in [(1.0, "as measured"), (0.75, "-25 % (packed arithmetic)")]:
        a, b = wq * scale, wk * scale
        w(f"| {a:.0f} / {b:.0f} clk ({label}) | {bwd(a, 512, 256, False):.2f} | {bwd(a, 512, 256, True):.2f} | "
          f"{bwd(b, 512, 512, False):.2f} | {bwd(b, 512, 512, True):.2f} |")
    w("\nReading: the default forward is bound by the LATENCY of one tile's softmax (the other tile's MMAs are all it can hide")
    w("behind); BN64 turns that into a THROUGHPUT bound.  The default backward is bound by the T -> element-wise -> accumulate")
    w("chain of a stage; splitting the element-wise work across both warpgroups halves the chain.  The utilisations are")
    w("upper bounds of the model (no issue-slot or MUFU contention between the two warpgroups of a sub-partition).")
    print("\n".join(out))


if __name__ == "__main__":
    sys.exit(main())
