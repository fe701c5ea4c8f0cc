"""Where do the warps of a kernel spend their time?  Reads the per-instruction source page exported by
tools/ncu_capture.sh (``<name>.source.csv.gz``), splits the SASS stream at synchronisation / TMEM / MMA instructions and
prints the share of warp-stall samples per region, plus the raw metrics that matter for these kernels.
  python tools/ncu_regions.py gpurun_out/r2d/ncu/fwd [min_pct]"""
import csv, gzip, sys

MARKS = ("LDTM", "STTM", "SYNCS", "UTCHMMA", "UTCBAR", "UTMALDG", "BAR", "EXIT", "WARPSYNC", "LDGSTS")
KEYS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum"]


def tables(path):
    rows = list(csv.reader(gzip.open(path, "rt")))
    his = [i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r]
    for t, hi in enumerate(his):
        end = his[t + 1] if t + 1 < len(his) else len(rows)
        ix = {h: i for i, h in enumerate(rows[hi])}
        data = []
        for r in rows[hi + 1:end]:
            if len(r) <= ix["# Samples"]:
                continue
            try:
                c = float(r[ix["# Samples"]])
            except ValueError:
                c = 0.0
            data.append((c, r[ix["Source"]], r[ix["Instructions Executed"]]))
        yield data


def main(base, min_pct=0.5):
    rows = list(csv.reader(open(base + ".raw.csv")))
    hdr = rows[0]
    names = []
    for r in rows[2:]:
        names.append(r[hdr.index("Kernel Name")])
        print("##", names[-1])
        for k in KEYS:
            if k in hdr:
                print(f"   {k} = {r[hdr.index(k)]}")
    seen = set()
    for t, data in enumerate(tables(base + ".source.csv.gz")):
        sig = (len(data), sum(d[0] for d in data))
        if sig in seen:
            continue
        seen.add(sig)
        tot = sum(d[0] for d in data) or 1.0
        print(f"\n== table {t}: {len(data)} instructions, {int(tot)} samples")
        acc, start = 0.0, 0
        for i, (c, s, n) in enumerate(data):
            acc += c
            if any(m in s for m in MARKS):
                if 100 * acc / tot >= min_pct:
                    print(f"{start:5d}-{i:5d} {100 * acc / tot:6.2f}%  ends at: {s.strip()[:72]}  exec={n}")
                acc, start = 0.0, i + 1


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0.5)
