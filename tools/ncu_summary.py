"""Summarise .ncu-rep captures into profiles/*.md (run here on the CPU box: ncu -i ... --page raw --csv)."""
import csv, io, subprocess, sys, collections

KEYS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.avg.per_second"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def source_top(rep, n=12):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = next(i for i, r in enumerate(rows[:10]) if "Source" in r)
    hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
    seen, agg = set(), collections.Counter()
    tot = 0
    for r in rows[hi + 1:]:
        if len(r) <= ix["# Samples"] or r[0] in seen:
            continue
        seen.add(r[0])
        try:
            c = float(r[ix["# Samples"]])
        except ValueError:
            continue
        parts = r[ix["Source"]].split()
        op = (parts[1] if parts and parts[0].startswith("@") and len(parts) > 1 else (parts[0] if parts else "")).split(".")[0]
        agg[op] += c; tot += c
    return [(op, c, 100 * c / max(tot, 1)) for op, c in agg.most_common(n)]


def main(rep, title, out_md):
    hdr, units, rows = raw(rep)
    with open(out_md, "w") as f:
        f.write(f"# {title}\n\nSource: `{rep}` (ncu --set full --clock-control none, B200).  Numbers under a profiler are for\n"
                "attribution only; timing claims come from CUDA-event runs (see bench / gpu_check logs).\n\n")
        for r in rows:
            name = r[hdr.index("Kernel Name")]
            f.write(f"## `{name}`\n\n| metric | value | unit |\n|---|---|---|\n")
            for k in KEYS:
                if k in hdr:
                    f.write(f"| {k} | {r[hdr.index(k)]} | {units[hdr.index(k)]} |\n")
            f.write("\n")
        f.write("## stall samples by opcode (first captured launch)\n\n| opcode | samples | % |\n|---|---|---|\n")
        for op, c, pct in source_top(rep):
            f.write(f"| {op} | {int(c)} | {pct:.1f} |\n")
    print(open(out_md).read()[:3000])


if __name__ == "__main__":
    main(*sys.argv[1:4])
