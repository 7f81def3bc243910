#!/usr/bin/env python
"""Markdown summary of an `ncu --set full` report: one column per captured launch (or per kernel name, averaged, with
--by-name), the metrics the roofline contract needs.
  ncu -i gpurun_out/x.ncu-rep --page raw --csv > x_raw.csv ; python tools/ncu_summary.py x_raw.csv [--by-name] [--filter substr]"""
import csv, sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size", "launch__cluster_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]


def short(name):
    name = name.replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("void ", "")
    return name.split("(")[0][:44]


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    by_name = "--by-name" in sys.argv
    filt = sys.argv[sys.argv.index("--filter") + 1] if "--filter" in sys.argv else None
    if filt in args:
        args.remove(filt)
    rows = list(csv.reader(open(args[0])))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    data = [r for r in data if (filt is None or filt in r[ki])]
    cols = []
    if by_name:
        names = []
        for r in data:
            if short(r[ki]) not in names:
                names.append(short(r[ki]))
        for n in names:
            cols.append((n + " (mean of %d)" % sum(short(r[ki]) == n for r in data), [r for r in data if short(r[ki]) == n]))
    else:
        cols = [("%d: %s" % (i, short(r[ki])), [r]) for i, r in enumerate(data)]
    print("| metric | " + " | ".join(c for c, _ in cols) + " |")
    print("|---|" + "---|" * len(cols))
    for m in WANT:
        if m not in hdr:
            continue
        j = hdr.index(m)
        vals = []
        for _, rs in cols:
            try:
                v = sum(float(r[j].replace(",", "")) for r in rs) / len(rs)
                vals.append("%.4g" % v)
            except ValueError:
                vals.append(rs[0][j])
        print("| %s (%s) | " % (m, units[j]) + " | ".join(vals) + " |")


if __name__ == "__main__":
    main()
