#!/usr/bin/env python
"""Turn the scratch ncu outputs under gpurun_out/ into the committed summaries under profiles/.
  python tools/ncu_summaries.py launches gpurun_out/launches_r1.csv profiles/launches_r1_summary.md "<command>"
  python tools/ncu_summaries.py raw /tmp/raw.csv   (ncu -i x.ncu-rep --page raw --csv > /tmp/raw.csv) -> metric table on stdout
"""
import collections
import csv
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum",
    "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
]


def launches(src, dst, command):
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows:
        name = r[kn].split("(")[0].replace("void ", "")
        n, t = agg.get(name, (0, 0.0))
        agg[name] = (n + 1, t + float(r[mv]) / 1e6)
    total = sum(t for _, t in agg.values())
    with open(dst, "w") as f:
        f.write("# ncu launch list (gpu__time_duration.sum, --clock-control none), `%s`\n\n" % command)
        f.write("Raw list: `profiles/%s` (warm-up step + timed step). Serialised, cold-cache times: shares, not absolutes.\n\n" % src.split("/")[-1])
        f.write("| kernel | launches | total ms | share |\n|---|---|---|---|\n")
        for name, (n, t) in agg.items():
            f.write("| `%s` | %d | %.3f | %.1f %% |\n" % (name, n, t, 100 * t / total))


def raw(src):
    rows = list(csv.reader(open(src)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    for m in METRICS:
        if m in hdr:
            i = hdr.index(m)
            print("| `%s` | %s |" % (m, " | ".join((d[i] + " " + units[i]).strip() for d in data)))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        raw(sys.argv[2])
