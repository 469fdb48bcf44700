#!/usr/bin/env python
"""Turn ncu outputs (gpurun_out/*.ncu-rep, launch-list csv) into the text summaries kept under profiles/."""
import csv, subprocess, sys, collections, io

def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__maximum_warps_per_active_cycle_pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__waves_per_multiprocessor", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]

def summarize_rep(rep):
    hdr, units, rows = raw(rep)
    lines = ["# ncu --set full --clock-control none : %s" % rep]
    for r in rows:
        lines.append("kernel: %s" % r[hdr.index("Kernel Name")])
        for w in WANT:
            if w in hdr:
                lines.append("  %-82s %s %s" % (w, r[hdr.index(w)], units[hdr.index(w)]))
    return "\n".join(lines) + "\n"

def summarize_launches(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot, cnt = collections.defaultdict(float), collections.defaultdict(int)
    for r in rows[1:]:
        name = r[ki].split("(")[0][:70]
        tot[name] += float(r[vi].replace(",", ""))
        cnt[name] += 1
    s = sum(tot.values())
    lines = ["# ncu --metrics gpu__time_duration.sum --clock-control none : %s (cold-cache, serialised: compare shares)" % path]
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        lines.append("%-72s n=%4d  %12.1f us  %5.1f%%" % (k, cnt[k], v / 1e3, 100 * v / s))
    return "\n".join(lines) + "\n"

if __name__ == "__main__":
    for a in sys.argv[1:]:
        print(summarize_launches(a) if a.endswith(".csv") else summarize_rep(a))
