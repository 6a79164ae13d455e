"""Helpers to summarise ncu outputs brought back in gpurun_out/ (run in the build container).

    python profiles/ncu_tools.py launches gpurun_out/launches.csv
    python profiles/ncu_tools.py raw gpurun_out/prof.ncu-rep
    python profiles/ncu_tools.py hot gpurun_out/prof.ncu-rep    # where the instructions go
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum.pct_of_peak_sustained_elapsed",
    "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed_op_shared_atom.sum", "smsp__inst_executed.sum",
    "smsp__cycles_active.avg", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "lts__t_sector_hit_rate.pct",
]


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
        a = agg.setdefault(r[ki].split("(")[0][:70], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(t for _, t in agg.values())
    print("%-72s %5s %12s %10s %6s" % ("kernel", "n", "total_us", "avg_us", "share"))
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-72s %5d %12.1f %10.1f %5.1f%%" % (k, n, t, t / n, 100 * t / tot))


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("---- " + r[hdr.index("Kernel Name")][:90])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print("  %-88s %16s %s" % (k, r[i], units[i]))


def hot(path):
    """Executed-instruction shares of contiguous SASS regions (source page of the FIRST
    kernel in the report): regions are runs of instructions with the same execution count."""
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    print("# " + rows[0][1][:100])
    hdr, data = rows[1], rows[2:]
    isrc, ie, iss = (hdr.index("Source"), hdr.index("Instructions Executed"),
                     hdr.index("# Samples"))
    data = [r for r in data if len(r) > ie and r[ie].isdigit()]
    tot = sum(int(r[ie]) for r in data)
    runs = []
    for i, r in enumerate(data):
        n = int(r[ie])
        if runs and abs(runs[-1][2] - n) <= 0.02 * max(n, 1):
            runs[-1][1], runs[-1][3], runs[-1][4] = i, runs[-1][3] + int(r[iss]), runs[-1][4] + n
        else:
            runs.append([i, i, n, int(r[iss]), n, r[isrc].strip()[:40]])
    print("total warp instructions executed: %.0f M" % (tot / 1e6))
    print("%-11s %4s %10s %6s %8s  %s" % ("sass", "len", "exec(M)", "share", "samples", "first"))
    for a, b, n, smp, t, first in runs:
        if t > tot * 0.004:
            print("%4d-%-6d %4d %10.2f %5.1f%% %8d  %s" % (a, b, b - a + 1, n / 1e6,
                                                        100.0 * t / tot, smp, first))


if __name__ == "__main__":
    {"launches": launches, "raw": raw, "hot": hot}[sys.argv[1]](sys.argv[2])
