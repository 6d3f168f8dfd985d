"""Summarise ncu outputs into profiles/ (tracked): a launch-list share table and the key raw metrics
of one full capture.   python tools/ncu_summary.py <tag> [launches.csv] [rep.ncu-rep]"""
import collections
import csv
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__instruction_throughput.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__cycles_elapsed.avg.per_second",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]


def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    tot = 0.0
    for r in rows[1:]:
        name = re.sub(r"<.*", "", r[ki]).split("(")[0].replace("void ", "")
        v = float(r[vi].replace(",", ""))
        v = v / 1e6 if r[ui] == "ns" else v / 1e3 if r[ui].startswith("us") else v
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    out.write("| kernel | launches | total ms | share |\n|---|---|---|---|\n")
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.write("| %s | %d | %.3f | %.1f%% |\n" % (k, c, v, 100 * v / tot))
    out.write("\n(gpu__time_duration.sum per launch, --clock-control none; cold-cache and serialised: compare shares)\n\n")


def full(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, zip(units, vals)))
        out.write("### %s\n\n| metric | value | unit |\n|---|---|---|\n" % d.get("Kernel Name", ("", "?"))[1][:120])
        for k in KEYS:
            if k in d:
                out.write("| %s | %s | %s |\n" % (k, d[k][1], d[k][0]))
        out.write("\n")


if __name__ == "__main__":
    tag = sys.argv[1]
    with open("profiles/%s.md" % tag, "w") as out:
        out.write("# ncu summary %s\n\n" % tag)
        if len(sys.argv) > 2 and sys.argv[2] != "-":
            out.write("## launch list (%s)\n\n" % sys.argv[2])
            launches(sys.argv[2], out)
        if len(sys.argv) > 3:
            out.write("## full capture (%s)\n\n" % sys.argv[3])
            full(sys.argv[3], out)
    print(open("profiles/%s.md" % tag).read())
