"""Turns ncu artefacts brought back in gpurun_out/ into the small tracked summaries under profiles/.

  python tools/summarize_profiles.py launches gpurun_out/launches_r01.csv profiles/launches_r01.md
  python tools/summarize_profiles.py report   gpurun_out/prof_naive_r01.ncu-rep profiles/naive_accel_r01.txt
"""
import collections
import csv
import re
import subprocess
import sys

KEY_METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.avg", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "l1tex__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__warps_eligible.avg.per_cycle_active",
]


def short_name(name):
    m = re.search(r"(\w+)(<[^(]*>)?\(", name)
    return m.group(1) if m else name[:48]


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e6 if unit == "ns" else v / 1e3 if unit == "us" else v * 1e3 if unit == "s" else v
        a = agg.setdefault(short_name(row["Kernel Name"]), [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write("# ncu launch list summary (`--metrics gpu__time_duration.sum --clock-control none`)\n\n")
        f.write("source: `%s` — per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n\n" % src)
        f.write("total %.1f ms over %d launches\n\n| kernel | launches | total ms | avg ms | share |\n|---|---:|---:|---:|---:|\n" %
                (tot, sum(a[0] for a in agg.values())))
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| %s | %d | %.3f | %.4f | %.2f%% |\n" % (k, c, t, t / c, 100 * t / tot))
    print(open(dst).read())


def report(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write("# ncu --set full --clock-control none: key metrics from %s\n" % src)
        for vals in rows[2:]:
            d = dict(zip(hdr, zip(units, vals)))
            f.write("\n## %s  (grid %s, block %s)\n" % (d.get("Kernel Name", ("", "?"))[1][:110],
                                                      d.get("launch__grid_size", ("", "?"))[1], d.get("launch__block_size", ("", "?"))[1]))
            for k in KEY_METRICS:
                if k in d:
                    f.write("%-90s %12s %s\n" % (k, d[k][1], d[k][0]))
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2], sys.argv[3])
