"""Turn ncu outputs brought back in gpurun_out/ into the small tracked summaries under profiles/.

    python profiles/summarize.py launches gpurun_out/launches_c3.csv profiles/r01_launches_c3.txt
    python profiles/summarize.py full gpurun_out/prof_c3_force.ncu-rep profiles/r01_ncu_c3_force.txt
"""
import collections
import csv
import subprocess
import sys

FULL_METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]


def launches(src, dst):
    with open(src) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    d = collections.defaultdict(list)
    for x in csv.DictReader(lines):
        if x.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(x["Metric Value"].replace(",", ""))
        val = {"ns": val / 1e3, "us": val, "ms": val * 1e3}.get(x["Metric Unit"], val)
        d[x["Kernel Name"].split("(")[0]].append(val)
    tot = sum(sum(v) for v in d.values())
    with open(dst, "w") as out:
        out.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none ; source: {src}\n")
        out.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
        out.write(f"{'kernel':58s} {'launches':>8s} {'mean_us':>10s} {'total_us':>12s} {'share':>7s}\n")
        for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
            out.write(f"{k[:58]:58s} {len(v):8d} {sum(v) / len(v):10.2f} {sum(v):12.1f} {100 * sum(v) / tot:6.1f}%\n")
    print(open(dst).read())


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open(dst, "w") as out:
        out.write(f"# ncu --set full --clock-control none ; source: {src}\n")
        for r in data:
            out.write(f"\n== {r[hdr.index('Kernel Name')][:100]}\n")
            for m in FULL_METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    out.write(f"{m:90s} {r[i]:>18s} {units[i]}\n")
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
