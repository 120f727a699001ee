"""Summarise ncu artefacts into small text files that can be committed under profiles/.

  python tools/ncu_summary.py launches <launches.csv> <out.md> [title]
      per-kernel launch count, total / mean duration and share of the captured window from a
      `ncu --metrics gpu__time_duration.sum --csv` launch list
  python tools/ncu_summary.py full <report.ncu-rep> <out.md> [title]
      the roofline-relevant counters of every kernel in a `ncu --set full` report
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio",
    "smsp__average_warp_latency_issue_stalled_barrier.ratio",
    "smsp__average_warp_latency_issue_stalled_mio_throttle.ratio",
    "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio",
    "smsp__average_warp_latency_issue_stalled_wait.ratio",
    "smsp__average_warp_latency_issue_stalled_sleeping.ratio",
]


def launches(path, out, title):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    total = 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        us = v / 1e3 if u.startswith("n") else v * 1e3 if u.startswith("m") else v
        k = row["Kernel Name"].split("(")[0].replace("void ", "")
        a = agg.setdefault(k, [0, 0.0, row["Grid Size"], row["Block Size"]])
        a[0] += 1
        a[1] += us
        total += us
    with open(out, "w") as f:
        f.write(f"# {title}\n\nsource: `{path}` (ncu --metrics gpu__time_duration.sum --clock-control none; per-launch times are "
                "cold-cache and serialised: read the SHARE column)\n\n")
        f.write(f"captured launches: {sum(a[0] for a in agg.values())}, total {total:.1f} us\n\n")
        f.write("| kernel | launches | total us | mean us | share | grid (first) | block |\n|---|---|---|---|---|---|---|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {a[0]} | {a[1]:.1f} | {a[1] / a[0]:.2f} | {100 * a[1] / total:.1f}% | {a[2]} | {a[3]} |\n")


def full(path, out, title):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader([l for l in raw.splitlines() if not l.startswith("==")]))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# {title}\n\nsource: `{path}` (ncu --set full --clock-control none), raw page excerpt\n")
        for row in rows[2:]:
            d = dict(zip(hdr, row))
            f.write(f"\n## `{d.get('Kernel Name', '?')}`  grid {d.get('Grid Size', '?')} block {d.get('Block Size', '?')}\n\n")
            f.write("| metric | unit | value |\n|---|---|---|\n")
            for i, h in enumerate(hdr):
                if h in KEYS and row[i] not in ("",):
                    f.write(f"| {h} | {units[i]} | {row[i]} |\n")


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    title = sys.argv[4] if len(sys.argv) > 4 else src
    (launches if mode == "launches" else full)(src, dst, title)
