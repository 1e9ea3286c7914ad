"""Turns the ncu outputs a gpurun call left in gpurun_out/ into the small tracked
summaries under profiles/ (launch shares + the key --set full metrics).

  python scripts/summarize_profiles.py <launches_raw.csv> <out_launches.csv> "<comment>"
  python scripts/summarize_profiles.py --full <ncu_raw.csv> <out_metrics.csv> "<comment>"
"""
import csv
import re
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime",
    "sm__pipe_tensor_subpipe_imma_cycles_active_realtime", "sm__inst_executed_pipe_alu", "sm__inst_executed_pipe_fp64",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
    "launch__cluster_size", "launch__cluster_max_active", "gpc__cycles_elapsed.max", "sm__cycles_active.avg",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_xu", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform", "smsp__average_warp_latency_issue_stalled", "sm__pipe_alu_cycles_active",
    "sm__pipe_fp64_cycles_active", "sm__inst_executed_pipe_lsu", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active", "smsp__average_warps_issue_stalled_wait_per_issue_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active", "launch__occupancy_limit_registers", "l1tex__t_sector_hit_rate.pct",
]


def launches(src, dst, comment):
    rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
    hdr = rows[0]
    k_name, k_val = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = OrderedDict()
    for r in rows[1:]:
        name = re.sub(r"\(.*", "", r[k_name]).replace("void ", "").replace("twkb::", "")
        ns = float(r[k_val].replace(",", ""))
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ns
    total = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"# {comment}\n# every launch of the process, cold-cache / serialised under ncu: compare SHARES\n")
        f.write("kernel,launches,total_ms,mean_ms,share\n")
        for name, (n, ns) in agg.items():
            f.write(f"{name},{n},{ns / 1e6:.3f},{ns / 1e6 / n:.3f},{ns / total:.4f}\n")


def full(src, dst, comment):
    rows = list(csv.reader(open(src)))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# {comment}\n")
        for r in rows[2:]:
            f.write(f"# kernel: {r[hdr.index('Kernel Name')]} grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}\n")
            f.write("metric,unit,value\n")
            for h, u, v in zip(hdr, units, r):
                if any(k in h for k in KEYS) and v != "" and ".min." not in h and ".max." not in h and "pipe_lsu.sum" not in h:
                    f.write(f"{h},{u},{v}\n")


if __name__ == "__main__":
    if sys.argv[1] == "--full":
        full(*sys.argv[2:5])
    else:
        launches(*sys.argv[1:4])
