#!/usr/bin/env python
"""Summarise ncu outputs into small tracked files under profiles/.

  python scripts/summarize_ncu.py launches gpurun_out/launches_r01.csv profiles/r01_launches_summary.csv
  python scripts/summarize_ncu.py full gpurun_out/prof_r01.ncu-rep profiles/r01_kernels_full.csv
"""
import csv
import re
import subprocess
import sys
from collections import OrderedDict, defaultdict

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_tag_requests.max.pct_of_peak_sustained_elapsed",
    "lts__t_requests_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
    "lts__t_requests_srcunit_tex_op_red.sum", "lts__t_requests_srcunit_tex_op_atom_dot_alu.sum",
    "lts__t_sectors_srcunit_tex_op_atom_dot_cas.sum", "lts__t_requests_srcunit_tex_op_write.sum",
    "lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__average_warp_latency_per_inst_issued.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__inst_executed_op_shared_atom.sum",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
]


def short(name):
    return re.sub(r"\(.*", "", name).replace("void ", "").strip()


def launches(src, dst):
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = OrderedDict()
    total = 0.0
    for r in rows[1:]:
        ns = float(r[vi].replace(",", ""))
        k = short(r[ki])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ns
        total += ns
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_ms", "avg_us", "share_pct"])
        for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k, n, f"{ns / 1e6:.3f}", f"{ns / n / 1e3:.1f}", f"{100 * ns / total:.1f}"])
        w.writerow(["TOTAL", sum(v[0] for v in agg.values()), f"{total / 1e6:.3f}", "", "100.0"])


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [(m, hdr.index(m)) for m in METRICS if m in hdr]
    ki = hdr.index("Kernel Name")
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + [f"{m} [{units[i]}]" for m, i in cols])
        for r in rows[2:]:
            w.writerow([short(r[ki])] + [r[i] for _, i in cols])


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
