"""Summarise ncu outputs into small text/JSON files for profiles/ (run in the dev container).

    python tools/ncu_summary.py launches gpurun_out/launches_v2.csv  profiles/r01_v2_launches_summary.txt
    python tools/ncu_summary.py full     gpurun_out/prof_v2.ncu-rep  profiles/r01_v2_ncu_full.json
"""
import csv
import json
import subprocess
import sys
from collections import defaultdict

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "sm__cycles_elapsed.max", "sm__cycles_active.avg",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
]


def launches(src, dst):
    rows = list(csv.reader(l for l in open(src) if not l.startswith("==")))
    hdr = rows[0]
    ni, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
        tot[r[ni]] += v * scale
        cnt[r[ni]] += 1
    total = sum(tot.values())
    with open(dst, "w") as fh:
        fh.write(f"# {src}: per-kernel totals over the whole process (ncu, cold-cache, serialised)\n")
        for name, ms in sorted(tot.items(), key=lambda kv: -kv[1]):
            fh.write(f"{name[:90]:90s} n={cnt[name]:5d} total_ms={ms:9.3f} share={ms/total:.3f} "
                     f"avg_us={1e3*ms/cnt[name]:9.1f}\n")


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                d[k] = {"value": r[i], "unit": units[i]}
        out.append(d)
    with open(dst, "w") as fh:
        json.dump({"source": src, "launches": out}, fh, indent=1)


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
