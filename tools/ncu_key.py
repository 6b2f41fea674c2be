#!/usr/bin/env python3
"""Print the key metrics of an ncu report (raw page): python tools/ncu_key.py REPORT.ncu-rep [kernel-index]"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]


def main():
    rep = sys.argv[1]
    idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, u, v = rows[0], rows[1], rows[2 + idx]
    col = {n: i for i, n in enumerate(h)}
    print("Kernel Name,,%s" % v[col["Kernel Name"]][:110])
    for n in WANT:
        if n in col:
            print("%s,%s,%s" % (n, u[col[n]], v[col[n]]))
    stalls = [(float(v[i] or 0), n) for n, i in col.items()
              if n.startswith("smsp__average_warp") and "issue_stalled" in n and n.endswith("_ratio") or
              n.startswith("smsp__average_warps_issue_stalled") and n.endswith("per_issue_active.ratio")]
    for val, n in sorted(stalls, reverse=True)[:8]:
        print("%s,ratio,%.3f" % (n, val))


if __name__ == "__main__":
    main()
