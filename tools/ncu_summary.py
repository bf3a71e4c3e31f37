#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): python tools/ncu_summary.py rep [out.json]"""
import csv, json, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__maximum_warps_per_active_cycle_pct",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active",
        "smsp__average_warps_issue_stalled_wait_per_issue_active", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "lts__t_bytes.sum", "sm__cycles_active.avg"]


def main():
    rep = sys.argv[1]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    out = []
    for row in rows[2:]:
        d = {"kernel": row[hdr.index("Kernel Name")][:60], "grid": row[hdr.index("Grid Size")], "block": row[hdr.index("Block Size")]}
        for k in KEYS:
            for i, h in enumerate(hdr):
                if h == k or h.endswith("." + k) or h.endswith(k):
                    d[k] = f"{row[i]} {units[i]}".strip()
                    break
        out.append(d)
    js = json.dumps(out, indent=1)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(js)
    print(js)


if __name__ == "__main__":
    main()
