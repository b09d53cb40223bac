#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU): python tools/ncu_summary.py gpurun_out/prof.ncu-rep"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_drain_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_selected_per_issue_active.ratio",
        "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_xu.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "smsp__cycles_active.avg", "lts__t_bytes.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__sass_inst_executed_op_shared_ld.sum"]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print(f"## {name[:100]}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"{k:86s} {r[i]:>18s} {units[i]}")
        print()


if __name__ == "__main__":
    main()
