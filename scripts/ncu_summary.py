#!/usr/bin/env python
"""Key counters of every launch in an ncu report (raw page), as a compact table for profiles/."""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "dur"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("lts__t_bytes.sum", "l2_bytes"),
    ("l1tex__t_bytes.sum", "l1_bytes"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fmacyc%"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64%"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("smsp__thread_inst_executed.sum", "thread_inst"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dsmem"),
    ("launch__occupancy_limit_shared_mem", "lim_smem"),
    ("launch__occupancy_limit_registers", "lim_regs"),
    ("launch__waves_per_multiprocessor", "waves"),
    ("smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "st_long"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stl_long"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stl_short"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stl_barrier"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stl_wait"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stl_math"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stl_mio"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stl_branch"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stl_noinst"),
    ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "stl_dispatch"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst"),
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("kernel:", r[hdr.index("Kernel Name")][:60])
        for k, short in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {short:14s} {r[i]:>16s} {units[i]}")


if __name__ == "__main__":
    main()
