#!/usr/bin/env python
"""Summarise an .ncu-rep (one or more kernel launches) into the handful of counters the roofline
argument needs.  Usage: python tools/ncu_summary.py gpurun_out/prof_c4.ncu-rep [...] > profiles/rNN/x.txt"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_lsu.sum", "smsp__inst_executed_op_shared_ld.sum", "sm__cycles_elapsed.max",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
]


def main():
    for path in sys.argv[1:]:
        out = subprocess.check_output(["ncu", "-i", path, "--page", "raw", "--csv"], text=True)
        rows = list(csv.reader(io.StringIO(out)))
        hdr, unit = rows[0], rows[1]
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            print(f"== {path} :: {d.get('Kernel Name', '?')[:90]}")
            for k in WANT:
                if k in d:
                    print(f"  {k:75s} {d[k]:>18s} {unit[hdr.index(k)]}")
            rd, wr = d.get("dram__bytes_read.sum"), d.get("dram__bytes_write.sum")
            print()


if __name__ == "__main__":
    main()
