#!/usr/bin/env python
"""Summarise an .ncu-rep (one row per profiled launch) into the handful of metrics DESIGN.md cites.
Usage: python benchmarks/ncu_summary.py gpurun_out/x.ncu-rep [more.ncu-rep ...]   (runs on the CPU box)"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg"]


def main():
    for rep in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        print(f"# {rep}")
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            print(f"## {d['Kernel Name'][:90]}  grid {d['Grid Size']} block {d['Block Size']}")
            for w in WANT:
                for k in hdr:
                    if k == w or k.endswith("." + w):
                        print(f"  {w} = {d[k]} {units[hdr.index(k)]}")
                        break


if __name__ == "__main__":
    main()
