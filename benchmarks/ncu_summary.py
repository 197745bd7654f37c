#!/usr/bin/env python
"""Summarise an .ncu-rep (one row per profiled launch) into the handful of metrics DESIGN.md cites.
Usage: python benchmarks/ncu_summary.py gpurun_out/x.ncu-rep [more.ncu-rep ...]   (runs on the CPU box)
       python benchmarks/ncu_summary.py --launches raw.csv clean.csv [frames]
           raw.csv = `ncu --metrics gpu__time_duration.sum --csv --log-file raw.csv ...`; writes clean.csv without the
           ==PROF== / ==WARNING== banner lines (valid CSV) and prints per-kernel totals and shares (per frame if given)."""
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


def launches(raw: str, clean: str, frames: int = 0):
    lines = [ln for ln in open(raw, errors="replace").read().splitlines() if ln.strip() and not ln.startswith("==")]
    rows = list(csv.reader(lines))
    hdr = rows[0]
    with open(clean, "w", newline="") as f:
        csv.writer(f, quoting=csv.QUOTE_ALL).writerows(rows)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = {}
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
        name = r[ki].split("(")[0].replace("void ", "").replace("moyolo::", "")[:70]
        n, t = tot.get(name, (0, 0.0))
        tot[name] = (n + 1, t + v)
    total = sum(t for _, t in tot.values())
    count = sum(n for n, _ in tot.values())
    per = f", {count / frames:.1f} launches and {total / frames:.1f} us per frame" if frames else ""
    print(f"# {raw}: {count} launches, {total:.1f} us summed kernel time (cold-cache, serialised){per}")
    print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
    for name, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"| {name} | {n} | {t:.1f} | {100 * t / total:.1f} % |")


def main():
    if len(sys.argv) >= 4 and sys.argv[1] == "--launches":
        launches(sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 0)
        return
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
