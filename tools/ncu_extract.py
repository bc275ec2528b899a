#!/usr/bin/env python
"""Reads .ncu-rep captures (brought back in gpurun_out/) on the CPU box and writes the compact per-launch tables
kept under profiles/:   python tools/ncu_extract.py gpurun_out/prof_step_r01b.ncu-rep profiles/step_kernels_r01b.csv"""
import csv
import io
import subprocess
import sys

KEEP = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg.per_second"]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = [(k, hdr.index(k)) for k in KEEP if k in hdr]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([f"{k} [{units[i]}]" if units[i] else k for k, i in idx])
        for r in rows[2:]:
            w.writerow([r[i].split("(")[0] if k == "Kernel Name" else r[i] for k, i in idx])
    print(out, len(rows) - 2, "launches")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
