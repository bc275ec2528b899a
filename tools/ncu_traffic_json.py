#!/usr/bin/env python
"""profiles/step_kernel_traffic_r02.json from a `--set full` capture of the step kernel: DRAM bytes per launch, read here
on the CPU box:   python tools/ncu_traffic_json.py gpurun_out/prof_step_persist_r02b.ncu-rep profiles/step_kernel_traffic_r02.json"""
import csv
import io
import json
import subprocess
import sys

ALG = 8262173824   # sml_predict_algorithmic_bytes of the headline model (bench.py roofline.algorithmic_bytes_per_launch)


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, last = rows[0], rows[1], rows[-1]
    g = lambda k: (last[hdr.index(k)], units[hdr.index(k)])
    rd, wr = to_bytes(*g("dram__bytes_read.sum")), to_bytes(*g("dram__bytes_write.sum"))
    d = {"kernel": last[hdr.index("Kernel Name")].split("(")[0],
         "source": f"ncu --set full --clock-control none, {rep}, last captured launch of bench.py --steps 3 --warmup 3 --no-train --no-cpu-baseline",
         "dram_bytes_read_per_launch": int(rd), "dram_bytes_write_per_launch": int(wr),
         "traffic_bytes_per_launch": int(rd + wr), "algorithmic_bytes_per_launch": ALG,
         "traffic_over_algorithmic": round((rd + wr) / ALG, 4),
         "gpu_time_duration_ms": float(g("gpu__time_duration.sum")[0]) * {"ms": 1, "us": 1e-3, "ns": 1e-6}.get(g("gpu__time_duration.sum")[1], 1),
         "l2_hit_rate_pct": float(g("lts__t_sector_hit_rate.pct")[0]),
         "registers_per_thread": int(float(g("launch__registers_per_thread")[0])),
         "note": "L2 cache-hint policies on (SML_L2_KEEP default): W_out evict-first, adjacency / W_in / state evict-last; traffic below "
                 "the algorithmic bytes = what stayed in L2 from the previous step"}
    json.dump(d, open(out, "w"), indent=1)
    print(json.dumps(d, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
