#!/usr/bin/env python
"""Turns `ncu -i X.ncu-rep --page raw --csv` output into the per-kernel summary bench.py reads (profiles/r01_traffic.json).

  ncu -i gpurun_out/r01_final.ncu-rep --page raw --csv > raw.csv ; python benchmarks/ncu_summary.py raw.csv C3 > profiles/r01_traffic.json"""
import csv
import json
import re
import sys

WANT = {
    "gpu__time_duration.sum": ("time_ms", 1.0),   # unit normalised below
    "dram__bytes_read.sum": ("dram_read_bytes", 1.0),
    "dram__bytes_write.sum": ("dram_write_bytes", 1.0),
    "smsp__issue_active.avg.pct_of_peak_sustained_active": ("issue_active_pct", 1.0),
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": ("fma_pipe_pct", 1.0),
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": ("dram_pct_of_peak", 1.0),
    "smsp__inst_executed.sum": ("warp_instructions", 1.0),
    "launch__registers_per_thread": ("registers", 1.0),
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": ("shared_wavefronts", 1.0),
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": ("shared_pipe_pct", 1.0),
}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def main():
    rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    out = {"config": sys.argv[2] if len(sys.argv) > 2 else "C3",
           "source": "ncu --set full --clock-control none, one launch each (%s)" % sys.argv[1], "kernels": {}}
    for r in rows[2:]:
        name = re.sub(r"^void ", "", r[ki])
        name = re.sub(r"<unnamed>::", "", name).split("(")[0]
        base = re.sub(r"<.*", "", name)
        k = {"full_name": name}
        for h, u, v in zip(hdr, units, r):
            if h in WANT:
                key, _ = WANT[h]
                try:
                    k[key] = float(v.replace(",", "")) * SCALE.get(u, 1.0)
                except ValueError:
                    pass
        if "dram_read_bytes" in k and "dram_write_bytes" in k:
            k["dram_bytes"] = k["dram_read_bytes"] + k["dram_write_bytes"]
        out["kernels"].setdefault(base, k)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
