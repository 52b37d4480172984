#!/usr/bin/env python
"""Summarise an .ncu-rep (read here on the CPU box): per captured launch, the metrics the roofline needs.
usage: python scripts/ncu_summary.py gpurun_out/prof_X.ncu-rep [> profiles/rNN_X.txt]"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "l1tex__t_bytes.sum", "sm__cycles_active.avg",
        "smsp__cycles_active.avg", "sm__pipe_tensor_subpipe", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("== %s  grid=%s block=%s" % (d["Kernel Name"][:110], d["Grid Size"], d["Block Size"]))
        for i, h in enumerate(hdr):
            if any(k in h for k in KEYS):
                print("   %-75s %s %s" % (h, r[i], units[i]))
        # ncu picks a unit per column (byte / Kbyte / Mbyte / Gbyte): convert both to bytes before adding
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}

        def in_bytes(col):
            if col not in hdr:
                return 0.0
            return float((d.get(col, "0") or "0").replace(",", "")) * scale.get(units[hdr.index(col)], 1.0)

        print("   traffic(dram read+write) = %.3f MB" % ((in_bytes("dram__bytes_read.sum") + in_bytes("dram__bytes_write.sum")) / 1e6))


if __name__ == "__main__":
    main(sys.argv[1])
