#!/usr/bin/env python
"""Top stall-sampled SASS instructions of an .ncu-rep (source page). usage: ncu_hot.py rep [topN]"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = [i for i, l in enumerate(lines) if l.startswith('"Address"')][0]
rows = list(csv.DictReader(lines[start:]))
tot = sum(int(r["# Samples"]) for r in rows)
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
print("total samples", tot, "instructions", len(rows))
stall_cols = [c for c in rows[0] if c.startswith("stall_") and "Not Issued" not in c]
for idx, r in sorted(enumerate(rows), key=lambda t: -int(t[1]["# Samples"]))[:top]:
    st = sorted(((int(r[c]), c) for c in stall_cols), reverse=True)[:2]
    print("%5d %6.2f%% exec=%-9s %-70s %s" % (idx, 100.0 * int(r["# Samples"]) / max(tot, 1), r["Instructions Executed"], r["Source"].strip()[:70],
                                     " ".join("%s=%d" % (c[6:], n) for n, c in st if n)))
