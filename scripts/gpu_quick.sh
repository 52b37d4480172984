#!/bin/bash
# usage: gpu_quick.sh "<pytest targets>" "<family names, comma separated>"  -- parity subset, then one bench run with the named families
set -u
mkdir -p gpurun_out
timeout -k 5 1200 python -m pytest $1 -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
FAMS="$2" python - <<'PY'
import json, os
d=json.loads(open("gpurun_out/bench_quick.json").read().strip().splitlines()[-1])
print("cfg2 %.3f ms"%d["ms_per_step"], "frac", round(d["roofline"]["frac"],3), {k:round(v["ms_per_step"],3) for k,v in d.get("configs",{}).items()}, d["clocks"])
f=d["families"]
print("   ", {k:(round(f[k]["ms_per_step"],3), f[k]["launches_per_step"]) for k in os.environ["FAMS"].split(",") if k in f})
PY
