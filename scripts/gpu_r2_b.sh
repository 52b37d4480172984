#!/bin/bash
set -u
mkdir -p gpurun_out
Q="--steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-extra-configs"
for V in 2 1; do
  VC_LSTM_SEQ=$V timeout 200 python bench.py --workload feats_normal_b256 $Q > gpurun_out/lstm_v${V}_n1280.json 2> gpurun_out/lstm_v${V}_n1280.err; echo "v$V n1280 rc=$?"
  VC_LSTM_SEQ=$V timeout 200 python bench.py --workload cfg3_feats_gmm_cv_b128 $Q > gpurun_out/lstm_v${V}_n640.json 2> gpurun_out/lstm_v${V}_n640.err; echo "v$V n640 rc=$?"
done
python - <<'PY'
import json
for f in ("lstm_v2_n1280","lstm_v1_n1280","lstm_v2_n640","lstm_v1_n640"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read())
        fam=d["families"]
        print(f, "ms/step %.3f"%d["ms_per_step"], {k:round(fam[k]["ms_per_step"],3) for k in fam if k.startswith("lstm")}, round(d["lstm_roofline"]["frac"],3))
    except Exception as e:
        print(f, "ERR", e)
PY
VC_LSTM_TRACE=1 timeout 200 python bench.py --workload feats_normal_b256 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-profile --no-extra-configs > gpurun_out/trace2_n1280.json 2> gpurun_out/trace2_n1280.err; echo "trace rc=$?"
grep -A4 "lstm trace" gpurun_out/trace2_n1280.err | head -24
VC_LSTM_TRACE=1 timeout 200 python bench.py --workload cfg3_feats_gmm_cv_b128 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-profile --no-extra-configs > gpurun_out/trace2_n640.json 2> gpurun_out/trace2_n640.err; echo "trace rc=$?"
grep -A4 "lstm trace" gpurun_out/trace2_n640.err | head -24
