#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_step_gpu.py tests/test_decode_gpu.py tests/test_main_gpu.py tests/test_dp_gpu.py tests/test_finetune_gpu.py -m gpu -q -x > gpurun_out/pytest_d.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_d.log
Q="--steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-extra-configs"
for W in cfg3_feats_gmm_cv_b128 feats_normal_b256; do
  timeout 200 python bench.py --workload $W $Q > gpurun_out/d_$W.json 2> gpurun_out/d_$W.err; echo "$W rc=$?"
done
python - <<'PY'
import json
for f in ("cfg3_feats_gmm_cv_b128","feats_normal_b256"):
    d=json.loads(open("gpurun_out/d_%s.json"%f).read())
    fam=d["families"]
    print(f, "ms/step %.3f"%d["ms_per_step"])
    for k,v in sorted(fam.items(), key=lambda kv:-kv[1]["ms_per_step"])[:22]: print("   %-18s %.3f ms x%.0f %s"%(k,v["ms_per_step"],v["launches_per_step"],("%.0f %s"%(v["achieved"],v["bound"])) if "achieved" in v else ""))
PY
