#!/bin/bash
# cfg 4 family profile + ncu --set full of the LSTM second-form kernels (cfg 2 caption model and cfg 3 shapes) + launch list
set -u
mkdir -p gpurun_out
timeout 300 python bench.py --workload cfg4_finetune_ag_cv_b256 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-extra-configs > gpurun_out/f_cfg4.json 2> gpurun_out/f_cfg4.err; echo "cfg4 rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/f_cfg4.json").read())
fam=d["families"]
print("cfg4 ms/step %.3f"%d["ms_per_step"])
for k,v in sorted(fam.items(), key=lambda kv:-kv[1]["ms_per_step"])[:40]: print("   %-18s %.3f ms x%.0f %s"%(k,v["ms_per_step"],v["launches_per_step"],("%.0f %s"%(v["achieved"],v["bound"])) if "achieved" in v else ""))
PY
# kernel instances: lstm_seq2 fwd kU=64 (encoder = 1st, decoder = 2nd launch of a step), bwd; skip the warm-up steps' launches
timeout 600 bash scripts/gpu_ncu.sh lstm2_fwd_u64:"lstm_seq2_kernel.*Seq2FwdEpi":7:feats_normal_b256 lstm2_bwd_u64:"lstm_seq2_kernel.*Seq2BwdEpi":6:feats_normal_b256 lstm2_fwd_u32:"lstm_seq2_kernel.*Seq2FwdEpi":7:cfg3_feats_gmm_cv_b128 lstm2_bwd_u32:"lstm_seq2_kernel.*Seq2BwdEpi":6:cfg3_feats_gmm_cv_b128
BENCH="python bench.py --warmup 3 --no-e2e --no-cpu-baseline --no-profile --no-extra-configs"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_cfg2.csv $BENCH --steps 2 > gpurun_out/launches_cfg2.log 2>&1; echo "ncu list rc=$?"; wc -l gpurun_out/launches_cfg2.csv
