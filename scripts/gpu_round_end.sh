#!/bin/bash
# Round-end evidence on ONE GPU: full GPU suite, smoke, reference arm, default bench (with cfg 3 / cfg 4), cfg 1 on the GPU,
# decode workload, ncu launch list of the bench command.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/bench_ref.json
timeout 900 python bench.py --gpus 1 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
B="timeout 200 python bench.py --gpus 1"
$B --workload cfg1_feats_normal_b32 > gpurun_out/bench_cfg1.json 2> gpurun_out/bench_cfg1.err; echo "cfg1 rc=$?"
$B --workload cfg3_feats_gmm_cv_b128 --no-extra-configs > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; echo "cfg3 rc=$?"
$B --workload cfg4_finetune_ag_cv_b256 --no-extra-configs > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err; echo "cfg4 rc=$?"
$B --workload cfg5_decode_greedy_beam5 --steps 5 > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err; echo "cfg5 rc=$?"
BENCH="python bench.py --warmup 3 --no-e2e --no-cpu-baseline --no-profile --no-extra-configs"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_cfg2.csv $BENCH --steps 2 > gpurun_out/launches_cfg2.log 2>&1; echo "ncu list rc=$?"; wc -l gpurun_out/launches_cfg2.csv
python - <<'PY'
import json
for f in ("bench_default","bench_cfg1","bench_cfg3","bench_cfg4","bench_cfg5"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read())
        r=d.get("roofline") or {}
        print(f, "value %.0f %s ms/step %.3f e2e %s"%(d["value"],d["unit"],d["ms_per_step"],(d.get("e2e") or {}).get("value")), "roofline", r.get("kernel"), r.get("frac"), "lstm", (d.get("lstm_roofline") or {}).get("frac"), "cpu", (d.get("cpu_baseline") or {}).get("value"), "tensor_frac", d.get("step_tensor_frac"))
        for k,v in (d.get("configs") or {}).items(): print("   ", k, v.get("value"), v.get("ms_per_step"), (v.get("e2e") or {}).get("value"))
    except Exception as e:
        print(f, "ERR", e)
PY
