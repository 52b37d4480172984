#!/bin/bash
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for BF in 0 1; do
VC_GRAD_BF16=$BF timeout 600 $TR --master-port 29520 scripts/dp_two_ranks.py > gpurun_out/dp_two_ranks_bf$BF.json 2> gpurun_out/dp_two_ranks_bf$BF.err; echo "dp_two_ranks bf16=$BF rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/dp_two_ranks_bf$BF.json'));print(d['dp_two_ranks'],{k:(v['grad_vs_alias_sum_max_rel'],v['crc_equal'],v['bytes']) for k,v in d['cases'].items()})"
VC_GRAD_BF16=$BF timeout 300 $TR --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 5 --workload cfg3_feats_gmm_cv_b128 --no-cpu-baseline --no-profile > gpurun_out/g_cfg3_n2_bf$BF.json 2> gpurun_out/g_cfg3_n2_bf$BF.err; python -c "
import json;d=json.load(open('gpurun_out/g_cfg3_n2_bf$BF.json'));a=d['allreduce'];print('cfg3 n2 bf16=$BF ms/step %.3f no_ar %.3f exposed %.3f unbucketed %.3f'%(d['ms_per_step'],a['ms_per_step_no_allreduce'],a['exposed_ms'],a['exposed_ms_unbucketed']),d['dp_check'],a['transport'],a['bytes'])"
done
