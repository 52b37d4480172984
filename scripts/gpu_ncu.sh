#!/bin/bash
# usage: gpu_ncu.sh name:regex:skip[:workload] ...  -- ncu --set full of one launch each (NCU_COUNT consecutive matches when set);
# text summaries into gpurun_out/
set -u
mkdir -p gpurun_out /tmp/ncu
for spec in "$@"; do
  IFS=: read name rx skip wl <<< "$spec"
  wl=${wl:-cfg2_vgg_normal_b256}
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$rx" -s $skip -c ${NCU_COUNT:-1} -f -o /tmp/ncu/$name \
     python bench.py --steps ${NCU_STEPS:-1} --warmup 3 --no-e2e --no-cpu-baseline --no-profile --workload $wl > gpurun_out/ncu_$name.log 2>&1
  echo "ncu $name rc=$?"
  python scripts/ncu_summary.py /tmp/ncu/$name.ncu-rep > gpurun_out/ncu_$name.txt 2>&1
  python scripts/ncu_hot.py /tmp/ncu/$name.ncu-rep 40 > gpurun_out/hot_$name.txt 2>&1
  ncu -i /tmp/ncu/$name.ncu-rep --page raw --csv > gpurun_out/raw_$name.csv 2>&1
  ncu -i /tmp/ncu/$name.ncu-rep --page details > gpurun_out/details_$name.txt 2>&1
  head -3 gpurun_out/ncu_$name.txt; grep -E "gpu__time_duration.sum|dram__bytes_read.sum |dram__bytes_write.sum " gpurun_out/ncu_$name.txt
done
ls -la /tmp/ncu
