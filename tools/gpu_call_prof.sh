#!/bin/bash
# profiling call: ncu launch list of one bench iteration, per-component launch counts, ncu --set full of the hot kernels (CSV pages only:
# the .ncu-rep stays on the box, gpurun_out/ is limited to 64 MiB).  Usage: tools/gpu_call_prof.sh <tag>
tag=${1:-r02p}
o=gpurun_out
mkdir -p $o
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $o/${tag}_launches.csv \
    python bench.py --steps 1 --warmup 17 --no-cpu-baseline --no-e2e --no-rooflines --profile-range > $o/${tag}_ncu_bench.log 2>&1
python tools/launches_summary.py $o/${tag}_launches.csv "ncu launch list of one bench.py iteration (Gmain + Dmain graphs), build $tag" > $o/${tag}_launches_summary.txt
timeout 300 python tools/count_launches.py > $o/${tag}_count_launches.txt 2>&1
python tools/ncu_targets.py > $o/${tag}_targets_plain.log 2>&1 && \
timeout 900 ncu --set full --profile-from-start off --clock-control none -k 'regex:conv_rows|conv_igemm|conv_wgrad|wgrad_reduce|upfirdn2d_tma|bulk|aug_' \
    -f -o /tmp/${tag}_full python tools/ncu_targets.py > $o/${tag}_ncu_full.log 2>&1
ncu -i /tmp/${tag}_full.ncu-rep --page raw --csv > $o/${tag}_full_raw.csv 2>/dev/null
ls -la /tmp/${tag}_full.ncu-rep $o | head -30
head -45 $o/${tag}_launches_summary.txt | cut -c1-200
cat $o/${tag}_count_launches.txt | tail -70
