#!/bin/bash
# evidence call: GPU test suite, default bench line, reference arm, no-igemm A/B, 512x512 line, conv microbenchmarks, ncu launch list of the
# bench command, ncu --set full of the hot kernels (the .ncu-rep stays on the box: gpurun_out/ is limited to 64 MiB; its raw page travels as CSV).
# Usage: tools/gpu_call_ev.sh <tag>
tag=${1:-r02z}
o=gpurun_out
mkdir -p $o
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $o/${tag}_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > $o/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $o/${tag}_pytest.log
timeout 600 python bench.py > $o/${tag}_bench.json 2> $o/${tag}_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $o/${tag}_bench_ref.json 2> $o/${tag}_bench_ref.err
timeout 400 python bench.py --no-cpu-baseline --no-e2e --no-rooflines --no-igemm > $o/${tag}_bench_noigemm.json 2> $o/${tag}_bench_noigemm.err
timeout 400 python bench.py --no-cpu-baseline --no-rooflines --res 512 --batch-gpu 16 > $o/${tag}_bench_512.json 2> $o/${tag}_bench_512.err
timeout 300 python tools/test_igemm.py --wgrad --time > $o/${tag}_conv_microbench.txt 2>&1
timeout 300 python tools/test_f16x3.py > $o/${tag}_f16x3.txt 2>&1
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $o/${tag}_launches.csv \
    python bench.py --steps 1 --warmup 17 --no-cpu-baseline --no-e2e --no-rooflines --profile-range > $o/${tag}_ncu_bench.log 2>&1
python tools/launches_summary.py $o/${tag}_launches.csv "ncu launch list of one bench.py iteration (Gmain + Dmain graphs), build $tag: ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv python bench.py --steps 1 --warmup 17 --no-cpu-baseline --no-e2e --no-rooflines --profile-range" > $o/${tag}_launches_summary.txt
python tools/ncu_targets.py > $o/${tag}_targets_plain.log 2>&1 && \
timeout 900 ncu --set full --profile-from-start off --clock-control none -k 'regex:conv_rows|conv_igemm|conv_wgrad|wgrad_reduce|upfirdn2d_tma|bulk|aug_' \
    -f -o /tmp/${tag}_full python tools/ncu_targets.py > $o/${tag}_ncu_full.log 2>&1
ncu -i /tmp/${tag}_full.ncu-rep --page raw --csv > $o/${tag}_full_raw.csv 2>/dev/null
tail -4 $o/${tag}_pytest.log; tail -2 $o/${tag}_bench.err
python - <<PY
import json
for f in ['bench', 'bench_noigemm', 'bench_ref', 'bench_512']:
    try:
        d = json.loads([l for l in open('$o/${tag}_%s.json' % f).read().splitlines() if l.startswith('{')][-1])
        print(f, {k: d.get(k) for k in ['value', 'ms_per_step', 'gpu_launches', 'e2e']}, (d.get('config') or {}).get('conv_routes'), (d.get('config') or {}).get('phase_ms'))
        if d.get('roofline'):
            print('  roofline', {k: d['roofline'].get(k) for k in ['kernel', 'case', 'achieved', 'peak', 'frac', 'traffic']})
    except Exception as e:
        print(f, 'parse failed', e)
PY
grep -E "FAIL|wgrad ours" $o/${tag}_conv_microbench.txt | cut -c1-330
du -sh $o
