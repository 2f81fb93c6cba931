#!/bin/bash
o=gpurun_out
mkdir -p $o
timeout 600 python -m pytest tests/test_gpu_conv_igemm.py -m gpu -q -x -p no:cacheprovider -k "fused_epilogue or forward" > $o/dev4_pytest.log 2>&1; echo "pytest rc=$?" >> $o/dev4_pytest.log
tail -5 $o/dev4_pytest.log
python tools/bench_fused_epilogue.py 32 > $o/dev4_fused_epilogue.txt 2>&1; cat $o/dev4_fused_epilogue.txt | tail -11
for f in "--no-fuse-bias-act" ""; do
  timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-rooflines $f 2>/dev/null > $o/dev4_ab.json
  python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/dev4_ab.json').read().splitlines() if l.startswith('{')][-1])
print('fuse_bias_act', d['config']['fuse_bias_act'], 'ms/step', round(d['ms_per_step'],3), d['config']['phase_ms'])
PY
done
