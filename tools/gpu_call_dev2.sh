#!/bin/bash
o=gpurun_out
mkdir -p $o
timeout 900 python -m pytest tests/test_gpu_modulated.py tests/test_gpu_model.py tests/test_wide_golden.py tests/test_loss_curve.py tests/test_training_step.py -m gpu -q -x -p no:cacheprovider > $o/dev2_pytest.log 2>&1; echo "pytest rc=$?" >> $o/dev2_pytest.log
tail -25 $o/dev2_pytest.log
timeout 400 python bench.py --no-cpu-baseline --no-e2e --no-rooflines > $o/dev2_bench.json 2> $o/dev2_bench.err
tail -3 $o/dev2_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/dev2_bench.json').read().splitlines() if l.startswith('{')][-1])
print({k:d[k] for k in ['value','ms_per_step','gpu_launches']}, d['config']['conv_routes'], d['config']['phase_ms'])
PY
GT_PROFILE_ROWS=400 timeout 300 python tools/profile_phase.py Greg $o/dev2_phase_Greg.txt > /dev/null 2>&1
head -3 $o/dev2_phase_Greg.txt | cut -c1-150
