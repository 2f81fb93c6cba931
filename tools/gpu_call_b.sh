#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_wide_golden.py > gpurun_out/r02g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02g_pytest.log
python -m pytest tests/test_wide_golden.py -m gpu -q -s -p no:cacheprovider > gpurun_out/r02g_pytest_wide.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02g_pytest_wide.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err
tail -25 gpurun_out/r02g_pytest.log; grep -E "^\[cuda|conv routes|passed|failed" gpurun_out/r02g_pytest_wide.log | cut -c1-600; tail -3 gpurun_out/r02g_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02g_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ['value','ms_per_step','gpu_launches']}, d['e2e'], d['config']['conv_routes'], d['config']['phase_ms'])
print(d['roofline']['case'], d['roofline']['frac'], d['roofline'].get('ms_per_step_by_case'))
PY
