#!/bin/bash
# checkpoint call: whole GPU test suite + a default-schedule bench line (no CPU baseline / rooflines)
o=gpurun_out
tag=${1:-chk}
mkdir -p $o
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $o/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $o/${tag}_pytest.log
tail -6 $o/${tag}_pytest.log
timeout 400 python bench.py --no-cpu-baseline --no-rooflines > $o/${tag}_bench.json 2> $o/${tag}_bench.err
python - <<PY
import json
d=json.loads([l for l in open('$o/${tag}_bench.json').read().splitlines() if l.startswith('{')][-1])
print({k:d[k] for k in ['value','ms_per_step','gpu_launches']}, d['e2e'], d['config']['conv_routes'], d['config']['phase_ms'])
PY
