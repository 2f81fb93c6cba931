#!/bin/bash
# last call of the round: whole GPU suite, smoke(), default bench line of the committed build
o=gpurun_out
mkdir -p $o
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > $o/fin_pytest.log 2>&1; echo "pytest rc=$?" >> $o/fin_pytest.log
tail -3 $o/fin_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $o/fin_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $o/fin_smoke.log
timeout 600 python bench.py > $o/fin_bench.json 2> $o/fin_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/fin_bench.json').read().splitlines() if l.startswith('{')][-1])
print({k:d[k] for k in ['value','ms_per_step','gpu_launches']}, d['e2e'], d['config']['conv_routes'], d['config']['phase_ms'], d['clocks'])
print({k:d['roofline'][k] for k in ['kernel','frac','achieved','traffic']}, d['cpu_baseline'])
PY
