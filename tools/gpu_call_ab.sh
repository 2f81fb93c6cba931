#!/bin/bash
# A/B of bench.py switches in ONE call (same box, back to back): tools/gpu_call_ab.sh "<flags A>" "<flags B>" ...
o=gpurun_out
mkdir -p $o
for f in "$@"; do
  timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-rooflines $f 2>/dev/null > $o/ab.json
  python - "$f" <<'PY'
import json, sys
d = json.loads([l for l in open('gpurun_out/ab.json').read().splitlines() if l.startswith('{')][-1])
print(repr(sys.argv[1]), 'ms/step', round(d['ms_per_step'], 3), {k: v['ms'] for k, v in d['config']['phase_ms'].items()})
PY
done
