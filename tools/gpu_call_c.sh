#!/bin/bash
# A/B: bias_act fused into the discriminator's convolution epilogue
mkdir -p gpurun_out
for f in "" "--fuse-bias-act"; do
  timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-rooflines $f 2>/dev/null > gpurun_out/r02k_ab.json
  python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02k_ab.json').read().strip().splitlines()[-1])
print('fuse_bias_act', d['config']['fuse_bias_act'], 'ms/step', round(d['ms_per_step'],3), d['config']['phase_ms'])
PY
done
