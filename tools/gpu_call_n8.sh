#!/bin/bash
# BASELINE configs 3 and 4 on one 8 x B200 node (+ the weak-scaling default line)
mkdir -p gpurun_out
run() { tag=$1; n=$2; shift; shift
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $n --no-cpu-baseline --no-rooflines --steps 16 "$@" > gpurun_out/r02m_$tag.out 2> gpurun_out/r02m_$tag.err; echo "$tag rc=$?"
  grep '^{' gpurun_out/r02m_$tag.out | tail -1 > gpurun_out/r02m_$tag.json
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02m_$tag.json').read())
    print('$tag', {k:d[k] for k in ['value','ms_per_step','scaling','n_gpus']}, 'e2e', d['e2e'] and round(d['e2e']['value'],4), 'batch_gpu', d['config']['batch_gpu'], d['config']['nccl_allreduce_alone'], d['config']['phase_ms'], d['clocks'])
except Exception as e:
    print('$tag parse failed', e)
PY
}
run n8_gb64_256 8 --global-batch 64 --no-e2e
run n8_gb64_512 8 --res 512 --global-batch 64 --no-e2e
run n8_weak_256 8
run n4_gb64_256 4 --global-batch 64 --no-e2e
