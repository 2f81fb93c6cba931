#!/bin/bash
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu-baseline --no-rooflines "$@" > gpurun_out/r02j_n2_$tag.json 2> gpurun_out/r02j_n2_$tag.err; echo "$tag rc=$?"; tail -c 1500 gpurun_out/r02j_n2_$tag.json | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in ['value','ms_per_step','scaling']}, d['e2e'] and d['e2e']['value'], d['config']['exchange'][:40], d['config']['nccl_allreduce_alone'], d['config']['phase_ms'])
except Exception as e: print('parse failed', e)
" ; tail -3 gpurun_out/r02j_n2_$tag.err; }
run overlap --steps 16
run nooverlap --steps 16 --no-overlap
run gb64 --steps 16 --global-batch 64
