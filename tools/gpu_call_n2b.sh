#!/bin/bash
mkdir -p gpurun_out
timeout 110 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --no-cpu-baseline --no-rooflines --steps 16 > gpurun_out/r02l_n2.json 2> gpurun_out/r02l_n2.err; echo "rc=$?"
grep '^{' gpurun_out/r02l_n2.json | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ['value','ms_per_step','scaling']}, d['e2e'] and round(d['e2e']['value'],4), d['config']['nccl_allreduce_alone'], d['config']['phase_ms'])"
tail -2 gpurun_out/r02l_n2.err | cut -c1-200
