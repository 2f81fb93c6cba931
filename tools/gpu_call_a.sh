#!/bin/bash
# round-2 GPU call A: full GPU test suite, UMMA rate probe, conv microbenchmarks, op sweep vs the reference plugin, bench A/B
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02a_smi.txt 2>&1
python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_wide_golden.py --deselect tests/test_gpu_ref_plugin.py > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
python -m pytest tests/test_wide_golden.py tests/test_gpu_ref_plugin.py -m gpu -q -s -p no:cacheprovider > gpurun_out/r02a_pytest_new.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest_new.log
timeout 120 ./tools/umma_probe > gpurun_out/r02a_umma_probe.txt 2>&1
timeout 300 python tools/test_igemm.py --time > gpurun_out/r02a_igemm_time.txt 2>&1
timeout 300 python tools/test_igemm.py --wgrad --time > gpurun_out/r02a_wgrad_time.txt 2>&1
timeout 900 python tools/op_sweep.py --iters 10 > gpurun_out/r02a_op_sweep.txt 2>&1
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-igemm > gpurun_out/r02a_bench_noigemm.json 2> gpurun_out/r02a_bench_noigemm.err
tail -3 gpurun_out/r02a_pytest.log gpurun_out/r02a_pytest_new.log; cat gpurun_out/r02a_umma_probe.txt; head -c 600 gpurun_out/r02a_bench.json
