#!/bin/bash
o=gpurun_out
mkdir -p $o
timeout 600 python -m pytest tests/test_gpu_conv_igemm.py -m gpu -q -x -p no:cacheprovider -k "wgrad or f16x3 or adjoint" > $o/dev3_pytest.log 2>&1; echo "pytest rc=$?" >> $o/dev3_pytest.log
tail -12 $o/dev3_pytest.log
timeout 300 python tools/test_igemm.py --wgrad --time > $o/dev3_wgrad_time.txt 2>&1
grep -E "FAIL|wgrad ours" $o/dev3_wgrad_time.txt | sed -e 's/fwd ours.*|| //' | cut -c1-200
timeout 200 python tools/test_f16x3.py > $o/dev3_f16x3.txt 2>&1; tail -12 $o/dev3_f16x3.txt | cut -c1-220
