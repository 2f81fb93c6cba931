#!/bin/bash
# dev call: stride-2 parity-plane wgrad kernel (tests + timing vs the per-tap-row kernel and the library), Greg / Dreg kernel breakdowns
o=gpurun_out
mkdir -p $o
timeout 600 python -m pytest tests/test_gpu_conv_igemm.py -m gpu -q -x -p no:cacheprovider -k "wgrad or f16x3" > $o/dev1_pytest.log 2>&1; echo "pytest rc=$?" >> $o/dev1_pytest.log
tail -15 $o/dev1_pytest.log
timeout 300 python tools/test_igemm.py --wgrad --time > $o/dev1_wgrad_time.txt 2>&1
grep -E "FAIL|wgrad ours" $o/dev1_wgrad_time.txt | sed -e 's/fwd ours.*|| //' | cut -c1-200
GT_PROFILE_ROWS=400 timeout 300 python tools/profile_phase.py Greg $o/dev1_phase_Greg.txt > /dev/null 2>&1
GT_PROFILE_ROWS=400 timeout 300 python tools/profile_phase.py Dreg $o/dev1_phase_Dreg.txt > /dev/null 2>&1
head -60 $o/dev1_phase_Greg.txt | cut -c1-190
