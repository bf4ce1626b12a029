#!/bin/bash
# round 2, call C: updated replay / periodic tests + ncu of the new Hellinger pair kernel
O=gpurun_out
mkdir -p $O
python -m pytest tests/test_gpu_dropin.py tests/test_gpu_fit.py -q --timeout 1500 > $O/pytest_r02c.log 2>&1; echo "pytest rc=$?"; tail -8 $O/pytest_r02c.log
C3="python tools/bench_configs.py --configs c3 --cpu-budget 0"
$C3 > $O/c3_r02c.log 2>&1; tail -1 $O/c3_r02c.log | cut -c1-400
ncu --set full --clock-control none --import-source on -k regex:hellinger_pair_kernel -s 1 -c 1 -f -o $O/prof_hpair_r02c $C3 > $O/ncu_hpair_r02c.log 2>&1
ls $O
