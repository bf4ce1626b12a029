#!/bin/bash
# round 2, call D: TMA probe (guarded by timeouts: an mbarrier that never completes would hang the GPU), then the updated
# replay / periodic tests and the Hellinger ncu capture.
O=gpurun_out
mkdir -p $O
timeout 180 python tools/tma_probe.py 1024 24 > $O/tma_probe_small.log 2>&1; echo "tma small rc=$?"; cat $O/tma_probe_small.log | tail -8
if grep -q "tma mode 3: bit-identical True" $O/tma_probe_small.log; then
  timeout 300 python tools/tma_probe.py 2048 256 > $O/tma_probe_c2.log 2>&1; echo "tma c2 rc=$?"; tail -7 $O/tma_probe_c2.log
else
  echo "TMA path not verified: switching the default off for the rest of this call"; export B200GP_TMA=0
fi
python -m pytest tests/test_gpu_dropin.py tests/test_gpu_fit.py -q --timeout 1500 > $O/pytest_r02d.log 2>&1; echo "pytest rc=$?"; tail -8 $O/pytest_r02d.log
C3="python tools/bench_configs.py --configs c3 --cpu-budget 0"
$C3 > $O/c3_r02d.log 2>&1; tail -1 $O/c3_r02d.log | cut -c1-400
ncu --set full --clock-control none --import-source on -k regex:hellinger_pair_kernel -s 1 -c 1 -f -o $O/prof_hpair_r02d $C3 > $O/ncu_hpair_r02d.log 2>&1
ls $O | tail -30
