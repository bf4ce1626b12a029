#!/bin/bash
O=gpurun_out
mkdir -p $O
C3="python tools/bench_configs.py --configs c3 --cpu-budget 0"
B200GP_HPAIR=0 ncu --set full --clock-control none --import-source on -k regex:hellinger_pair_dmma -s 1 -c 1 -f -o $O/prof_hdmma_r02h $C3 > $O/ncu_hdmma_r02h.log 2>&1
tail -3 $O/ncu_hdmma_r02h.log
