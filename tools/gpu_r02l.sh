#!/bin/bash
# round 2, call L: ncu --set full with source of the Hellinger pair kernel in use (C3)
O=gpurun_out
mkdir -p $O
timeout 300 python tools/hp_probe.py 1 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hellinger_pair -s 1 -c 1 -f -o $O/prof_hpair_r02l python tools/hp_probe.py 1 > $O/ncu_hpair_r02l.log 2>&1; echo "ncu rc=$?"; tail -2 $O/ncu_hpair_r02l.log
