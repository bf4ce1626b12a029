#!/bin/bash
# round 2, call Q: ncu --set full with source of the look-ahead diagonal-tile kernel (32 C2 kernels, N = 2048)
O=gpurun_out
mkdir -p $O
PROBE="python tools/perf_probe.py 2048 32 1"
$PROBE > $O/probe_plain_r02q.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:potrf_diag_la -s 3 -c 1 -f -o $O/prof_potrf_r02q $PROBE > $O/ncu_potrf_r02q.log 2>&1; echo "ncu rc=$?"
