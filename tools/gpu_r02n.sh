#!/bin/bash
# round 2, call N: ncu --set full with source of the <= 4-leaf gradient and build kernels (32 C2 kernels, N = 2048)
O=gpurun_out
mkdir -p $O
PROBE="python tools/perf_probe.py 2048 32 1"
$PROBE > $O/probe_plain_r02n.log 2>&1 || exit 1
for k in grad build; do
ncu --set full --clock-control none --import-source on -k regex:poly_${k}_kernel -s 0 -c 1 -f -o $O/prof_${k}_r02n $PROBE > $O/ncu_${k}_r02n.log 2>&1; echo "ncu $k rc=$?"
done
