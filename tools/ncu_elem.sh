#!/bin/bash
# ncu --set full of the element-wise kernels (after a plain run of the same command exits 0)
TAG=${1:-elem}
O=gpurun_out
PROBE="python tools/perf_probe.py 2048 32 1"
$PROBE > $O/probe_plain_$TAG.log 2>&1 || { tail -5 $O/probe_plain_$TAG.log; exit 1; }
tail -1 $O/probe_plain_$TAG.log
for k in poly_grad_kernel poly_build_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o $O/prof_${k}_$TAG $PROBE > $O/ncu_${k}_$TAG.log 2>&1
done
ls -la $O | grep $TAG
