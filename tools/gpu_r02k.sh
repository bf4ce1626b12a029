#!/bin/bash
# round 2, call K: look-ahead Hellinger pair kernel against the one without (C3), its parity tests, dependent-issue latencies
O=gpurun_out
mkdir -p $O
./tools/lat_probe > $O/lat_probe_r02k.txt 2>&1; cat $O/lat_probe_r02k.txt
B200GP_HP_VARIANT=0 timeout 300 python tools/hp_probe.py 3 2>&1 | tail -1 | tee $O/hp_probe_v0.json
B200GP_HP_VARIANT=1 timeout 300 python tools/hp_probe.py 3 2>&1 | tail -1 | tee $O/hp_probe_v1.json
timeout 600 python -m pytest tests/test_gpu_hellinger.py tests/test_gpu_distances_reference.py -x -q -m gpu > $O/pytest_hell_r02k.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_hell_r02k.log
