#!/bin/bash
# round 2, call K: Hellinger pair kernel variants on C3 (0 = no look-ahead, 1 = look-ahead 16 x 8, 2 = look-ahead 8 x 8) and their parity tests
O=gpurun_out
mkdir -p $O
for v in ${VARIANTS:-1 2}; do
B200GP_HP_VARIANT=$v timeout 300 python tools/hp_probe.py 3 2>&1 | tail -1 | tee $O/hp_probe_v$v.json
done
B200GP_HP_VARIANT=${TESTV:-2} timeout 600 python -m pytest tests/test_gpu_hellinger.py tests/test_gpu_distances_reference.py -x -q -m gpu > $O/pytest_hell_r02k.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_hell_r02k.log
