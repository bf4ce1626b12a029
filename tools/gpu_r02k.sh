#!/bin/bash
# round 2, call K: C3 Hellinger pair timing (tools/hp_probe.py) and the Hellinger / distance parity tests.
# (During development the loop below ran once per kernel variant through an environment switch of the library:
# profiles/experiments/hellinger_la_r02.md.)
O=gpurun_out
mkdir -p $O
timeout 300 python tools/hp_probe.py 3 2>&1 | tail -1 | tee $O/hp_probe.json
timeout 600 python -m pytest tests/test_gpu_hellinger.py tests/test_gpu_distances_reference.py -x -q -m gpu > $O/pytest_hell_r02k.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_hell_r02k.log
