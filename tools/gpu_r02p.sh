#!/bin/bash
# round 2, call P: look-ahead diagonal-tile factorisation: parity tests, then timing against the blocked one
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_potrf.py -x -q -m gpu > $O/pytest_potrf_r02p.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_potrf_r02p.log
timeout 600 python tools/potrf_la_probe.py > $O/potrf_la_probe_r02p.txt 2>&1; tail -9 $O/potrf_la_probe_r02p.txt
