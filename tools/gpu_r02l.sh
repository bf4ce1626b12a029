#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_distances_reference.py tests/test_gpu_hellinger.py tests/test_gpu_boms.py tests/test_gpu_vector_distances.py -q > $O/pytest_r02l.log 2>&1; echo "pytest rc=$?"; tail -12 $O/pytest_r02l.log
timeout 300 python tools/bench_configs.py --configs c3 --cpu-budget 0 --out $O/configs_c3_r02l.json > $O/configs_c3_r02l.log 2>&1; python -c "import json; d=json.load(open('$O/configs_c3_r02l.json'))['c3']; print('pairs_seconds', d['pairs_seconds'], 'pre', d['precompute_seconds'])"
