#!/bin/bash
# round 2, call M: element-wise kernels after a change: parity tests of the scoring path + the bench phases
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_lml.py tests/test_gpu_abi.py tests/test_gpu_fullsize.py -x -q -m gpu > $O/pytest_elem_r02m.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_elem_r02m.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-fit > $O/bench_r02m.json 2> $O/bench_r02m.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_r02m.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'phases', {k: round(v, 2) for k, v in d['phases_ms'].items()})
PY
