#!/bin/bash
# round 2, call B: Hellinger pair kernel variants (A/B) + parity
O=gpurun_out
mkdir -p $O
python -m pytest tests/test_gpu_hellinger.py tests/test_gpu_boms.py tests/test_gpu_vector_distances.py -q > $O/pytest_hell_r02b.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_hell_r02b.log
for V in 0 1 2; do
  B200GP_HPAIR=$V python tools/bench_configs.py --configs c3 --cpu-budget 0 --out $O/configs_c3_r02b_v$V.json > $O/configs_c3_r02b_v$V.log 2>&1
  python -c "import json; d=json.load(open('$O/configs_c3_r02b_v$V.json'))['c3']; print('variant $V pairs_seconds', d['pairs_seconds'], 'pre', d['precompute_seconds'], 'nan', d['nan_distances'])"
done
