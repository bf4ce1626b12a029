#!/bin/bash
# round 2, call G: DMMA Hellinger pair kernel: parity + timing per variant (guarded by timeouts)
O=gpurun_out
mkdir -p $O
for V in 0 2 3 1; do
  B200GP_HPAIR=$V timeout 300 python -m pytest tests/test_gpu_hellinger.py tests/test_gpu_boms.py -q -x > $O/pytest_hell_r02g_v$V.log 2>&1; echo "variant $V pytest rc=$?"; tail -3 $O/pytest_hell_r02g_v$V.log
  B200GP_HPAIR=$V timeout 300 python tools/bench_configs.py --configs c3 --cpu-budget 0 --out $O/configs_c3_r02g_v$V.json > $O/configs_c3_r02g_v$V.log 2>&1
  python -c "import json; d=json.load(open('$O/configs_c3_r02g_v$V.json'))['c3']; print('variant $V pairs_seconds', d['pairs_seconds'], 'pre', d['precompute_seconds'], 'nan', d['nan_distances'], 'bad', d['bad_status'])"
done
