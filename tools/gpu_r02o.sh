#!/bin/bash
# round 2, call O: compute-sanitizer (racecheck, memcheck) over the Hellinger pair kernels at every size class; C4 phases
O=gpurun_out
mkdir -p $O
cat > /tmp/hp_small.py <<'PY'
import sys
sys.path.insert(0, '.')
import numpy as np
from tests.test_gpu_hellinger import _setup
for n in (17, 70, 90, 100, 110, 120):
    rng = np.random.default_rng(n)
    x = rng.random((n, 3)); x = (x - x.mean(0)) / x.std(0)
    am, builder, prob, noise_prior, pool = _setup(x, 4, 2, ['SE', 'RQ'], 3)
    for i in range(len(pool)):
        builder.compute_distance(am, [i], list(range(i + 1, len(pool))))
    D = builder.get_kernel(len(pool))
    print(n, float(np.nansum(D)), int((builder.last_status != 0).sum()))
PY
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python /tmp/hp_small.py > $O/sanitizer_race_r02o.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|ERROR SUMMARY|hazard" $O/sanitizer_race_r02o.log | sort | uniq -c | head
timeout 900 compute-sanitizer --tool memcheck python /tmp/hp_small.py > $O/sanitizer_mem_r02o.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY" $O/sanitizer_mem_r02o.log | head -3
timeout 300 python tools/c4_phases.py > $O/c4_phases_r02o.log 2>&1; tail -4 $O/c4_phases_r02o.log
