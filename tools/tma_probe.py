"""TMA-fed tile GEMMs against the cp.async ones: bit-identical results in every swizzle mode, and phase timings.
    python tools/tma_probe.py [N] [items]"""
import sys
import numpy as np
sys.path.insert(0, '.')
import torch
from ms_project_b200 import workloads
from ms_project_b200.engine import get_engine
from ms_project_b200.program import ProgramBatch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
m = int(sys.argv[2]) if len(sys.argv) > 2 else 256
X, y, kernels = workloads.config_c2(n=n, n_kernels=m)
eng = get_engine(0)
batch = ProgramBatch(kernels)
slots = eng.slots_for(n, 1, m, batch.max_params)
eng.bind(X, y, slots, batch.max_params)
eng.upload_programs(batch)
rng = np.random.default_rng(0)
theta = batch.pack_theta([np.exp(rng.normal(size=int(p)) * 0.3) for p in batch.n_params], np.full(m, 0.1))
ref = None
eng.set_profiling(True)
for mode in (0, 3, 4, 0, 3, 4):
    eng.set_tma(mode)
    best = None
    for r in range(3):
        lml, dlml, st, tr = [a.copy() for a in eng.lml_grad(batch, theta)]
        ph = eng.last_phase_ms()
        if best is None or sum(ph) < sum(best):
            best = ph
    if ref is None:
        ref = (lml, dlml)
    same = np.array_equal(lml, ref[0]) and np.array_equal(dlml, ref[1])
    print('tma mode %d: bit-identical %s  status %s  phases(ms) build %.2f chol %.2f inv %.2f solve %.2f kinv %.2f grad %.2f  dmma %.2f'
          % ((mode, same, np.bincount(st, minlength=4).tolist()) + tuple(best) + (best[1] + best[2] + best[4],)), flush=True)
    if not same:
        d = np.abs(lml - ref[0]) / np.maximum(1, np.abs(ref[0]))
        print('   max rel lml diff %.3g at item %d' % (d.max(), int(d.argmax())), flush=True)
