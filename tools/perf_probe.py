"""Quick phase-level timing of the batched LML+grad path (development aid; bench.py is the contract)."""
import sys, time, json
import numpy as np
sys.path.insert(0, '.')
import torch
from ms_project_b200 import workloads
from ms_project_b200.engine import get_engine
from ms_project_b200.program import ProgramBatch

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    m = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    X, y, kernels = workloads.config_c2(n=n, n_kernels=m)
    eng = get_engine(0)
    import os
    if os.environ.get('PANEL'): eng.set_panel_tiles(int(os.environ['PANEL']))
    if os.environ.get('QUIRKS'): eng.set_quirks(int(os.environ['QUIRKS']))
    if os.environ.get('FUSED'): eng.set_fused_grad(int(os.environ['FUSED']))
    if os.environ.get('TILE64'): eng.set_tile64(int(os.environ['TILE64']))
    batch = ProgramBatch(kernels)
    slots = eng.slots_for(n, 1, m, batch.max_params)
    print('slots', slots, 'max_params', batch.max_params, 'tokens', batch.tokens.shape)
    eng.bind(X, y, slots, batch.max_params)
    eng.upload_programs(batch)
    rng = np.random.default_rng(0)
    theta = batch.pack_theta([np.exp(rng.normal(size=int(p)) * 0.3) for p in batch.n_params], np.full(m, 0.1))
    eng.set_profiling(True)
    for r in range(reps):
        torch.cuda.synchronize(); t0 = time.time()
        lml, dlml, st, tr = eng.lml_grad(batch, theta)
        torch.cuda.synchronize(); dt = time.time() - t0
        ph = eng.last_phase_ms()
        print('rep', r, 'wall %.1f ms' % (dt * 1e3), 'evals/s %.1f' % (m / dt), 'TFLOP/s %.2f' % (m * n ** 3 / dt / 1e12),
              'phases(ms) build %.1f chol %.1f inv %.1f solve %.1f kinv %.1f grad %.1f' % tuple(ph), 'status', np.bincount(st, minlength=4))
    eng.set_profiling(False)

main()
