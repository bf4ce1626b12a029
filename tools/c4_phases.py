import sys, time, numpy as np
sys.path.insert(0, '.')
import torch
from ms_project_b200 import workloads
from ms_project_b200.engine import get_engine
from ms_project_b200.program import ProgramBatch
X, y, kernels = workloads.config_c4()
kernels = kernels[:128]
eng = get_engine(0)
batch = ProgramBatch(kernels)
print('leaves per kernel histogram', np.bincount([len([t for t in k.flattened_parameters()]) for k in kernels])[:30])
eng.bind(X, y, len(kernels), batch.max_params)
eng.upload_programs(batch)
theta = batch.initial_theta(np.full(len(kernels), 0.1))
eng.set_profiling(True)
for r in range(3):
    torch.cuda.synchronize(); t0 = time.time()
    lml, g, st, tr = eng.lml_grad(batch, theta)
    torch.cuda.synchronize(); dt = time.time() - t0
    print('wall %.1f ms' % (dt * 1e3), 'phases build %.1f chol %.1f inv %.1f solve %.1f kinv %.1f grad %.1f' % tuple(eng.last_phase_ms()), np.bincount(st, minlength=4))
