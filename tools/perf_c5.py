"""C5 development probe: single large fit N=16384 (blocked Cholesky alone + full LML+grad)."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
import torch
from ms_project_b200 import workloads
from ms_project_b200.engine import get_engine
from ms_project_b200.program import ProgramBatch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
X, y, k, noise = workloads.config_c5(n=n)
eng = get_engine(0)
import os
if os.environ.get('PANEL'): eng.set_panel_tiles(int(os.environ['PANEL']))
batch = ProgramBatch([k])
eng.bind(X, y, 1, batch.max_params)
eng.upload_programs(batch)
theta = batch.pack_theta([k.param_array], [noise])
eng.set_profiling(True)
for r in range(2):
    torch.cuda.synchronize(); t0 = time.time()
    lml, dlml, st, tr = eng.lml_grad(batch, theta)
    torch.cuda.synchronize(); dt = time.time() - t0
    ph = eng.last_phase_ms()
    print('lml_grad N=%d wall %.1f ms  TFLOP/s %.2f  phases build %.1f chol %.1f inv %.1f solve %.1f kinv %.1f grad %.1f' % ((n, dt * 1e3, n ** 3 / dt / 1e12) + tuple(ph)),
          'chol TF/s %.2f' % (n ** 3 / 3 / (ph[1] * 1e-3) / 1e12), 'lml', lml[0], 'status', st[0])
# standalone potrf on K_y
Ky = torch.from_numpy(np.zeros(1)).to('cuda')
K = torch.empty((n, n), dtype=torch.float64, device='cuda')
import ctypes
from ms_project_b200 import _lib
rc = eng.lib.b200gp_build_K(eng.ctx, batch.d_tokens.data_ptr(), batch.d_prog_off.data_ptr(), batch.d_theta.data_ptr(),
                            batch.d_theta_off.data_ptr(), batch.d_all_ids.data_ptr(), 1, 1, K.data_ptr())
_lib.check(eng.ctx, rc, 'build_K')
K0 = K.clone()
for r in range(2):
    K.copy_(K0)
    torch.cuda.synchronize(); t0 = time.time()
    ld, st = eng.potrf(K)
    torch.cuda.synchronize(); dt = time.time() - t0
    print('potrf N=%d wall %.1f ms (event %.1f ms) TFLOP/s %.2f logdet %.6f status %d' % (n, dt * 1e3, eng.last_phase_ms()[0], n ** 3 / 3 / dt / 1e12, ld, st))
if n <= 4096:
    ref = 2 * np.sum(np.log(np.diag(np.linalg.cholesky(K0.cpu().numpy()))))
    print('logdet ref', ref, 'rel err', abs(ref - ld) / abs(ref))
