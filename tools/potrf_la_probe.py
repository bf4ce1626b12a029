"""Diagonal-tile factorisation, look-ahead (b200gp_set_potrf_la 1) against blocked (0): latency of one NLML+grad evaluation of
small batches at N = 512 / 2048, the C2 bulk phases, and the N = 16384 potrf (C5)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from ms_project_b200 import workloads
from ms_project_b200.engine import get_engine
from ms_project_b200.program import ProgramBatch

eng = get_engine(0)
out = {}
for N in (512, 2048):
    X, y, kernels = workloads.config_c2(n=N, n_kernels=64)
    for items in (1, 8, 64):
        batch = ProgramBatch(kernels[:items])
        eng.bind(X, y, items, batch.max_params)
        eng.upload_programs(batch)
        theta = batch.initial_theta(np.full(items, 0.1))
        res = {}
        for mode in (0, 1, 2):
            eng.set_potrf_la(mode)
            for _ in range(3):
                lml, g, st, tr = eng.lml_grad(batch, theta)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(20):
                lml, g, st, tr = eng.lml_grad(batch, theta)
            torch.cuda.synchronize()
            res[mode] = ((time.perf_counter() - t0) / 20 * 1e3, float(np.sum(lml)), int(np.sum(st != 0)))
        out['N%d_items%d' % (N, items)] = res
        print(N, items, 'blocked %.3f ms  look-ahead %.3f ms  split barrier %.3f ms' % (res[0][0], res[1][0], res[2][0]), 'lml rel diff %.2e %.2e' % (abs(res[0][1] - res[1][1]) / abs(res[0][1]), abs(res[2][1] - res[1][1]) / abs(res[1][1])), res[0][2], res[1][2], res[2][2])
# C5 potrf
n = 16384
rng = np.random.default_rng(0)
x = np.sort(rng.uniform(0, 50, n))
A0 = torch.from_numpy(np.exp(-0.5 * (x[:, None] - x[None, :]) ** 2) + 0.1 * np.eye(n)).cuda()
for mode in (0, 1, 2):
    eng.set_potrf_la(mode)
    ts = []
    for r in range(4):
        A = A0.clone()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ld, st = eng.potrf(A)
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    print('C5 potrf mode', mode, ['%.2f' % t for t in ts], 'logdet', ld, st)
eng.set_potrf_la(1)
