"""Development aid: where does a lock-step fit (fit.fit_models) spend its time?

  python tools/fit_trace.py [n_candidates] [N]

Prints (1) the latency of one NLML+grad call as a function of the number of active items, (2) the lock-step trace of a
C2 fit: per iteration the active count, the device-call wall time and the host time between calls."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, '.')
import torch  # noqa: E402
from ms_project_b200 import fit as fitmod, workloads  # noqa: E402
from ms_project_b200.engine import get_engine  # noqa: E402
from ms_project_b200.gpy_compat.likelihoods import Gaussian  # noqa: E402
from ms_project_b200.program import ProgramBatch  # noqa: E402


def main():
    n_cand = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
    X, y, kernels = workloads.config_c2(n=n, n_kernels=n_cand)
    eng = get_engine(0)
    R = 3
    batch = ProgramBatch([k for k in kernels for _ in range(R)])
    slots = eng.slots_for(n, 1, batch.n_items, batch.max_params)
    eng.bind(X, y, slots, batch.max_params)
    eng.upload_programs(batch)
    theta = batch.pack_theta([np.asarray(k.param_array) for k in kernels for _ in range(R)], np.full(batch.n_items, 0.1))

    out = {'N': n, 'latency_ms': {}}
    m = 1
    while m <= batch.n_items:
        ids = np.arange(m, dtype=np.int32)
        eng.lml_grad(batch, theta, ids=ids)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 5 if m <= 64 else 2
        for _ in range(reps):
            eng.lml_grad(batch, theta, ids=ids)
        dt = (time.perf_counter() - t0) / reps
        out['latency_ms'][m] = dt * 1e3
        print('active %4d: %.2f ms per call, %.1f evals/s' % (m, dt * 1e3, m / dt), flush=True)
        m *= 2

    # ---- lock-step trace
    noise_prior = workloads.boms_hyperpriors()['GP']['variance']
    liks = []
    for _ in kernels:
        lik = Gaussian()
        lik.variance.set_prior(noise_prior, warning=False)
        liks.append(lik)
    trace = []
    orig = eng.lml_grad
    state = {'t_last': None}

    def timed(b, th, ids=None, with_grad=True):
        t0 = time.perf_counter()
        host = (t0 - state['t_last']) if state['t_last'] is not None else 0.0
        r = orig(b, th, ids=ids, with_grad=with_grad)
        t1 = time.perf_counter()
        state['t_last'] = t1
        trace.append((b.n_items if ids is None else int(len(ids)), (t1 - t0) * 1e3, host * 1e3))
        return r

    eng.lml_grad = timed
    np.random.seed(0)
    t0 = time.perf_counter()
    res = fitmod.fit_models(eng, [k.copy() for k in kernels], liks, num_restarts=R)
    dt = time.perf_counter() - t0
    eng.lml_grad = orig
    tr = np.asarray(trace)
    evals = int(tr[:, 0].sum())
    print('fit: %d candidates x %d restarts, %.2f s, %d evals, %.1f evals/s, %.2f candidates/s' %
          (n_cand, R, dt, evals, evals / dt, n_cand / dt))
    print('iterations %d; device %.2f s; host %.2f s' % (len(tr), tr[:, 1].sum() / 1e3, tr[:, 2].sum() / 1e3))
    edges = [0, 1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 100000]
    for lo, hi in zip(edges[:-1], edges[1:]):
        sel = (tr[:, 0] > lo) & (tr[:, 0] <= hi)
        if sel.any():
            print('  active in (%d, %d]: %4d iterations, device %.2f s, host %.2f s, evals %d' %
                  (lo, hi, sel.sum(), tr[sel, 1].sum() / 1e3, tr[sel, 2].sum() / 1e3, int(tr[sel, 0].sum())))
    ne = np.asarray([r.n_evals for r in res])
    print('evals per candidate: min %d median %d mean %.1f max %d' % (ne.min(), np.median(ne), ne.mean(), ne.max()))
    out['fit'] = {'candidates': n_cand, 'seconds': dt, 'evals': evals, 'iterations': len(tr),
                  'device_s': tr[:, 1].sum() / 1e3, 'host_s': tr[:, 2].sum() / 1e3}
    out['trace'] = tr.tolist()
    json.dump(out, open('gpurun_out/fit_trace.json', 'w'))


main()
