"""CPU, world_size 2 over gloo: sharding, record all_gather and the sharded scoring seam (with a stub scorer --
the device path itself is covered by the -m gpu tests)."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import from_string, ROOT

NAMES = ['SE_1', 'RQ_1', 'SE_1 + RQ_1', 'PER_1 * SE_1', 'LIN_1', 'SE_1 * SE_1 + RQ_1 + PER_1', 'RQ_1 * RQ_1']


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _stub_score_fn(models, x, y, scoring_func, optimizer=None, n_restarts=10, engine=None, model_dict=None,
                   all_models=None, shard=None):
    """Stands in for the GPU scorer: 'optimum' = restart draw with the smallest first coordinate."""
    from ms_project_b200.fit import draw_starts, FitResult
    from ms_project_b200.gpy_compat.likelihoods import Gaussian
    for m in all_models:
        if m.likelihood is None:
            m.likelihood = Gaussian()
    starts = draw_starts([m.covariance.raw_kernel for m in all_models], [m.likelihood for m in all_models], n_restarts)
    for m, st in zip(models, starts[shard[0]:shard[1]]):
        best = min(st, key=lambda v: v[0])
        m.covariance.raw_kernel[:] = best[:-1]
        m.likelihood[:] = best[-1:]
        r = FitResult()
        r.lml = -float(np.sum(best))
        r.n_evals = 7
        m.fit_result = r
        m.score = r.lml / 10.0


def _worker(rank, world_size, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import torch.distributed as dist
    from ms_project_b200 import parallel, workloads
    from ms_project_b200.covariance import Covariance
    from ms_project_b200.gp_model import GPModel
    dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=world_size)
    try:
        hp = workloads.boms_hyperpriors()
        models = []
        for s in NAMES:
            k = from_string(s)
            k.traverse(lambda q: [getattr(q, n).set_prior(pr, warning=False) for n, pr in hp[q.op_name].items()]
                       if getattr(q, 'op_name', None) in hp else None)
            models.append(GPModel(Covariance(k)))
        np.random.seed(123)
        x = np.zeros((50, 1))
        scores = parallel.score_models_sharded(models, x, x, None, n_restarts=3, score_fn=_stub_score_fn)
        rec = np.array([[m.score, float(m.likelihood.variance)] + list(m.covariance.raw_kernel.param_array[:1]) for m in models])
        got = parallel.all_gather_records(np.full((rank + 1, 2), float(rank)), [r + 1 for r in range(world_size)])
        np.save(os.path.join(out_dir, 'rank%d.npy' % rank), rec)
        np.save(os.path.join(out_dir, 'gather%d.npy' % rank), got)
        assert all(m.evaluated for m in models) and len(scores) == len(NAMES)
    finally:
        dist.destroy_process_group()


def test_shard_bounds_balance_and_cover():
    from ms_project_b200.parallel import shard_bounds
    b = shard_bounds(np.ones(10), 4)
    assert b[0] == 0 and b[-1] == 10 and all(b[i] <= b[i + 1] for i in range(4))
    assert max(np.diff(b)) - min(np.diff(b)) <= 1
    b = shard_bounds([10, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1], 2)
    assert b == [0, 1, 11]
    assert shard_bounds([1.0], 8)[-1] == 1 and len(shard_bounds([1.0], 8)) == 9
    assert shard_bounds([], 2) == [0, 0, 0]


@pytest.mark.timeout(180)
def test_sharded_scoring_matches_single_process(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / 'rank0.npy'), np.load(tmp_path / 'rank1.npy')
    assert np.array_equal(r0, r1)                              # every rank ends with identical model state
    g0, g1 = np.load(tmp_path / 'gather0.npy'), np.load(tmp_path / 'gather1.npy')
    assert np.array_equal(g0, g1) and g0.shape == (3, 2) and list(g0[:, 0]) == [0.0, 1.0, 1.0]
    # single-process reference of the same stub
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from ms_project_b200 import parallel, workloads
    from ms_project_b200.covariance import Covariance
    from ms_project_b200.gp_model import GPModel
    hp = workloads.boms_hyperpriors()
    models = []
    for s in NAMES:
        k = from_string(s)
        k.traverse(lambda q: [getattr(q, n).set_prior(pr, warning=False) for n, pr in hp[q.op_name].items()]
                   if getattr(q, 'op_name', None) in hp else None)
        models.append(GPModel(Covariance(k)))
    np.random.seed(123)
    x = np.zeros((50, 1))
    parallel.score_models_sharded(models, x, x, None, n_restarts=3, score_fn=_stub_score_fn)
    rec = np.array([[m.score, float(m.likelihood.variance)] + list(m.covariance.raw_kernel.param_array[:1]) for m in models])
    assert np.allclose(rec, r0, rtol=0, atol=0)
