"""CPU, world_size 2 over gloo: shard assignment, the record all_gather, and the REAL sharded scoring seam
(gp_model.score_models with a Comm) with the oracle standing in for the device (tests/oracle_engine.py) -- the device
path itself is covered by the -m gpu tests."""
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


def _worker(rank, world_size, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import torch.distributed as dist
    from ms_project_b200 import parallel
    dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=world_size)
    try:
        got = parallel.all_gather_records(np.full((rank + 1, 2), float(rank)), [r + 1 for r in range(world_size)])
        comm = parallel.Comm()
        union = comm.union([5, 1] if rank == 1 else [])
        mine = comm.shard([10, 11, 12, 13, 14], [5.0, 1.0, 1.0, 1.0, 2.0])
        np.save(os.path.join(out_dir, 'gather%d.npy' % rank), got)
        np.save(os.path.join(out_dir, 'union%d.npy' % rank), np.array(union))
        np.save(os.path.join(out_dir, 'mine%d.npy' % rank), np.array(mine))
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


def test_lpt_assignment_balances_cost():
    from ms_project_b200.parallel import lpt_assign
    owner = lpt_assign([5, 1, 1, 1, 2], 2)
    assert list(owner) == [0, 1, 1, 1, 1]
    costs = np.random.RandomState(0).rand(64) ** 3
    owner = lpt_assign(costs, 8)
    load = np.array([costs[owner == r].sum() for r in range(8)])
    assert load.max() <= 1.15 * load.mean()
    assert list(lpt_assign([], 3)) == []


@pytest.mark.timeout(180)
def test_record_exchange_primitives(tmp_path):
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    g0, g1 = np.load(tmp_path / 'gather0.npy'), np.load(tmp_path / 'gather1.npy')
    assert np.array_equal(g0, g1) and g0.shape == (3, 2) and list(g0[:, 0]) == [0.0, 1.0, 1.0]
    assert list(np.load(tmp_path / 'union0.npy')) == [1, 5] == list(np.load(tmp_path / 'union1.npy'))
    assert list(np.load(tmp_path / 'mine0.npy')) == [10] and list(np.load(tmp_path / 'mine1.npy')) == [11, 12, 13, 14]


# ---- the REAL scoring seam (gp_model.score_models: lock-step fit, retry-once, write-back) sharded over two gloo ranks,
#      with the oracle standing in for the device (tests/oracle_engine.py) ------------------------------------------------
FIT_NAMES = ['SE_1', 'RQ_1', 'SE_1 + RQ_1', 'LIN_1 + SE_1', 'SE_1 * RQ_1', 'LIN_1']
FAIL_ONCE = 'SE_1 + RQ_1'        # its first fit reports LML = -inf => retry-once on the rank that owns it


def _fit_problem():
    from ms_project_b200 import workloads
    from ms_project_b200.covariance import Covariance
    from ms_project_b200.gp_model import GPModel
    from ms_project_b200.gp_models import gp_regression
    hp = workloads.boms_hyperpriors()
    models = []
    for s in FIT_NAMES:
        k = from_string(s)
        k.traverse(lambda q: [getattr(q, n).set_prior(pr, warning=False) for n, pr in hp[q.op_name].items()]
                   if getattr(q, 'op_name', None) in hp else None)
        models.append(GPModel(Covariance(k)))
    models.append(GPModel(models[0].covariance))       # the SAME kernel object again (random grammars do this): a second wave
    rng = np.random.RandomState(3)
    x = np.sort(rng.rand(30, 1) * 6, axis=0)
    y = np.sin(x) + 0.1 * rng.randn(30, 1)
    x, y = (x - x.mean()) / x.std(), (y - y.mean()) / y.std()
    return models, x, y, gp_regression(likelihood_hyperprior={'variance': hp['GP']['variance']})


def _failing_engine():
    from oracle_engine import OracleEngine
    from ms_project_b200 import program

    class Eng(OracleEngine):          # a fresh class per call, so `tripped` starts False for every fit
        tripped = False

        def lml_grad(self, batch, theta, ids=None, with_grad=True):
            out = OracleEngine.lml_grad(self, batch, theta, ids, with_grad)
            if not with_grad and not Eng.tripped:          # the evaluation at the optimum that yields model.log_likelihood()
                for i, k in enumerate(batch.kernels):
                    if program.infix_string(k) == FAIL_ONCE:
                        out[0][i] = -np.inf
                        Eng.tripped = True
            return out
    return Eng()


def _fit_state(models):
    rec = [[m.score, float(m.failed_fitting), m.fit_result.n_evals if m.fit_result is not None else -1,
            float(m.likelihood.variance)] + list(m.covariance.raw_kernel.param_array) for m in models]
    width = max(len(r) for r in rec)
    return np.array([r + [0.0] * (width - len(r)) for r in rec])


def _fit_worker(rank, world_size, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import torch.distributed as dist
    from ms_project_b200 import fitness, parallel
    dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=world_size)
    try:
        for policy in ('lpt', 'contiguous'):
            models, x, y, md = _fit_problem()
            np.random.seed(7)
            parallel.score_models_sharded(models, x, y, fitness.negative_bic, n_restarts=2, engine=_failing_engine(), model_dict=md,
                                          policy=policy)
            np.save(os.path.join(out_dir, 'fit_%s%d.npy' % (policy, rank)), _fit_state(models))
            np.save(os.path.join(out_dir, 'rng_%s%d.npy' % (policy, rank)), np.random.get_state()[1])
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_fit_with_retry_matches_single_process(tmp_path):
    """Two ranks, real score_models, both shard policies: identical model state AND identical np.random state on both ranks
    after a retry on one of them, and both equal to the single-process run (retry draws used to diverge the ranks'
    streams)."""
    import torch.multiprocessing as mp
    mp.spawn(_fit_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from ms_project_b200 import fitness
    from ms_project_b200.gp_model import score_models
    models, x, y, md = _fit_problem()
    np.random.seed(7)
    score_models(models, x, y, fitness.negative_bic, n_restarts=2, engine=_failing_engine(), model_dict=md, overlap=False)
    single = _fit_state(models)
    rng_single = np.random.get_state()[1]
    for policy in ('lpt', 'contiguous'):
        f0, f1 = np.load(tmp_path / ('fit_%s0.npy' % policy)), np.load(tmp_path / ('fit_%s1.npy' % policy))
        assert np.array_equal(f0, f1, equal_nan=True)
        assert np.array_equal(np.load(tmp_path / ('rng_%s0.npy' % policy)), np.load(tmp_path / ('rng_%s1.npy' % policy)))
        assert np.array_equal(single[:, :2], f0[:, :2], equal_nan=True)            # scores and failed flags: same bits
        assert np.array_equal(single[:, 3:], f0[:, 3:])                            # optimised parameters: same bits
        assert np.array_equal(rng_single, np.load(tmp_path / ('rng_%s0.npy' % policy)))
        i = FIT_NAMES.index(FAIL_ONCE)
        assert np.isfinite(f0[i, 0]) and f0[i, 1] == 0.0                           # the retry recovered the model
        assert f0[i, 2] > 0
        # the second occurrence of the shared kernel object started from the first one's optimum: fewer evaluations, same fit
        assert f0[-1, 2] < f0[0, 2] and abs(f0[-1, 0] - f0[0, 0]) < 1e-3 * abs(f0[0, 0])
