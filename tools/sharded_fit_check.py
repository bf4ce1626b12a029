"""Multi-GPU check of the sharded scoring seam on real hardware (the CPU suite covers it over gloo):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 tools/sharded_fit_check.py

Every rank scores its shard of 12 candidates (N = 400, 2 restarts, nBIC) over NCCL and all-gathers the records; rank 0 also
fits the whole list on its own GPU and the two results must agree (same restart draws, same per-instance optimiser runs)."""
import os
import sys

import numpy as np

sys.path.insert(0, '.')
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from ms_project_b200 import fitness, workloads  # noqa: E402
from ms_project_b200.covariance import Covariance  # noqa: E402
from ms_project_b200.engine import Engine  # noqa: E402
from ms_project_b200.gp_model import GPModel, score_models  # noqa: E402
from ms_project_b200.gp_models import gp_regression  # noqa: E402
from ms_project_b200.parallel import score_models_sharded  # noqa: E402


def main():
    rank, lr = int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(lr)
    dist.init_process_group('nccl', device_id=torch.device('cuda', lr))
    X, y, kernels = workloads.config_c2(n=400, n_kernels=12)
    md = gp_regression(likelihood_hyperprior={'variance': workloads.boms_hyperpriors()['GP']['variance']})
    eng = Engine(lr)
    models = [GPModel(Covariance(k.copy())) for k in kernels]
    np.random.seed(3)
    scores = score_models_sharded(models, X, y, fitness.negative_bic, n_restarts=2, engine=eng, model_dict=md)
    # every rank must hold identical results
    t = torch.tensor(scores, dtype=torch.float64, device='cuda')
    ref = t.clone()
    dist.broadcast(ref, 0)
    assert torch.equal(t, ref), 'ranks disagree after the all_gather'
    if rank == 0:
        single = [GPModel(Covariance(k.copy())) for k in kernels]
        np.random.seed(3)
        s1 = score_models(single, X, y, fitness.negative_bic, n_restarts=2, engine=eng, model_dict=md)
        err = float(np.max(np.abs(np.asarray(s1) - np.asarray(scores)) / np.maximum(1.0, np.abs(s1))))
        same_params = all(np.allclose(a.covariance.raw_kernel.param_array, b.covariance.raw_kernel.param_array, rtol=1e-9)
                          for a, b in zip(single, models))
        print('world %d: sharded vs single-GPU scores max rel diff %.3e, identical parameters: %s, ranking equal: %s'
              % (dist.get_world_size(), err, same_params, list(np.argsort(s1)) == list(np.argsort(scores))))
        assert err <= 1e-9 and same_params
    dist.barrier()
    dist.destroy_process_group()


main()
