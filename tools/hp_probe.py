"""C3 Hellinger pair timing only (500 models, n = 100, 20 samples, 124 750 pairs): `python tools/hp_probe.py [reps]`.
While the pair kernel was being developed the library read B200GP_HP_VARIANT to pick a variant (profiles/experiments/
hellinger_la_r02.md); the final library dispatches on the subset size only."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from ms_project_b200 import workloads
from ms_project_b200.engine import get_engine
from ms_project_b200.covariance import Covariance
from ms_project_b200.distance import ActiveModels, HellingerDistanceBuilder, probability_samples
from ms_project_b200.gpy_compat.priors import Gaussian
from tools.bench_configs import _Model

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
X, Y, pool, xsub = workloads.config_c3()
eng = get_engine(0)
n = len(pool)
am = ActiveModels(max_n_models=n)
for k in pool:
    am.add(_Model(Covariance(k)))
prob = probability_samples(40, 20)
noise_prior = Gaussian(np.log(0.01), 1)
t0 = time.perf_counter()
builder = HellingerDistanceBuilder(noise_prior, 20, 40, n, am, list(range(n)), xsub, probability_samples_matrix=prob, engine=eng)
torch.cuda.synchronize()
t_pre = time.perf_counter() - t0
ii, jj = np.triu_indices(n, 1)
builder.pair_distances(ii[:1000], jj[:1000])
torch.cuda.synchronize()
ts = []
for _ in range(reps):
    t0 = time.perf_counter()
    dist, status = builder.pair_distances(ii, jj)
    torch.cuda.synchronize()
    ts.append(time.perf_counter() - t0)
print(json.dumps({'precompute_s': t_pre, 'pairs_s': ts,
                  'nan': int(np.isnan(dist).sum()), 'bad': int((status != 0).sum()), 'checksum': float(np.nansum(dist))}))
