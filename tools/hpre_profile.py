"""cProfile of HellingerDistanceBuilder construction (C3 precompute: 500 models x 20 prior samples): where the host time goes."""
import os, sys, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from ms_project_b200 import workloads
from ms_project_b200.engine import get_engine
from ms_project_b200.covariance import Covariance
from ms_project_b200.distance import ActiveModels, HellingerDistanceBuilder, probability_samples
from ms_project_b200.gpy_compat.priors import Gaussian
from tools.bench_configs import _Model
X, Y, pool, xsub = workloads.config_c3()
eng = get_engine(0)
n = len(pool)
prob = probability_samples(40, 20)
noise_prior = Gaussian(np.log(0.01), 1)
def build():
    am = ActiveModels(max_n_models=n)
    for k in pool:
        am.add(_Model(Covariance(k)))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    b = HellingerDistanceBuilder(noise_prior, 20, 40, n, am, list(range(n)), xsub, probability_samples_matrix=prob, engine=eng)
    torch.cuda.synchronize()
    return time.perf_counter() - t0
print('warm-up', build())
print('second', build())
pr = cProfile.Profile()
pr.enable()
t = build()
pr.disable()
print('profiled', t)
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
