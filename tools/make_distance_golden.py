#!/usr/bin/env python
"""Generates tests/golden/distances_reference.json: the reference's OWN distance builders
(/root/reference/src/autoks/distance/distance.py: HellingerDistanceBuilder, FrobeniusDistanceBuilder,
CorrelationDistanceBuilder -- unmodified, imported through tests/refshim.py) over a pool of grammar kernels with the BOMS
hyperpriors.  `kern.K` is served by the oracle standing in for the device; the prior sampling (distance/util.py), the
Cholesky / `chol_safe` log-determinants, the asymmetric log-determinant skip rule and the averaging are the reference's code.
tests/test_gpu_distances_reference.py feeds the same kernels and the same probability samples to the device builders.

Runs only in the build container (needs /root/reference):   python tools/make_distance_golden.py
"""
import json
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
warnings.filterwarnings('ignore')

import refshim  # noqa: E402

refshim.install()
import oracle_engine  # noqa: E402
from src.autoks.core.covariance import Covariance  # noqa: E402
from src.autoks.core.gp_model import GPModel  # noqa: E402
from src.autoks.core.grammar import BomsGrammar  # noqa: E402
from src.autoks.core.hyperprior import boms_hyperpriors  # noqa: E402
from src.autoks.distance.distance import ActiveModels, HellingerDistanceBuilder, FrobeniusDistanceBuilder, \
    CorrelationDistanceBuilder  # noqa: E402
from GPy.core.parameterization.priors import Gaussian  # noqa: E402


def main():
    n_models, n_points, S, max_hyp = 14, 40, 20, 40
    rng = np.random.default_rng(5)
    x = rng.random((n_points, 2))
    x = (x - x.mean(0)) / x.std(0)
    grammar = BomsGrammar(base_kernel_names=['SE', 'RQ'], hyperpriors=boms_hyperpriors())
    grammar.build(2)
    covs = grammar.expand_full_brute_force(2, 60)[:n_models]
    out = {'x': x.tolist(), 'S': S, 'max_num_hyperparameters': max_hyp, 'noise_prior': [float(np.log(0.01)), 1.0],
           'kernels': [], 'builders': {},
           'generator': 'tools/make_distance_golden.py: the reference\'s own distance builders; kern.K from the oracle'}
    with oracle_engine.installed():
        for name, cls in (('hellinger', HellingerDistanceBuilder), ('frobenius', FrobeniusDistanceBuilder),
                          ('correlation', CorrelationDistanceBuilder)):
            am = ActiveModels(n_models)
            kernels = [c.raw_kernel.copy() for c in covs]
            idx = am.update([GPModel(Covariance(k)) for k in kernels])
            assert idx == list(range(n_models))
            if not out['kernels']:
                for k in kernels:
                    out['kernels'].append({'kernel': k.to_dict(),
                                           'priors': [[type(p.prior).__name__, float(p.prior.mu), float(p.prior.sigma)]
                                                      for p in k.flattened_parameters()]})
            builder = cls(Gaussian(np.log(0.01), 1.0), S, max_hyp, n_models, am, idx, x)
            builder.compute_distance(am, idx, idx)
            D = builder.get_kernel(n_models)
            out['probability_samples'] = np.asarray(builder.probability_samples).tolist()
            out['builders'][name] = {'distance': np.asarray(D).tolist()}
            if name == 'hellinger':
                out['builders'][name]['log_det'] = [np.asarray(am.models[i].info[0]).tolist() for i in idx]
            print(name, 'distance matrix', D.shape, 'min/max off-diagonal', float(np.nanmin(D + np.eye(n_models))), float(np.nanmax(D)))
    path = os.path.join(ROOT, 'tests', 'golden', 'distances_reference.json')
    json.dump(out, open(path, 'w'))
    print('->', path, '%.0f KB' % (os.path.getsize(path) / 1e3))


if __name__ == '__main__':
    main()
