#!/usr/bin/env python
"""Generates tests/golden/dropin_*.json: the reference's OWN selectors (/root/reference, unmodified, imported through
tests/refshim.py) run a search over the product's GPy-compatible objects with the CPU oracle standing in for the device
(tests/oracle_engine.py); every batch the search hands to the scoring seam is recorded -- candidate kernels with their
hyperparameters and priors, which candidates share one kernel object, the np.random seed set just before the batch, and
what came back (scores, LMLs, optimised parameters).  tests/test_gpu_dropin.py replays every recorded batch on the B200
through the real engine and must reproduce the scores and the selected kernel of every search step.

Runs only in the build container (needs /root/reference; takes minutes at N = 500):
    python tools/make_dropin_golden.py [cks_nbic cks_loglikn cks_scg random evolutionary]
"""
import json
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
warnings.filterwarnings('ignore')

import refshim  # noqa: E402

refshim.install()
import oracle_engine  # noqa: E402
from ms_project_b200 import integration, gp_model  # noqa: E402
from src.autoks.core.model_selection import CKSModelSelector, EvolutionaryModelSelector, GeomRandomModelSelector  # noqa: E402
from src.autoks.core.hyperprior import boms_hyperpriors  # noqa: E402
from src.autoks.core.grammar import CKSGrammar  # noqa: E402


def prior_rec(p):
    return None if p is None else [type(p).__name__, float(p.mu), float(p.sigma)]


def model_rec(m):
    k = m.covariance.raw_kernel
    lik = m.likelihood
    return {'kernel': k.to_dict(), 'priors': [prior_rec(p.prior) for p in k.flattened_parameters()],
            'noise': float(lik.variance), 'noise_prior': prior_rec(lik.variance.prior), 'expr': str(m)}


class Recorder(object):
    def __init__(self, seed_base):
        self.calls = []
        self.seed_base = seed_base
        self.real = gp_model.score_models

    def __call__(self, models, x, y, scoring_func, optimizer=None, n_restarts=10, engine=None, model_dict=None, overlap=None,
                 comm=None):
        seed = self.seed_base + len(self.calls)
        np.random.seed(seed)
        first = {}
        share = [first.setdefault(id(m.covariance.raw_kernel), i) for i, m in enumerate(models)]
        before = [model_rec(m) for m in models]
        t0 = time.time()
        scores = self.real(models, x, y, scoring_func, optimizer=optimizer, n_restarts=n_restarts, engine=engine,
                           model_dict=model_dict, overlap=False, comm=comm)
        self.calls.append({
            'seed': seed, 'n_restarts': int(n_restarts), 'optimizer': optimizer, 'fitness': scoring_func.__name__,
            'models': before, 'shared_with': share,
            'scores': [float(s) for s in scores],
            'lml': [float(m.fit_result.lml) for m in models],
            'failed': [bool(m.failed_fitting) for m in models],
            'n_evals': [int(m.fit_result.n_evals) for m in models],
            'theta': [np.concatenate([m.covariance.raw_kernel.param_array, m.likelihood.param_array]).tolist() for m in models],
            'oracle_seconds': time.time() - t0})
        print('  batch %d: %d models, %.1f s, best %s' % (len(self.calls), len(models), time.time() - t0,
                                                          str(models[int(np.nanargmax(scores))]) if models else '-'), flush=True)
        return scores


def c1_data(n):
    """C1: x = 8 U(0,1), y = sin(x) + 0.05 N(0,1), seed 0 (Sinusoid1Dataset, synthetic_data.py:16-23)."""
    rng = np.random.default_rng(0)
    x = 8.0 * rng.random((n, 1))
    return x, np.sin(x) + 0.05 * rng.standard_normal((n, 1))


def d2_data(n):
    rng = np.random.default_rng(3)
    x = rng.standard_normal((n, 2))
    return x, (np.sin(x[:, :1]) * x[:, 1:] + 0.1 * rng.standard_normal((n, 1)))


def cks(fitness, optimizer, n, budget):
    grammar = CKSGrammar(base_kernel_names=['SE', 'RQ', 'LIN', 'PER'], hyperpriors=boms_hyperpriors())
    return (lambda: CKSModelSelector(grammar=grammar, fitness_fn=fitness, optimizer=optimizer, n_restarts_optimizer=3),
            c1_data(n), budget, 3)


CONFIGS = {
    # C1 as stated: CKS depth 3, SE/RQ/LIN/PER, N = 500, 3 restarts; both fitness functions (base.py:29-32)
    'cks_nbic': lambda: cks('nbic', None, 500, 64),
    'cks_loglikn': lambda: cks('loglikn', None, 500, 64),
    'cks_scg': lambda: cks('nbic', 'scg', 120, 40),           # CKSModelSelector's default optimiser (cks_model_selector.py:21)
    'random': lambda: (lambda: GeomRandomModelSelector(base_kernel_names=['SE', 'RQ', 'LIN', 'PER'], n_restarts_optimizer=3),
                       c1_data(150), 24, None),
    'evolutionary': lambda: (lambda: EvolutionaryModelSelector(base_kernel_names=['SE', 'RQ'], n_restarts_optimizer=2, pop_size=12),
                             d2_data(150), 40, 12),
}


def generate(name):
    make, (x, y), budget, max_gen = CONFIGS[name]()
    rec = Recorder(seed_base=1000)
    integration.score_models = rec
    with oracle_engine.installed():
        np.random.seed(0)
        sel = integration.use_batched_evaluation(make())
        t0 = time.time()
        sel.train(x, y, eval_budget=budget, max_generations=max_gen, verbose=0)
        best = sel.best_model()
        out = {'config': name, 'selector': type(sel).__name__, 'eval_budget': budget, 'n_evals': int(sel.n_evals),
               'x': sel._x_train.tolist(), 'y': sel._y_train.tolist(),
               'best': str(best), 'best_score': float(best.score),
               'selected': [[str(m), float(m.score)] for m in sel.selected_models],
               'calls': rec.calls,
               'generator': 'tools/make_dropin_golden.py: reference selector (unmodified) over gpy_compat, oracle numerics',
               'seconds': time.time() - t0}
    integration.score_models = rec.real
    path = os.path.join(ROOT, 'tests', 'golden', 'dropin_%s.json' % name)
    json.dump(out, open(path, 'w'))
    print('%s: %d batches, %d evaluations, best %s (%.6f), %.0f s -> %s (%.0f KB)'
          % (name, len(rec.calls), sel.n_evals, best, best.score, out['seconds'], path, os.path.getsize(path) / 1e3), flush=True)


if __name__ == '__main__':
    for name in (sys.argv[1:] or list(CONFIGS)):
        generate(name)
