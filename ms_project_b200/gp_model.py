"""`GPModel` with the reference's scoring API (src/autoks/core/gp_model.py:14-82) and the batched evaluation
seam of `ModelSelector._evaluate_models` (src/autoks/core/model_selection/base.py:314-389).

`score_model` / `fit` / `build_model` keep the signatures and side effects (optimised kernel + likelihood
written back, `failed_fitting`, `score`, `evaluated`); numerical failure never raises.  `evaluate_models`
first works out which models the sequential reference loop WOULD evaluate (budget cut-off, visited set,
duplicates inside the batch: first occurrence wins), scores exactly those in one GPU batch, then replays the
loop's bookkeeping and callbacks in the original order."""
import copy
from time import time

import numpy as np

from .covariance import Covariance
from .fit import fit_models, fit_models_overlapped, check_optimizer
from .gpy_compat.gp import GP
from .gpy_compat.likelihoods import Likelihood

RawGPModelType = GP


class GPModel(object):
    def __init__(self, covariance, likelihood=None, evaluated=False):
        self.covariance = covariance
        self.evaluated = evaluated
        self._score = np.nan
        self.model_input_dict = dict()
        self.likelihood = likelihood
        self.failed_fitting = False
        self.fit_result = None

    # -- reference API -----------------------------------------------------------------------------
    def score_model(self, x, y, scoring_func, **fit_kwargs):
        model = self.fit(x, y, **fit_kwargs)
        self.score = scoring_func(model) if not self.failed_fitting else np.nan
        return self.score

    def build_model(self, x, y):
        input_dict = self.model_input_dict.copy()
        input_dict['X'] = x
        input_dict['Y'] = y
        input_dict['kernel'] = self.covariance.raw_kernel
        input_dict['likelihood'] = self.likelihood
        return RawGPModelType(**input_dict)

    def fit(self, x, y, optimizer=None, n_restarts=10):
        model = self.build_model(x, y)
        model.optimize_restarts(ipython_notebook=False, optimizer=optimizer, num_restarts=n_restarts, verbose=False,
                                robust=True, messages=False)
        if _is_bad(model.log_likelihood()):
            model.optimize_restarts(ipython_notebook=False, optimizer=optimizer, num_restarts=n_restarts, verbose=False,
                                    robust=True, messages=False)
            if _is_bad(model.log_likelihood()):
                self.failed_fitting = True
        self.covariance.raw_kernel = model.kern
        self.likelihood = model.likelihood
        return model

    @property
    def score(self):
        return self._score

    @score.setter
    def score(self, score):
        self._score = score
        self.evaluated = True

    def to_dict(self):
        return {'__class__': 'GPModel', '__module__': 'src.autoks.core.gp_model',
                'likelihood': None if self.likelihood is None else self.likelihood.to_dict(),
                'covariance': self.covariance.to_dict(), 'evaluated': self.evaluated, 'score': self.score}

    @classmethod
    def from_dict(cls, d):
        m = cls(Covariance.from_dict(d['covariance']),
                None if d.get('likelihood') is None else Likelihood.from_dict(d['likelihood']), d.get('evaluated', False))
        m._score = d.get('score', np.nan)
        return m

    def __str__(self):
        return str(self.covariance.symbolic_expr)


def _is_bad(lml):
    """`model.log_likelihood() in (float('-inf'), float('inf'), np.nan)` (gp_model.py:67-69).  Tuple membership
    tests identity first and equality second, so a computed NaN is NOT caught by the reference (only the
    np.nan singleton would be); +-inf are.  Reproduced as is."""
    return lml is np.nan or lml == float('inf') or lml == float('-inf')


# -- batched seam ------------------------------------------------------------------------------------
def plan_evaluation(models, n_evals, eval_budget, visited, skip_evaluated=True):
    """Simulate base.py:330-347 without fitting: returns (to_fit, order) where `to_fit` lists the models the
    sequential loop would pass to `_evaluate_model` and `order` lists (model, action) in loop order with action
    in {'fit', 'keep', 'skip'}; the loop stops at the budget."""
    visited = set(visited)
    n = n_evals
    to_fit, order = [], []
    for m in models:
        if n >= eval_budget:
            order.append((m, 'stop'))
            break
        if m.covariance.symbolic_expr_expanded in visited and skip_evaluated:
            order.append((m, 'skip'))
            continue
        if not m.evaluated:
            to_fit.append(m)
            order.append((m, 'fit'))
            n += 1
            visited.add(m.covariance.symbolic_expr_expanded)
        else:
            order.append((m, 'keep'))
    return to_fit, order


def score_models(models, x, y, scoring_func, optimizer=None, n_restarts=10, engine=None, model_dict=None,
                 all_models=None, shard=None, overlap=None):
    """GPModel.score_model for a whole list in ONE lock-step GPU batch (same per-model results and side effects as
    calling score_model on each, including the retry-once rule and failed_fitting)."""
    check_optimizer(optimizer)
    if all_models is not None:
        # sharded run: draw the restart points of the WHOLE list (identical on every rank), keep our slice
        all_gps = []
        for m in all_models:
            _prepare(m, model_dict)
            all_gps.append(m.build_model(x, y))
        from .fit import draw_starts
        starts_all = draw_starts([g.kern for g in all_gps], [g.likelihood for g in all_gps], n_restarts)
        starts = starts_all[shard[0]:shard[1]]
    else:
        starts = None
    if not models:
        return []
    if engine is None:
        from .engine import get_engine
        engine = get_engine(0)
    gps = []
    for m in models:
        _prepare(m, model_dict)
        gps.append(m.build_model(x, y))
    maxp = max(g.kern.size for g in gps)
    n_inst = len(gps) * max(int(n_restarts), 1)
    if overlap is None:
        overlap = len(gps) >= 16         # enough instances for two lock-step drivers to keep the GPU busy
    if overlap:
        # two host threads, two engines on the same GPU (fit.fit_models_overlapped): host L-BFGS-B work of one group
        # hides behind the device batch of the other
        engines = [engine, engine.peer()]
        for e in engines:
            e.ensure_bound(x, y, (n_inst + 1) // 2, maxp)
        results = fit_models_overlapped(engines, [g.kern for g in gps], [g.likelihood for g in gps], num_restarts=n_restarts,
                                        robust=True, starts=starts, optimizer=optimizer)
    else:
        engine.ensure_bound(x, y, n_inst, maxp)
        results = fit_models(engine, [g.kern for g in gps], [g.likelihood for g in gps], num_restarts=n_restarts, robust=True,
                             starts=starts, optimizer=optimizer)
    for g, r in zip(gps, results):
        g._lml, g._status = r.lml, r.status
        g.optimization_runs += r.runs
    retry = [i for i, g in enumerate(gps) if _is_bad(g.log_likelihood())]
    if retry:
        engine.ensure_bound(x, y, len(retry) * max(int(n_restarts), 1), maxp)
        rr = fit_models(engine, [gps[i].kern for i in retry], [gps[i].likelihood for i in retry], num_restarts=n_restarts,
                        robust=True, optimizer=optimizer)
        for i, r in zip(retry, rr):
            gps[i]._lml, gps[i]._status = r.lml, r.status
            if _is_bad(gps[i].log_likelihood()):
                models[i].failed_fitting = True
    for m, g, r in zip(models, gps, results):
        m.covariance.raw_kernel = g.kern
        m.likelihood = g.likelihood
        m.fit_result = r
        m.score = scoring_func(g) if not m.failed_fitting else np.nan
    return [m.score for m in models]


def _prepare(m, model_dict):
    if model_dict:
        m.model_input_dict = {k: copy.deepcopy(v) for k, v in model_dict.items() if k != 'likelihood'}
        if m.likelihood is None and model_dict.get('likelihood') is not None:
            m.likelihood = model_dict['likelihood'].copy()


class BatchedEvaluator(object):
    """Drop-in body for ModelSelector._evaluate_models: keeps n_evals / visited / total_eval_time and fires the
    per-model callbacks in the original order (base.py:329-389)."""

    def __init__(self, fitness_fn, optimizer=None, n_restarts_optimizer=10, engine=None, model_dict=None):
        self.fitness_fn = fitness_fn
        self.optimizer = optimizer
        self.n_restarts_optimizer = n_restarts_optimizer
        self.engine = engine
        self.model_dict = model_dict
        self.n_evals = 0
        self.total_eval_time = 0
        self.visited = set()

    def evaluate_models(self, models, x, y, eval_budget, callbacks=None, skip_evaluated=True):
        if callbacks is not None:
            callbacks.on_evaluate_all_begin(logs={'gp_models': models})
        to_fit, order = plan_evaluation(models, self.n_evals, eval_budget, self.visited, skip_evaluated)
        t0 = time()
        score_models(to_fit, x, y, self.fitness_fn, optimizer=self.optimizer, n_restarts=self.n_restarts_optimizer,
                     engine=self.engine, model_dict=self.model_dict)
        self.total_eval_time += time() - t0
        evaluated = []
        for m, action in order:
            if callbacks is not None:
                callbacks.on_evaluate_begin(logs={'gp_model': m})
            if action == 'stop':
                break
            if action == 'skip':
                continue
            if action == 'fit':
                self.n_evals += 1
                self.visited.add(m.covariance.symbolic_expr_expanded)
            evaluated.append(m)
            if callbacks is not None:
                callbacks.on_evaluate_end({'gp_model': m, 'x': x})
        if callbacks is not None:
            callbacks.on_evaluate_all_end({'gp_models': evaluated, 'x': x})
        return evaluated
