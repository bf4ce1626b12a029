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


def _waves(models):
    """Models that share ONE kernel object (the random grammars hand out `grammar.base_kernels[i]` itself, so a candidate
    list may hold `RQ_1` several times as the same object) are a sequential chain in the reference: the second occurrence
    starts its restart 0 from the optimum the first one wrote into the shared object (gp_model.py:61-79).  Wave w holds the
    w-th occurrence of every kernel object; waves run one after the other, everything inside a wave in one batch."""
    seen, waves = {}, []
    for i, m in enumerate(models):
        w = seen.get(id(m.covariance.raw_kernel), 0)
        seen[id(m.covariance.raw_kernel)] = w + 1
        while len(waves) <= w:
            waves.append([])
        waves[w].append(i)
    return waves


def score_models(models, x, y, scoring_func, optimizer=None, n_restarts=10, engine=None, model_dict=None, overlap=None,
                 comm=None, max_iters=1000):
    """GPModel.score_model for a whole list in ONE lock-step GPU batch (same per-model results and side effects as
    calling score_model on each, including the retry-once rule and failed_fitting).

    Restart points come from the global np.random in the reference's order (model-major, restart-minor), all drawn before
    the batch runs.  Retry-once (gp_model.py:67-77): the models whose LML came back +-inf are re-fitted in a second batch
    whose restart points are drawn AFTER the first batch, in list order -- the reference draws them right after the failing
    model, so the streams differ from the first retry on (hazard 5).

    `comm` (parallel.Comm, None = this process alone) shards every batch over the ranks of a torch.distributed group:
    each rank fits its slice, one all_gather of fixed-width records writes scores and optimised parameters into EVERY
    rank's model objects.  Every rank draws the restart points of the WHOLE list and -- after the failure flags have been
    all-gathered -- the retry points of EVERY failing model, so np.random stays in the same state on all ranks and the search
    loop continues identically everywhere.  `max_iters` is paramz's optimiser budget (GPModel.fit leaves it at 1000)."""
    check_optimizer(optimizer)
    from .fit import draw_starts
    if not models:
        return []
    R = max(int(n_restarts), 1)
    gps = []
    for m in models:
        _prepare(m, model_dict)
        gps.append(m.build_model(x, y))
    starts = draw_starts([g.kern for g in gps], [g.likelihood for g in gps], n_restarts)
    if engine is None:
        from .engine import get_engine
        engine = get_engine(0)
    n_data = np.asarray(x).shape[0]

    def fit(idx, st):
        if not idx:
            return []
        ks, ls = [gps[i].kern for i in idx], [gps[i].likelihood for i in idx]
        maxp = max(k.size for k in ks)
        n_inst = len(idx) * R
        two = (len(idx) >= 16) if overlap is None else (overlap and len(idx) >= 4)
        if two:
            # two host threads, two engines on the same GPU (fit.fit_models_overlapped): host L-BFGS-B work of one group
            # hides behind the device batch of the other
            engines = [engine, engine.peer()]
            for e in engines:
                e.ensure_bound(x, y, (n_inst + 1) // 2, maxp)
            return fit_models_overlapped(engines, ks, ls, num_restarts=n_restarts, max_iters=max_iters, robust=True, starts=st,
                                         optimizer=optimizer)
        engine.ensure_bound(x, y, n_inst, maxp)
        return fit_models(engine, ks, ls, num_restarts=n_restarts, max_iters=max_iters, robust=True, starts=st,
                          optimizer=optimizer)

    for w, wave in enumerate(_waves(models)):
        if w > 0:                       # restart 0 = the parameters the previous occurrence of the shared object left behind
            for i in wave:
                starts[i][0] = np.concatenate([gps[i].kern.param_array, gps[i].likelihood.param_array])
        mine = wave if comm is None else comm.shard(wave, [model_cost(gps[i].kern, n_data) for i in wave])
        results = dict(zip(mine, fit(mine, [starts[i] for i in mine])))
        for i, r in results.items():
            gps[i]._lml, gps[i]._status = r.lml, r.status
            gps[i].optimization_runs += r.runs
        bad = [i for i in mine if _is_bad(gps[i].log_likelihood())]
        if comm is not None:
            bad = comm.union(bad)       # who failed, over the whole wave, identical on every rank
        # retry points of every failing model of the wave, in list order, drawn on every rank (restart 0 = the model's
        # current, i.e. optimised, parameters: only the owner holds them and only the owner uses them)
        retry_starts = {i: draw_starts([gps[i].kern], [gps[i].likelihood], n_restarts)[0] for i in sorted(bad)}
        retry = [i for i in sorted(bad) if i in results]
        for i, r in zip(retry, fit(retry, [retry_starts[i] for i in retry])):
            gps[i]._lml, gps[i]._status = r.lml, r.status
            gps[i].optimization_runs += r.runs
            r.n_evals += results[i].n_evals
            results[i] = r                                     # the record (lml, theta) of the fit that stands
            if _is_bad(gps[i].log_likelihood()):
                models[i].failed_fitting = True
        for i in mine:
            m, g = models[i], gps[i]
            m.covariance.raw_kernel = g.kern
            m.likelihood = g.likelihood
            m.fit_result = results[i]
            m.score = scoring_func(g) if not m.failed_fitting else np.nan
        if comm is not None:
            comm.exchange(models, wave, mine)
    return [m.score for m in models]


# L-BFGS-B evaluations a leaf of each family adds to a 3-restart fit, relative to SE: least squares over the per-model counts of
# the C2 fit (256 candidates, profiles/fit_evals_r02.json: 89 / 132 / 314 / 108 evaluations per SE / RQ / PER / LIN leaf,
# R^2 = 0.52 -- periodic surfaces need about three times the iterations; the rest of the 50x spread is not predictable from the
# kernel's shape)
_EVALS_PER_LEAF = {'SE': 1.0, 'RQ': 1.5, 'PER': 3.5, 'LIN': 1.2}


def model_cost(kernel, n_data):
    """Relative cost of fitting one model, for balancing shards: (cost of one evaluation: N^3 for the factorisation + inverse
    plus the element-wise passes, which scale with the number of leaves) x (expected number of evaluations,
    which grows with the leaves and their families).  With this cost LPT leaves the busiest of 8 ranks 10 % above the mean on
    C2 (17 % with a parameter-count cost, 19 % with equal costs, 0.1 % if the counts were known in advance)."""
    leaves = []

    def visit(q):                      # leaf kernels carry a device opcode family; sums, products and parameters do not
        family = getattr(q, 'op_name', None)
        if family is not None:
            leaves.append(family)
    kernel.traverse(visit)
    leaves = leaves or ['SE']
    n_evals = 0.5 + sum(_EVALS_PER_LEAF.get(f, 1.0) for f in leaves)
    return (float(n_data) ** 3 + 150.0 * len(leaves) * float(n_data) ** 2) * n_evals


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
