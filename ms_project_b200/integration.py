"""The batch seam as code: replaces the body of the reference's
`ModelSelector._evaluate_models` (src/autoks/core/model_selection/base.py:314-369) by the batched evaluator, for a
selector instance or for a whole selector class, without touching the reference's sources.

The replaced method keeps the loop's observable behaviour -- budget cut-off (:333-336), visited skip incl. duplicates
inside the batch (:337-342), per-model callbacks in list order (:331,353), `n_evals` / `visited` / `total_eval_time`
(:383-385), the progress bar (:386-387) -- but every model the loop would have fitted is fitted in ONE lock-step GPU batch
(`gp_model.score_models`), or sharded over the ranks of an initialised torch.distributed group
(`parallel.score_models_sharded`)."""
from time import time

from .gp_model import plan_evaluation, score_models


def batched_evaluate_models(self, models, eval_budget, callbacks, verbose=0, skip_evaluated=True, _engine=None):
    """Drop-in for ModelSelector._evaluate_models(self, models, eval_budget, callbacks, verbose, skip_evaluated)."""
    from . import parallel
    comm = parallel.Comm() if parallel.world()[1] > 1 else None
    callbacks.on_evaluate_all_begin(logs={'gp_models': models})
    to_fit, order = plan_evaluation(models, self.n_evals, eval_budget, self.visited, skip_evaluated)
    t0 = time()
    score_models(to_fit, self._x_train, self._y_train, self.fitness_fn, optimizer=self.optimizer,
                 n_restarts=self.n_restarts_optimizer, engine=_engine, comm=comm)
    self.total_eval_time += time() - t0
    evaluated = []
    for m, action in order:
        callbacks.on_evaluate_begin(logs={'gp_model': m})
        if action == 'stop':
            break
        if action == 'skip':
            continue
        if action == 'fit':
            self.n_evals += 1
            self.visited.add(m.covariance.symbolic_expr_expanded)
            if verbose and getattr(self, 'pbar', None) is not None:
                self.pbar.update()
        evaluated.append(m)
        callbacks.on_evaluate_end({'gp_model': m, 'x': self._x_train})
    callbacks.on_evaluate_all_end({'gp_models': evaluated, 'x': self._x_train})
    return evaluated


def use_batched_evaluation(target, engine=None):
    """`target`: a ModelSelector instance or class of the reference.  Returns `target`."""
    def _evaluate_models(self, models, eval_budget, callbacks, verbose=0, skip_evaluated=True):
        return batched_evaluate_models(self, models, eval_budget, callbacks, verbose, skip_evaluated, _engine=engine)
    if isinstance(target, type):
        target._evaluate_models = _evaluate_models
    else:
        import types
        target._evaluate_models = types.MethodType(_evaluate_models, target)
    return target
