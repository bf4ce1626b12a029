"""Batched `optimize_restarts`: every (model, restart) pair is an independent L-BFGS-B instance, and all
instances advance in lock-step against ONE batched LML+gradient launch per iteration.

Replaces the sequential hot loops of the reference:
  ModelSelector._evaluate_models      src/autoks/core/model_selection/base.py:330   (candidates, one at a time)
  GPModel.fit -> optimize_restarts    src/autoks/core/gp_model.py:58-82            (restarts, one at a time)
  paramz Model.optimize / _objective_grads (per function evaluation: GP.parameters_changed)
while keeping their arithmetic: MAP objective -LML - log p(theta) - log|J| over Logexp-transformed parameters,
paramz's failure budget, restart draws in the reference's RNG order, argmin over restarts.

Host work per iteration is vectorised over the flat theta layout the device uses; only the compiled
L-BFGS-B step (scipy's `setulb`) is called per instance."""
import numpy as np

from .gpy_compat.priors import Gaussian as _GaussianPrior, LogGaussian as _LogGaussianPrior
from .gpy_compat.transformations import Logexp, Logistic, Fixed
from .optimizer import make_optimizer
from .program import ProgramBatch

F_MAX = np.finfo(np.float64).max
ALLOWED_FAILURES = 10          # paramz Model._allowed_failures
_LOGEXP = Logexp()


def check_optimizer(optimizer):
    """paramz resolves `optimizer=None` to its preferred optimiser, L-BFGS-B, which is what every experiment
    JSON of the reference uses (training/experiments/*/*.json: "optimizer": null); 'scg' is the default of
    CKSModelSelector (cks_model_selector.py:21).  Both have batched ask/tell drivers (optimizer.py); other paramz
    optimisers (tnc, simplex, ...) are rejected loudly."""
    make_optimizer(optimizer, np.zeros(1))


class _PriorTable(object):
    """Flat per-parameter prior description in device theta layout.  Gaussian / LogGaussian (the only families the
    reference's hyperprior tables use, core/hyperprior.py:54-112) are vectorised; any other prior object is applied
    through its own lnpdf / lnpdf_grad, grouped per instance."""

    def __init__(self, priors):
        n = len(priors)
        self.kind = np.zeros(n, dtype=np.int8)     # 0 none, 1 LogGaussian, 2 Gaussian, 3 other
        self.mu = np.zeros(n)
        self.s2 = np.ones(n)
        self.const = np.zeros(n)
        other = {}
        for i, p in enumerate(priors):
            if p is None:
                continue
            if type(p) is _LogGaussianPrior:
                self.kind[i] = 1
            elif type(p) is _GaussianPrior:
                self.kind[i] = 2
            elif hasattr(p, 'lnpdf') and hasattr(p, 'lnpdf_grad'):
                self.kind[i] = 3
                other.setdefault(id(p), (p, []))[1].append(i)
                continue
            else:
                raise NotImplementedError('prior %s is not supported on the scoring path' % type(p).__name__)
            self.mu[i], self.s2[i], self.const[i] = p.mu, p.sigma2, p.constant
        self.other = [(p, np.asarray(ix, dtype=int)) for p, ix in other.values()]
        self.has = self.kind != 0
        self.jac = self.has.copy()       # priored parameters whose constraint has a log-Jacobian (Logexp)
        self.lg = self.kind == 1
        self.ga = self.kind == 2

    def log_prior_terms(self, theta):
        """Per-parameter lnpdf + log-Jacobian of Logexp (priored parameters only), and their derivatives
        (GPy priorizable.py log_prior / _log_prior_gradients)."""
        lp = np.zeros_like(theta)
        dlp = np.zeros_like(theta)
        with np.errstate(all='ignore'):
            if self.lg.any():
                t = theta[self.lg]
                lt = np.log(t)
                lp[self.lg] = self.const[self.lg] - 0.5 * np.square(lt - self.mu[self.lg]) / self.s2[self.lg] - lt
                dlp[self.lg] = -((lt - self.mu[self.lg]) / self.s2[self.lg] + 1.) / t
            if self.ga.any():
                t = theta[self.ga]
                lp[self.ga] = self.const[self.ga] - 0.5 * np.square(t - self.mu[self.ga]) / self.s2[self.ga]
                dlp[self.ga] = -(t - self.mu[self.ga]) / self.s2[self.ga]
            for p, ix in self.other:
                lp[ix] = p.lnpdf(theta[ix])
                dlp[ix] = p.lnpdf_grad(theta[ix])
            if self.jac.any():
                t = theta[self.jac]
                lp[self.jac] += _LOGEXP.log_jacobian(t)
                dlp[self.jac] += _LOGEXP.log_jacobian_grad(t)
        return lp, dlp


class _TransformTable(object):
    """Flat per-parameter constraint description in device theta layout: Logexp (every parameter the grammar creates),
    Logistic(lower, upper) (`constrain_bounded`) or fixed (`constrain_fixed`: the optimiser still carries the coordinate,
    with a zero gradient, so it never moves and never influences the other coordinates' L-BFGS-B updates)."""

    def __init__(self, transforms, values):
        n = len(transforms)
        self.kind = np.zeros(n, dtype=np.int8)          # 0 Logexp, 1 Logistic, 2 fixed
        self.lo = np.zeros(n)
        self.hi = np.ones(n)
        self.fixed_value = np.asarray(values, dtype=np.float64).copy()
        for i, t in enumerate(transforms):
            if isinstance(t, Logistic):
                self.kind[i], self.lo[i], self.hi[i] = 1, t.lower, t.upper
            elif isinstance(t, Fixed):
                self.kind[i] = 2
            elif not isinstance(t, Logexp):
                raise NotImplementedError('constraint %s is not supported on the scoring path' % type(t).__name__)
        self.plain = not self.kind.any()
        self.lg = self.kind == 1
        self.fx = self.kind == 2

    def f(self, x):
        th = _LOGEXP.f(x)
        if not self.plain:
            if self.lg.any():
                xx = np.clip(x[self.lg], -700.0, 700.0)
                th[self.lg] = self.lo[self.lg] + (self.hi[self.lg] - self.lo[self.lg]) / (1. + np.exp(-xx))
            th[self.fx] = self.fixed_value[self.fx]
        return th

    def finv(self, theta):
        x = np.asarray(_LOGEXP.finv(np.where(self.kind == 0, theta, 1.0)), dtype=np.float64).copy()
        if not self.plain:
            if self.lg.any():
                t, lo, hi = theta[self.lg], self.lo[self.lg], self.hi[self.lg]
                x[self.lg] = np.log(np.clip(t - lo, 1e-10, np.inf) / np.clip(hi - t, 1e-10, np.inf))
            x[self.fx] = 0.0
        return x

    def gradfactor(self, theta, g):
        out = _LOGEXP.gradfactor(np.where(self.kind == 0, theta, 1.0), g)
        if not self.plain:
            if self.lg.any():
                t, lo, hi = theta[self.lg], self.lo[self.lg], self.hi[self.lg]
                out[self.lg] = g[self.lg] * (t - lo) * (hi - t) / (hi - lo)
            out[self.fx] = 0.0
        return out


class FitResult(object):
    __slots__ = ('lml', 'status', 'n_evals', 'runs', 'f_opt', 'theta')

    def __init__(self):
        self.lml = np.nan
        self.status = 0
        self.n_evals = 0
        self.runs = []          # per completed restart: dict(x_opt, f_opt, funcalls, nit, warnflag)
        self.f_opt = np.nan
        self.theta = None


def _model_priors(kernel, likelihood):
    return [p.prior for p in kernel.flattened_parameters()] + [p.prior for p in likelihood.flattened_parameters()]


def _model_transforms(kernel, likelihood):
    return [p.constraint for p in kernel.flattened_parameters()] + [p.constraint for p in likelihood.flattened_parameters()]


def draw_starts(kernels, likelihoods, num_restarts):
    """Starting points of every (model, restart) instance in the reference's RNG order (model-major,
    restart-minor): restart 0 = the model's current parameters, restart r > 0 = `randomize()` from the global
    np.random.  Returns a list (per model) of lists (per restart) of constrained parameter vectors."""
    out = []
    for k, lik in zip(kernels, likelihoods):
        cur = np.concatenate([k.param_array, lik.param_array])
        mine = [cur.copy()]
        for _ in range(1, int(num_restarts)):
            mine.append(_randomized(k, lik))
        out.append(mine)
    return out


def fit_models(engine, kernels, likelihoods, num_restarts=10, max_iters=1000, robust=True, verbose=False,
               on_iteration=None, starts=None, optimizer=None):
    """optimize_restarts for M models at once on `engine` (already bound to X, y).

    kernels / likelihoods are modified in place (the optimum is written back, as GPModel.fit does with
    model.kern / model.likelihood).  Returns a list of FitResult.  Restart r > 0 of model i starts from
    `randomize()` drawn from the global np.random in the reference's order (model-major, restart-minor)."""
    M = len(kernels)
    if M == 0:
        return []
    R = int(num_restarts)
    # ---- starting points, in the reference's RNG order (or pre-drawn by the caller, e.g. a sharded run)
    per_model = draw_starts(kernels, likelihoods, R) if starts is None else starts
    starts = []                     # per instance: constrained parameter vector (kernel params + noise)
    owner = []                      # instance -> model
    inst_kernels = []
    for i, k in enumerate(kernels):
        assert len(per_model[i]) == R
        for r in range(R):
            starts.append(np.asarray(per_model[i][r], dtype=np.float64))
            owner.append(i)
            inst_kernels.append(k)
    owner = np.asarray(owner)
    n_inst = len(starts)
    batch = ProgramBatch(inst_kernels)
    engine.upload_programs(batch)
    off = batch.theta_off
    priors = _PriorTable([p for i in range(M) for _ in range(R) for p in _model_priors(kernels[i], likelihoods[i])])

    theta = np.concatenate(starts)
    tf = _TransformTable([t for i in range(M) for _ in range(R) for t in _model_transforms(kernels[i], likelihoods[i])], theta)
    priors.jac &= tf.kind == 0          # the log-Jacobian term exists for Logexp only (paramz: Logistic has none)
    x0 = tf.finv(theta)
    opts = [make_optimizer(optimizer, x0[off[j]:off[j + 1]], max_iters=max_iters) for j in range(n_inst)]
    x_all = x0.copy()
    stale = np.zeros_like(theta)            # last successfully computed dLML/dtheta per instance
    fail_count = np.zeros(n_inst, dtype=int)
    aborted = np.zeros(n_inst, dtype=bool)
    n_evals = np.zeros(n_inst, dtype=int)
    active = []
    for j, o in enumerate(opts):
        x = o.ask()
        if x is not None:
            x_all[off[j]:off[j + 1]] = x
            active.append(j)
    seg_start = off[:-1]
    it = 0
    while active:
        ids = np.asarray(active, dtype=np.int32)
        theta = tf.f(x_all)
        lml, dlml, status, _tries = engine.lml_grad(batch, theta, ids=ids, with_grad=True)
        lp, dlp = priors.log_prior_terms(theta)
        lp_sum = np.add.reduceat(lp, seg_start)
        ok = status[ids] != 2                                  # 2 = NOT_PD == GPy raising LinAlgError
        good = ids[ok]
        if good.size:
            mask = _segment_mask(off, good, theta.size)
            stale[mask] = dlml[mask]
        g_theta = -(stale + dlp)
        g_x = tf.gradfactor(theta, g_theta)
        nxt_active = []
        for j, is_ok in zip(active, ok):
            a, b = off[j], off[j + 1]
            n_evals[j] += 1
            if is_ok:
                fail_count[j] = 0
                f, g = -lml[j] - lp_sum[j], g_x[a:b]
            else:
                if fail_count[j] >= ALLOWED_FAILURES:          # paramz re-raises: the restart is lost
                    aborted[j] = True
                    if not robust:
                        raise np.linalg.LinAlgError('not positive definite, even with jitter.')
                    continue
                fail_count[j] += 1
                f, g = opts[j].fail_value, np.clip(g_x[a:b], -1e10, 1e10)
            o = opts[j]
            o.tell(f, g)
            x = o.ask()
            if x is not None:
                x_all[a:b] = x
                nxt_active.append(j)
        active = nxt_active
        it += 1
        if on_iteration is not None:
            on_iteration(it, len(active))
        if verbose:
            print('lock-step iteration %d: %d / %d instances active' % (it, len(active), n_inst))

    # ---- argmin over each model's completed restarts; write the optimum back
    results = [FitResult() for _ in range(M)]
    best_theta = []
    for i in range(M):
        res = results[i]
        best = None
        for j in np.nonzero(owner == i)[0]:
            res.n_evals += int(n_evals[j])
            if aborted[j]:
                continue
            run = opts[j].result()
            run['x_opt'] = run.pop('x')
            run['f_opt'] = run.pop('f')
            res.runs.append(run)
            if best is None or run['f_opt'] < best['f_opt']:
                best = run
        if best is not None:
            j0 = int(np.nonzero(owner == i)[0][0])
            x_full = x_all.copy()
            x_full[off[j0]:off[j0 + 1]] = best['x_opt']                # through instance j0's slice of the transform table
            th = tf.f(x_full)[off[j0]:off[j0 + 1]]                    # (all restarts of a model share their constraints)
            res.f_opt = best['f_opt']
        else:                                                   # every restart failed: keep the initial parameters
            j0 = int(np.nonzero(owner == i)[0][0])
            th = starts[j0]
        res.theta = np.asarray(th, dtype=np.float64)
        best_theta.append(res.theta)
        kernels[i][:] = res.theta[:-1]
        likelihoods[i][:] = res.theta[-1:]

    # ---- GP.parameters_changed at the optimum: the model's log_likelihood()
    final = ProgramBatch(list(kernels))
    engine.upload_programs(final)
    th = final.pack_theta([t[:-1] for t in best_theta], [t[-1] for t in best_theta])
    lml, _g, status, _t = engine.lml_grad(final, th, with_grad=False)
    for i in range(M):
        results[i].lml = float(lml[i])
        results[i].status = int(status[i])
    return results


def fit_models_overlapped(engines, kernels, likelihoods, num_restarts=10, max_iters=1000, robust=True, starts=None,
                          optimizer=None):
    """fit_models with the models split into len(engines) contiguous groups, each driven by its own host thread against
    its own engine (same GPU, private streams): while one group's L-BFGS-B steps run on the host (scipy holds the GIL),
    the other group's LML+gradient batch runs on the device (ctypes releases the GIL), and the tails of the two groups
    fill each other's idle SMs.  Results are identical to fit_models: every instance is an independent optimisation, the
    restart points are drawn up front in the reference's order, and no reduction depends on the batch composition."""
    M = len(kernels)
    engines = list(engines)
    if len(engines) < 2 or M < 2 * len(engines):
        return fit_models(engines[0], kernels, likelihoods, num_restarts=num_restarts, max_iters=max_iters, robust=robust,
                          starts=starts, optimizer=optimizer)
    import threading
    per_model = draw_starts(kernels, likelihoods, int(num_restarts)) if starts is None else starts
    bounds = np.linspace(0, M, len(engines) + 1).astype(int)
    results = [None] * len(engines)
    errors = []

    def run(g):
        lo, hi = int(bounds[g]), int(bounds[g + 1])
        try:
            results[g] = fit_models(engines[g], kernels[lo:hi], likelihoods[lo:hi], num_restarts=num_restarts,
                                    max_iters=max_iters, robust=robust, starts=per_model[lo:hi], optimizer=optimizer)
        except BaseException as exc:          # re-raised in the caller's thread
            errors.append(exc)

    threads = [threading.Thread(target=run, args=(g,)) for g in range(1, len(engines))]
    for t in threads:
        t.start()
    run(0)
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return [r for part in results for r in part]


def _segment_mask(off, items, size):
    mask = np.zeros(size, dtype=bool)
    for j in items:
        mask[off[j]:off[j + 1]] = True
    return mask


def _randomized(kernel, likelihood):
    """Model-level Priorizable.randomize over [kernel params..., noise]: N(0,1) in optimiser space for every
    parameter first (one draw of size P+1), then one rvs(count) per distinct prior in order of first
    appearance."""
    params = kernel.flattened_parameters() + likelihood.flattened_parameters()
    free = [i for i, p in enumerate(params) if not p.is_fixed]
    x = np.random.normal(size=len(free))                  # paramz draws `_size_transformed()` numbers
    vals = np.array([float(p) for p in params], dtype=np.float64)
    for i, xi in zip(free, x):
        vals[i] = float(params[i].constraint.f(xi))
    seen = []
    for p in params:
        if p.prior is not None and not any(p.prior is q for q in seen):
            seen.append(p.prior)
    for pr in seen:
        ind = [i for i, p in enumerate(params) if p.prior is pr]
        vals[ind] = pr.rvs(len(ind))
    for i, p in enumerate(params):
        if p.is_fixed:
            vals[i] = float(p)
    return vals


def map_objective(engine, kernel, likelihood, xs):
    """MAP objective and gradient in optimiser space (what paramz `_objective_grads` returns) of ONE model at several
    optimiser-space points `xs` (n x (P+1)), evaluated in one batch.  The model objects are not modified."""
    xs = np.atleast_2d(np.asarray(xs, dtype=np.float64))
    n = xs.shape[0]
    batch = ProgramBatch([kernel] * n)
    engine.upload_programs(batch)
    priors = _PriorTable(_model_priors(kernel, likelihood) * n)
    cur = np.concatenate([kernel.param_array, likelihood.param_array])
    tf = _TransformTable(_model_transforms(kernel, likelihood) * n, np.tile(cur, n))
    priors.jac &= tf.kind == 0
    theta = tf.f(xs.reshape(-1))
    lml, dlml, status, _ = engine.lml_grad(batch, theta, with_grad=True)
    lp, dlp = priors.log_prior_terms(theta)
    f = -lml[:n] - lp.reshape(n, -1).sum(1)
    g = tf.gradfactor(theta, -(dlml[:theta.size] + dlp)).reshape(n, -1)
    return f.copy(), g.copy(), status[:n].copy()
