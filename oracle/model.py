"""Oracle: the MAP objective and `optimize_restarts` as paramz / GPy run them.  TEST INFRASTRUCTURE.

Restates, from the public upstream source as recalled (paramz / GPy are not importable here):
  paramz/transformations.py:Logexp; GPy/core/parameterization/priors.py:{Gaussian,LogGaussian};
  GPy/core/parameterization/priorizable.py:{log_prior,_log_prior_gradients,randomize};
  paramz/model.py:{_objective_grads,optimize,optimize_restarts};
  paramz/optimization/optimization.py:opt_lbfgsb (scipy fmin_l_bfgs_b, maxfun=maxiter=max_iters).
Reference call site: src/autoks/core/gp_model.py:58-82 (`fit`).
"""
import numpy as np
from scipy import optimize

from . import gp, kernels

_LIM_VAL = 36.0
_LOG_LIM_VAL = np.log(np.finfo(np.float64).max)
F_MAX = np.finfo(np.float64).max


# ----------------------------------------------------------------------------- Logexp transform
def logexp_f(x):
    x = np.asarray(x, dtype=np.float64)
    return np.where(x > _LIM_VAL, x, np.log1p(np.exp(np.clip(x, -_LOG_LIM_VAL, _LIM_VAL))))


def logexp_finv(f):
    f = np.asarray(f, dtype=np.float64)
    with np.errstate(over='ignore'):
        return np.where(f > _LIM_VAL, f, np.log(np.expm1(f)))


def logexp_gradfactor(f, df):
    return df * np.where(f > _LIM_VAL, 1., -np.expm1(-f))


def logexp_log_jacobian(f):
    with np.errstate(over='ignore'):
        return np.where(f > _LIM_VAL, f, np.log(np.expm1(f))) - f


def logexp_log_jacobian_grad(f):
    return 1. / np.expm1(f)


# ----------------------------------------------------------------------------- priors
# A prior is a tuple (kind, mu, sigma), kind in {'lognormal', 'gaussian'}.  Equal tuples play the
# role of GPy's per-(mu, sigma) singleton prior objects.
def prior_lnpdf(prior, x):
    kind, mu, sigma = prior
    s2 = sigma * sigma
    const = -0.5 * np.log(2 * np.pi * s2)
    if kind == 'lognormal':
        return const - 0.5 * np.square(np.log(x) - mu) / s2 - np.log(x)
    if kind == 'gaussian':
        return const - 0.5 * np.square(x - mu) / s2
    raise ValueError(kind)


def prior_lnpdf_grad(prior, x):
    kind, mu, sigma = prior
    s2 = sigma * sigma
    if kind == 'lognormal':
        return -((np.log(x) - mu) / s2 + 1.) / x
    if kind == 'gaussian':
        return -(x - mu) / s2
    raise ValueError(kind)


def prior_rvs(prior, n, rng):
    kind, mu, sigma = prior
    z = rng.randn(int(n)) * sigma + mu
    return np.exp(z) if kind == 'lognormal' else z


def boms_priors(expr, data_noise=0.01):
    """Per-parameter priors from `boms_hyperpriors` (src/autoks/core/hyperprior.py:54-112), in
    param_array order, and the likelihood-noise prior."""
    pl = ('lognormal', np.log(0.1), 1.0)
    ps = ('lognormal', np.log(0.4), 1.0)
    pa = ('lognormal', np.log(0.05), np.sqrt(0.5))
    plp = ('lognormal', np.log(2), np.sqrt(0.5))
    ps0 = ('lognormal', 0.0, 1.0)
    pp = ('lognormal', np.log(0.1), np.sqrt(0.5))
    fam_map = {
        'SE': {'variance': ps, 'lengthscale': pl},
        'RQ': {'variance': ps, 'lengthscale': pl, 'power': pa},
        'PER': {'variance': ps, 'lengthscale': plp, 'period': pp},
        'LIN': {'variances': ps, 'shifts': ps0},
    }
    kp = [fam_map[f][n] for f, _, n in kernels.param_names(expr)]
    return kp, ('lognormal', np.log(data_noise), np.sqrt(0.1))


# ----------------------------------------------------------------------------- the model
class OracleGP:
    """GPy `GP(X, Y, kernel, likelihood=Gaussian())` + paramz `Model`, for one compositional kernel.

    param_array = [kernel params..., Gaussian_noise.variance]; every parameter is Logexp
    constrained (true for all five leaf families and for the Gaussian likelihood)."""

    allowed_failures = 10

    def __init__(self, expr, X, y, theta=None, noise_var=1.0, kern_priors=None, noise_prior=None,
                 dist_mode='gpy', quirks=False):
        self.expr = expr
        self.X = np.asarray(X, dtype=np.float64)
        self.y = np.asarray(y, dtype=np.float64).reshape(-1, 1)
        th = kernels.default_theta(expr) if theta is None else np.asarray(theta, dtype=np.float64)
        self.param_array = np.concatenate([th, [noise_var]])
        P = self.param_array.size
        pri = list(kern_priors) if kern_priors is not None else [None] * (P - 1)
        self.priors = pri + [noise_prior]
        assert len(self.priors) == P
        self.dist_mode, self.quirks = dist_mode, quirks
        self._fail_count = 0
        self._lml = np.nan
        self._grad = np.zeros(P)          # dLML/dparam (constrained space), last successful
        self.n_evals = 0
        self.jitter_tries = 0
        self.optimization_runs = []
        self._update()

    # -- sizes / accessors mirroring backend/model.py:95-108
    @property
    def num_data(self):
        return self.X.shape[0]

    def _size_transformed(self):
        return self.param_array.size

    def log_likelihood(self):
        return self._lml

    @property
    def optimizer_array(self):
        return logexp_finv(self.param_array)

    def _set_optimizer_array(self, x):
        self.param_array = logexp_f(x)
        self._update()

    def _update(self):
        """GP.parameters_changed."""
        self.n_evals += 1
        lml, g, inf = gp.lml_and_grad(self.expr, self.param_array[:-1], self.param_array[-1],
                                      self.X, self.y, self.dist_mode, self.quirks)
        self._lml, self._grad, self.jitter_tries = lml, g, inf['tries']

    # -- prior terms (priorizable.py)
    def log_prior(self):
        lp = 0.
        for p, x in zip(self.priors, self.param_array):
            if p is not None:
                lp += prior_lnpdf(p, x) + logexp_log_jacobian(x)
        return float(lp)

    def _log_prior_gradients(self):
        g = np.zeros(self.param_array.size)
        for i, (p, x) in enumerate(zip(self.priors, self.param_array)):
            if p is not None:
                g[i] = prior_lnpdf_grad(p, x) + logexp_log_jacobian_grad(x)
        return g

    def objective_function(self):
        return -float(self._lml) - self.log_prior()

    def objective_function_gradients(self):
        return -(self._grad + self._log_prior_gradients())

    def _transform_gradients(self, g):
        return logexp_gradfactor(self.param_array, g)

    def _objective_grads(self, x):
        """paramz Model._objective_grads, including the failure budget."""
        try:
            self._set_optimizer_array(x)
            f = self.objective_function()
            g = self._transform_gradients(self.objective_function_gradients())
            self._fail_count = 0
        except (np.linalg.LinAlgError, ZeroDivisionError, ValueError):
            if self._fail_count >= self.allowed_failures:
                raise
            self._fail_count += 1
            f = F_MAX
            g = np.clip(self._transform_gradients(self.objective_function_gradients()), -1e10, 1e10)
        return f, g

    def _objective(self, x):
        """paramz Model._objective (what SCG calls for function values)."""
        try:
            self._set_optimizer_array(x)
            obj = self.objective_function()
            self._fail_count = 0
        except (np.linalg.LinAlgError, ZeroDivisionError, ValueError):
            if self._fail_count >= self.allowed_failures:
                raise
            self._fail_count += 1
            return np.inf
        return obj

    def _grads(self, x):
        """paramz Model._grads (what SCG calls for gradients)."""
        try:
            self._set_optimizer_array(x)
            g = self._transform_gradients(self.objective_function_gradients())
            self._fail_count = 0
        except (np.linalg.LinAlgError, ZeroDivisionError, ValueError):
            if self._fail_count >= self.allowed_failures:
                raise
            self._fail_count += 1
            g = np.clip(self._transform_gradients(self.objective_function_gradients()), -1e10, 1e10)
        return g

    # -- optimisation (paramz Model.optimize / opt_lbfgsb / opt_SCG)
    def optimize(self, max_iters=1000, optimizer=None):
        x0 = self.optimizer_array.copy()
        if optimizer is not None and str(optimizer).lower() == 'scg':
            from .scg import scg
            x_opt, flog, nfev, status = scg(self._objective, self._grads, x0, maxiters=max_iters)
            run = dict(x_opt=x_opt, f_opt=flog[-1], funcalls=nfev, warnflag=status, nit=len(flog) - 1)
            self._set_optimizer_array_quiet(x_opt)
            self.optimization_runs.append(run)
            return run
        x_opt, _, info = optimize.fmin_l_bfgs_b(self._objective_grads, x0, maxfun=max_iters, maxiter=max_iters)
        f_opt = self._objective_grads(x_opt)[0]
        run = dict(x_opt=x_opt, f_opt=f_opt, funcalls=info['funcalls'], warnflag=info['warnflag'], nit=info['nit'])
        self._set_optimizer_array_quiet(x_opt)
        self.optimization_runs.append(run)
        return run

    def _set_optimizer_array_quiet(self, x):
        try:
            self._set_optimizer_array(x)
        except np.linalg.LinAlgError:
            self._lml = np.nan

    def randomize(self, rng):
        """Priorizable.randomize: N(0,1) in optimiser space, then prior draws where a prior exists.
        Distinct priors draw in order of first appearance in param_array, one rvs(count) call each
        (GPy iterates `self.priors.items()`; that dict's order is not recoverable here -- documented
        choice, shared with the product's host shim)."""
        x = rng.normal(size=self._size_transformed())
        params = logexp_f(x)
        seen = []
        for p in self.priors:
            if p is not None and p not in seen:
                seen.append(p)
        for p in seen:
            ind = [i for i, q in enumerate(self.priors) if q == p]
            params[ind] = prior_rvs(p, len(ind), rng)
        self.param_array = params
        try:
            self._update()
        except np.linalg.LinAlgError:
            self._lml = np.nan

    def optimize_restarts(self, num_restarts=10, rng=None, robust=True, max_iters=1000, optimizer=None):
        """paramz Model.optimize_restarts (sequential branch)."""
        rng = np.random if rng is None else rng
        initial_length = len(self.optimization_runs)
        initial_parameters = self.optimizer_array.copy()
        for i in range(num_restarts):
            try:
                if i > 0:
                    self.randomize(rng)
                self.optimize(max_iters=max_iters, optimizer=optimizer)
            except Exception:
                if not robust:
                    raise
        if len(self.optimization_runs) > initial_length:
            i = int(np.argmin([o['f_opt'] for o in self.optimization_runs[initial_length:]]))
            self._set_optimizer_array_quiet(self.optimization_runs[initial_length + i]['x_opt'])
        else:
            self._set_optimizer_array_quiet(initial_parameters)
        return self.optimization_runs


def fit(expr, X, y, n_restarts=10, rng=None, optimizer=None, **kw):
    """GPModel.fit (src/autoks/core/gp_model.py:58-82): optimise; if LML is +-inf/NaN retry once;
    returns (model, failed_fitting)."""
    m = OracleGP(expr, X, y, **kw)
    m.optimize_restarts(n_restarts, rng=rng, optimizer=optimizer)
    failed = False
    if not np.isfinite(m.log_likelihood()):
        m.optimize_restarts(n_restarts, rng=rng, optimizer=optimizer)
        if not np.isfinite(m.log_likelihood()):
            failed = True
    return m, failed
