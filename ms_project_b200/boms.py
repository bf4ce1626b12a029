"""BOMS meta-level: the kernel-kernel surrogate GP over model indices, expected improvement and
the query step that consumes the distance matrix of `distance.HellingerDistanceBuilder`.

Mirrors, with the same names and argument meaning:
  KernelKernelGPModel        src/autoks/gp_regression_models.py:11-140
  get_quantiles              src/autoks/acquisition/util.py:5-23
  expected_improvement       src/autoks/acquisition/ei.py:7-17
  BayesOptGPStrategy.query   src/autoks/core/strategies/bayes_opt_gp_strategy.py:15-115
The surrogate is a (tiny) GP whose kernel is RBFDistanceBuilderKernelKernel -- the OP_LOOKUP leaf on the device -- so its
fit (optimize_restarts), LML and predictions run through the same libb200gp pipeline as every candidate kernel."""
import numpy as np
from scipy.special import erfc

from .gpy_compat.gp import GPRegression
from .gpy_compat.kern import RBFDistanceBuilderKernelKernel


def set_priors(param, priors, in_place=False):
    """set_priors (src/autoks/backend/kernel.py:105-116); `priors` maps parameter names to priors (PriorDist.raw_prior
    objects are unwrapped like the reference does)."""
    target = param if in_place else param.copy()
    for name, prior in priors.items():
        if name in target.parameter_names():
            getattr(target, name).set_prior(getattr(prior, 'raw_prior', prior), warning=False)
    return target


class KernelKernelGPModel(object):
    def __init__(self, kernel_kernel=None, noise_var=None, exact_f_eval=False, optimizer='lbfgsb', max_iters=1000,
                 optimize_restarts=5, verbose=True, kernel_kernel_hyperpriors=None):
        self.noise_var = noise_var
        self.exact_f_eval = exact_f_eval
        self.optimize_restarts = optimize_restarts
        self.optimizer = optimizer
        self.max_iters = max_iters
        self.verbose = verbose
        self.covariance = kernel_kernel
        self.kernel_hyperpriors = kernel_kernel_hyperpriors
        self.model = None

    def train(self):
        """Train (optimize) the model (gp_regression_models.py:41-52)."""
        if self.max_iters > 0:
            if self.optimize_restarts == 1:
                self.model.optimize(optimizer=self.optimizer, max_iters=self.max_iters, messages=False, ipython_notebook=False)
            else:
                self.model.optimize_restarts(num_restarts=self.optimize_restarts, optimizer=self.optimizer,
                                             max_iters=self.max_iters, ipython_notebook=False, verbose=self.verbose,
                                             robust=True, messages=False)

    def _create_model(self, x, y):
        """gp_regression_models.py:54-104: x = 2-d array of model indices, y = fitness scores."""
        assert np.issubdtype(x.dtype, np.integer) and x.min() >= 0
        self.input_dim = x.shape[1]
        assert self.covariance is not None
        kern = self.covariance.raw_kernel
        self.covariance = None
        noise_var = y.var() * 0.01 if self.noise_var is None else self.noise_var
        normalize = x.size > 1            # only normalize if more than 1 observation
        self.model = GPRegression(x, y, kern, noise_var=noise_var, normalizer=normalize)
        if self.kernel_hyperpriors is not None:
            if 'GP' in self.kernel_hyperpriors:
                set_priors(self.model.likelihood, self.kernel_hyperpriors['GP'], in_place=True)
            if 'SE' in self.kernel_hyperpriors:
                set_priors(self.model.kern, self.kernel_hyperpriors['SE'], in_place=True)
        # gp_regression_models.py:93-102: exact evaluations fix the noise at 1e-6; otherwise the noise stays positive when the
        # model carries priors (paramz has no log-Jacobian for Logistic) and is bounded to [1e-9, 1e6] when it does not
        if self.exact_f_eval:
            self.model.Gaussian_noise.constrain_fixed(1e-6, warning=False)
        elif self.model.priors.size > 0:
            self.model.Gaussian_noise.constrain_positive(warning=False)
        else:
            self.model.Gaussian_noise.constrain_bounded(1e-9, 1e6, warning=False)

    def update(self, x_all, y_all, x_new, y_new):
        """Update model with new observations (gp_regression_models.py:106-113)."""
        if self.model is None:
            self._create_model(x_all, y_all)
        else:
            self.model.set_XY(x_all, y_all)
        self.train()

    def _predict(self, x, full_cov, include_likelihood):
        if x.ndim == 1:
            x = x[None, :]
        m, v = self.model.predict(x, full_cov=full_cov, include_likelihood=include_likelihood)
        v = np.clip(v, 1e-10, np.inf)
        return m, v

    def predict(self, x, with_noise=True):
        m, v = self._predict(x, False, with_noise)
        return m, np.sqrt(v)

    def get_f_max(self):
        """The maximum of the posterior mean over the evaluated models (gp_regression_models.py:130-134)."""
        return self.model.predict(self.model.X)[0].max()


def get_quantiles(acquisition_par, f_max, m, s):
    """acquisition/util.py:5-23."""
    if isinstance(s, np.ndarray):
        s[s < 1e-10] = 1e-10
    elif s < 1e-10:
        s = 1e-10
    u = (m - f_max - acquisition_par) / s
    phi = np.exp(-0.5 * u ** 2) / np.sqrt(2 * np.pi)
    Phi = 0.5 * erfc(-u / np.sqrt(2))
    return phi, Phi, u


def expected_improvement(x, model, jitter=0.01):
    """acquisition/ei.py:7-17."""
    m, s = model.predict(x)
    f_max = model.get_f_max()
    phi, Phi, u = get_quantiles(jitter, f_max, m, s)
    return s * (u * Phi + phi)


class BayesOptGPStrategy(object):
    """Bayesian-optimisation strategy in model space (bayes_opt_gp_strategy.py:15-115).  `active_models` needs
    `.models`, `.selected_indices`, `__len__` and `get_candidate_indices()` like the reference's ActiveSet.  The inner
    evolutionary search for better acquisition scores (BoemsSurrogateSelector, a search driver outside the scoring
    path) is an optional hook: `expansion_fn(strategy, x_train, y_train)` may add models to the active set."""

    def __init__(self, active_models, acquisition_fn, kernel_builder, kernel_kernel_hyperpriors, expansion_fn=None,
                 optimize_restarts=10):
        from .covariance import Covariance
        self.active_models = active_models
        self.acquisition_fn = acquisition_fn
        self.kernel_builder = kernel_builder
        kernel_kernel = Covariance(RBFDistanceBuilderKernelKernel(self.kernel_builder, n_models=len(self.active_models)))
        self.kernel_kernel_gp_model = KernelKernelGPModel(kernel_kernel, verbose=False, exact_f_eval=False,
                                                          kernel_kernel_hyperpriors=kernel_kernel_hyperpriors,
                                                          optimize_restarts=optimize_restarts)
        self.kernel_kernel_gp_model_train_freq = 3
        self.expansion_freq = 3
        self.expansion_fn = expansion_fn

    def query(self, selected_models, fitness_scores, x_train, y_train):
        """The next model to evaluate (bayes_opt_gp_strategy.py:35-115).  Acquisition scores are left in `.score` of every
        candidate of the active set, as the reference does; the RETURNED model is a fresh, un-evaluated `GPModel` around
        the chosen covariance ("Reinitialize to get rid of acquisition score", :97-101), so the evaluation loop fits it
        instead of mistaking its acquisition value for a fitness.  Its active-set index is `self.last_selected_index`
        (also appended to `active_models.selected_indices`)."""
        from .gp_model import GPModel
        generation = len(selected_models)
        meta_x = np.array(self.active_models.selected_indices)[:, None]
        meta_y = np.array(fitness_scores, dtype=np.float64)[:, None]
        assert meta_x.shape == (len(selected_models), 1) and meta_x.shape == meta_y.shape
        self.kernel_kernel_gp_model.update(meta_x, meta_y, None, None)
        newly_selected = [self.active_models.selected_indices[-1]]
        self.kernel_builder.compute_distance(self.active_models, newly_selected, self.active_models.get_candidate_indices())
        selected = self.active_models.selected_indices
        candidates = sorted(set(range(len(self.active_models))) - set(selected))
        self.kernel_kernel_gp_model.model.kern.n_models = len(self.active_models)
        if generation % self.kernel_kernel_gp_model_train_freq == 0:
            self.kernel_kernel_gp_model.train()
        x_test = np.array(candidates)[:, None]
        acq = self.acquisition_fn(x_test, self.kernel_kernel_gp_model)
        assert not np.isnan(acq).any()
        for i, score in zip(candidates, acq[:, 0]):
            self.active_models.models[i].score = float(score)
        if self.expansion_fn is not None and (generation == 1 or generation % self.expansion_freq == 0):
            self.expansion_fn(self, x_train, y_train)
            candidates = sorted(set(range(len(self.active_models))) - set(selected))
        scores = [self.active_models.models[i].score for i in candidates]
        nxt = candidates[int(np.nanargmax([scores]))]
        chosen = self.active_models.models[nxt]
        next_model = GPModel(chosen.covariance, getattr(chosen, 'likelihood', None))
        self.active_models.selected_indices += [nxt]
        self.last_selected_index = nxt
        return next_model
