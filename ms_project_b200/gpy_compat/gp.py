"""`GPy.core.GP` stand-in: the model object `GPModel.build_model` constructs
(src/autoks/core/gp_model.py:44-56, RawGPModelType at src/autoks/backend/model.py:10) with the attributes the
reference reads: `optimize_restarts`, `log_likelihood`, `kern`, `likelihood`, `num_data`,
`_size_transformed`, `param_array`, `predict`, `set_XY`, `X`, `priors` (gp_model.py:64-80;
backend/model.py:86-108).  Numbers come from libb200gp; optimisation is the batched lock-step driver in
ms_project_b200.fit run with a batch of one."""
import numpy as np

from .likelihoods import Gaussian
from .param import Parameterized


class GP(Parameterized):
    def __init__(self, X, Y, kernel, likelihood=None, mean_function=None, inference_method=None, name='gp',
                 Y_metadata=None, normalizer=False, device=0):
        super(GP, self).__init__(name)
        if mean_function is not None:
            raise NotImplementedError('mean functions are outside the scoring path')
        X = np.asarray(X, dtype=np.float64)
        Y = np.asarray(Y, dtype=np.float64)
        assert X.ndim == 2, 'X must be N x D'
        if Y.ndim == 1:
            Y = Y[:, None]
        assert Y.shape == (X.shape[0], 1), 'the scoring path is single-output GP regression (Y is N x 1)'
        # GPy's `normalizer=True` (GPRegression(..., normalizer=True) in KernelKernelGPModel._create_model,
        # src/autoks/gp_regression_models.py:79-80) is the Standardize normaliser: the GP sees (Y - mean) / std and
        # predictions are mapped back (mean * std + mean, var * std^2)
        self.normalizer = bool(normalizer)
        self.Y_raw = Y
        self._set_Y(Y)
        self.X = X
        self.num_data, self.input_dim = X.shape
        self.kern = kernel
        self.likelihood = Gaussian() if likelihood is None else likelihood
        # 'laplace' is accepted because the experiment JSONs pass it; with a Gaussian likelihood its mode is
        # the exact posterior, and the exact path is what the library implements.
        self.inference_method = inference_method
        self.link_parameters(self.kern, self.likelihood)
        self.optimization_runs = []
        self._device = device
        self._role = 'search'          # engine cache key: the BOMS surrogate uses its own so the search's bound data stays
        self._lml = None
        self._status = 0

    def _set_Y(self, Y):
        if self.normalizer:
            self._y_mean, self._y_std = float(Y.mean()), float(Y.std())
            self.Y = (Y - self._y_mean) / self._y_std
        else:
            self._y_mean, self._y_std = 0.0, 1.0
            self.Y = Y

    # ------------------------------------------------------------------ device evaluation
    def _engine(self):
        from ..engine import get_engine
        return get_engine(self._device, self._role)

    def _evaluate(self, with_grad=True):
        from ..program import ProgramBatch
        eng = self._engine()
        batch = ProgramBatch([self.kern])
        eng.ensure_bound(self.X, self.Y, 1, batch.max_params)
        eng.upload_programs(batch)
        theta = batch.pack_theta([self.kern.param_array], [float(self.likelihood.variance)])
        lml, dlml, status, _ = eng.lml_grad(batch, theta, with_grad=with_grad)
        self._lml, self._status = float(lml[0]), int(status[0])
        if with_grad:
            for p, g in zip(self.flattened_parameters(), dlml[:self.size]):
                p.gradient[0] = g
        return self._lml

    def parameters_changed(self):
        self._lml = None

    def __setitem__(self, item, value):
        super(GP, self).__setitem__(item, value)
        self._lml = None

    def log_likelihood(self):
        """GP.log_likelihood(): the log marginal likelihood at the current parameters."""
        if self._lml is None:
            self._evaluate(with_grad=True)
        if self._status == 2:
            return np.nan
        return self._lml

    def log_prior(self):
        from ..fit import _PriorTable
        t = _PriorTable([p.prior for p in self.flattened_parameters()])
        return float(np.sum(t.log_prior_terms(self.param_array)[0]))

    def objective_function(self):
        return -float(self.log_likelihood()) - self.log_prior()

    # ------------------------------------------------------------------ optimisation
    @property
    def optimizer_array(self):
        return np.array([float(p.constraint.finv(float(p))) for p in self.flattened_parameters()])

    def randomize(self, rand_gen=None, *args, **kwargs):
        super(GP, self).randomize(rand_gen, *args, **kwargs)
        self._lml = None

    def optimize(self, optimizer=None, start=None, messages=False, max_iters=1000, ipython_notebook=True,
                 clear_after_finish=False, **kwargs):
        return self.optimize_restarts(num_restarts=1, optimizer=optimizer, max_iters=max_iters, verbose=False)

    def optimize_restarts(self, num_restarts=10, robust=False, verbose=True, parallel=False, num_processes=None,
                          optimizer=None, max_iters=1000, ipython_notebook=False, messages=False, **kwargs):
        from ..fit import fit_models, check_optimizer
        check_optimizer(optimizer)
        eng = self._engine()
        eng.ensure_bound(self.X, self.Y, max(int(num_restarts), 1), self.kern.size)
        res = fit_models(eng, [self.kern], [self.likelihood], num_restarts=num_restarts, max_iters=max_iters,
                         robust=robust, optimizer=optimizer)[0]
        self.optimization_runs += res.runs
        self._lml, self._status = res.lml, res.status
        return self.optimization_runs

    def set_XY(self, X=None, Y=None):
        if X is not None:
            self.X = np.asarray(X, dtype=np.float64)
            self.num_data = self.X.shape[0]
        if Y is not None:
            self.Y_raw = np.asarray(Y, dtype=np.float64).reshape(-1, 1)
            self._set_Y(self.Y_raw)
        self._lml = None

    def predict(self, Xnew, full_cov=False, Y_metadata=None, kern=None, likelihood=None, include_likelihood=True):
        """GP.predict: posterior mean and marginal variance at Xnew, each Ns x 1 (GPy Posterior._raw_predict +
        Gaussian.predictive_values; callers: base.py:231-237, gp_regression_models.py:113-134), through
        b200gp_posterior.  NOT_PD even after the jitter ladder raises LinAlgError like GPy."""
        if full_cov:
            raise NotImplementedError('full_cov=True is not used by the reference path and is not implemented')
        if kern is not None or likelihood is not None:
            raise NotImplementedError('predict with a replacement kernel / likelihood is not implemented')
        from ..program import ProgramBatch
        Xnew = np.asarray(Xnew, dtype=np.float64)
        if Xnew.ndim == 1:
            Xnew = Xnew[None, :]
        eng = self._engine()
        batch = ProgramBatch([self.kern])
        eng.ensure_bound(self.X, self.Y, 1, batch.max_params)
        eng.upload_programs(batch)
        theta = batch.pack_theta([self.kern.param_array], [float(self.likelihood.variance)])
        mean, var, lml, status = eng.posterior(batch, theta, 0, Xnew, include_likelihood=include_likelihood)
        self._lml, self._status = lml, status
        if status == 2:
            raise np.linalg.LinAlgError('not positive definite, even with jitter.')
        return (mean * self._y_std + self._y_mean)[:, None], (var * self._y_std ** 2)[:, None]


class GPRegression(GP):
    """`GPy.models.GPRegression(X, Y, kernel, normalizer=, noise_var=)`: a GP with a Gaussian likelihood, as
    KernelKernelGPModel._create_model builds it (src/autoks/gp_regression_models.py:79-80)."""

    def __init__(self, X, Y, kernel=None, Y_metadata=None, normalizer=None, noise_var=1., mean_function=None):
        if kernel is None:
            from .kern import RBF
            kernel = RBF(np.asarray(X).shape[1])
        super(GPRegression, self).__init__(X, Y, kernel, Gaussian(variance=noise_var), mean_function=mean_function,
                                           name='GP regression', Y_metadata=Y_metadata, normalizer=bool(normalizer))
        self._role = 'surrogate'       # its own engine: a surrogate fitted in the middle of a search leaves the search's bound data alone

    @property
    def Gaussian_noise(self):
        return self.likelihood
