"""paramz `Logexp` transform (the only constraint the scoring path uses: every leaf-kernel parameter and
the Gaussian noise variance are `Param(..., Logexp())`; rational-quadratic.ipynb#2, periodic.ipynb#2,
linear.ipynb#2)."""
import numpy as np

_lim_val = 36.0
_log_lim_val = np.log(np.finfo(np.float64).max)


class Transformation(object):
    domain = None


class Logexp(Transformation):
    domain = 'positive'
    _instance = None

    def __new__(cls):
        if cls._instance is None:
            cls._instance = super(Logexp, cls).__new__(cls)
        return cls._instance

    def f(self, x):
        x = np.asarray(x, dtype=np.float64)
        return np.where(x > _lim_val, x, np.log1p(np.exp(np.clip(x, -_log_lim_val, _lim_val))))

    def finv(self, f):
        f = np.asarray(f, dtype=np.float64)
        with np.errstate(over='ignore', divide='ignore'):
            return np.where(f > _lim_val, f, np.log(np.expm1(f)))

    def gradfactor(self, f, df):
        return df * np.where(f > _lim_val, 1., -np.expm1(-f))

    def log_jacobian(self, model_param):
        with np.errstate(over='ignore', divide='ignore'):
            return np.where(model_param > _lim_val, model_param, np.log(np.expm1(model_param))) - model_param

    def log_jacobian_grad(self, model_param):
        return 1. / (np.expm1(model_param))

    def __str__(self):
        return '+ve'


class Logistic(Transformation):
    """paramz `Logistic(lower, upper)`: f(x) = lower + (upper - lower) / (1 + exp(-x)) -- what `constrain_bounded` installs
    (KernelKernelGPModel bounds the surrogate's noise to [1e-9, 1e6] when it has no hyperpriors,
    src/autoks/gp_regression_models.py:102).  Like in paramz it has no log-Jacobian, so it cannot carry a prior."""
    domain = 'bounded'

    def __init__(self, lower, upper):
        assert lower < upper
        self.lower, self.upper = float(lower), float(upper)
        self.difference = self.upper - self.lower

    def f(self, x):
        x = np.asarray(x, dtype=np.float64)
        return self.lower + self.difference / (1. + np.exp(-np.clip(x, -_log_lim_val, _log_lim_val)))

    def finv(self, f):
        f = np.asarray(f, dtype=np.float64)
        with np.errstate(divide='ignore', invalid='ignore'):
            return np.log(np.clip(f - self.lower, 1e-10, np.inf) / np.clip(self.upper - f, 1e-10, np.inf))

    def gradfactor(self, f, df):
        return df * (f - self.lower) * (self.upper - f) / self.difference

    def __str__(self):
        return '{},{}'.format(self.lower, self.upper)


class Fixed(Transformation):
    """Marker for `constrain_fixed`: the parameter keeps its value and is not an optimisation variable."""
    domain = 'fixed'

    def __str__(self):
        return 'fixed'
