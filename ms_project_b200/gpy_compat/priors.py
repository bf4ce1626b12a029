"""Priors with GPy's surface (`GPy.core.parameterization.priors`), as the reference uses them:
`PRIOR_DICT` (src/autoks/backend/prior.py:5-17), class identity + `.mu` / `.sigma` in
`prior_sample*` (src/autoks/distance/util.py:21-71), `set_prior` (backend/kernel.py:105-115).
Like GPy, equal (mu, sigma) pairs share ONE instance, so parameters with the same hyperprior are
grouped under one key of `kernel.priors`."""
import weakref

import numpy as np


class Prior(object):
    domain = None
    _instances = []

    def pdf(self, x):
        return np.exp(self.lnpdf(x))

    def plot(self):
        raise NotImplementedError('plotting is outside the scoring path')

    def __repr__(self):
        return self.__str__()


class Gaussian(Prior):
    domain = 'real'
    _instances = []

    def __new__(cls, mu=0, sigma=1):
        if cls._instances:
            cls._instances[:] = [i for i in cls._instances if i()]
            for inst in cls._instances:
                if inst().mu == mu and inst().sigma == sigma:
                    return inst()
        o = super(Prior, cls).__new__(cls)
        cls._instances.append(weakref.ref(o))
        return cls._instances[-1]()

    def __init__(self, mu=0, sigma=1):
        self.mu = float(mu)
        self.sigma = float(sigma)
        self.sigma2 = np.square(self.sigma)
        self.constant = -0.5 * np.log(2 * np.pi * self.sigma2)

    def __getnewargs__(self):
        return self.mu, self.sigma

    def __str__(self):
        return 'N({:.2g}, {:.2g})'.format(self.mu, self.sigma)

    def lnpdf(self, x):
        return self.constant - 0.5 * np.square(x - self.mu) / self.sigma2

    def lnpdf_grad(self, x):
        return -(x - self.mu) / self.sigma2

    def rvs(self, n):
        return np.random.randn(int(n)) * self.sigma + self.mu

    def to_dict(self):
        return {'class': 'GPy.core.parameterization.priors.Gaussian', 'mu': self.mu, 'sigma': self.sigma}


class LogGaussian(Gaussian):
    domain = 'positive'
    _instances = []

    def __str__(self):
        return 'lnN({:.2g}, {:.2g})'.format(self.mu, self.sigma)

    def lnpdf(self, x):
        return self.constant - 0.5 * np.square(np.log(x) - self.mu) / self.sigma2 - np.log(x)

    def lnpdf_grad(self, x):
        return -((np.log(x) - self.mu) / self.sigma2 + 1.) / x

    def rvs(self, n):
        return np.exp(np.random.randn(int(n)) * self.sigma + self.mu)

    def to_dict(self):
        return {'class': 'GPy.core.parameterization.priors.LogGaussian', 'mu': self.mu, 'sigma': self.sigma}


# ---- the remaining names of PRIOR_DICT (src/autoks/backend/prior.py:5-17).  The reference's own hyperprior tables
# (core/hyperprior.py:54-112) use Gaussian / LogGaussian only; the scalar families below are restated from GPy's
# public priors.py so that a user-built PriorDist keeps working (host-side lnpdf / lnpdf_grad / rvs, applied by fit.py
# per parameter); the multivariate / DGPLVM ones are not scoring-path priors and refuse construction.
class Uniform(Prior):
    domain = 'real'

    def __init__(self, lower=0, upper=1):
        self.lower, self.upper = float(lower), float(upper)
        assert self.lower < self.upper, 'Lower needs to be strictly smaller than upper.'
        if self.lower >= 0:
            self.domain = 'positive'
        elif self.upper <= 0:
            self.domain = 'negative'

    def __str__(self):
        return '[{:.2g}, {:.2g}]'.format(self.lower, self.upper)

    def lnpdf(self, x):
        region = (np.asarray(x) >= self.lower) * (np.asarray(x) <= self.upper)
        with np.errstate(divide='ignore'):
            return np.log(region * np.e)

    def lnpdf_grad(self, x):
        return np.zeros(np.shape(x))

    def rvs(self, n):
        return np.random.uniform(self.lower, self.upper, size=int(n))

    def to_dict(self):
        return {'class': 'GPy.core.parameterization.priors.Uniform', 'lower': self.lower, 'upper': self.upper}


class Gamma(Prior):
    """Gamma(a, b) with rate b: p(x) = b^a / Gamma(a) x^(a-1) exp(-b x)."""
    domain = 'positive'

    def __init__(self, a=1, b=.5):
        from scipy.special import gammaln
        self.a, self.b = float(a), float(b)
        self.constant = -gammaln(self.a) + self.a * np.log(self.b)

    def __str__(self):
        return 'Ga({:.2g}, {:.2g})'.format(self.a, self.b)

    def lnpdf(self, x):
        return self.constant + (self.a - 1) * np.log(x) - self.b * x

    def lnpdf_grad(self, x):
        return (self.a - 1.) / x - self.b

    def rvs(self, n):
        return np.random.gamma(scale=1. / self.b, shape=self.a, size=int(n))

    def to_dict(self):
        return {'class': 'GPy.core.parameterization.priors.Gamma', 'a': self.a, 'b': self.b}


class InverseGamma(Gamma):
    domain = 'positive'

    def __str__(self):
        return 'iGa({:.2g}, {:.2g})'.format(self.a, self.b)

    def lnpdf(self, x):
        return self.constant - (self.a + 1) * np.log(x) - self.b / x

    def lnpdf_grad(self, x):
        return -(self.a + 1.) / x + self.b / np.square(x)

    def rvs(self, n):
        return 1. / np.random.gamma(scale=1. / self.b, shape=self.a, size=int(n))

    def to_dict(self):
        return {'class': 'GPy.core.parameterization.priors.InverseGamma', 'a': self.a, 'b': self.b}


class Exponential(Prior):
    domain = 'positive'

    def __init__(self, l=1):
        self.l = float(l)

    def __str__(self):
        return 'Exp({:.2g})'.format(self.l)

    def lnpdf(self, x):
        return np.log(self.l) - self.l * np.asarray(x)

    def lnpdf_grad(self, x):
        return -self.l * np.ones(np.shape(x))

    def rvs(self, n):
        return np.random.exponential(scale=1. / self.l, size=int(n))

    def to_dict(self):
        return {'class': 'GPy.core.parameterization.priors.Exponential', 'l': self.l}


class StudentT(Prior):
    domain = 'real'

    def __init__(self, mu=0, sigma=1, nu=4):
        self.mu, self.sigma, self.nu = float(mu), float(sigma), float(nu)
        self.sigma2 = np.square(self.sigma)

    def __str__(self):
        return 'St({:.2g}, {:.2g}, {:.2g})'.format(self.mu, self.sigma, self.nu)

    def lnpdf(self, x):
        from scipy.stats import t
        return t.logpdf(x, self.nu, self.mu, self.sigma)

    def lnpdf_grad(self, x):
        return -(self.nu + 1.) * (x - self.mu) / (self.nu * self.sigma2 + np.square(x - self.mu))

    def rvs(self, n):
        from scipy.stats import t
        return t.rvs(self.nu, loc=self.mu, scale=self.sigma, size=int(n))

    def to_dict(self):
        return {'class': 'GPy.core.parameterization.priors.StudentT', 'mu': self.mu, 'sigma': self.sigma, 'nu': self.nu}


class HalfT(Prior):
    """Half-Student-t(A, nu) on the positive axis."""
    domain = 'positive'

    def __init__(self, A=1, nu=4):
        from scipy.special import gammaln
        self.A, self.nu = float(A), float(nu)
        self.constant = gammaln(.5 * (self.nu + 1.)) - gammaln(.5 * self.nu) - .5 * np.log(np.pi * self.A * self.nu)

    def __str__(self):
        return 'hT({:.2g}, {:.2g})'.format(self.A, self.nu)

    def lnpdf(self, theta):
        theta = np.asarray(theta, dtype=np.float64)
        return (theta > 0) * (self.constant - .5 * (self.nu + 1) * np.log(1. + (1. / self.nu) * np.square(theta / self.A)))

    def lnpdf_grad(self, theta):
        theta = np.asarray(theta, dtype=np.float64)
        return (theta > 0) * (-(self.nu + 1.) * theta / (self.nu * self.A ** 2 + np.square(theta)))

    def rvs(self, n):
        from scipy.stats import t
        return np.abs(t.rvs(df=self.nu, loc=0, scale=self.A, size=int(n)))

    def to_dict(self):
        return {'class': 'GPy.core.parameterization.priors.HalfT', 'A': self.A, 'nu': self.nu}


class _NotOnScoringPath(Prior):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError('%s is a latent-variable-model prior of GPy, not a hyperprior of the GP scoring path'
                                  % type(self).__name__)


class MultivariateGaussian(_NotOnScoringPath):
    pass


class DGPLVM_KFDA(_NotOnScoringPath):
    pass


class DGPLVM_Lamda(_NotOnScoringPath):
    pass
