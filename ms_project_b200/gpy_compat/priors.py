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
