"""`GPy.likelihoods.Gaussian` stand-in (the only likelihood `gp_regression` builds,
src/autoks/core/gp_models/gp_regression.py:10-27)."""
from .param import Param, Parameterized


class Likelihood(Parameterized):
    @staticmethod
    def from_dict(input_dict):
        import copy
        d = copy.deepcopy(input_dict)
        cls = d.pop('class')
        if cls != 'GPy.likelihoods.Gaussian':
            raise NotImplementedError('only the Gaussian likelihood is on the scoring path (got %s)' % cls)
        d.pop('gp_link_dict', None)
        name = d.pop('name', 'Gaussian_noise')
        var = d.pop('variance', [1.0])
        return Gaussian(variance=var, name=str(name))


class Gaussian(Likelihood):
    """y = f + eps, eps ~ N(0, variance); `variance` is Logexp-constrained, default 1.0."""

    def __init__(self, gp_link=None, variance=1., name='Gaussian_noise'):
        super(Gaussian, self).__init__(name)
        self.link_parameter(Param('variance', variance))

    def gaussian_variance(self, Y_metadata=None):
        return float(self.variance)

    def to_dict(self):
        return {'class': 'GPy.likelihoods.Gaussian', 'name': self.name,
                'gp_link_dict': {'class': 'GPy.likelihoods.link_functions.Identity'},
                'variance': self.variance.values.tolist()}
