"""`gp_fn` factories -- the reference's own plug-in point for the GP backend
(src/autoks/core/gp_models/gp_regression.py:10-27, resolved by name at base.py:99-110)."""
from .gpy_compat import likelihoods


def gp_regression(inference_method=None, likelihood_hyperprior=None):
    """kwargs for GP(...): Gaussian likelihood (+ hyperprior) and the inference method tag."""
    likelihood = likelihoods.Gaussian()
    if likelihood_hyperprior is not None:
        for name, prior in likelihood_hyperprior.items():
            if name in likelihood.parameter_names():
                getattr(likelihood, name).set_prior(getattr(prior, 'raw_prior', prior), warning=False)
    return {'likelihood': likelihood, 'inference_method': inference_method}
