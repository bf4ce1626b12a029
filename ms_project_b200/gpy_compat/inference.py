"""`GPy.inference.latent_function_inference` names the reference passes around (`Laplace()` in
src/autoks/core/gp_models/gp_regression.py:23-25, the BOMS default at boms_model_selector.py:29-33).  With the Gaussian
likelihood of the scoring path the Laplace mode IS the exact posterior, so both are markers for the same device path
(hazard 4)."""


class LatentFunctionInference(object):
    def to_dict(self):
        return {'class': 'GPy.inference.latent_function_inference.%s' % type(self).__name__}


class ExactGaussianInference(LatentFunctionInference):
    pass


class Laplace(LatentFunctionInference):
    pass
