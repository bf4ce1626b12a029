"""GPy-compatible object surface for the autoks scoring path (seam B3).

GPy (the fork pinned at Pipfile.lock:236-239) cannot be installed here; these lightweight stand-ins
present the attributes the reference's backend shim uses and delegate every number to libb200gp."""
from . import kern, priors, transformations  # noqa: F401
from .kern import Kern, RBF, RationalQuadratic, StandardPeriodic, LinScaleShift, Bias, Add, Prod, \
    RBFDistanceBuilderKernelKernel  # noqa: F401
from .priors import Prior, Gaussian, LogGaussian  # noqa: F401
from .likelihoods import Likelihood  # noqa: F401,E402
from . import likelihoods  # noqa: F401,E402
from .gp import GP, GPRegression  # noqa: F401,E402
from .inference import Laplace, ExactGaussianInference  # noqa: F401,E402
