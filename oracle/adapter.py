"""Oracle-side adapter: any GPy-style kernel object (duck-typed: `.parts`, class names Add / Prod,
`.active_dims`, and the family read from the class name) -> oracle expression tree.  TEST INFRASTRUCTURE.
It imports nothing from the product package."""

_FAMILY = {'RBF': 'SE', 'RationalQuadratic': 'RQ', 'StandardPeriodic': 'PER', 'LinScaleShift': 'LIN', 'Bias': 'CONST',
           'RBFDistanceBuilderKernelKernel': 'KK'}


def to_oracle_expr(kernel):
    name = type(kernel).__name__
    if name == 'Add':
        return ('+', [to_oracle_expr(p) for p in kernel.parts])
    if name == 'Prod':
        return ('*', [to_oracle_expr(p) for p in kernel.parts])
    return (_FAMILY[name], int(kernel.active_dims[0]))


def prior_tuple(prior):
    """GPy-style prior object -> oracle prior tuple."""
    if prior is None:
        return None
    kind = {'LogGaussian': 'lognormal', 'Gaussian': 'gaussian'}[type(prior).__name__]
    return (kind, float(prior.mu), float(prior.sigma))
