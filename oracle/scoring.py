"""Oracle: model-selection scores.  TEST INFRASTRUCTURE.
Follows src/autoks/core/fitness_functions.py:7-27 and src/autoks/model_selection_criteria.py:9-46."""
import numpy as np


def BIC(lml, n, k):
    """model_selection_criteria.py:9-20 -- ln(n) k - 2 ln(L)."""
    return np.log(n) * k - 2 * lml


def AIC(lml, k):
    """model_selection_criteria.py:23-33."""
    return 2 * k - 2 * lml


def pl2(lml, n, k):
    """model_selection_criteria.py:36-46."""
    return -lml / n + k / (2 * n)


def negative_bic(lml, n, k):
    """fitness_functions.py:25-27."""
    return -BIC(lml, n, k)


def log_likelihood_normalized(lml, n):
    """fitness_functions.py:7-13."""
    return lml / n


def standardize(x):
    """ModelSelector._prepare_data (src/autoks/core/model_selection/base.py:206-220):
    population std (ddof=0), per column."""
    x = np.asarray(x, dtype=np.float64)
    return (x - np.mean(x, axis=0)) / np.std(x, ddof=0, axis=0)
