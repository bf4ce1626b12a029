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


# ---- kernel-matrix alignment metrics (src/autoks/core/covariance.py:278-338), restated verbatim in NumPy --------------
def inner_frob(m, n):
    return np.trace(m.T.conjugate() @ n)


def alignment(k1, k2):
    return inner_frob(k1, k2) / np.sqrt(inner_frob(k1, k1) * inner_frob(k2, k2))


def center_kernel(k):
    m = k.shape[0]
    ones = np.ones((m, 1))
    centering = np.eye(m) - (ones @ ones.T) / m
    return centering @ k @ centering


def centered_alignment(k1, k2):
    return alignment(center_kernel(k1), center_kernel(k2))


def pairwise_centered_alignments(kernel_matrices):
    n = len(kernel_matrices)
    dists = np.zeros((n, n))
    for i in range(n):
        for j in range(i + 1, n):
            dists[i, j] = centered_alignment(kernel_matrices[i], kernel_matrices[j])
    return (dists + dists.T) / 2.
