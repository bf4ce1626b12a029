"""Oracle: BOMS Hellinger kernel-kernel distance.  TEST INFRASTRUCTURE.

Restates src/autoks/distance/distance.py:126-194 (HellingerDistanceBuilder), :110-118
(compute_distance), :300-312 (chol_safe) and src/autoks/distance/util.py:21-71 (prior_sample*).
`cov_nearest` (statsmodels, absent here) is the reference's non-PD repair; the oracle raises
instead and callers treat that as the status-flag path."""
import numpy as np
from scipy.stats import norm

from . import kernels


def prior_sample(priors, prob_samples):
    """distance/util.py:21-71.  priors: list of (kind, mu, sigma); prob_samples: S x max_hyp.
    Column i of the result uses column i of prob_samples (`prob_samples[:, :num_hyp]` is applied
    separately to the gaussian and the log-gaussian sub-lists, each starting at column 0 --
    util.py:60-62 -- so with mixed priors the columns are re-used; restated exactly)."""
    priors = list(priors)
    S = prob_samples.shape[0]
    hyps = np.full((S, len(priors)), np.nan)
    g_ind = [i for i, p in enumerate(priors) if p[0] == 'gaussian']
    lg_ind = [i for i, p in enumerate(priors) if p[0] == 'lognormal']
    if len(g_ind) + len(lg_ind) < len(priors):
        raise ValueError('Unsupported prior found in priors')

    def gaussian(sub):
        mu = np.array([p[1] for p in sub])
        sd = np.array([p[2] for p in sub])
        return norm.ppf(prob_samples[:, :len(sub)], np.tile(mu, (S, 1)), np.tile(sd, (S, 1)))

    if g_ind:
        hyps[:, g_ind] = gaussian([priors[i] for i in g_ind])
    if lg_ind:
        hyps[:, lg_ind] = np.exp(gaussian([priors[i] for i in lg_ind]))
    return hyps


def noise_samples(noise_prior, prob_samples):
    """DistanceBuilder.__init__ (distance.py:41-48): exp(ppf) of a *Gaussian* noise prior."""
    assert noise_prior[0] == 'gaussian'
    return np.exp(prior_sample([noise_prior], prob_samples))[:, 0]


def chol_safe(k):
    """distance.py:300-312 without the cov_nearest repair (raises LinAlgError instead)."""
    return np.linalg.cholesky(k).T


def create_precomputed_info(expr, kern_priors, data_X, prob_samples, lam, dist_mode='gpy'):
    """distance.py:170-194 -> (log_det[S], mini_gram[n, n, S])."""
    n = data_X.shape[0]
    S = prob_samples.shape[0]
    log_det = np.full(S, np.nan)
    mini = np.full((n, n, S), np.nan)
    hyps = prior_sample(kern_priors, prob_samples)
    for i in range(S):
        k = kernels.K(expr, hyps[i], data_X, data_X, dist_mode)
        k = k + lam[i] * np.eye(n)
        mini[:, :, i] = k
        log_det[i] = 2 * np.sum(np.log(np.diag(chol_safe(k))), axis=0)
    return log_det, mini


def hellinger_distance(log_det_i, mini_i, log_det_j, mini_j, tol=0.02):
    """distance.py:135-168, including the skip of samples whose logdets differ by <= tol."""
    are_different = np.abs(log_det_i - log_det_j) > tol
    logdet_pq = log_det_i.copy()
    for s in np.arange(are_different.size)[are_different]:
        pq = 0.5 * (mini_i[:, :, s] + mini_j[:, :, s])
        logdet_pq[s] = 2 * np.sum(np.log(np.diag(chol_safe(pq))), axis=0)
    log_h = 0.25 * (log_det_i + log_det_j) - 0.5 * logdet_pq
    h = 1 - np.exp(log_h)
    d = np.clip(np.mean(h, axis=0), 0, 1)
    return float(np.sqrt(d))


def distance_matrix(infos, max_n=None):
    """DistanceBuilder._average_distance for all pairs (distance.py:50-51,110-118)."""
    m = len(infos)
    max_n = max_n or m
    D = np.full((max_n, max_n), np.nan)
    np.fill_diagonal(D, 0)
    for i in range(m):
        for j in range(i + 1, m):
            D[i, j] = D[j, i] = hellinger_distance(*infos[i], *infos[j])
    return D


# ---- Frobenius / correlation builders (src/autoks/distance/distance.py:197-284), float32 like the reference ----------
def create_vectors(expr, kern_priors, data_X, prob_samples, lam, dist_mode='gpy'):
    """create_precomputed_info of FrobeniusDistanceBuilder / CorrelationDistanceBuilder: (n^2 x S) float32."""
    _ld, G = create_precomputed_info(expr, kern_priors, data_X, prob_samples, lam, dist_mode)
    n = G.shape[0]
    return G.reshape(n * n, G.shape[2]).astype(np.float32)


def frobenius_distance(a, b):
    return float(np.mean(np.sqrt(np.sum((a - b) ** 2, axis=0))))


def correlation_distance(a, b):
    a_centered = a - np.mean(a, axis=0)
    b_centered = b - np.mean(b, axis=0)
    dot_prod = np.einsum('ij,ji->i', a_centered.T, b_centered)
    correlation = dot_prod / (np.linalg.norm(a_centered, axis=0) * np.linalg.norm(b_centered, axis=0))
    correlation = np.clip(correlation, 0, 1)
    return float(np.mean(np.sqrt(0.5 * (1 - correlation)), axis=0))
