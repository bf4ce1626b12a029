"""Oracle: exact Gaussian GP inference as GPy performs it.  TEST INFRASTRUCTURE.

Restates (from the public upstream source as recalled; GPy is not importable here)
`GPy/inference/latent_function_inference/exact_gaussian_inference.py:inference`,
`GPy/util/linalg.py:{jitchol,pdinv,dtrtri,dpotri,dpotrs}` and `GP.parameters_changed`.
Reference call sites: src/autoks/core/gp_model.py:54,64-65 (build + optimise),
src/autoks/backend/model.py:95-108 (n_data / n_params / log_likelihood).
"""
import numpy as np
from scipy.linalg import lapack

from . import kernels

LOG_2_PI = np.log(2 * np.pi)
EXACT_INFERENCE_JITTER = 1e-8     # `diag.add(Ky, variance + 1e-8)` in exact_gaussian_inference.py


class NotPD(np.linalg.LinAlgError):
    pass


def jitchol(A, maxtries=5):
    """GPy util.linalg.jitchol: dpotrf; on failure add mean(diag)*1e-6*10^k, k = 0..maxtries-1.
    Returns (L, n_jitter_tries) -- tries is 0 when the plain factorisation succeeded."""
    A = np.ascontiguousarray(A)
    L, info = lapack.dpotrf(A, lower=1)
    if info == 0:
        return L, 0
    diagA = np.diag(A)
    if np.any(diagA <= 0.):
        raise NotPD('not pd: non-positive diagonal elements')
    jitter = diagA.mean() * 1e-6
    num_tries = 1
    while num_tries <= maxtries and np.isfinite(jitter):
        L, info = lapack.dpotrf(np.ascontiguousarray(A + np.eye(A.shape[0]) * jitter), lower=1)
        if info == 0:
            return L, num_tries
        jitter *= 10
        num_tries += 1
    raise NotPD('not positive definite, even with jitter.')


def pdinv(A):
    """GPy util.linalg.pdinv -> (Ai, L, Li, logdet, tries)."""
    L, tries = jitchol(A)
    logdet = 2. * np.sum(np.log(np.diag(L)))
    Li, _ = lapack.dtrtri(L, lower=1)
    Ai, _ = lapack.dpotri(L, lower=1)
    Ai = np.tril(Ai) + np.tril(Ai, -1).T     # symmetrify
    return Ai, L, Li, logdet, tries


def inference(Kmat, y, noise_var):
    """ExactGaussianInference.inference for a single-output GP with zero mean.
    Returns dict(lml, dL_dK, dL_dnoise, alpha, L, logdet, tries)."""
    y = np.asarray(y, dtype=np.float64).reshape(-1, 1)
    Ky = Kmat.copy()
    Ky[np.diag_indices_from(Ky)] += noise_var + EXACT_INFERENCE_JITTER
    Wi, LW, _, W_logdet, tries = pdinv(Ky)
    alpha, _ = lapack.dpotrs(LW, y, lower=1)
    lml = 0.5 * (-y.size * LOG_2_PI - y.shape[1] * W_logdet - np.sum(alpha * y))
    dL_dK = 0.5 * (alpha @ alpha.T - y.shape[1] * Wi)
    return dict(lml=float(lml), dL_dK=dL_dK, dL_dnoise=float(np.sum(np.diag(dL_dK))),
                alpha=alpha, L=LW, logdet=float(W_logdet), tries=tries)


def lml_and_grad(expr, theta, noise_var, X, y, dist_mode='gpy', quirks=False):
    """LML and its gradient wrt the constrained parameters [theta_kernel..., noise_var]
    (GP links kern first, likelihood second => noise is last; GP.parameters_changed).
    Raises NotPD exactly where GPy would raise LinAlgError."""
    Kmat = kernels.K(expr, theta, X, None, dist_mode)
    inf = inference(Kmat, y, noise_var)
    g_k = kernels.grad_theta(expr, theta, X, inf['dL_dK'], dist_mode, quirks)
    return inf['lml'], np.concatenate([g_k, [inf['dL_dnoise']]]), inf


def predict(expr, theta, noise_var, X, y, Xs, dist_mode='gpy', include_likelihood=True):
    """GP.predict: posterior mean and variance (diag) at Xs (GPy Posterior._raw_predict + Gaussian
    likelihood predictive_values), as recalled."""
    Kmat = kernels.K(expr, theta, X, None, dist_mode)
    inf = inference(Kmat, y, noise_var)
    Kxs = kernels.K(expr, theta, X, Xs, dist_mode)
    mu = Kxs.T @ inf['alpha']
    Kss = np.diag(kernels.K(expr, theta, Xs, Xs, 'direct')).copy()
    tmp, _ = lapack.dtrtrs(inf['L'], Kxs, lower=1)
    var = (Kss - np.sum(np.square(tmp), 0)).reshape(-1, 1)
    var = np.clip(var, 1e-15, np.inf)
    if include_likelihood:
        var = var + noise_var
    return mu, var
