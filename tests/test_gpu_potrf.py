"""GPU parity of the standalone blocked Cholesky (b200gp_potrf: in-register 128x128 diagonal factorisation, DMMA
triangular solves and trailing updates) against LAPACK dpotrf -- what GPy's jitchol calls
(reference call site src/autoks/core/gp_model.py:64-65 via ExactGaussianInference)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def spd(n, seed, cond_boost=1e-3):
    rng = np.random.RandomState(seed)
    x = np.sort(rng.uniform(0, 8, size=n))
    d = x[:, None] - x[None, :]
    return np.exp(-0.5 * d * d / 0.6 ** 2) + np.cos(0.3 * d) * 0.2 + (0.2 + cond_boost) * np.eye(n)


@pytest.fixture(params=[1, 2, 0], ids=['lookahead', 'lookahead-split-barrier', 'blocked'])
def diag_mode(engine, request):
    """Both diagonal-tile factorisations (b200gp_set_potrf_la): look-ahead [A | I] elimination (default) and blocked."""
    engine.set_potrf_la(request.param)
    yield request.param
    engine.set_potrf_la(1)


@pytest.mark.parametrize('n', [128, 256, 640, 2048])
def test_potrf_matches_lapack(engine, n, diag_mode):
    import torch
    K = spd(n, n)
    A = torch.from_numpy(K.copy()).to(engine.device)
    logdet, status = engine.potrf(A)
    assert status == 0
    L = np.tril(A.cpu().numpy())
    ref = np.linalg.cholesky(K)
    assert abs(logdet - 2 * np.sum(np.log(np.diag(ref)))) <= 1e-10 * max(1.0, abs(logdet))
    assert np.max(np.abs(L - ref)) <= 1e-10 * np.max(np.abs(ref))
    assert np.max(np.abs(L @ L.T - K)) <= 1e-12 * n


def test_potrf_reports_first_bad_pivot(engine, diag_mode):
    import torch
    n = 256
    K = spd(n, 3)
    K[150, 150] = -1.0                         # leading minor 151 is not positive definite
    A = torch.from_numpy(K.copy()).to(engine.device)
    logdet, status = engine.potrf(A)
    assert status == 151 and np.isnan(logdet)
    K = spd(n, 4)
    K[7, 7] = np.nan
    A = torch.from_numpy(K.copy()).to(engine.device)
    logdet, status = engine.potrf(A)
    assert status == 8 and np.isnan(logdet)
