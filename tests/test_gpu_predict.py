"""GPU parity of posterior prediction (GP.predict as ModelSelector.predict / the BOMS surrogate
use it) against the oracle's restatement of GPy's Posterior._raw_predict + Gaussian.predictive_values."""
import numpy as np
import pytest

from conftest import to_oracle_expr, from_string, make_data

pytestmark = pytest.mark.gpu

KERNELS = ['SE_1', 'RQ_2 + LIN_1', 'SE_1 * PER_1 + RQ_2', '(SE_1 + RQ_2) * (PER_1 + LIN_2 + SE_2) * RQ_1',
           '(SE_1 + RQ_1 + PER_1) * (SE_2 + RQ_2 + LIN_2) * (PER_2 + LIN_1 + SE_2)']


@pytest.mark.parametrize('n,ns', [(5, 3), (100, 300), (300, 77), (513, 200)])
def test_predict_matches_oracle(engine, n, ns):
    from ms_project_b200.program import ProgramBatch
    from oracle import gp
    X, y = make_data(n, 2, seed=n)
    rng = np.random.RandomState(ns)
    Xs = rng.randn(ns, 2) * 1.2
    kernels = [from_string(s) for s in KERNELS]
    batch = ProgramBatch(kernels)
    engine.bind(X, y, slots=2, max_params=batch.max_params)
    engine.upload_programs(batch)
    ks = [np.exp(rng.uniform(-0.7, 0.7, size=int(p))) for p in batch.n_params]
    nz = np.exp(rng.uniform(-3, -0.5, size=batch.n_items))
    theta = batch.pack_theta(ks, nz)
    for i, k in enumerate(kernels):
        for incl in (True, False):
            mean, var, lml, status = engine.posterior(batch, theta, i, Xs, include_likelihood=incl)
            mu_ref, var_ref = gp.predict(to_oracle_expr(k), ks[i], nz[i], X, y, Xs, include_likelihood=incl)
            ref_lml = gp.lml_and_grad(to_oracle_expr(k), ks[i], nz[i], X, y)[0]
            assert status == 0
            assert abs(lml - ref_lml) <= 1e-8 * max(1.0, abs(ref_lml))
            scale = max(1.0, np.max(np.abs(mu_ref)))
            assert np.max(np.abs(mean - mu_ref[:, 0])) <= 1e-8 * scale, (KERNELS[i], np.max(np.abs(mean - mu_ref[:, 0])))
            # var = k** - |L^-1 k*|^2 cancels: absolute error ~ eps * cond-ish * k**; compare on the prior-variance scale
            vscale = max(1.0, np.max(np.abs(var_ref)))
            assert np.max(np.abs(var - var_ref[:, 0])) <= 1e-8 * vscale, (KERNELS[i], np.max(np.abs(var - var_ref[:, 0])))
            assert np.all(var >= 1e-15)


def test_gp_predict_interface_and_normalizer(engine):
    """GP.predict through the GPy stand-in: shapes, include_likelihood, the Standardize normaliser of
    GPRegression(normalizer=True) (gp_regression_models.py:79-80), ModelSelector.predict's call pattern."""
    from ms_project_b200.gpy_compat import RBF, GP
    from ms_project_b200.gpy_compat.likelihoods import Gaussian
    from oracle import gp as ogp
    X, y = make_data(60, 1, seed=4)
    y = 3.0 * y + 5.0
    Xs = np.linspace(-2, 2, 25)[:, None]
    k = RBF(1, variance=1.3, lengthscale=0.7, active_dims=[0])
    lik = Gaussian(variance=0.05)
    m = GP(X, y, k, likelihood=lik, normalizer=True)
    mu, var = m.predict(Xs)
    assert mu.shape == (25, 1) and var.shape == (25, 1)
    ym, ys = y.mean(), y.std()
    mu_ref, var_ref = ogp.predict(to_oracle_expr(k), k.param_array, 0.05, X, (y - ym) / ys, Xs)
    assert np.allclose(mu, mu_ref * ys + ym, rtol=1e-8, atol=1e-8)
    assert np.allclose(var, var_ref * ys ** 2, rtol=1e-7, atol=1e-9)
    mu2, var2 = m.predict(Xs, include_likelihood=False)
    assert np.allclose(var - var2, 0.05 * ys ** 2, rtol=1e-9)
    with pytest.raises(NotImplementedError):
        m.predict(Xs, full_cov=True)
    # f_max of the surrogate: prediction at the training inputs (gp_regression_models.py:130-134)
    assert np.isfinite(m.predict(m.X)[0].max())
