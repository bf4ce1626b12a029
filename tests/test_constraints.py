"""CPU: parameter constraints beyond Logexp -- Logistic (`constrain_bounded`) and fixed (`constrain_fixed`) -- through the
host fit path, as KernelKernelGPModel needs them (src/autoks/gp_regression_models.py:93-102).  The oracle stands in for the
device (tests/oracle_engine.py)."""
import numpy as np

from conftest import from_string, make_data


def test_transform_table_round_trip_and_gradfactor():
    from ms_project_b200.fit import _TransformTable
    from ms_project_b200.gpy_compat.transformations import Logexp, Logistic, Fixed
    tf = _TransformTable([Logexp(), Logistic(1e-3, 5.0), Fixed(), Logexp()], np.array([0.7, 2.0, 0.25, 40.0]))
    theta = np.array([0.7, 2.0, 0.25, 40.0])
    x = tf.finv(theta)
    assert np.allclose(tf.f(x), theta, rtol=1e-12)
    h = 1e-6
    g = np.array([1.3, -0.4, 2.0, 0.9])
    num = np.array([(tf.f(x + h * e) - tf.f(x - h * e)) @ g / (2 * h) for e in np.eye(4)])
    assert np.allclose(tf.gradfactor(theta, g), num, rtol=1e-6, atol=1e-9)
    assert tf.gradfactor(theta, g)[2] == 0.0 and tf.f(x + 3.0)[2] == 0.25          # the fixed coordinate never moves
    lo_hi = tf.f(np.array([0.0, -800.0, 0.0, 0.0]))[1], tf.f(np.array([0.0, 800.0, 0.0, 0.0]))[1]
    assert lo_hi[0] >= 1e-3 and lo_hi[1] <= 5.0


def test_fit_with_fixed_and_bounded_noise():
    import oracle_engine
    from ms_project_b200.fit import fit_models
    from ms_project_b200.gpy_compat.gp import GP
    from ms_project_b200.gpy_compat.likelihoods import Gaussian
    X, y = make_data(40, 1, seed=2)
    eng = oracle_engine.OracleEngine()
    eng.bind(X, y, 4, 4)
    # fixed noise: stays where it was put, does not count as a parameter, the kernel is still fitted
    k, lik = from_string('SE_1'), Gaussian(variance=1.0)
    lik.variance.constrain_fixed(0.05)
    np.random.seed(0)
    res = fit_models(eng, [k], [lik], num_restarts=2)[0]
    assert float(lik.variance) == 0.05 and np.isfinite(res.lml)
    gp = GP(X, y, k, lik)
    assert gp._size_transformed() == 2 and gp.size == 3
    assert not np.allclose(k.param_array, [1.0, 1.0])
    # bounded noise: the optimum respects the bounds and equals the free (Logexp) optimum when that lies inside them
    k1, l1 = from_string('SE_1'), Gaussian(variance=1.0)
    l1.variance.constrain_bounded(1e-9, 1e6)
    k2, l2 = from_string('SE_1'), Gaussian(variance=1.0)
    np.random.seed(1)
    r1 = fit_models(eng, [k1], [l1], num_restarts=2)[0]
    np.random.seed(1)
    r2 = fit_models(eng, [k2], [l2], num_restarts=2)[0]
    assert 1e-9 <= float(l1.variance) <= 1e6
    assert abs(r1.lml - r2.lml) <= 1e-5 * abs(r2.lml)
    l3 = Gaussian(variance=1.0)
    l3.variance.constrain_bounded(0.5, 2.0)
    np.random.seed(1)
    fit_models(eng, [from_string('SE_1')], [l3], num_restarts=2)
    assert 0.5 <= float(l3.variance) <= 2.0


def test_kernel_kernel_model_noise_branches():
    """gp_regression_models.py:93-102: fixed at 1e-6 for exact evaluations, positive with hyperpriors, bounded without."""
    import oracle_engine
    from ms_project_b200.boms import KernelKernelGPModel
    from ms_project_b200.covariance import Covariance
    from ms_project_b200.gpy_compat import RBFDistanceBuilderKernelKernel
    from ms_project_b200.gpy_compat.transformations import Logexp, Logistic, Fixed
    from ms_project_b200 import workloads

    class Builder(object):
        def __init__(self, n):
            rng = np.random.RandomState(0)
            pts = rng.rand(n, 2)
            self.D = np.sqrt(((pts[:, None, :] - pts[None, :, :]) ** 2).sum(-1))

        def get_kernel(self, n):
            return self.D[:n, :n]
    x = np.arange(6)[:, None]
    yv = np.random.RandomState(1).randn(6, 1)
    hp = workloads.boms_hyperpriors()
    with oracle_engine.installed():
        for kwargs, kind in ((dict(exact_f_eval=True), Fixed), (dict(kernel_kernel_hyperpriors={'SE': hp['SE'], 'GP': hp['GP']}), Logexp),
                             (dict(), Logistic)):
            kk = Covariance(RBFDistanceBuilderKernelKernel(Builder(6), n_models=6))
            m = KernelKernelGPModel(kk, verbose=False, optimize_restarts=2, **kwargs)
            np.random.seed(3)
            m.update(x, yv, None, None)
            noise = m.model.Gaussian_noise.variance
            assert isinstance(noise.constraint, kind), (kwargs, type(noise.constraint))
            if kind is Fixed:
                assert float(noise) == 1e-6
            mu, sd = m.predict(x)
            assert mu.shape == (6, 1) and np.all(np.isfinite(mu)) and np.all(sd > 0)
