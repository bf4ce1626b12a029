"""CPU: pin the oracle against the reference's own golden vectors (GPML cross-checks in
notebooks/exploratory/test custom kernels/*.ipynb) and against finite differences / metric axioms."""
import json
import os

import numpy as np
import pytest

from oracle import kernels as ok, gp, model, hellinger, scoring

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'gpml_kernels.json')))
X5 = np.array(GOLD['X'], dtype=np.float64).reshape(-1, 1)
Y5 = np.array(GOLD['y'], dtype=np.float64).reshape(-1, 1)


@pytest.mark.parametrize('case', GOLD['cases'], ids=[c['kernel'] for c in GOLD['cases']])
def test_gpml_kernel_matrices(case):
    for mode in ('gpy', 'direct'):
        K = ok.K(ok.parse(case['kernel']), case['theta'], X5, dist_mode=mode)
        assert np.allclose(K, np.array(case['K']))          # the notebooks' own assertion (np.allclose defaults)
        assert np.max(np.abs(K - np.array(case['K']))) < 5e-8   # printed precision of the golden matrices


@pytest.mark.parametrize('name,theta,restarts', [('SE_1', [6.25, 1.5], 10), ('RQ_1', [6.25, 1.5, 2.1], 10),
                                                 ('PER_1', [6.25, 2.1, 1.5], 60)])
def test_gpml_nlml(name, theta, restarts):
    """rbf.ipynb#2 / rational-quadratic.ipynb#4 / periodic.ipynb#4: optimised NLML == GPML's at rtol 1e-3."""
    with np.errstate(all='ignore'):
        m = model.OracleGP(ok.parse(name), X5, Y5, theta=theta, noise_var=GOLD['nlml']['noise_init'])
        m.optimize_restarts(restarts, rng=np.random.RandomState(1))
    assert np.isclose(GOLD['nlml'][name], -m.log_likelihood(), rtol=GOLD['nlml']['rtol'])


@pytest.mark.parametrize('expr', ['SE_1', 'RQ_2', 'PER_1', 'LIN_2', 'CONST_1 + LIN_1', 'SE_1 * PER_1 + RQ_2',
                                  '(SE_1 + RQ_2) * LIN_2 * PER_1', 'SE_1 * SE_1 + SE_2'])
def test_gradients_match_finite_differences(expr):
    rng = np.random.RandomState(0)
    X = rng.randn(40, 2)
    y = rng.randn(40, 1)
    e = ok.parse(expr)
    th = np.exp(rng.randn(ok.n_params(e)) * 0.3)
    nv = 0.3
    _, g, _ = gp.lml_and_grad(e, th, nv, X, y)
    p = np.concatenate([th, [nv]])
    for i in range(p.size):
        h = 1e-6 * p[i]
        pp, pm = p.copy(), p.copy()
        pp[i] += h
        pm[i] -= h
        fd = (gp.lml_and_grad(e, pp[:-1], pp[-1], X, y)[0] - gp.lml_and_grad(e, pm[:-1], pm[-1], X, y)[0]) / (2 * h)
        assert abs(fd - g[i]) <= 1e-6 * max(1.0, abs(g[i])), (expr, i, fd, g[i])


def test_quirk_mode_reproduces_notebook_formulas():
    rng = np.random.RandomState(1)
    X = rng.randn(30, 1)
    y = rng.randn(30, 1)
    for expr, idx, factor in [('PER_1', 2, 0.25), ('LIN_1', 1, None)]:
        e = ok.parse(expr)
        th = np.array([1.3, 0.7, 0.9])[:ok.n_params(e)]
        g = gp.lml_and_grad(e, th, 0.2, X, y)[1]
        gq = gp.lml_and_grad(e, th, 0.2, X, y, quirks=True)[1]
        f = factor if factor is not None else 1.0 / th[0]
        assert np.isclose(gq[idx], g[idx] * f, rtol=1e-12)
        others = [i for i in range(g.size) if i != idx]
        assert np.allclose(gq[others], g[others], rtol=0, atol=0)


def test_parse_flattens_like_gpy():
    assert ok.parse('SE_1 + (RQ_1 + PER_1)') == ('+', [('SE', 0), ('RQ', 0), ('PER', 0)])
    assert ok.parse('SE_1 * RQ_1 * PER_2 + LIN_1') == ('+', [('*', [('SE', 0), ('RQ', 0), ('PER', 1)]), ('LIN', 0)])
    assert ok.to_postfix(ok.parse('SE_1 * SE_1 + RQ_1')) == [('SE', 0), ('SE', 0), '*', ('RQ', 0), '+']   # test_gp_model.py:16-21
    assert ok.to_infix(ok.parse('(SE_1 + RQ_2) * LIN_2')) == '(SE_1 + RQ_2) * LIN_2'


def test_jitchol_ladder():
    A = np.ones((6, 6)) * 1e12
    with pytest.raises(np.linalg.LinAlgError):
        np.linalg.cholesky(A)
    L, tries = gp.jitchol(A)
    assert tries == 1 and np.allclose(L @ L.T, A + np.eye(6) * 1e6)
    with pytest.raises(gp.NotPD):
        gp.jitchol(-np.eye(3))


def test_logexp_and_prior_terms():
    x = np.array([-800., -5., 0., 3., 40.])
    f = model.logexp_f(x)
    assert np.all(f >= 0) and f[-1] == 40.
    xr = model.logexp_finv(f[1:])
    assert np.allclose(xr, x[1:])
    p = ('lognormal', np.log(0.4), 1.0)
    v = 0.7
    h = 1e-6
    assert np.isclose((model.prior_lnpdf(p, v + h) - model.prior_lnpdf(p, v - h)) / (2 * h), model.prior_lnpdf_grad(p, v))
    assert np.isclose((model.logexp_log_jacobian(v + h) - model.logexp_log_jacobian(v - h)) / (2 * h),
                      model.logexp_log_jacobian_grad(v))


def test_objective_gradient_with_priors():
    rng = np.random.RandomState(2)
    X = rng.randn(30, 1)
    y = np.sin(X) + 0.1 * rng.randn(30, 1)
    e = ok.parse('SE_1 + PER_1 * LIN_1')
    kp, npri = model.boms_priors(e)
    m = model.OracleGP(e, X, y, kern_priors=kp, noise_prior=npri)
    x0 = m.optimizer_array.copy() + 0.1
    f0, g0 = m._objective_grads(x0)
    for i in range(x0.size):
        xp, xm = x0.copy(), x0.copy()
        xp[i] += 1e-6
        xm[i] -= 1e-6
        fd = (m._objective_grads(xp)[0] - m._objective_grads(xm)[0]) / 2e-6
        assert abs(fd - g0[i]) <= 1e-5 * max(1.0, abs(g0[i]))


def test_scores():
    """model_selection_criteria.py:9-20, fitness_functions.py:7-27."""
    assert scoring.BIC(-10.0, 100, 4) == np.log(100) * 4 + 20.0
    assert scoring.negative_bic(-10.0, 100, 4) == -scoring.BIC(-10.0, 100, 4)
    assert scoring.log_likelihood_normalized(-10.0, 100) == -0.1
    x = np.arange(12.).reshape(4, 3)
    s = scoring.standardize(x)
    assert np.allclose(s.mean(0), 0) and np.allclose(s.std(0), 1)


# ---- Hellinger: the metric axioms the reference tests (src/tests/unit/autoks/test_distance.py:19-98,305-318)
def _pool():
    fams = ['SE', 'RQ']
    base = [(f, d) for f in fams for d in range(2)]
    pool = list(base)
    for a in base:
        for b in base:
            pool.append(('+', [a, b]))
            pool.append(('*', [a, b]))
    return pool[:24]


def _halton(n, d):
    primes = [2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53, 59, 61, 67, 71]
    out = np.zeros((n, d))
    for j in range(d):
        b = primes[j]
        for i in range(n):
            f, r, k = 1.0, 0.0, i + 1
            while k > 0:
                f /= b
                r += f * (k % b)
                k //= b
            out[i, j] = r
    return out


def test_hellinger_metric_axioms():
    x = np.array([[1, 2], [4, 5], [6, 7], [8, 9], [10, 11]], dtype=np.float64)    # test_distance.py:19
    prob = _halton(20, 12)
    lam = hellinger.noise_samples(('gaussian', np.log(0.01), 1.0), prob)
    infos = []
    for e in _pool():
        kp, _ = model.boms_priors(e)
        infos.append(hellinger.create_precomputed_info(e, kp, x, prob, lam))
    D = hellinger.distance_matrix(infos)
    n = len(infos)
    assert np.all(D >= 0) and np.all(D <= 1)
    assert np.allclose(np.diag(D), 0, atol=1e-3)
    assert np.allclose(D, D.T, atol=1e-2)
    for i in range(n):
        for j in range(n):
            assert np.all(D[i, j] <= D[i, :] + D[:, j] + 1e-3)
    assert D[0, 2] > 0.01          # SE_1 vs RQ_1: different kernels are at positive distance


def test_prior_sample_column_reuse():
    """util.py:60-62: each prior family indexes prob_samples from column 0."""
    prob = _halton(5, 4)
    pri = [('lognormal', 0.0, 1.0), ('gaussian', 1.0, 2.0), ('lognormal', 0.5, 0.5)]
    h = hellinger.prior_sample(pri, prob)
    from scipy.stats import norm
    assert np.allclose(h[:, 1], norm.ppf(prob[:, 0], 1.0, 2.0))
    assert np.allclose(h[:, 0], np.exp(norm.ppf(prob[:, 0], 0.0, 1.0)))
    assert np.allclose(h[:, 2], np.exp(norm.ppf(prob[:, 1], 0.5, 0.5)))


def test_oracle_lml_and_gradient_match_scikit_learn():
    """Independent pin of the oracle's exact-inference arithmetic: scikit-learn's GaussianProcessRegressor evaluates the
    same log marginal likelihood (Rasmussen & Williams alg. 2.1) and its gradient for kernels that coincide with the
    fork's definitions -- RBF = SE, RationalQuadratic = RQ (alpha = power), ExpSineSquared = PER (exp(-2 sin^2(pi d / p) / l^2),
    periodic.ipynb#2), sums and products -- with WhiteKernel as the Gaussian noise.  sklearn differentiates w.r.t. log(theta),
    so its gradient is theta * dLML/dtheta.  (GPy adds 1e-8 to the diagonal, sklearn 1e-10: the noise level absorbs the
    difference here.)"""
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import RBF, RationalQuadratic, ExpSineSquared, ConstantKernel as C, WhiteKernel
    from oracle import gp as ogp, kernels as ok
    rng = np.random.RandomState(7)
    X = rng.randn(60, 2)
    y = np.sin(X[:, :1]) + 0.3 * X[:, 1:] + 0.1 * rng.randn(60, 1)
    # 1-D leaves act on one column: sklearn kernels are isotropic over all columns, so use the column as the input
    x0 = X[:, :1]
    cases = [
        ('SE_1', [1.7, 0.6], lambda t: C(t[0]) * RBF(t[1])),
        ('RQ_1', [0.8, 1.3, 0.7], lambda t: C(t[0]) * RationalQuadratic(length_scale=t[1], alpha=t[2])),
        ('PER_1', [1.2, 2.1, 0.9], lambda t: C(t[0]) * ExpSineSquared(length_scale=t[2], periodicity=t[1])),
        ('SE_1 + RQ_1', [1.7, 0.6, 0.8, 1.3, 0.7],
         lambda t: C(t[0]) * RBF(t[1]) + C(t[2]) * RationalQuadratic(length_scale=t[3], alpha=t[4])),
        ('SE_1 * PER_1', [1.7, 0.6, 1.2, 2.1, 0.9],
         lambda t: (C(t[0]) * RBF(t[1])) * (C(t[2]) * ExpSineSquared(length_scale=t[4], periodicity=t[3]))),
    ]
    noise = 0.05
    for name, th, make in cases:
        expr = ok.parse(name)
        lml, grad, _ = ogp.lml_and_grad(expr, np.array(th), noise, x0, y, dist_mode='direct')
        kernel = make(th) + WhiteKernel(noise - 1e-10 + 1e-8)
        gpr = GaussianProcessRegressor(kernel=kernel, alpha=1e-10, optimizer=None).fit(x0, y)
        ref, ref_g = gpr.log_marginal_likelihood(gpr.kernel_.theta, eval_gradient=True)
        assert abs(lml - ref) <= 1e-9 * max(1.0, abs(ref)), (name, lml, ref)
        # map sklearn's log-space gradient (kernel_.theta order) onto the oracle's parameter order by name
        names = [h.name for h in gpr.kernel_.hyperparameters]
        vals = np.exp(gpr.kernel_.theta)
        sk = {}
        for n_, v_, g_ in zip(names, vals, ref_g):
            sk.setdefault(round(float(v_), 12), []).append(g_ / v_)
        got = np.concatenate([grad[:-1], grad[-1:]])
        want_vals = list(th) + [noise - 1e-10 + 1e-8]
        for gv, val in zip(got, want_vals):
            cands = sk[round(float(val), 12)]
            assert any(abs(gv - c) <= 1e-6 * max(1.0, abs(c)) for c in cands), (name, val, gv, cands)
