"""GPU parity of the Frobenius / correlation kernel-kernel distances (what SMBOModelSelector wires,
src/autoks/core/model_selection/smbo_model_selector.py:91) against the oracle's float32 restatement of
src/autoks/distance/distance.py:197-284, plus the reference tests' axioms (test_distance.py:100-182)."""
import numpy as np
import pytest

from oracle import hellinger as ohell
from oracle.adapter import to_oracle_expr, prior_tuple

pytestmark = pytest.mark.gpu


class _Model(object):
    def __init__(self, covariance):
        self.covariance = covariance
        self.info = None


def _setup(cls, x, n_models, names, n_dims):
    from ms_project_b200 import workloads
    from ms_project_b200.covariance import Covariance
    from ms_project_b200.distance import ActiveModels, probability_samples
    from ms_project_b200.gpy_compat.priors import Gaussian
    hp = workloads.boms_hyperpriors()
    bases = workloads.get_all_1d_kernels(names, n_dims, hp)
    pool = (bases + workloads.expand_full_brute_force(bases, 2, n_models))[:n_models]
    am = ActiveModels(max_n_models=n_models)
    for k in pool:
        am.add(_Model(Covariance(k)))
    prob = probability_samples(40, 20)
    noise_prior = Gaussian(np.log(0.01), 1)
    builder = cls(noise_prior, 20, 40, n_models, am, list(range(len(pool))), x, probability_samples_matrix=prob)
    return am, builder, prob, noise_prior, pool


@pytest.mark.parametrize('which', ['frobenius', 'correlation'])
def test_pair_distances_match_oracle(which):
    from ms_project_b200.distance import FrobeniusDistanceBuilder, CorrelationDistanceBuilder
    cls = FrobeniusDistanceBuilder if which == 'frobenius' else CorrelationDistanceBuilder
    ref_fn = ohell.frobenius_distance if which == 'frobenius' else ohell.correlation_distance
    rng = np.random.default_rng(3)
    x = rng.random((64, 3))
    x = (x - x.mean(0)) / x.std(0)
    am, builder, prob, noise_prior, pool = _setup(cls, x, 28, ['SE', 'RQ'], 3)
    lam = ohell.noise_samples(prior_tuple(noise_prior), prob)
    vecs = []
    for k, m in zip(pool, am.models):
        kp = [prior_tuple(p.prior) for p in k.flattened_parameters()]
        v_ref = ohell.create_vectors(to_oracle_expr(k), kp, x, prob, lam)
        v = np.asarray(m.info)
        assert v.dtype == np.float32 and v.shape == v_ref.shape == (64 * 64, 20)
        assert np.allclose(v, v_ref, rtol=2e-6, atol=1e-7)
        vecs.append(v_ref)
    n = len(pool)
    idx = list(range(n))
    builder.compute_distance(am, idx, idx)
    D = builder.get_kernel(n)
    D_ref = np.array([[ref_fn(vecs[i], vecs[j]) for j in range(n)] for i in range(n)])
    assert not np.isnan(D).any() and np.allclose(D, D.T)
    off = ~np.eye(n, dtype=bool)
    # (1) the same formulas evaluated in float64 on the reference's float32 vectors: what the device computes, tightly
    v64 = [v.astype(np.float64) for v in vecs]
    if which == 'frobenius':
        D64 = np.array([[np.mean(np.sqrt(np.sum((vecs[i] - vecs[j]).astype(np.float64) ** 2, axis=0))) for j in range(n)]
                        for i in range(n)])
    else:
        D64 = np.array([[ref_fn(v64[i], v64[j]) for j in range(n)] for i in range(n)])
    assert np.allclose(D[off], D64[off], rtol=1e-9, atol=1e-7), np.max(np.abs(D[off] - D64[off]))
    # (2) the reference's own float32 arithmetic: agreement down to ITS rounding noise
    if which == 'frobenius':
        assert np.allclose(D[off], D_ref[off], rtol=1e-5, atol=1e-5), np.max(np.abs(D - D_ref))
    else:
        # the reference centres and reduces 4096-element float32 vectors in float32: its corr carries ~3e-5 of rounding,
        # which d = sqrt(1/2 (1 - corr)) amplifies by 1 / (4 d) (up to ~1e-3 per sample for near-identical samples)
        assert np.allclose(D[off], D_ref[off], rtol=1e-3, atol=5e-4), np.max(np.abs(D[off] - D_ref[off]))
        assert np.all(D >= 0) and np.all(D <= 1.0 / np.sqrt(2) + 1e-12)
    # the static reference entry point on explicit vectors, and d(x, x) = 0
    d05 = cls.metric(vecs[0], vecs[5])
    assert abs(d05 - D_ref[0, 5]) <= 2e-3 * max(1.0, D_ref[0, 5])
    self_d, _ = builder.pair_distances([2], [2])
    assert abs(self_d[0]) < 1e-3


def test_correlation_distance_on_more_than_128_points():
    """The whole training set as the builders' data_X (n = 300): Gram matrices from b200gp_gram_logdet."""
    from ms_project_b200.distance import CorrelationDistanceBuilder
    rng = np.random.default_rng(4)
    x = rng.random((300, 2))
    x = (x - x.mean(0)) / x.std(0)
    am, builder, prob, noise_prior, pool = _setup(CorrelationDistanceBuilder, x, 10, ['SE', 'RQ'], 2)
    assert builder._large
    lam = ohell.noise_samples(prior_tuple(noise_prior), prob)
    vecs = []
    for k, m in zip(pool, am.models):
        kp = [prior_tuple(p.prior) for p in k.flattened_parameters()]
        v_ref = ohell.create_vectors(to_oracle_expr(k), kp, x, prob, lam)
        v = np.asarray(m.info)
        assert v.shape == v_ref.shape == (300 * 300, 20)
        assert np.allclose(v, v_ref, rtol=2e-6, atol=1e-7)
        vecs.append(v_ref.astype(np.float64))
    n = len(pool)
    builder.compute_distance(am, list(range(n)), list(range(n)))
    D = builder.get_kernel(n)
    D64 = np.array([[ohell.correlation_distance(vecs[i], vecs[j]) for j in range(n)] for i in range(n)])
    off = ~np.eye(n, dtype=bool)
    assert np.allclose(D[off], D64[off], rtol=1e-9, atol=1e-7), np.max(np.abs(D[off] - D64[off]))
