"""GPU parity of the Hellinger kernel-kernel distance against the oracle's restatement of distance.py, plus the
metric axioms the reference's own tests assert (src/tests/unit/autoks/test_distance.py:19-98,214-318)."""
import numpy as np
import pytest

from oracle import hellinger as ohell
from oracle.adapter import to_oracle_expr, prior_tuple

pytestmark = pytest.mark.gpu


class _Model(object):
    def __init__(self, covariance):
        self.covariance = covariance
        self.info = None


def _setup(x, n_models, level, names, n_dims):
    from ms_project_b200 import workloads
    from ms_project_b200.covariance import Covariance
    from ms_project_b200.distance import ActiveModels, HellingerDistanceBuilder, probability_samples
    from ms_project_b200.gpy_compat.priors import Gaussian
    hp = workloads.boms_hyperpriors()
    bases = workloads.get_all_1d_kernels(names, n_dims, hp)
    pool = (bases + workloads.expand_full_brute_force(bases, level, n_models))[:n_models]
    am = ActiveModels(max_n_models=n_models)
    for k in pool:
        am.add(_Model(Covariance(k)))
    prob = probability_samples(40, 20)
    noise_prior = Gaussian(np.log(0.01), 1)           # metrics.py:80
    builder = HellingerDistanceBuilder(noise_prior, 20, 40, n_models, am, list(range(len(pool))), x,
                                       probability_samples_matrix=prob)
    return am, builder, prob, noise_prior, pool


def _oracle_infos(pool, x, prob, noise_prior):
    lam = ohell.noise_samples(prior_tuple(noise_prior), prob)
    infos = []
    for k in pool:
        kp = [prior_tuple(p.prior) for p in k.flattened_parameters()]
        infos.append(ohell.create_precomputed_info(to_oracle_expr(k), kp, x, prob, lam))
    return infos, lam


def test_precompute_and_distances_match_oracle():
    rng = np.random.default_rng(2)
    x = rng.random((100, 4))
    x = (x - x.mean(0)) / x.std(0)
    am, builder, prob, noise_prior, pool = _setup(x, 40, 2, ['SE', 'RQ'], 4)
    infos, lam = _oracle_infos(pool, x, prob, noise_prior)
    assert np.allclose(builder.hyperparameter_data_noise_samples[:, 0], lam, rtol=1e-14)
    for m, (ld_ref, G_ref) in zip(am.models, infos):
        ld, G = tuple(m.info)
        assert np.all(m.info.status == 0)
        assert G.shape == G_ref.shape == (100, 100, 20)
        # the oracle forms r^2 as x^2 + x'^2 - 2xx' like GPy (absolute error ~1e-16 before the 1/l^2 scaling), the device
        # uses (x - x')^2: entries agree to ~1e-16 / l^2 relative
        assert np.allclose(G, G_ref, rtol=1e-8, atol=1e-13), np.max(np.abs(G - G_ref) / (np.abs(G_ref) + 1e-13))
        assert np.allclose(ld, ld_ref, rtol=1e-9, atol=1e-8)
    # NB the reference's rule is order dependent (a skipped sample re-uses logdet_i, distance.py:147): evaluate
    # each unordered pair once, as (i, j) with i < j, on both sides
    for i in range(len(pool)):
        builder.compute_distance(am, [i], list(range(i + 1, len(pool))))
    D = builder.get_kernel(len(pool))
    D_ref = ohell.distance_matrix(infos)
    off = ~np.eye(len(pool), dtype=bool)
    # d = sqrt(clip(mean_s(1 - exp(log_h)))) is ill-conditioned at d ~ 0 (cancellation in 1 - exp, then a sqrt): compare
    # d^2 -- the quantity the arithmetic produces -- tightly, and d itself at the reference's own tolerance for
    # near-identical models (test_distance.py:37: abs 1e-3)
    assert np.allclose(D[off] ** 2, D_ref[off] ** 2, rtol=0, atol=1e-7), np.max(np.abs(D[off] ** 2 - D_ref[off] ** 2))
    big = off & (D_ref > 1e-2)
    assert np.allclose(D[big], D_ref[big], rtol=1e-5, atol=0), np.max(np.abs(D[big] - D_ref[big]))
    assert np.allclose(D[off], D_ref[off], rtol=0, atol=1e-3)
    assert np.all(builder.last_status == 0)
    # one explicit pair through the static reference entry point
    d01 = builder.hellinger_distance(*tuple(am.models[0].info), *tuple(am.models[5].info))
    assert abs(d01 ** 2 - D_ref[0, 5] ** 2) < 1e-7


@pytest.mark.parametrize('n', [17, 64, 65, 80, 81, 96, 97, 104, 105, 112, 113, 128])
def test_pair_distances_at_every_subset_size_class(n):
    """The pair kernel is instantiated per size class (8 x 8 thread grid with 8, 10, 12, 13 block rows up to n = 104, the
    16 x 8 grid up to 112, the 16 x 16 grid up to 128): every class, at its edges, against the oracle."""
    rng = np.random.default_rng(n)
    x = rng.random((n, 3))
    x = (x - x.mean(0)) / x.std(0)
    am, builder, prob, noise_prior, pool = _setup(x, 10, 2, ['SE', 'RQ'], 3)
    infos, _lam = _oracle_infos(pool, x, prob, noise_prior)
    for m, (ld_ref, _G) in zip(am.models, infos):
        assert np.allclose(m.info.log_det, ld_ref, rtol=1e-9, atol=1e-8)
    for i in range(len(pool)):
        builder.compute_distance(am, [i], list(range(i + 1, len(pool))))
    D, D_ref = builder.get_kernel(len(pool)), ohell.distance_matrix(infos)
    off = ~np.eye(len(pool), dtype=bool)
    assert np.allclose(D[off] ** 2, D_ref[off] ** 2, rtol=0, atol=1e-7), np.max(np.abs(D[off] ** 2 - D_ref[off] ** 2))
    assert np.all(builder.last_status == 0)


def test_metric_axioms_like_reference_tests():
    x = np.array([[1, 2], [4, 5], [6, 7], [8, 9], [10, 11]], dtype=np.float64)     # test_distance.py:19
    am, builder, prob, noise_prior, pool = _setup(x, 50, 2, ['SE', 'RQ'], 2)
    n = len(pool)
    idx = list(range(n))
    builder.compute_distance(am, idx, idx)
    D = builder.get_kernel(n).copy()
    assert D.shape == (n, n) and not np.isnan(D).any()
    assert np.all(D >= 0) and np.all(D <= 1)                      # non-negativity and the Hellinger bound
    assert np.allclose(D, D.T)                                     # symmetry (test_distance.py:250-257)
    for i in range(n):                                             # d(x, x) ~ 0: recomputed, not the preset diagonal
        dist, _ = builder.pair_distances([i], [i])
        assert abs(dist[0]) < 1e-3
    for i in range(n):
        assert np.all(D[i][:, None] + D <= D[i][None, :] + 2.0)    # trivial upper bound guard
        assert np.all(D[i][None, :] <= D[i][:, None] + D + 1e-3)   # triangle inequality (tolerance of the reference)
    # NaN-initialised matrix outside the computed block, zero diagonal
    assert builder._average_distance.shape == (50, 50)


def test_update_only_touches_requested_pairs():
    x = np.linspace(-1, 1, 30).reshape(-1, 1)
    am, builder, prob, noise_prior, pool = _setup(x, 12, 1, ['SE', 'RQ', 'LIN', 'PER'], 1)
    assert np.isnan(builder._average_distance[0, 1])
    builder.compute_distance(am, [0], [1, 2])
    A = builder._average_distance
    assert not np.isnan(A[0, 1]) and A[0, 1] == A[1, 0] and not np.isnan(A[2, 0]) and np.isnan(A[1, 2])
    builder.update(am, new_candidates_indices=[5], all_candidates_indices=[3, 4, 5], selected_indices=[0, 1], data_X=x)
    assert not np.isnan(A[1, 3]) and not np.isnan(A[1, 4]) and not np.isnan(A[0, 5]) and not np.isnan(A[1, 5])
    assert np.isnan(A[0, 3])


def test_large_subset_matches_oracle_and_small_path():
    """More than 128 points (the reference hands the builders the whole training set, smbo_model_selector.py:91-97): Gram
    matrices through the blocked DMMA Cholesky of the scoring path.  Parity against the oracle at n = 200, and the two
    device paths against each other where both apply (n = 128 through the blocked entry points directly)."""
    rng = np.random.default_rng(5)
    x = rng.random((200, 2))
    x = (x - x.mean(0)) / x.std(0)
    am, builder, prob, noise_prior, pool = _setup(x, 12, 2, ['SE', 'RQ'], 2)
    assert builder._large
    infos, lam = _oracle_infos(pool, x, prob, noise_prior)
    for m, (ld_ref, G_ref) in zip(am.models, infos):
        ld, G = tuple(m.info)
        assert np.all(m.info.status == 0)
        assert G.shape == G_ref.shape == (200, 200, 20)
        assert np.allclose(G, G_ref, rtol=1e-8, atol=1e-13)
        assert np.allclose(ld, ld_ref, rtol=1e-9, atol=1e-8)
    for i in range(len(pool)):
        builder.compute_distance(am, [i], list(range(i + 1, len(pool))))
    D = builder.get_kernel(len(pool))
    D_ref = ohell.distance_matrix(infos)
    off = ~np.eye(len(pool), dtype=bool)
    assert np.allclose(D[off] ** 2, D_ref[off] ** 2, rtol=0, atol=1e-7), np.max(np.abs(D[off] ** 2 - D_ref[off] ** 2))
    assert np.all(builder.last_status == 0)
    d03 = builder.hellinger_distance(*tuple(am.models[0].info), *tuple(am.models[3].info))
    assert abs(d03 ** 2 - D_ref[0, 3] ** 2) < 1e-7
    # the same 128-point problem through both device paths
    x2 = x[:128]
    am_s, b_small, _, _, pool_s = _setup(x2, 8, 2, ['SE', 'RQ'], 2)
    assert not b_small._large
    for i in range(len(pool_s)):
        b_small.compute_distance(am_s, [i], list(range(i + 1, len(pool_s))))
    D_small = b_small.get_kernel(len(pool_s)).copy()
    am_l, b_large, _, _, _ = _setup(x2, 8, 2, ['SE', 'RQ'], 2)
    b_large._large = True                   # force the blocked path on the same subset
    b_large.precompute_information(am_l, list(range(len(pool_s))), x2)
    for ms, ml in zip(am_s.models, am_l.models):
        assert np.allclose(tuple(ms.info)[0], tuple(ml.info)[0], rtol=1e-12, atol=1e-10)
        assert np.allclose(tuple(ms.info)[1], tuple(ml.info)[1], rtol=1e-12, atol=1e-14)
    for i in range(len(pool_s)):
        b_large.compute_distance(am_l, [i], list(range(i + 1, len(pool_s))))
    D_large = b_large.get_kernel(len(pool_s))
    offs = ~np.eye(len(pool_s), dtype=bool)
    assert np.allclose(D_small[offs] ** 2, D_large[offs] ** 2, rtol=0, atol=1e-10)
