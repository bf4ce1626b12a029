"""GPU: BOMS meta-level -- the kernel-kernel surrogate GP over model indices (OP_LOOKUP leaf),
expected improvement and the query step -- against the oracle's NumPy restatement of the same GP."""
import numpy as np
import pytest

from oracle import gp as ogp, kernels as ok, model as omodel
from oracle.adapter import to_oracle_expr, prior_tuple

pytestmark = pytest.mark.gpu


class _Model(object):
    def __init__(self, covariance):
        self.covariance = covariance
        self.info = None
        self.score = None


class _Active(object):
    """The slice of ActiveSet (src/autoks/core/active_set.py) the strategy touches."""

    def __init__(self, models):
        self.models = models
        self.selected_indices = []

    def __len__(self):
        return len(self.models)

    def get_candidate_indices(self):
        return [i for i in range(len(self.models)) if i not in self.selected_indices]


def _pool(n_models):
    from ms_project_b200 import workloads
    from ms_project_b200.covariance import Covariance
    from ms_project_b200.distance import HellingerDistanceBuilder, probability_samples
    from ms_project_b200.gpy_compat.priors import Gaussian
    rng = np.random.default_rng(5)
    x = rng.random((60, 2))
    x = (x - x.mean(0)) / x.std(0)
    hp = workloads.boms_hyperpriors()
    bases = workloads.get_all_1d_kernels(['SE', 'RQ'], 2, hp)
    pool = (bases + workloads.expand_full_brute_force(bases, 2, n_models))[:n_models]
    am = _Active([_Model(Covariance(k)) for k in pool])
    builder = HellingerDistanceBuilder(Gaussian(np.log(0.01), 1), 20, 40, n_models, am, list(range(n_models)), x,
                                       probability_samples_matrix=probability_samples(40, 20))
    idx = list(range(n_models))
    builder.compute_distance(am, idx, idx)
    return am, builder, hp


def test_lookup_kernel_lml_grad_predict_match_oracle(engine):
    from ms_project_b200.gpy_compat import RBFDistanceBuilderKernelKernel
    from ms_project_b200.program import ProgramBatch
    am, builder, hp = _pool(24)
    D = builder.get_kernel(24).copy()
    assert not np.isnan(D).any()
    ok.set_lookup_table(D)
    rng = np.random.RandomState(0)
    idx = rng.permutation(24)[:14]
    X = idx[:, None].astype(np.float64)
    y = rng.randn(14, 1)
    k = RBFDistanceBuilderKernelKernel(builder, n_models=24, variance=1.7, lengthscale=0.35)
    # kern.K through the device equals the table look-up
    Kdev = k.K(X)
    assert np.allclose(Kdev, 1.7 * np.exp(-0.5 * (D[np.ix_(idx, idx)] / 0.35) ** 2), rtol=1e-13, atol=1e-15)
    batch = ProgramBatch([k])
    engine.bind(X, y, slots=1, max_params=batch.max_params)
    engine.upload_programs(batch)
    theta = batch.pack_theta([k.param_array], [0.07])
    lml, dlml, status, _ = engine.lml_grad(batch, theta)
    ref, g, _ = ogp.lml_and_grad(to_oracle_expr(k), k.param_array, 0.07, X, y)
    assert status[0] == 0
    assert abs(lml[0] - ref) <= 1e-8 * max(1.0, abs(ref))
    assert np.max(np.abs(dlml[:3] - g)) <= 1e-6 * max(1.0, np.max(np.abs(g)))
    Xs = np.setdiff1d(np.arange(24), idx)[:, None].astype(np.float64)
    mean, var, _, st = engine.posterior(batch, theta, 0, Xs)
    mu_ref, var_ref = ogp.predict(to_oracle_expr(k), k.param_array, 0.07, X, y, Xs)
    assert st == 0
    assert np.allclose(mean, mu_ref[:, 0], rtol=1e-8, atol=1e-9) and np.allclose(var, var_ref[:, 0], rtol=1e-8, atol=1e-9)


def test_surrogate_fit_and_expected_improvement_match_oracle(engine):
    from ms_project_b200.boms import KernelKernelGPModel, expected_improvement, get_quantiles
    from ms_project_b200.covariance import Covariance
    from ms_project_b200.gpy_compat import RBFDistanceBuilderKernelKernel
    am, builder, hp = _pool(30)
    D = builder.get_kernel(30).copy()
    ok.set_lookup_table(D)
    rng = np.random.RandomState(3)
    sel = rng.permutation(30)[:12]
    x = sel[:, None]
    y = (np.sin(sel / 3.0) + 0.1 * rng.randn(12))[:, None]
    kk = KernelKernelGPModel(Covariance(RBFDistanceBuilderKernelKernel(builder, n_models=30)), verbose=False,
                             kernel_kernel_hyperpriors={'SE': hp['SE'], 'GP': hp['GP']}, optimize_restarts=3)
    np.random.seed(21)
    kk.update(x, y, None, None)
    # oracle: the same GP (Standardize normaliser, 1 % noise start, SE / GP hyperpriors), same RNG stream
    ym, ys = y.mean(), y.std()
    expr = ('KK', 0)
    om, failed = omodel.fit(expr, x.astype(np.float64), (y - ym) / ys, n_restarts=3, rng=np.random.RandomState(21),
                            theta=np.array([1.0, 1.0]), noise_var=float(y.var() * 0.01),
                            kern_priors=[prior_tuple(hp['SE']['variance']), prior_tuple(hp['SE']['lengthscale'])],
                            noise_prior=prior_tuple(hp['GP']['variance']))
    assert not failed
    got = np.concatenate([kk.model.kern.param_array, kk.model.likelihood.param_array])
    assert abs(kk.model.log_likelihood() - om.log_likelihood()) <= 1e-6 * max(1.0, abs(om.log_likelihood()))
    assert np.allclose(got, om.param_array, rtol=1e-3, atol=1e-6), (got, om.param_array)
    cand = np.setdiff1d(np.arange(30), sel)[:, None]
    ei = expected_improvement(cand, kk)
    mu, var = ogp.predict(expr, om.param_array[:-1], om.param_array[-1], x.astype(np.float64), (y - ym) / ys,
                          cand.astype(np.float64))
    mu, sd = mu * ys + ym, np.sqrt(np.clip(var * ys ** 2, 1e-10, np.inf))
    fmax = (ogp.predict(expr, om.param_array[:-1], om.param_array[-1], x.astype(np.float64), (y - ym) / ys,
                        x.astype(np.float64))[0] * ys + ym).max()
    phi, Phi, u = get_quantiles(0.01, fmax, mu, sd)
    ei_ref = sd * (u * Phi + phi)
    assert ei.shape == (18, 1) and np.all(ei >= 0)
    assert np.allclose(ei, ei_ref, rtol=1e-4, atol=1e-9), np.max(np.abs(ei - ei_ref))
    assert int(np.argmax(ei)) == int(np.argmax(ei_ref))


def test_query_step_selects_an_unevaluated_model(engine):
    from ms_project_b200.boms import BayesOptGPStrategy, expected_improvement
    am, builder, hp = _pool(20)
    strategy = BayesOptGPStrategy(am, expected_improvement, builder, {'SE': hp['SE'], 'GP': hp['GP']}, optimize_restarts=2)
    rng = np.random.RandomState(1)
    np.random.seed(1)
    am.selected_indices = [0, 3]
    fitness = [0.2, 0.5]
    chosen = []
    for step in range(3):
        model = strategy.query([am.models[i] for i in am.selected_indices], fitness, None, None)
        nxt = strategy.last_selected_index
        assert nxt not in am.selected_indices[:-1] and am.selected_indices[-1] == nxt
        assert all(am.models[i].score is not None for i in am.get_candidate_indices())
        # the returned model is fresh (bayes_opt_gp_strategy.py:97-101): the evaluation loop must FIT it, not keep its EI value
        assert model.covariance is am.models[nxt].covariance and not model.evaluated and np.isnan(model.score)
        chosen.append(nxt)
        fitness.append(float(rng.rand()))
    assert len(set(chosen)) == 3


def test_queried_model_is_fitted_by_the_evaluation_loop(engine):
    """query -> evaluate_models: the chosen model goes through plan_evaluation as 'fit' and ends with a fitness (nBIC of a
    fitted GP), not with the acquisition value that query left on the active set's entry."""
    from ms_project_b200 import fitness as F
    from ms_project_b200.boms import BayesOptGPStrategy, expected_improvement
    from ms_project_b200.gp_model import BatchedEvaluator, plan_evaluation
    am, builder, hp = _pool(12)
    strategy = BayesOptGPStrategy(am, expected_improvement, builder, {'SE': hp['SE'], 'GP': hp['GP']}, optimize_restarts=2)
    np.random.seed(2)
    am.selected_indices = [0, 3]
    model = strategy.query([am.models[i] for i in am.selected_indices], [0.2, 0.5], None, None)
    ei = am.models[strategy.last_selected_index].score
    to_fit, order = plan_evaluation([model], 0, 10, set())
    assert to_fit == [model] and order[0][1] == 'fit'
    rng = np.random.RandomState(0)
    X = rng.rand(40, 2)
    y = rng.randn(40, 1)
    ev = BatchedEvaluator(F.negative_bic, n_restarts_optimizer=2, engine=engine)
    out = ev.evaluate_models([model], X, y, eval_budget=10)
    assert out == [model] and model.evaluated and np.isfinite(model.score) and model.score != ei
    assert model.fit_result is not None and model.fit_result.n_evals > 0
