"""GPU parity of the batched optimize_restarts / GPModel scoring seam against the oracle's sequential
restatement of paramz (same RNG seed => same restart draws; same L-BFGS-B routine on both sides)."""
import numpy as np
import pytest

from conftest import from_string, make_data
from oracle import model as omodel, scoring as oscoring
from oracle.adapter import to_oracle_expr, prior_tuple

pytestmark = pytest.mark.gpu


def _with_priors(kernel):
    from ms_project_b200 import workloads
    hp = workloads.boms_hyperpriors()
    fam = {'SE': 'SE', 'RQ': 'RQ', 'PER': 'PER', 'LIN': 'LIN'}

    def visit(k):
        if getattr(k, 'op_name', None) in fam:
            for name, prior in hp[fam[k.op_name]].items():
                getattr(k, name).set_prior(prior, warning=False)
    kernel.traverse(lambda k: visit(k) if hasattr(k, 'op_name') else None)
    return kernel


def _oracle_fit(kernels, X, y, n_restarts, seed, noise_prior=None, optimizer=None):
    rng = np.random.RandomState(seed)
    out = []
    for k in kernels:
        kp = [prior_tuple(p.prior) for p in k.flattened_parameters()]
        m, failed = omodel.fit(to_oracle_expr(k), X, y, n_restarts=n_restarts, rng=rng, optimizer=optimizer,
                               theta=k.param_array.copy(), kern_priors=kp, noise_prior=prior_tuple(noise_prior))
        out.append((m, failed))
    return out


def test_map_objective_matches_oracle_along_its_trajectory(engine):
    """Every (f, g) the oracle's L-BFGS-B run requests -- MAP objective with BOMS hyperpriors in optimiser space --
    is reproduced by the device path at the same points: NLML rel 1e-8, gradient rel 1e-6."""
    from ms_project_b200 import workloads
    from ms_project_b200.fit import map_objective
    from ms_project_b200.gpy_compat.likelihoods import Gaussian
    X, y = make_data(150, 2, seed=3)
    noise_prior = workloads.boms_hyperpriors()['GP']['variance']
    for name in ['PER_1 + SE_1', 'RQ_1 * LIN_2 + SE_2', 'PER_2 * SE_1']:
        k = _with_priors(from_string(name))
        lik = Gaussian()
        lik.variance.set_prior(noise_prior)
        kp = [prior_tuple(p.prior) for p in k.flattened_parameters()]
        om = omodel.OracleGP(to_oracle_expr(k), X, y, theta=k.param_array.copy(), kern_priors=kp,
                             noise_prior=prior_tuple(noise_prior))
        trace = []
        orig = om._objective_grads

        def rec(x, orig=orig, trace=trace):
            f, g = orig(x)
            trace.append((x.copy(), f, g.copy()))
            return f, g
        om._objective_grads = rec
        om.randomize(np.random.RandomState(4))
        om.optimize(max_iters=60)
        assert len(trace) > 10
        engine.ensure_bound(X, y, len(trace), k.size)
        f, g, status = map_objective(engine, k, lik, np.array([t[0] for t in trace]))
        for i, (x, fr, gr) in enumerate(trace):
            assert status[i] == 0
            assert abs(f[i] - fr) <= 1e-8 * max(1.0, abs(fr)), (name, i, f[i], fr)
            assert np.max(np.abs(g[i] - gr)) <= 1e-6 * max(1.0, np.max(np.abs(gr))), (name, i)


@pytest.mark.parametrize('use_priors', [False, True])
def test_fit_matches_oracle(engine, use_priors):
    from ms_project_b200 import workloads, fitness
    from ms_project_b200.covariance import Covariance
    from ms_project_b200.gp_model import GPModel, score_models
    from ms_project_b200.gp_models import gp_regression
    X, y = make_data(120, 2, seed=7)
    # (periodic kernels are left to the trajectory test above: their MAP surface is so multimodal that last-digit
    #  differences legitimately send L-BFGS-B to different local optima)
    names = ['SE_1', 'RQ_1 + SE_2', 'SE_1 * SE_2', 'LIN_1 + SE_2', 'RQ_2 * LIN_1']
    kernels = [from_string(s) for s in names]
    noise_prior = None
    if use_priors:
        kernels = [_with_priors(k) for k in kernels]
        noise_prior = workloads.boms_hyperpriors()['GP']['variance']
    ref = _oracle_fit([k.copy() for k in kernels], X, y, 3, seed=11, noise_prior=noise_prior)
    md = gp_regression(likelihood_hyperprior={'variance': noise_prior} if use_priors else None)
    models = [GPModel(Covariance(k)) for k in kernels]
    np.random.seed(11)
    scores = score_models(models, X, y, fitness.negative_bic, n_restarts=3, engine=engine, model_dict=md)
    ref_scores = []
    for (om, failed), m, name in zip(ref, models, names):
        assert not failed and not m.failed_fitting
        lml_ref = om.log_likelihood()
        ref_scores.append(oscoring.negative_bic(lml_ref, X.shape[0], om._size_transformed()))
        # optimiser trajectories are chaotic in the last digits; the optimum itself is compared at 1e-6
        assert abs(m.fit_result.lml - lml_ref) <= 1e-6 * max(1.0, abs(lml_ref)), (name, m.fit_result.lml, lml_ref)
        if use_priors:      # without hyperpriors switched-off components leave flat, unidentifiable directions
            got = np.concatenate([m.covariance.raw_kernel.param_array, m.likelihood.param_array])
            assert np.allclose(got, om.param_array, rtol=1e-2, atol=1e-6), (name, got, om.param_array)
    assert np.allclose(scores, ref_scores, rtol=1e-6)
    assert list(np.argsort(scores)) == list(np.argsort(ref_scores))          # same ranking => same selected kernels


def test_scg_fit_matches_oracle(engine):
    """optimizer='scg' (the CKSModelSelector default, cks_model_selector.py:21): the batched ask/tell SCG against the
    oracle's sequential restatement of paramz's SCG, same restart draws."""
    from ms_project_b200 import workloads, fitness
    from ms_project_b200.covariance import Covariance
    from ms_project_b200.gp_model import GPModel, score_models
    from ms_project_b200.gp_models import gp_regression
    X, y = make_data(120, 2, seed=7)
    names = ['SE_1', 'RQ_1 + SE_2', 'SE_1 * SE_2', 'LIN_1 + SE_2']
    kernels = [_with_priors(from_string(s)) for s in names]
    noise_prior = workloads.boms_hyperpriors()['GP']['variance']
    ref = _oracle_fit([k.copy() for k in kernels], X, y, 2, seed=5, noise_prior=noise_prior, optimizer='scg')
    md = gp_regression(likelihood_hyperprior={'variance': noise_prior})
    models = [GPModel(Covariance(k)) for k in kernels]
    np.random.seed(5)
    scores = score_models(models, X, y, fitness.negative_bic, optimizer='scg', n_restarts=2, engine=engine, model_dict=md)
    for (om, failed), m, name in zip(ref, models, names):
        assert not failed and not m.failed_fitting
        lml_ref = om.log_likelihood()
        assert abs(m.fit_result.lml - lml_ref) <= 1e-5 * max(1.0, abs(lml_ref)), (name, m.fit_result.lml, lml_ref)
        assert all(isinstance(r['warnflag'], str) for r in m.fit_result.runs)           # SCG status strings
    ref_scores = [oscoring.negative_bic(om.log_likelihood(), X.shape[0], om._size_transformed()) for om, _ in ref]
    assert list(np.argsort(scores)) == list(np.argsort(ref_scores))


def test_first_evaluations_are_bitwise_reproducible(engine):
    """Same item, different batch composition / slot => identical bits (no atomics, fixed reduction order)."""
    from ms_project_b200.program import ProgramBatch
    X, y = make_data(300, 2, seed=5)
    ks = [from_string(s) for s in ['SE_1 * PER_1 + RQ_2', 'SE_2', 'LIN_1 + RQ_1']]
    engine.bind(X, y, slots=3, max_params=9)
    b1 = engine.upload_programs(ProgramBatch(ks))
    th1 = b1.initial_theta(np.full(3, 0.3))
    l1, g1, _, _ = [a.copy() for a in engine.lml_grad(b1, th1)]
    b2 = engine.upload_programs(ProgramBatch([ks[2], ks[0]]))
    th2 = b2.initial_theta(np.full(2, 0.3))
    l2, g2, _, _ = [a.copy() for a in engine.lml_grad(b2, th2)]
    assert l2[1] == l1[0] and l2[0] == l1[2]
    assert np.array_equal(g2[b2.theta_off[1]:b2.theta_off[2]], g1[b1.theta_off[0]:b1.theta_off[1]])


def test_batched_evaluator_replays_reference_loop(engine):
    from ms_project_b200 import fitness
    from ms_project_b200.covariance import Covariance
    from ms_project_b200.gp_model import GPModel, BatchedEvaluator
    X, y = make_data(80, 1, seed=9)
    names = ['SE_1', 'RQ_1', 'SE_1', 'SE_1 + RQ_1', 'RQ_1 + SE_1', 'PER_1', 'LIN_1']
    models = [GPModel(Covariance(from_string(s))) for s in names]
    ev = BatchedEvaluator(fitness.log_likelihood_normalized, n_restarts_optimizer=2, engine=engine)
    log = []

    class CB(object):
        def on_evaluate_all_begin(self, logs=None): log.append('all_begin')
        def on_evaluate_begin(self, logs=None): log.append('begin:' + str(logs['gp_model']))
        def on_evaluate_end(self, logs=None): log.append('end:' + str(logs['gp_model']))
        def on_evaluate_all_end(self, logs=None): log.append('all_end:%d' % len(logs['gp_models']))
    np.random.seed(0)
    out = ev.evaluate_models(models, X, y, eval_budget=4, callbacks=CB())
    # SE, RQ fitted; 2nd SE skipped (visited); SE+RQ fitted; RQ+SE skipped (same expanded expr); PER fitted -> budget 4
    assert [str(m) for m in out] == ['SE_1', 'RQ_1', 'RQ_1 + SE_1', 'PER_1']
    assert ev.n_evals == 4 and len(ev.visited) == 4
    assert all(m.evaluated and np.isfinite(m.score) for m in out)
    assert not models[6].evaluated and not models[2].evaluated
    assert log[0] == 'all_begin' and log[-1] == 'all_end:4' and log.count('begin:LIN_1') == 1 and 'end:LIN_1' not in log
    # a second call is a no-op within the budget
    assert ev.evaluate_models(models, X, y, eval_budget=4) == []


def test_overlapped_fit_is_identical_to_single_driver(engine):
    """Two host threads / two engines on one GPU (fit.fit_models_overlapped) give bit-identical fits: every instance is an
    independent optimisation and no device reduction depends on the batch composition."""
    from ms_project_b200 import fitness, workloads
    from ms_project_b200.covariance import Covariance
    from ms_project_b200.gp_model import GPModel, score_models
    from ms_project_b200.gp_models import gp_regression
    X, y = make_data(140, 2, seed=13)
    hp = workloads.boms_hyperpriors()
    bases = workloads.get_all_1d_kernels(['SE', 'RQ'], 2, hp)
    pool = (bases + workloads.expand_full_brute_force(bases, 1, 40))[:18]
    md = gp_regression(likelihood_hyperprior={'variance': hp['GP']['variance']})
    out = []
    for overlap in (False, True):
        models = [GPModel(Covariance(k.copy())) for k in pool]
        np.random.seed(5)
        scores = score_models(models, X, y, fitness.negative_bic, n_restarts=2, engine=engine, model_dict=md, overlap=overlap)
        out.append((np.array(scores), [m.fit_result.theta.copy() for m in models], [m.fit_result.n_evals for m in models]))
    assert np.array_equal(out[0][0], out[1][0])
    assert all(np.array_equal(a, b) for a, b in zip(out[0][1], out[1][1]))
    assert out[0][2] == out[1][2]


def test_singular_score_model_equals_the_batched_seam(engine):
    """`GPModel.score_model()` / `GPModel.fit()` called one model at a time -- the API the reference's
    `_evaluate_model` uses (base.py:381-382) -- against `score_models` on the whole list: same np.random stream, same
    restart points, every instance an independent optimisation => the same bits."""
    from ms_project_b200 import workloads, fitness
    from ms_project_b200.covariance import Covariance
    from ms_project_b200.gp_model import GPModel, score_models
    from ms_project_b200.gp_models import gp_regression
    X, y = make_data(110, 2, seed=21)
    names = ['SE_1', 'PER_1 + SE_2', 'RQ_1 * SE_2', 'LIN_2 + RQ_1', 'SE_1 * PER_2 + RQ_2']
    noise_prior = workloads.boms_hyperpriors()['GP']['variance']
    md = gp_regression(likelihood_hyperprior={'variance': noise_prior})

    def models():
        out = []
        for s in names:
            m = GPModel(Covariance(_with_priors(from_string(s))))
            m.model_input_dict = {k: v for k, v in md.items() if k != 'likelihood'}
            m.likelihood = md['likelihood'].copy()
            out.append(m)
        return out

    one_by_one = models()
    np.random.seed(31)
    s1 = [m.score_model(X, y, fitness.negative_bic, optimizer=None, n_restarts=3) for m in one_by_one]
    batched = models()
    np.random.seed(31)
    s2 = score_models(batched, X, y, fitness.negative_bic, n_restarts=3, engine=engine)
    assert s1 == s2
    for a, b in zip(one_by_one, batched):
        assert a.evaluated and b.evaluated and not a.failed_fitting and not b.failed_fitting
        assert np.array_equal(a.covariance.raw_kernel.param_array, b.covariance.raw_kernel.param_array)
        assert np.array_equal(a.likelihood.param_array, b.likelihood.param_array)
    # fit() alone: returns the GP stand-in whose log_likelihood() / fitness is the score
    m = models()[1]
    np.random.seed(31)
    _ = np.random.normal(size=0)
    models()[0]                                               # (no RNG use: construction draws nothing)
    np.random.seed(5)
    gp = m.fit(X, y, optimizer=None, n_restarts=2)
    assert np.isfinite(gp.log_likelihood()) and fitness.negative_bic(gp) == pytest.approx(
        -(np.log(X.shape[0]) * gp._size_transformed() - 2 * gp.log_likelihood()), rel=1e-14)
    assert m.covariance.raw_kernel is gp.kern and m.likelihood is gp.likelihood


def test_periodic_fits_vs_oracle_restart_by_restart(engine):
    """Periodic kernels back in the fit-level comparison: every restart of every model is compared with the
    oracle's run from the same starting point.  A PER surface is multimodal; two L-BFGS-B runs whose objectives differ in
    the 13th digit can part ways at a line search and settle in different local optima -- the oracle does that to ITSELF
    under 1e-13 noise (tests/test_gpu_dropin.py).  So the fit is run six times on the device, once as is and five times under
    1e-13 / 1e-12 noise (tests/noisy_engine.py): a restart whose six results agree is STABLE and must equal the oracle's run from
    the same start at 1e-5; the others are counted and reported (gpurun_out/per_fit_report.json)."""
    import json
    import os
    from noisy_engine import NoisyEngine
    from ms_project_b200 import workloads, fitness
    from ms_project_b200.covariance import Covariance
    from ms_project_b200.gp_model import GPModel, score_models
    from ms_project_b200.gp_models import gp_regression
    X, y = make_data(160, 1, seed=17)
    names = ['PER_1', 'PER_1 + SE_1', 'PER_1 * SE_1', 'PER_1 + LIN_1', 'PER_1 * RQ_1', 'PER_1 * SE_1 + RQ_1', 'PER_1 + PER_1',
             'PER_1 * LIN_1 + SE_1', 'SE_1 + RQ_1 * PER_1', 'PER_1 * PER_1']
    noise_prior = workloads.boms_hyperpriors()['GP']['variance']
    ref = _oracle_fit([_with_priors(from_string(s)) for s in names], X, y, 3, seed=23, noise_prior=noise_prior)

    def device_fit(eng):
        models = [GPModel(Covariance(_with_priors(from_string(s)))) for s in names]
        np.random.seed(23)
        score_models(models, X, y, fitness.negative_bic, n_restarts=3, engine=eng,
                     model_dict=gp_regression(likelihood_hyperprior={'variance': noise_prior}))
        return models
    # Noise levels: 1e-13 (a BLAS change) and 1e-12 (the size of the device - oracle difference in the objective: other
    # summation orders, factored polynomial forms).  PER_1 * LIN_1 + SE_1, restart 1, sits on such a bifurcation: f = 132.05
    # under every 1e-13 seed tried, 133.23 (the oracle's value) under 1e-12 seed 2, 129.65 under 1e-11 (tools/noise_probe.py).
    runs = [device_fit(engine)] + [device_fit(NoisyEngine(engine, lvl, seed)) for lvl, seed in ((1e-13, 1), (1e-13, 2), (1e-12, 1), (1e-12, 2), (1e-12, 3))]

    def close(a, b):
        return abs(a - b) <= 1e-5 * max(1.0, abs(b))
    rows, n_restarts, n_stable, n_stable_equal, n_equal = [], 0, 0, 0, 0
    for i, ((om, failed), name) in enumerate(zip(ref, names)):
        f_ref = [r['f_opt'] for r in om.optimization_runs]
        f_dev = [[r['f_opt'] for r in run[i].fit_result.runs] for run in runs]
        assert len(f_ref) == 3 and all(len(f) == 3 for f in f_dev) and not failed and not runs[0][i].failed_fitting
        # a run that ends in ABNORMAL_TERMINATION_IN_LNSRCH (warnflag 2) stopped wherever its last line search gave up: no
        # optimum to compare; such restarts never win the argmin over restarts
        conv = [om.optimization_runs[r]['warnflag'] in (0, 1) and runs[0][i].fit_result.runs[r]['warnflag'] in (0, 1) for r in range(3)]
        stable = [conv[r] and all(close(f[r], f_dev[0][r]) for f in f_dev[1:]) for r in range(3)]
        equal = [close(f_dev[0][r], f_ref[r]) for r in range(3)]
        n_restarts += 3
        n_stable += sum(stable)
        n_equal += sum(equal)
        n_stable_equal += sum(s and e for s, e in zip(stable, equal))
        rows.append({'kernel': name, 'f_opt_oracle': f_ref, 'f_opt_device': f_dev[0], 'f_opt_device_noise': f_dev[1:],
                     'stable': stable, 'equal_to_oracle': equal, 'lml_oracle': om.log_likelihood(),
                     'lml_device': runs[0][i].fit_result.lml})
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out')
    os.makedirs(out, exist_ok=True)
    json.dump({'N': 160, 'restarts': n_restarts, 'stable': n_stable, 'stable_and_equal_to_oracle': n_stable_equal,
               'equal_to_oracle': n_equal, 'rows': rows}, open(os.path.join(out, 'per_fit_report.json'), 'w'), indent=1)
    bad = [(r['kernel'], k) for r in rows for k in range(3) if r['stable'][k] and not r['equal_to_oracle'][k]]
    assert not bad, bad
    assert n_stable >= 0.4 * n_restarts, (n_stable, n_restarts)
