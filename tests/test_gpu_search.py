"""GPU: a CKS-style greedy search (src/autoks/core/model_selection/cks_model_selector.py:28-52 + CKSGrammar.expand,
grammar.py:206-242) driven once through the product's batched evaluator and once through the oracle's sequential
restatement of the reference path: the kernel selected at every depth must be the same (north star: "the same selected
kernel sequence per search step")."""
import numpy as np
import pytest

from conftest import make_data
from oracle import model as omodel, scoring as oscoring
from oracle.adapter import to_oracle_expr, prior_tuple

pytestmark = pytest.mark.gpu


def _expand(best, bases, workloads):
    """CKSGrammar.expand for one parent: S+B, S*B for every base kernel B, de-duplicated on the symbolic form."""
    out, seen = [], set()
    for k in workloads.expand_single_kernel(best, bases):
        key = workloads.polynomial_key(k)
        if key not in seen:
            seen.add(key)
            out.append(k)
    return out


def test_cks_selected_sequence_matches_oracle(engine):
    from ms_project_b200 import fitness, workloads, program
    from ms_project_b200.covariance import Covariance
    from ms_project_b200.gp_model import GPModel, BatchedEvaluator
    from ms_project_b200.gp_models import gp_regression
    rng = np.random.default_rng(0)
    x = 8.0 * rng.random((90, 1))
    y = np.sin(x) + 0.3 * x + 0.05 * rng.standard_normal((90, 1))
    X, Y = workloads.standardize(x), workloads.standardize(y)
    hp = workloads.boms_hyperpriors()
    noise_prior = hp['GP']['variance']
    bases = workloads.get_all_1d_kernels(['SE', 'RQ', 'LIN'], 1, hp)
    depth, restarts = 3, 2

    # ---- product: batched evaluator
    np.random.seed(42)
    ev = BatchedEvaluator(fitness.negative_bic, n_restarts_optimizer=restarts, engine=engine,
                          model_dict=gp_regression(likelihood_hyperprior={'variance': noise_prior}))
    cands = [GPModel(Covariance(k.copy())) for k in bases]
    got_seq, got_scores = [], []
    for d in range(depth):
        ev.evaluate_models(cands, X, Y, eval_budget=10 ** 6)
        scored = [m for m in cands if m.evaluated]
        best = max(scored, key=lambda m: m.score)
        got_seq.append(program.infix_string(best.covariance.raw_kernel))
        got_scores.append(best.score)
        cands = [GPModel(Covariance(k)) for k in _expand(best.covariance.raw_kernel, bases, workloads)]

    # ---- oracle: sequential loop, same candidate lists, same RNG stream
    orng = np.random.RandomState(42)
    visited = set()
    ocands = [k.copy() for k in bases]
    ref_seq, ref_scores = [], []
    for d in range(depth):
        scored = []
        for k in ocands:
            key = workloads.polynomial_key(k)
            if key in visited:
                continue
            visited.add(key)
            kp = [prior_tuple(p.prior) for p in k.flattened_parameters()]
            m, failed = omodel.fit(to_oracle_expr(k), X, Y, n_restarts=restarts, rng=orng, theta=k.param_array.copy(),
                                   kern_priors=kp, noise_prior=prior_tuple(noise_prior))
            score = oscoring.negative_bic(m.log_likelihood(), X.shape[0], m._size_transformed())
            scored.append((score, k, m))
        score, k, m = max(scored, key=lambda t: t[0])
        ref_seq.append(program.infix_string(k))
        ref_scores.append(score)
        best = k.copy()
        best[:] = m.param_array[:-1]
        ocands = _expand(best, bases, workloads)
    assert got_seq == ref_seq, (got_seq, ref_seq)
    # L-BFGS-B stops on a relative decrease of 2.2e-9 per step; the two runs end within ~1e-6 of each other
    assert np.allclose(got_scores, ref_scores, rtol=1e-5), (got_scores, ref_scores)
