"""GPU: replay of the searches the reference's OWN selectors ran (tests/golden/dropin_*.json, recorded by
tools/make_dropin_golden.py: unmodified /root/reference selectors over gpy_compat with the CPU oracle as the numerics).

Every batch a recorded search handed to the scoring seam is re-scored here through the real engine from the same state
(kernels, hyperparameters, priors, shared kernel objects, np.random seed): the scores must agree with the oracle's and --
the north star's "same selected kernel sequence per search step" -- the best candidate of every batch must be the same
kernel.  C1 as stated is covered by dropin_cks_nbic / dropin_cks_loglikn: CKS depth 3, SE/RQ/LIN/PER, N = 500, 3 restarts.

What "the same" can mean.  Both sides run the same optimiser from the same starts on objectives that agree to ~1e-13.  On a
multimodal surface (periodic kernels, over-parameterised sums) that is not enough to pin the RESULT: perturbing the ORACLE's
own LML / gradients by a relative 1e-13 (tests/noisy_engine.py) sends its SCG fit of PER_1 at N = 120 from nBIC 226.1 to
-377.4 -- the reference cannot reproduce itself across BLAS builds there.  The test therefore measures that sensitivity
on the device: every batch is scored three times, once as is and twice under 1e-13 noise.
  * a candidate whose three device scores agree (RTOL = 1e-5) is STABLE: its score must equal the oracle's (RTOL_MATCH =
    1e-4: the flat optima of over-parameterised sums such as LIN*SE^2 + LIN end up to 3e-5 apart between two correct runs);
  * a candidate whose score moves under the noise is CHAOTIC: reported (gpurun_out/dropin_report_*.json), not compared;
  * the selected kernel of a batch must be the oracle's whenever the oracle's two best candidates are stable and further
    apart than the tolerance.
Measured on B200 (round 2): C1 as stated (N = 500), fitness nbic: the selected kernel of all five search steps is the
reference's; 45 of 64 candidate scores equal the oracle's at 1e-5, 40 are stable, all of those within 3e-5.  Fitness loglikn
(the search walks into periodic kernels): again the same selected kernel at all five steps; 49 of 64 equal, 24 stable.  With SCG
(cks_scg, N = 120) only 5 of 40 candidates are stable -- and the oracle's own PER_1 fit flips under the same noise."""
import glob
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(glob.glob(os.path.join(HERE, 'golden', 'dropin_*.json')))
RTOL = 1e-5          # two scores closer than this are 'the same' (stability under noise, ties)
RTOL_MATCH = 1e-4    # a stable candidate's score against the oracle's: flat optima of over-parameterised sums end up to 3e-5 apart
NOISE = ((1e-13, 1), (1e-13, 2), (1e-12, 1), (1e-12, 2), (1e-12, 3))     # (relative noise level, seed) of the stability runs


def _prior(rec):
    from ms_project_b200.gpy_compat import priors
    return None if rec is None else getattr(priors, rec[0])(rec[1], rec[2])


def _rebuild(call):
    from ms_project_b200.covariance import Covariance
    from ms_project_b200.gp_model import GPModel
    from ms_project_b200.gpy_compat.kern import Kern
    from ms_project_b200.gpy_compat.likelihoods import Gaussian
    models = []
    for i, (rec, share) in enumerate(zip(call['models'], call['shared_with'])):
        if share != i:
            cov = models[share].covariance                 # the same kernel OBJECT, as in the recorded search
        else:
            k = Kern.from_dict(rec['kernel'])
            for p, pr in zip(k.flattened_parameters(), rec['priors']):
                p.set_prior(_prior(pr), warning=False)
            cov = Covariance(k)
        lik = Gaussian(variance=rec['noise'])
        lik.variance.set_prior(_prior(rec['noise_prior']), warning=False)
        models.append(GPModel(cov, lik))
    return models


def _has_per(expr):
    return 'PER' in expr


def _score(call, X, Y, engine):
    from ms_project_b200 import fitness
    from ms_project_b200.gp_model import score_models
    models = _rebuild(call)
    np.random.seed(call['seed'])
    return np.array(score_models(models, X, Y, getattr(fitness, call['fitness']), optimizer=call['optimizer'],
                                 n_restarts=call['n_restarts'], engine=engine))


def _close(a, b, rtol=RTOL):
    with np.errstate(invalid='ignore'):
        return (np.abs(a - b) <= rtol * np.maximum(1.0, np.abs(b))) | (np.isnan(a) & np.isnan(b))


@pytest.mark.parametrize('path', FIXTURES, ids=[os.path.basename(p)[7:-5] for p in FIXTURES])
def test_recorded_reference_search_replays_on_the_device(engine, path):
    from noisy_engine import NoisyEngine
    fx = json.load(open(path))
    X, Y = np.array(fx['x']), np.array(fx['y'])
    report = {'config': fx['config'], 'selector': fx['selector'], 'N': int(X.shape[0]), 'rtol': RTOL, 'batches': []}
    n_cand = n_stable = n_stable_match = n_match = 0
    selections = []
    problems = []
    for b, call in enumerate(fx['calls']):
        if not call['models']:
            continue
        ref = np.array(call['scores'])
        got = _score(call, X, Y, engine)
        # 1e-13: a BLAS change; 1e-12: the size of the device - oracle difference in the objective (summation orders, factored
        # polynomial forms, another diagonal-tile factorisation): LIN_1*SE_1 + PER_1 ends in one of two optima depending on it
        noisy = [_score(call, X, Y, NoisyEngine(engine, lvl, seed)) for lvl, seed in NOISE]
        stable = np.ones(len(got), dtype=bool)
        for nz in noisy:
            stable &= _close(nz, got)
        match = _close(got, ref, RTOL_MATCH)
        exprs = [m['expr'] for m in call['models']]
        n_cand += len(exprs)
        n_stable += int(stable.sum())
        n_stable_match += int((stable & match).sum())
        n_match += int(match.sum())
        for i in np.nonzero(stable & ~match)[0]:
            problems.append('batch %d: %s is insensitive to 1e-13 / 1e-12 noise on the device (%.9g) but the oracle has %.9g'
                            % (b, exprs[i], got[i], ref[i]))
        order = np.argsort(-np.nan_to_num(ref, nan=-np.inf), kind='stable')
        best_ref, best_got = int(order[0]), int(np.nanargmax(got))
        decided = len(order) == 1 or (stable[order[0]] and stable[order[1]] and
                                      (ref[order[0]] - ref[order[1]]) > 2 * RTOL_MATCH * max(1.0, abs(ref[order[0]])))
        if decided and best_got != best_ref:
            problems.append('batch %d selects %s, the reference %s' % (b, exprs[best_got], exprs[best_ref]))
        selections.append({'batch': b, 'n': len(exprs), 'reference': exprs[best_ref], 'device': exprs[best_got],
                           'decided': bool(decided), 'same': best_got == best_ref})
        report['batches'].append({
            'batch': b, 'n': len(exprs), 'stable': int(stable.sum()), 'stable_and_equal': int((stable & match).sum()),
            'chaotic': [{'expr': exprs[i], 'oracle': float(ref[i]), 'device': float(got[i]),
                         'device_noise': [float(nz[i]) for nz in noisy]}
                        for i in np.nonzero(~stable)[0]]})
    report.update({'candidates': n_cand, 'stable': n_stable, 'stable_and_equal_to_oracle': n_stable_match,
                   'equal_to_oracle': n_match, 'selections': selections, 'problems': problems})
    out = os.path.join(os.path.dirname(HERE), 'gpurun_out')
    os.makedirs(out, exist_ok=True)
    json.dump(report, open(os.path.join(out, 'dropin_report_%s.json' % fx['config']), 'w'), indent=1)
    assert not problems, problems
    assert all(s['same'] for s in selections if s['decided'])
    if all(c['optimizer'] is None for c in fx['calls']):      # L-BFGS-B; SCG is chaotic on nearly every periodic candidate
        # enough stable candidates for the comparison to mean something (a search that walks into periodic territory, like
        # cks_loglikn, has 19 of 64 under the 1e-13 / 1e-12 criterion; cks_nbic 33 of 64)
        assert n_stable >= 0.25 * n_cand, 'too few stable candidates: %d of %d' % (n_stable, n_cand)
