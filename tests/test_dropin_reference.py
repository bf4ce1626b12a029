"""CPU, build container only: the reference's OWN selectors (`/root/reference/src/autoks/core/model_selection/*.py`,
unmodified, imported through tests/refshim.py) run over the product's GPy-compatible object surface -- B3 of SURVEY.md 8b --
once as they are (`ModelSelector._evaluate_models` -> `GPModel.score_model` per candidate, base.py:314-389) and once with
the batch seam of INTEGRATION.md section 3 (`integration.use_batched_evaluation`).  Numbers come from the CPU oracle
standing in for the device (tests/oracle_engine.py); the same searches are replayed against the real engine on the B200 by
tests/test_gpu_dropin.py from the fixtures tools/make_dropin_golden.py records.

Skipped where /root/reference does not exist (the GPU box)."""
import warnings

import numpy as np
import pytest

import refshim

pytestmark = pytest.mark.skipif(not refshim.available(), reason='/root/reference is only present in the build container')


@pytest.fixture(scope='module')
def ref():
    warnings.filterwarnings('ignore')
    refshim.install()
    from src.autoks.core import model_selection
    return model_selection


def _data(n=24, d=1, seed=0):
    rng = np.random.default_rng(seed)
    x = 8 * rng.random((n, d))
    return x, np.sin(x[:, :1]) + 0.05 * rng.standard_normal((n, 1))


def _run(make, batched, budget=10, seed=0, data=None):
    import oracle_engine
    from ms_project_b200.integration import use_batched_evaluation
    x, y = data if data is not None else _data()
    with oracle_engine.installed():
        np.random.seed(seed)
        sel = make()
        if batched:
            use_batched_evaluation(sel)
        log = []

        class Spy(object):          # a callback of the reference's CallbackList protocol (callbacks.py)
            def __getattr__(self, name):
                if name.startswith('on_'):
                    return lambda *a, **k: log.append(name)
                raise AttributeError(name)

            def set_params(self, p): pass

            def set_model(self, m): pass
        sel.train(x, y, eval_budget=budget, verbose=0, callbacks=[Spy()])
    return sel, [(str(m), m.score) for m in sel.selected_models], log


SELECTORS = {
    'cks': lambda ms: ms.CKSModelSelector(base_kernel_names=['SE', 'RQ', 'LIN', 'PER'], optimizer=None, n_restarts_optimizer=2),
    'cks_scg': lambda ms: ms.CKSModelSelector(base_kernel_names=['SE', 'LIN'], n_restarts_optimizer=1),     # optimizer='scg' default
    'geometric_random': lambda ms: ms.GeomRandomModelSelector(base_kernel_names=['SE', 'RQ'], n_restarts_optimizer=2),
    'uniform_random': lambda ms: ms.UniformRandomModelSelector(base_kernel_names=['SE', 'RQ'], n_restarts_optimizer=2),
    'evolutionary': lambda ms: ms.EvolutionaryModelSelector(base_kernel_names=['SE', 'RQ', 'LIN'], n_restarts_optimizer=2,
                                                            pop_size=8),
    # BOMS as the reference wires it: CorrelationDistanceBuilder (the reference's own, kern.K served by the product),
    # BayesOptGPStrategy, KernelKernelGPModel over RBFDistanceBuilderKernelKernel (the OP_LOOKUP leaf), expected improvement
    # and the inner BoemsSurrogateSelector search (smbo_model_selector.py:43-109)
    'smbo': lambda ms: ms.SMBOModelSelector(base_kernel_names=['SE', 'RQ'], eval_budget_low_level=8),
}


@pytest.mark.parametrize('name', sorted(SELECTORS))
def test_reference_selector_runs_unchanged_and_batch_seam_is_equivalent(ref, name):
    make = lambda: SELECTORS[name](ref)       # noqa: E731
    budget = 6
    data = _data(24, 2, seed=1) if name == 'smbo' else None
    sel_a, got_a, log_a = _run(make, batched=False, budget=budget, data=data)
    sel_b, got_b, log_b = _run(make, batched=True, budget=budget, data=data)
    assert sel_a.n_evals == sel_b.n_evals == budget
    assert len(got_a) >= 1 and all(np.isfinite(s) for _, s in got_a)
    # same candidates, same scores TO THE BIT, same order: the batch seam replays the sequential loop exactly
    assert got_a == got_b
    assert str(sel_a.best_model()) == str(sel_b.best_model())
    assert sel_a.visited == sel_b.visited
    assert log_a == log_b                      # every callback fires the same number of times in the same order
    if name == 'smbo':          # SMBOModelSelector holds live strategy / builder objects: no to_dict round trip in the reference either
        return
    # serialisation round trip of the selector and of its models (base.py:409-443, gp_model.py:98-125)
    sel_c = type(sel_a).from_dict(sel_a.to_dict())
    assert [str(m) for m in sel_c.selected_models] == [str(m) for m in sel_a.selected_models]
    assert [m.score for m in sel_c.selected_models] == [m.score for m in sel_a.selected_models]
    th_a = [m.covariance.raw_kernel.param_array for m in sel_a.selected_models]
    th_c = [m.covariance.raw_kernel.param_array for m in sel_c.selected_models]
    assert all(np.array_equal(a, c) for a, c in zip(th_a, th_c))


def test_reference_gpmodel_api_singular(ref):
    """The reference's own GPModel (core/gp_model.py:14-82) over gpy_compat: score_model / fit / build_model, predict and
    evaluate through the selector (base.py:231-251)."""
    import oracle_engine
    from src.autoks.core.covariance import Covariance
    from src.autoks.core.gp_model import GPModel
    from src.autoks.core.fitness_functions import negative_bic
    from src.autoks.core.gp_models import gp_regression
    from GPy.kern import RBF, StandardPeriodic
    x, y = _data(30)
    x, y = (x - x.mean()) / x.std(), (y - y.mean()) / y.std()
    with oracle_engine.installed():
        np.random.seed(1)
        m = GPModel(Covariance(RBF(1, active_dims=[0]) + StandardPeriodic(1, active_dims=[0])))
        md = gp_regression()
        m.model_input_dict, m.likelihood = md, md['likelihood'].copy()
        score = m.score_model(x, y, negative_bic, optimizer=None, n_restarts=2)
        assert np.isfinite(score) and m.evaluated and not m.failed_fitting
        gp = m.build_model(x, y)
        assert negative_bic(gp) == pytest.approx(score, rel=1e-12)
        mu, var = gp.predict(x[:5])
        assert mu.shape == (5, 1) and var.shape == (5, 1) and np.all(var > 0)
        assert np.abs(mu - y[:5]).max() < 0.5
        assert m.covariance.to_binary_tree().postfix_tokens() == ['SE_1', 'PER_1', '+']     # test_gp_model.py:16-21


# the reference's own unit tests that cannot pass here for reasons outside the product: optional third-party samplers that are
# not installed (sobol_seq; ghalton's published EA permutation table -- the Halton stand-in of refshim.py is exact for the
# plain sequence only)
EXPECTED_REFERENCE_FAILURES = {
    'test_halton.py::test_generate_generalized_halton',
    'test_sample.py::test_sample_sobol',
    'test_sobol.py::TestSobol::test_gen_sobol',
    'test_sobol.py::TestSobol::test_sobol_sample',
}


@pytest.mark.timeout(600)
def test_reference_unit_test_suite_passes_over_gpy_compat():
    """`/root/reference/src/tests/unit/autoks/*.py` -- the reference's OWN unit tests, unmodified -- collected and run in a
    subprocess with GPy replaced by ms_project_b200.gpy_compat (tests/refshim_plugin.py): kernels, priors, Covariance
    tokens / symbolic forms, GPModel (de)serialisation, grammars, active sets, hyperpriors, the distance builders' own
    axioms with kern.K served by the product, model selectors.  Everything passes except the allow-list above."""
    import os
    import re
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([here, os.path.dirname(here)]))
    cmd = [sys.executable, '-m', 'pytest', '-p', 'refshim_plugin', '-q', '--no-header', '-p', 'no:cacheprovider', '-W', 'ignore',
           '--rootdir=/tmp', os.path.join(refshim.REFERENCE, 'src', 'tests', 'unit', 'autoks'),
           '--deselect', os.path.join(refshim.REFERENCE, 'src', 'tests', 'unit', 'autoks', 'test_matlab.py')]
    out = subprocess.run(cmd, cwd='/tmp', env=env, capture_output=True, text=True).stdout
    failed = set(re.findall(r'^(?:FAILED|ERROR) (?:.*/)?(\S+?::\S+)', out, flags=re.M))
    summary = out.strip().splitlines()[-1]
    n_passed = int(re.search(r'(\d+) passed', summary).group(1))
    assert failed <= EXPECTED_REFERENCE_FAILURES, (sorted(failed - EXPECTED_REFERENCE_FAILURES), summary)
    assert n_passed >= 205, summary
