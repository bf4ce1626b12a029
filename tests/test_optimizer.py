"""CPU: the ask/tell L-BFGS-B driver reproduces scipy.optimize.fmin_l_bfgs_b bit for bit."""
import numpy as np
import pytest
from scipy import optimize

from ms_project_b200.optimizer import LBFGSB, minimize


def rosen(x):
    return optimize.rosen(x), optimize.rosen_der(x)


def quartic(x):
    f = np.sum((x - 1.5) ** 4) + np.sum(np.sin(3 * x))
    g = 4 * (x - 1.5) ** 3 + 3 * np.cos(3 * x)
    return f, g


def plateau(x):            # pgtol convergence at start
    return 1.0, np.zeros_like(x)


@pytest.mark.parametrize('fun,x0,kw', [
    (rosen, np.array([-1.2, 1.0, 0.5, 2.0]), {}),
    (rosen, np.zeros(7), dict(maxfun=15, maxiter=15)),
    (rosen, np.full(3, 3.0), dict(maxfun=1000, maxiter=4)),
    (quartic, np.linspace(-2, 2, 9), {}),
    (plateau, np.ones(3), {}),
])
def test_identical_to_fmin_l_bfgs_b(fun, x0, kw):
    calls = []

    def f(x):
        calls.append(x.copy())
        return fun(x)

    ref_x, ref_f, ref_d = optimize.fmin_l_bfgs_b(f, x0.copy(), maxfun=kw.get('maxfun', 1000), maxiter=kw.get('maxiter', 1000))
    ref_calls = list(calls)
    calls.clear()
    res = minimize(f, x0.copy(), **kw)
    assert np.array_equal(res['x'], ref_x)
    assert res['f'] == ref_f
    assert res['funcalls'] == ref_d['funcalls'] and res['nit'] == ref_d['nit'] and res['warnflag'] == ref_d['warnflag']
    assert len(calls) == len(ref_calls) and all(np.array_equal(a, b) for a, b in zip(calls, ref_calls))


def test_lockstep_instances_are_independent():
    rng = np.random.RandomState(0)
    starts = [rng.randn(5) for _ in range(6)]
    refs = [minimize(rosen, s) for s in starts]
    opts = [LBFGSB(s) for s in starts]
    pending = {i: o.ask() for i, o in enumerate(opts)}
    while pending:
        for i in list(pending):
            f, g = rosen(pending[i].copy())
            opts[i].tell(f, g)
            nxt = opts[i].ask()
            if nxt is None:
                del pending[i]
            else:
                pending[i] = nxt
    for o, r in zip(opts, refs):
        assert np.array_equal(o.x, r['x']) and o.nfev == r['funcalls']


# ---- SCG: the ask/tell driver against the sequential restatement of paramz's SCG (oracle/scg.py)
@pytest.mark.parametrize('fun,x0,maxiter', [
    (rosen, np.array([-1.2, 1.0, 0.5, 2.0]), 1000),
    (rosen, np.zeros(7), 12),
    (quartic, np.linspace(-2, 2, 9), 1000),
    (quartic, np.full(3, 0.3), 1000),
])
def test_scg_identical_to_sequential_restatement(fun, x0, maxiter):
    from ms_project_b200.optimizer import SCG
    from oracle.scg import scg
    ref_x, flog, ref_nfe, ref_status = scg(lambda x: fun(x)[0], lambda x: fun(x)[1], x0.copy(), maxiters=maxiter)
    opt = SCG(x0.copy(), maxiter=maxiter)
    while True:
        x = opt.ask()
        if x is None:
            break
        f, g = fun(x.copy())
        opt.tell(f, g)
    res = opt.result()
    assert np.array_equal(res['x'], ref_x)
    assert res['f'] == flog[-1]
    assert res['funcalls'] == ref_nfe and res['warnflag'] == ref_status and res['nit'] == len(flog) - 1
    assert opt.nfev < ref_nfe                      # the gradient at an accepted point rides on its function evaluation


def test_scg_failed_evaluations_are_rejected_steps():
    """paramz returns inf for an objective that raised LinAlgError: the step is rejected and beta grows."""
    from ms_project_b200.optimizer import SCG
    opt = SCG(np.array([2.0, -1.0]), maxiter=50)
    n = 0
    while True:
        x = opt.ask()
        if x is None:
            break
        f, g = quartic(x.copy())
        n += 1
        if n in (3, 4):                            # two consecutive failures on function-value requests
            f = SCG.fail_value
        opt.tell(f, g)
    assert np.isfinite(opt.result()['f']) and opt.result()['f'] < quartic(np.array([2.0, -1.0]))[0]


def test_make_optimizer_names():
    from ms_project_b200.optimizer import SCG, make_optimizer
    assert isinstance(make_optimizer(None, np.zeros(2)), LBFGSB)
    assert isinstance(make_optimizer('lbfgsb', np.zeros(2)), LBFGSB)
    assert isinstance(make_optimizer('scg', np.zeros(2)), SCG)
    with pytest.raises(NotImplementedError):
        make_optimizer('tnc', np.zeros(2))
