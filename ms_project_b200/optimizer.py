"""Ask/tell optimisers: the reference's L-BFGS-B and SCG, re-entrant.

paramz `opt_lbfgsb` (the optimiser `optimize_restarts` uses when `optimizer=None`, reached from
src/autoks/core/gp_model.py:64-65) calls `scipy.optimize.fmin_l_bfgs_b(f_fp, x0, maxfun=max_iters,
maxiter=max_iters)`.  That driver owns the loop, so it can only score one candidate at a time.  This class
drives the SAME compiled routine (`scipy.optimize._lbfgsb.setulb`, reverse communication) with the same
loop logic as scipy's `_minimize_lbfgsb`, but hands control back whenever an objective value is needed, so
hundreds of independent instances advance in lock-step against one batched GPU evaluation.
tests/test_optimizer.py checks bit-identical trajectories against `fmin_l_bfgs_b`."""
import numpy as np
import scipy
from scipy.optimize import _lbfgsb

try:                                    # scipy >= 1.15 carries the ILP64 switch next to the driver
    from scipy.optimize._lbfgsb_py import HAS_ILP64 as _ILP64
except Exception:                       # pragma: no cover
    _ILP64 = False
_INT = np.int64 if _ILP64 else np.int32


def _check_setulb():
    """`setulb` is a PRIVATE scipy routine whose argument list changed with the C rewrite of L-BFGS-B (scipy 1.15: integer
    task arrays, `ln_task`, no `iprint` / `csave`).  This module is written against that signature (tested with scipy
    1.15 - 1.18); anything else fails HERE, at import, with a message -- not as a TypeError inside the first fit."""
    doc = getattr(_lbfgsb.setulb, '__doc__', None) or ''
    ver = tuple(int(p) for p in scipy.__version__.split('.')[:2] if p.isdigit())
    if ver < (1, 15) or ('ln_task' not in doc and 'iprint' in doc):
        raise ImportError('ms_project_b200.optimizer drives scipy.optimize._lbfgsb.setulb with the scipy >= 1.15 signature '
                          '(m, x, l, u, nbd, f, g, factr, pgtol, wa, iwa, task, lsave, isave, dsave, maxls, ln_task); '
                          'found scipy %s with setulb%s' % (scipy.__version__, doc.splitlines()[0] if doc else ' (no signature)'))


_check_setulb()


class LBFGSB(object):
    """One unbounded L-BFGS-B minimisation (scipy defaults: m=10, factr=1e7, pgtol=1e-5, maxls=20)."""
    fail_value = np.finfo(np.float64).max        # paramz Model._objective_grads on LinAlgError

    def __init__(self, x0, maxfun=1000, maxiter=1000, m=10, factr=1e7, pgtol=1e-5, maxls=20):
        x0 = np.asarray(x0, dtype=np.float64).ravel()
        n = x0.size
        self.n, self.m = n, m
        self.maxfun, self.maxiter, self.factr, self.pgtol, self.maxls = maxfun, maxiter, factr, pgtol, maxls
        self.x = np.array(x0, dtype=np.float64)
        self.f = np.array(0.0, dtype=np.float64)
        self.g = np.zeros(n, dtype=np.float64)
        self._nbd = np.zeros(n, dtype=_INT)
        self._low = np.zeros(n, dtype=np.float64)
        self._up = np.zeros(n, dtype=np.float64)
        self._wa = np.zeros(2 * m * n + 5 * n + 11 * m * m + 8 * m, np.float64)
        self._iwa = np.zeros(3 * n, dtype=_INT)
        self._task = np.zeros(2, dtype=_INT)
        self._ln_task = np.zeros(2, dtype=_INT)
        self._lsave = np.zeros(4, dtype=_INT)
        self._isave = np.zeros(44, dtype=_INT)
        self._dsave = np.zeros(29, dtype=np.float64)
        self.nfev = 0
        self.nit = 0
        self.done = False
        self.warnflag = None
        self._waiting = False

    def ask(self):
        """Advance until the routine needs f, g at `self.x` (returns that x) or stops (returns None)."""
        if self.done:
            return None
        assert not self._waiting, 'tell() the pending evaluation first'
        while True:
            _lbfgsb.setulb(self.m, self.x, self._low, self._up, self._nbd, self.f, self.g, self.factr, self.pgtol,
                           self._wa, self._iwa, self._task, self._lsave, self._isave, self._dsave, self.maxls,
                           self._ln_task)
            t = self._task[0]
            if t == 3:                                  # FG: evaluate at x
                self._waiting = True
                return self.x
            elif t == 1:                                # NEW_X: an iteration finished
                self.nit += 1
                if self.nit >= self.maxiter:
                    self._task[0], self._task[1] = 5, 504
                elif self.nfev > self.maxfun:
                    self._task[0], self._task[1] = 5, 502
            else:
                break
        if self._task[0] == 4:
            self.warnflag = 0
        elif self.nfev > self.maxfun or self.nit >= self.maxiter:
            self.warnflag = 1
        else:
            self.warnflag = 2
        self.done = True
        return None

    def tell(self, f, g):
        assert self._waiting
        self.f = float(f)
        self.g = np.asarray(g, dtype=np.float64).copy()
        self.nfev += 1
        self._waiting = False

    def result(self):
        return dict(x=self.x.copy(), f=float(self.f), funcalls=self.nfev, nit=self.nit, warnflag=self.warnflag)


def minimize(fun_and_grad, x0, **kw):
    """Sequential convenience wrapper (same signature role as fmin_l_bfgs_b with a combined f/g callable)."""
    opt = LBFGSB(x0, **kw)
    while True:
        x = opt.ask()
        if x is None:
            break
        f, g = fun_and_grad(x.copy())
        opt.tell(f, g)
    return opt.result()


class SCG(object):
    """One scaled-conjugate-gradient minimisation as paramz runs it for `optimizer='scg'` -- the default optimiser of the
    reference's CKSModelSelector (src/autoks/core/model_selection/cks_model_selector.py:21); paramz
    optimization/scg.py:SCG driven by opt_SCG with maxiters = max_iters, max_f_eval = 1e4, xtol = ftol = 1e-6,
    gtol = 1e-5.

    paramz calls f (Model._objective) and gradf (Model._grads) separately; the device returns both from one evaluation,
    so the gradient paramz would fetch at an accepted point is taken from the evaluation that produced its function value
    (same numbers, one device evaluation fewer per successful iteration).  `funcalls` still counts what paramz would
    have counted.  Same ask / tell / result protocol as LBFGSB; tests/test_optimizer.py checks identical iterates against
    the sequential restatement oracle/scg.py."""
    fail_value = np.inf                           # paramz Model._objective on LinAlgError

    def __init__(self, x0, maxfun=1e4, maxiter=1000, xtol=1e-6, ftol=1e-6, gtol=1e-5):
        self.x = np.array(x0, dtype=np.float64).ravel()
        self.maxiter, self.max_f_eval = int(maxiter), maxfun
        self.xtol, self.ftol, self.gtol = xtol, ftol, gtol
        self.f = np.nan
        self.nfev = 0          # device evaluations
        self.funcalls = 0      # paramz's function_eval
        self.nit = 0
        self.status = 'Not converged'
        self.done = False
        self._waiting = False
        self._reply = None
        self._gen = self._run()
        self._next = next(self._gen)

    def ask(self):
        if self.done:
            return None
        assert not self._waiting, 'tell() the pending evaluation first'
        self._waiting = True
        return self._next

    def tell(self, f, g):
        assert self._waiting
        self._waiting = False
        self.nfev += 1
        try:
            self._next = self._gen.send((float(f), np.asarray(g, dtype=np.float64).copy()))
        except StopIteration:
            self.done = True
            self._next = None

    def result(self):
        return dict(x=self.x.copy(), f=float(self.f), funcalls=int(self.funcalls), nit=self.nit, warnflag=self.status)

    def _run(self):
        x = self.x
        sigma0 = 1.0e-7
        fold, gradnew = yield x                    # f(x) and gradf(x): two paramz calls, one evaluation
        self.funcalls = 2
        fnow = fold
        self.f = fnow
        gradold = gradnew.copy()
        d = -gradnew
        success = True
        nsuccess = 0
        beta, betamin, betamax = 1.0, 1.0e-15, 1.0e15
        mu = kappa = theta = 0.0
        iteration = 0
        status = None
        while iteration < self.maxiter:
            if success:
                mu = np.dot(d, gradnew)
                if mu >= 0:
                    d = -gradnew
                    mu = np.dot(d, gradnew)
                kappa = np.dot(d, d)
                sigma = sigma0 / np.sqrt(kappa)
                _, gplus = yield x + sigma * d     # gradf(xplus)
                self.funcalls += 1
                theta = np.dot(d, (gplus - gradnew)) / sigma
            delta = theta + beta * kappa
            if delta <= 0:
                delta = beta * kappa
                beta = beta - theta / kappa
            alpha = -mu / delta
            xnew = x + alpha * d
            fnew, gnew = yield xnew                # f(xnew); its gradient is what gradf(x) returns if the step is accepted
            self.funcalls += 1
            if self.funcalls >= self.max_f_eval:
                status = 'maximum number of function evaluations exceeded'
                break
            Delta = 2. * (fnew - fold) / (alpha * mu)
            if Delta >= 0.:
                success = True
                nsuccess += 1
                x = xnew
                fnow = fnew
            else:
                success = False
                fnow = fold
            self.x, self.f = x, fnow
            iteration += 1
            self.nit = iteration
            if success:
                if np.abs(fnew - fold) < self.ftol:
                    status = 'converged - relative reduction in objective'
                    break
                elif np.max(np.abs(alpha * d)) < self.xtol:
                    status = 'converged - relative stepsize'
                    break
                else:
                    gradold = gradnew
                    gradnew = gnew
                    self.funcalls += 1
                    fold = fnew
                    if np.dot(gradnew, gradnew) <= self.gtol:
                        status = 'converged - relative reduction in gradient'
                        break
            if Delta < 0.25:
                beta = min(4.0 * beta, betamax)
            if Delta > 0.75:
                beta = max(0.25 * beta, betamin)
            if nsuccess == x.size:
                d = -gradnew
                beta = 1.
                nsuccess = 0
            elif success:
                Gamma = np.dot(gradold - gradnew, gradnew) / mu
                d = Gamma * d - gradnew
        self.status = status if status is not None else 'maxiter exceeded'


def make_optimizer(name, x0, max_iters=1000):
    """The ask/tell optimiser paramz would pick for `optimizer=name` (None = its preferred one, L-BFGS-B)."""
    if name is None or str(name).lower() in ('lbfgsb', 'lbfgs', 'bfgs', 'l-bfgs-b'):
        return LBFGSB(x0, maxfun=max_iters, maxiter=max_iters)
    if str(name).lower() == 'scg':
        return SCG(x0, maxiter=max_iters)
    raise NotImplementedError("optimizer %r: only L-BFGS-B (optimizer=None / 'lbfgsb') and 'scg' are implemented" % (name,))
