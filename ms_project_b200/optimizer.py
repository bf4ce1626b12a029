"""Ask/tell L-BFGS-B: the reference's optimiser, re-entrant.

paramz `opt_lbfgsb` (the optimiser `optimize_restarts` uses when `optimizer=None`, reached from
src/autoks/core/gp_model.py:64-65) calls `scipy.optimize.fmin_l_bfgs_b(f_fp, x0, maxfun=max_iters,
maxiter=max_iters)`.  That driver owns the loop, so it can only score one candidate at a time.  This class
drives the SAME compiled routine (`scipy.optimize._lbfgsb.setulb`, reverse communication) with the same
loop logic as scipy's `_minimize_lbfgsb`, but hands control back whenever an objective value is needed, so
hundreds of independent instances advance in lock-step against one batched GPU evaluation.
tests/test_optimizer.py checks bit-identical trajectories against `fmin_l_bfgs_b`."""
import numpy as np
from scipy.optimize import _lbfgsb

try:                                    # scipy >= 1.15 carries the ILP64 switch next to the driver
    from scipy.optimize._lbfgsb_py import HAS_ILP64 as _ILP64
except Exception:                       # pragma: no cover
    _ILP64 = False
_INT = np.int64 if _ILP64 else np.int32


class LBFGSB(object):
    """One unbounded L-BFGS-B minimisation (scipy defaults: m=10, factr=1e7, pgtol=1e-5, maxls=20)."""

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
