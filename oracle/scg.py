"""Oracle: scaled conjugate gradients as paramz runs it (`optimizer='scg'`, the default of the reference's
CKSModelSelector: src/autoks/core/model_selection/cks_model_selector.py:21).  TEST INFRASTRUCTURE.

Restates paramz/optimization/scg.py:SCG (Moller 1993 as implemented by Nabney / Lawrence / Hensman) and
paramz/optimization/optimization.py:opt_SCG from the public upstream source as recalled -- paramz is not importable
here, so this is "parity unpinned" like the rest of the paramz control flow.  f and gradf are separate
callables, exactly as paramz passes Model._objective and Model._grads."""
import numpy as np


def scg(f, gradf, x, maxiters=1000, max_f_eval=1e4, xtol=None, ftol=None, gtol=None):
    """Returns (x, flog, function_eval, status)."""
    xtol = 1e-6 if xtol is None else xtol
    ftol = 1e-6 if ftol is None else ftol
    gtol = 1e-5 if gtol is None else gtol
    x = np.array(x, dtype=np.float64)
    sigma0 = 1.0e-7
    fold = f(x)
    function_eval = 1
    fnow = fold
    gradnew = gradf(x)
    function_eval += 1
    current_grad = np.dot(gradnew, gradnew)
    gradold = gradnew.copy()
    d = -gradnew
    success = True
    nsuccess = 0
    beta = 1.0
    betamin = 1.0e-15
    betamax = 1.0e15
    status = 'Not converged'
    flog = [fold]
    iteration = 0
    while iteration < maxiters:
        if success:
            mu = np.dot(d, gradnew)
            if mu >= 0:
                d = -gradnew
                mu = np.dot(d, gradnew)
            kappa = np.dot(d, d)
            sigma = sigma0 / np.sqrt(kappa)
            xplus = x + sigma * d
            gplus = gradf(xplus)
            function_eval += 1
            theta = np.dot(d, (gplus - gradnew)) / sigma
        delta = theta + beta * kappa
        if delta <= 0:
            delta = beta * kappa
            beta = beta - theta / kappa
        alpha = -mu / delta
        xnew = x + alpha * d
        fnew = f(xnew)
        function_eval += 1
        if function_eval >= max_f_eval:
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
        flog.append(fnow)
        iteration += 1
        if success:
            if np.abs(fnew - fold) < ftol:
                status = 'converged - relative reduction in objective'
                break
            elif np.max(np.abs(alpha * d)) < xtol:
                status = 'converged - relative stepsize'
                break
            else:
                gradold = gradnew
                gradnew = gradf(x)
                function_eval += 1
                current_grad = np.dot(gradnew, gradnew)
                fold = fnew
                if current_grad <= gtol:
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
    else:
        status = 'maxiter exceeded'
    return x, flog, function_eval, status
