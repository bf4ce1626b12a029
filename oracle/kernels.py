"""Oracle: compositional kernel evaluation and hyperparameter gradients (NumPy).  TEST INFRASTRUCTURE.

Restates, per leaf, the GPy-fork kernels the reference binds in
`src/autoks/backend/kernel.py:17-22` (KERNEL_DICT = SE:RBF, RQ:RationalQuadratic,
LIN:LinScaleShift, PER:StandardPeriodic) plus GPy `Bias` (used by linear.ipynb#4), and the
`Add` / `Prod` combination rules.  Sources (all under /root/reference):

* RationalQuadratic  -- notebooks/exploratory/test custom kernels/rational-quadratic.ipynb#2
* StandardPeriodic   -- .../periodic.ipynb#2
* LinScaleShift      -- .../linear.ipynb#2
* RBF / Stationary   -- .../rbf.ipynb#1-2 says upstream GPy RBF is used unchanged; upstream
  `GPy/kern/src/stationary.py` (`_unscaled_dist`, `update_gradients_full`) and `rbf.py` are
  restated from the public source as recalled (not verifiable in this container).
* Add / Prod         -- upstream `GPy/kern/src/add.py`, `prod.py` as recalled.

Expression format (oracle-local, deliberately NOT shared with the product):
    leaf  := (FAMILY, dim)            FAMILY in {'SE','RQ','PER','LIN','CONST'}, dim 0-based
    node  := ('+', [expr, ...]) | ('*', [expr, ...])
theta is the flat GPy `param_array` of the kernel: leaves in in-order (== postfix operand order),
per leaf SE [variance, lengthscale]; RQ [variance, lengthscale, power];
PER [variance, period, lengthscale]; LIN [variances, shifts]; CONST [variance].
"""
import re
from functools import reduce

import numpy as np

N_PARAMS = {'SE': 2, 'RQ': 3, 'PER': 3, 'LIN': 2, 'CONST': 1, 'KK': 2}
PARAM_NAMES = {
    'SE': ['variance', 'lengthscale'],
    'RQ': ['variance', 'lengthscale', 'power'],
    'PER': ['variance', 'period', 'lengthscale'],
    'LIN': ['variances', 'shifts'],
    'CONST': ['variance'],
    'KK': ['variance', 'lengthscale'],
}
# GPy constructor defaults (all 1.0, RQ power 2.0): rational-quadratic.ipynb#2 `power=2.`,
# periodic.ipynb#2 / linear.ipynb#2 `np.ones(1)`.
DEFAULTS = {'SE': [1., 1.], 'RQ': [1., 1., 2.], 'PER': [1., 1., 1.], 'LIN': [1., 1.], 'CONST': [1.], 'KK': [1., 1.]}

# 'KK': the BOMS "kernel kernel" (notebooks/exploratory/test custom kernels/kernel-kernel.ipynb#2; fork class
# RBFDistanceBuilderKernelKernel, src/autoks/core/strategies/bayes_opt_gp_strategy.py:27-28): a Stationary RBF whose
# unscaled distance between two inputs (model indices) is looked up in a distance matrix.  The table is oracle-global.
LOOKUP_TABLE = [None]


def set_lookup_table(D):
    LOOKUP_TABLE[0] = None if D is None else np.asarray(D, dtype=np.float64)


def _lookup_dist(x, x2):
    D = LOOKUP_TABLE[0]
    i = np.asarray(x).astype(int)
    j = i if x2 is None else np.asarray(x2).astype(int)
    return D[np.ix_(i, j)]


# ----------------------------------------------------------------------------- expression utils
def parse(s):
    """'SE_1 * PER_1 + RQ_2' -> expression tree.  Leaf names are FAMILY_<dim+1> as produced by
    `subkernel_expression` (src/autoks/backend/kernel.py:141-150); '*' binds tighter than '+'
    (src/evalg/encoding.py:421-424); parentheses allowed.  Same-type nesting is flattened the way
    GPy's Add/Prod constructors flatten their parts."""
    toks = re.findall(r'[A-Z]+_\d+|[()+*]', s)
    pos = [0]

    def peek():
        return toks[pos[0]] if pos[0] < len(toks) else None

    def take():
        pos[0] += 1
        return toks[pos[0] - 1]

    def atom():
        t = take()
        if t == '(':
            e = summ()
            assert take() == ')'
            return e
        fam, d = t.rsplit('_', 1)
        assert fam in N_PARAMS, t
        return (fam, int(d) - 1)

    def prod():
        parts = [atom()]
        while peek() == '*':
            take()
            parts.append(atom())
        return _combine('*', parts)

    def summ():
        parts = [prod()]
        while peek() == '+':
            take()
            parts.append(prod())
        return _combine('+', parts)

    e = summ()
    assert pos[0] == len(toks), 'trailing tokens in %r' % s
    return e


def _combine(op, parts):
    if len(parts) == 1:
        return parts[0]
    flat = []
    for p in parts:
        if p[0] == op:
            flat.extend(p[1])
        else:
            flat.append(p)
    return (op, flat)


def leaves(expr):
    if expr[0] in ('+', '*'):
        return [l for c in expr[1] for l in leaves(c)]
    return [expr]


def n_params(expr):
    return sum(N_PARAMS[f] for f, _ in leaves(expr))


def default_theta(expr):
    return np.array([v for f, _ in leaves(expr) for v in DEFAULTS[f]], dtype=np.float64)


def param_names(expr):
    return [(f, d, n) for f, d in leaves(expr) for n in PARAM_NAMES[f]]


def to_postfix(expr):
    """Postfix token list, n-ary nodes emitted as left-assoc binary chains -- what
    `infix_tokens_to_postfix_tokens` (src/evalg/encoding.py:415-447) yields for the flattened infix."""
    if expr[0] in ('+', '*'):
        out = to_postfix(expr[1][0])
        for c in expr[1][1:]:
            out = out + to_postfix(c) + [expr[0]]
        return out
    return [expr]


def to_infix(expr, top=True):
    if expr[0] in ('+', '*'):
        s = (' %s ' % expr[0]).join(to_infix(c, False) for c in expr[1])
        return s if top else '(' + s + ')'
    return '%s_%d' % (expr[0], expr[1] + 1)


# ----------------------------------------------------------------------------- distances
def _unscaled_dist(x, x2, mode):
    """Stationary._unscaled_dist for ONE active dimension (autoks kernels are 1-D,
    src/autoks/backend/kernel.py:60).  mode='gpy': r^2 = x^2 + x'^2 - 2xx' (tdot), diagonal forced
    to 0 when X2 is None, clipped at 0, sqrt (upstream stationary.py, as recalled).
    mode='direct': |x - x'| (what the CUDA path computes)."""
    if mode == 'direct':
        xx2 = x if x2 is None else x2
        return np.abs(x[:, None] - xx2[None, :])
    if x2 is None:
        xsq = np.square(x)
        r2 = -2. * np.outer(x, x) + (xsq[:, None] + xsq[None, :])
        np.fill_diagonal(r2, 0.)
    else:
        r2 = -2. * np.outer(x, x2) + (np.square(x)[:, None] + np.square(x2)[None, :])
    return np.sqrt(np.clip(r2, 0, np.inf))


# ----------------------------------------------------------------------------- leaf kernels
def _leaf_K(fam, th, x, x2, mode):
    if fam == 'SE':
        v, l = th
        r = _unscaled_dist(x, x2, mode) / l
        return v * np.exp(-0.5 * r ** 2)                      # rbf.py K_of_r
    if fam == 'RQ':
        v, l, a = th
        r = _unscaled_dist(x, x2, mode) / l
        return v * np.exp(-a * np.log1p(np.square(r) / (2. * a)))   # rational-quadratic.ipynb#2 K_of_r
    if fam == 'PER':
        v, p, l = th
        xx2 = x if x2 is None else x2
        base = np.pi * (x[:, None] - xx2[None, :]) / p
        return v * np.exp(-2. * np.square(np.sin(base) / l))  # periodic.ipynb#2 K
    if fam == 'LIN':
        v, c = th
        xx2 = x if x2 is None else x2
        return np.outer(x - c, xx2 - c) * v                   # linear.ipynb#2 K (non-ARD)
    if fam == 'KK':
        v, l = th
        r = _lookup_dist(x, x2) / l                           # kernel-kernel.ipynb#2 _scaled_dist
        return v * np.exp(-0.5 * r ** 2)                      # K_of_r
    if fam == 'CONST':
        n2 = x.shape[0] if x2 is None else x2.shape[0]
        return np.full((x.shape[0], n2), th[0], dtype=np.float64)   # GPy Bias.K
    raise ValueError(fam)


def _leaf_grad(fam, th, x, dL_dK, mode, quirks):
    """update_gradients_full(dL_dK, X) for one leaf (X2 = None)."""
    if fam in ('SE', 'RQ'):
        v, l = th[0], th[1]
        r = _unscaled_dist(x, None, mode) / l
        K = _leaf_K(fam, th, x, None, mode)
        g_v = np.sum(K * dL_dK) / v                           # stationary.py: variance.gradient
        if fam == 'SE':
            dK_dr = -r * K                                    # rbf.py dK_dr
        else:
            a = th[2]
            dK_dr = -v * r * np.exp(-(a + 1) * np.log1p(np.square(r) / (2. * a)))
        g_l = -np.sum(dK_dr * dL_dK * r) / l                  # stationary.py non-ARD lengthscale.gradient
        if fam == 'SE':
            return [g_v, g_l]
        r2 = np.square(r) + np.spacing(1)
        dK_dpow = K * (np.exp(np.log(r2) - np.log(r2 + 2. * a)) - np.log1p(r2 / (2. * a)))
        return [g_v, g_l, np.sum(dL_dK * dK_dpow)]
    if fam == 'KK':                                           # Stationary.update_gradients_full with the looked-up r
        v, l = th
        r = _lookup_dist(x, None) / l
        K = _leaf_K(fam, th, x, None, mode)
        return [np.sum(K * dL_dK) / v, -np.sum(-r * K * dL_dK * r) / l]
    if fam == 'PER':
        v, p, l = th
        base = np.pi * (x[:, None] - x[None, :]) / p
        sin_base = np.sin(base)
        exp_dist = np.exp(-2. * np.square(sin_base / l))
        dwl = v * (4.0 / np.square(l)) * sin_base * np.cos(base) * (base / p)
        # periodic.ipynb#2 has dl = v*sin^2/l^3, which is the derivative of GPy's ORIGINAL
        # exp(-1/2 ...) kernel; with the fork's factor 2 the true derivative is 4x that.
        dl = v * np.square(sin_base) / np.power(l, 3) * (1.0 if quirks else 4.0)
        return [np.sum(exp_dist * dL_dK), np.sum(dwl * exp_dist * dL_dK), np.sum(dl * exp_dist * dL_dK)]
    if fam == 'LIN':
        v, c = th
        dLs = (dL_dK + dL_dK.T) / 2
        xs = x - c
        g_v = np.sum(np.outer(xs, xs) * dLs)
        # linear.ipynb#2 omits the `variances` factor on the shift gradient.
        g_c = np.sum((2 * c - x[:, None] - x[None, :]) * dLs) * (1.0 if quirks else v)
        return [g_v, g_c]
    if fam == 'CONST':
        return [np.sum(dL_dK)]
    raise ValueError(fam)


# ----------------------------------------------------------------------------- tree evaluation
def K(expr, theta, X, X2=None, dist_mode='gpy'):
    """kern.K(X, X2) for a compositional kernel (Add.K = sum of parts, Prod.K = product of parts)."""
    theta = np.asarray(theta, dtype=np.float64)
    X = np.asarray(X, dtype=np.float64)
    X2 = None if X2 is None else np.asarray(X2, dtype=np.float64)
    off = [0]

    def rec(e):
        if e[0] in ('+', '*'):
            mats = [rec(c) for c in e[1]]
            return reduce(np.add if e[0] == '+' else np.multiply, mats)
        fam, d = e
        th = theta[off[0]:off[0] + N_PARAMS[fam]]
        off[0] += N_PARAMS[fam]
        return _leaf_K(fam, th, X[:, d], None if X2 is None else X2[:, d], dist_mode)

    out = rec(expr)
    assert off[0] == theta.size, 'theta length mismatch'
    return out


def grad_theta(expr, theta, X, dL_dK, dist_mode='gpy', quirks=False):
    """kern.update_gradients_full(dL_dK, X): returns d/dtheta sum(dL_dK * K) in param_array order.
    Add fans dL_dK out to its parts; Prod hands each part dL_dK * prod(sibling K's) (prod.py)."""
    theta = np.asarray(theta, dtype=np.float64)
    X = np.asarray(X, dtype=np.float64)
    out = np.zeros(theta.size)

    # annotate leaves with theta offsets
    def annotate(e, off):
        if e[0] in ('+', '*'):
            kids = []
            for c in e[1]:
                k, off = annotate(c, off)
                kids.append(k)
            return (e[0], kids), off
        return (e[0], e[1], off), off + N_PARAMS[e[0]]

    ann, total = annotate(expr, 0)
    assert total == theta.size

    def val(e):
        if e[0] in ('+', '*'):
            return reduce(np.add if e[0] == '+' else np.multiply, [val(c) for c in e[1]])
        fam, d, o = e
        return _leaf_K(fam, theta[o:o + N_PARAMS[fam]], X[:, d], None, dist_mode)

    def rec(e, dL):
        if e[0] == '+':
            for c in e[1]:
                rec(c, dL)
        elif e[0] == '*':
            vals = [val(c) for c in e[1]]
            for i, c in enumerate(e[1]):
                others = [vals[j] for j in range(len(vals)) if j != i]
                rec(c, dL * reduce(np.multiply, others))
        else:
            fam, d, o = e
            out[o:o + N_PARAMS[fam]] = _leaf_grad(fam, theta[o:o + N_PARAMS[fam]], X[:, d], dL, dist_mode, quirks)

    rec(ann, np.asarray(dL_dK, dtype=np.float64))
    return out
