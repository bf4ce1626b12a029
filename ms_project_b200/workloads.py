"""Seeded synthetic workloads for BASELINE.json's configurations and the grammar
enumeration they draw their candidates from.

`expand_full_brute_force` restates CKSGrammar.expand_full_brute_force / expand_single_kernel
(src/autoks/core/grammar.py:230-242,275-321) over the product's kernel objects; the de-duplication key is
the expanded polynomial of the kernel expression, the same equivalence the reference gets from
`Covariance.symbolic_expr_expanded` (src/autoks/core/covariance.py:44-45) without needing sympy."""
from collections import Counter

import numpy as np

from .gpy_compat import kern as _k
from .gpy_compat.priors import LogGaussian

KERNEL_DICT = {'SE': _k.RBF, 'RQ': _k.RationalQuadratic, 'LIN': _k.LinScaleShift, 'PER': _k.StandardPeriodic}


def boms_hyperpriors(data_noise=0.01):
    """Constants of `boms_hyperpriors` (src/autoks/core/hyperprior.py:54-112; pinned by
    src/tests/unit/autoks/test_hyperprior.py:41-66)."""
    prior_l = LogGaussian(np.log(0.1), 1)
    prior_sigma = LogGaussian(np.log(0.4), 1)
    prior_sigma_n = LogGaussian(np.log(data_noise), np.sqrt(0.1))
    prior_alpha = LogGaussian(np.log(0.05), np.sqrt(0.5))
    prior_l_p = LogGaussian(np.log(2), np.sqrt(0.5))
    prior_sigma_0 = LogGaussian(0, 1)
    prior_p = LogGaussian(np.log(0.1), np.sqrt(0.5))
    return {
        'SE': {'variance': prior_sigma, 'lengthscale': prior_l},
        'RQ': {'variance': prior_sigma, 'lengthscale': prior_l, 'power': prior_alpha},
        'PER': {'variance': prior_sigma, 'lengthscale': prior_l_p, 'period': prior_p},
        'LIN': {'variances': prior_sigma, 'shifts': prior_sigma_0},
        'GP': {'variance': prior_sigma_n},
    }


def create_1d_kernel(family, active_dim, hyperpriors=None):
    """create_1d_kernel (src/autoks/backend/kernel.py:41-65)."""
    k = KERNEL_DICT[family](input_dim=1, active_dims=[active_dim])
    if hyperpriors is not None:
        for name, prior in hyperpriors[family].items():
            getattr(k, name).set_prior(prior, warning=False)
    return k


def get_all_1d_kernels(base_kernel_names, n_dims, hyperpriors=None):
    """get_all_1d_kernels (backend/kernel.py:68-89): family-major, dimension-minor."""
    return [create_1d_kernel(f, d, hyperpriors) for f in base_kernel_names for d in range(n_dims)]


def default_base_kernel_names(n_dims):
    """CKSGrammar.default_base_kernel_names (grammar.py:194-200)."""
    return ['SE', 'RQ'] if n_dims > 1 else ['SE', 'RQ', 'LIN', 'PER']


def polynomial_key(kernel):
    """Expanded sum-of-products form as a hashable multiset: equal keys <=> equal expanded expressions."""
    def poly(k):
        if isinstance(k, _k.Add):
            out = Counter()
            for p in k.parts:
                out.update(poly(p))
            return out
        if isinstance(k, _k.Prod):
            acc = Counter({(): 1})
            for p in k.parts:
                nxt = Counter()
                pp = poly(p)
                for m1, c1 in acc.items():
                    for m2, c2 in pp.items():
                        nxt[tuple(sorted(m1 + m2))] += c1 * c2
                acc = nxt
            return acc
        return Counter({('%s_%d' % (k.op_name, int(k.active_dims[0]) + 1),): 1})
    return frozenset(poly(kernel).items())


def expand_single_kernel(kernel, base_kernels):
    out = []
    for b in base_kernels:
        out.append(kernel + b)
        out.append(kernel * b)
    return out


def expand_full_brute_force(base_kernels, level, max_n_models):
    current = list(base_kernels)
    if level == 0:
        return current
    all_kernels = []
    while level > 0:
        this_level, keys = [], set()
        number_of_models = len(all_kernels)
        for ck in current:
            fresh = []
            for nk in expand_single_kernel(ck, base_kernels):
                key = polynomial_key(nk)
                if key not in keys:          # duplicates are only removed against this level, like the reference
                    keys.add(key)
                    fresh.append(nk)
            this_level += fresh
            number_of_models += len(fresh)
            if number_of_models > max_n_models:
                all_kernels += this_level
                return all_kernels[:max_n_models]
        current = this_level
        all_kernels += this_level
        level -= 1
    return all_kernels


def standardize(a):
    """ModelSelector._prepare_data (src/autoks/core/model_selection/base.py:206-220): ddof = 0."""
    a = np.asarray(a, dtype=np.float64)
    return (a - np.mean(a, axis=0)) / np.std(a, ddof=0, axis=0)


def config_c1(n=500, seed=0):
    """C1: CKS, 1-D, N = 500, x = 8 U(0,1), y = sin(x) + 0.05 eps (Sinusoid1Dataset shape)."""
    rng = np.random.default_rng(seed)
    x = 8.0 * rng.random((n, 1))
    y = np.sin(x) + 0.05 * rng.standard_normal((n, 1))
    return standardize(x), standardize(y)


def config_c2(n=2048, n_kernels=256, seed=1, with_priors=True):
    """C2: 256 kernels drawn with replacement from the level-3 brute-force CKS expansion over
    SE/RQ/LIN/PER (<= 4 leaves, <= 12 parameters), 1-D, N = 2048, BOMS hyperpriors."""
    rng = np.random.default_rng(seed)
    x = rng.uniform(0.0, 8.0, size=(n, 1))
    y = np.sin(x) + 0.5 * np.sin(3.0 * x) + 0.05 * rng.standard_normal((n, 1))
    hp = boms_hyperpriors() if with_priors else None
    pool = expand_full_brute_force(get_all_1d_kernels(['SE', 'RQ', 'LIN', 'PER'], 1, hp), 3, 1000)
    idx = rng.integers(0, len(pool), size=n_kernels)
    kernels = [pool[i].copy() for i in idx]
    return standardize(x), standardize(y), kernels


def config_c3(n=1024, n_sub=100, n_candidates=500, seed=2):
    """C3: BOMS Hellinger, D = 4, N = 1024, 500 candidates from the level-2 expansion over SE/RQ per dim."""
    rng = np.random.default_rng(seed)
    x = rng.random((n, 4))
    y = np.sum(np.sin(2 * np.pi * x), axis=1, keepdims=True) + 0.1 * rng.standard_normal((n, 1))
    hp = boms_hyperpriors()
    pool = expand_full_brute_force(get_all_1d_kernels(['SE', 'RQ'], 4, hp), 2, n_candidates)
    xs, ys = standardize(x), standardize(y)
    return xs, ys, pool[:n_candidates], xs[:n_sub]


def config_c4(n=4096, pop=512, seed=3):
    """C4: evolutionary population, D = 8, N = 4096, random SE/RQ trees (<= 8 leaves)."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, 8))
    y = (np.sin(x[:, 0]) * x[:, 1] + 0.5 * np.cos(2 * x[:, 2]) + 0.1 * rng.standard_normal(n)).reshape(-1, 1)
    bases = get_all_1d_kernels(['SE', 'RQ'], 8, boms_hyperpriors())
    kernels = []
    for _ in range(pop):
        k = bases[rng.integers(len(bases))].copy()
        for _ in range(int(rng.integers(0, 7))):
            b = bases[rng.integers(len(bases))]
            k = (k + b) if rng.random() < 0.5 else (k * b)
        kernels.append(k)
    return standardize(x), standardize(y), kernels


def config_c5(n=16384, seed=4):
    """C5: one large fit, D = 2, SE_1 * PER_1 + RQ_2 at fixed hyperparameters, noise 0.05."""
    rng = np.random.default_rng(seed)
    x = rng.uniform(0.0, 10.0, size=(n, 2))
    y = (np.sin(x[:, 0]) * np.exp(-x[:, 0] / 10.0) + 0.3 * np.cos(x[:, 1]) + 0.05 * rng.standard_normal(n)).reshape(-1, 1)
    k = _k.RBF(1, variance=1.0, lengthscale=1.5, active_dims=[0]) * \
        _k.StandardPeriodic(1, variance=1.0, period=2.1, lengthscale=1.5, active_dims=[0]) + \
        _k.RationalQuadratic(1, variance=1.0, lengthscale=2.0, power=1.2, active_dims=[1])
    return standardize(x), standardize(y), k, 0.05
