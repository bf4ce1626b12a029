"""CPU: host-side logic of the product (kernel objects, program compiler, workloads) and the C-ABI surface."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import from_string, ROOT
from ms_project_b200 import _lib, program, workloads
from ms_project_b200.gpy_compat import (RBF, RationalQuadratic, StandardPeriodic, LinScaleShift, Bias, Add, Prod, Kern,
                                        LogGaussian, Gaussian)


def test_library_exports_every_declared_symbol():
    """The C-ABI library loads and exports exactly what include/b200gp.h declares (no compute calls)."""
    hdr = open(os.path.join(ROOT, 'include', 'b200gp.h')).read()
    declared = set(re.findall(r'\b(b200gp_[a-z_0-9A-Z]+)\s*\(', hdr))
    assert declared, 'no declarations found'
    l = _lib.lib()
    for name in sorted(declared):
        assert hasattr(l, name), 'libb200gp.so does not export %s' % name
    assert set(_lib.EXPORTS) == declared          # the ctypes binding covers the whole header
    assert b'sm_100a' in l.b200gp_version()


def test_workspace_bytes_is_monotone_and_rejects_bad_geometry():
    l = _lib.lib()
    assert l.b200gp_workspace_bytes(0, 1, 1, 4) == 0
    a = l.b200gp_workspace_bytes(2048, 1, 1, 12)
    b = l.b200gp_workspace_bytes(2048, 1, 2, 12)
    assert b - a >= 2 * 2048 * 2048 * 8          # two N x N FP64 matrices per slot
    assert l.b200gp_potrf_work_bytes(100) == 0 and l.b200gp_potrf_work_bytes(256) > 0


def test_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    ctx = ctypes.c_void_p()
    assert _lib.lib().b200gp_create(0, None, ctypes.byref(ctx)) != 0
    from ms_project_b200.engine import Engine
    with pytest.raises(_lib.B200GPError):
        Engine(0)


def test_tokens_follow_reference_postfix():
    k = from_string('SE_1 * SE_1 + RQ_1')
    assert program.postfix_string(k) == 'SE_1 SE_1 * RQ_1 +'            # test_gp_model.py:16-21
    assert program.infix_string(k) == '( SE_1 * SE_1 ) + RQ_1'
    toks, p = program.compile_kernel(k)
    assert p == 7
    assert toks.tolist() == [[0, 0, 0, 0], [0, 0, 2, 0], [6, 0, 1, 0], [1, 0, 4, 0], [5, 2, 3, 0]]
    k2 = from_string('(SE_1 + RQ_2) * LIN_2 * PER_1')
    assert program.postfix_string(k2) == 'SE_1 RQ_2 + LIN_2 * PER_1 *'


def test_add_prod_flatten_and_param_order():
    a, b, c = RBF(1, active_dims=[0]), RationalQuadratic(1, active_dims=[1]), StandardPeriodic(1, active_dims=[0])
    s = a + b + c
    assert isinstance(s, Add) and len(s.parts) == 3
    p = a * b * c
    assert isinstance(p, Prod) and len(p.parts) == 3
    assert s.size == 8 and np.allclose(s.param_array, [1, 1, 1, 1, 2, 1, 1, 1])
    assert [n.split('.')[-1] for n in s.parameter_names()] == ['variance', 'lengthscale', 'variance', 'lengthscale', 'power',
                                                               'variance', 'period', 'lengthscale']
    s[:] = np.arange(1., 9.)
    assert np.allclose(s.param_array, np.arange(1., 9.))
    assert float(s.parts[2].period) == 7.0
    assert a.variance.values[0] == 1.0                     # operands are copied, not aliased


def test_to_dict_round_trip_matches_fork_shape():
    k = from_string('SE_1 * PER_1 + RQ_2 + LIN_1')
    k[:] = np.linspace(0.5, 2.0, k.size)
    d = k.to_dict()
    assert d['class'] == 'GPy.kern.Add' and d['parts'][0]['class'] == 'GPy.kern.Prod'
    assert d['parts'][1]['class'] == 'GPy.kern.RationalQuadratic' and 'power' in d['parts'][1]
    assert set(d['parts'][0]['parts'][1]) >= {'variance', 'period', 'lengthscale', 'ARD1', 'ARD2'}
    assert set(d['parts'][2]) >= {'variances', 'shifts', 'ARD'}
    k2 = Kern.from_dict(eval(str(d)))                       # decode_kernel(encode_kernel(k)), backend/kernel.py:473-498
    assert program.infix_string(k2) == program.infix_string(k)
    assert np.allclose(k2.param_array, k.param_array)


def test_priors_are_singletons_and_indexed_in_param_order():
    assert LogGaussian(0.1, 2) is LogGaussian(0.1, 2)
    assert LogGaussian(0.1, 2) is not Gaussian(0.1, 2)
    hp = workloads.boms_hyperpriors()
    k = workloads.create_1d_kernel('SE', 0, hp) + workloads.create_1d_kernel('RQ', 0, hp)
    assert k.priors.size == k.size == 5                    # get_priors' precondition (backend/kernel.py:96-97)
    pri = np.empty(k.priors.size, dtype=object)
    for prior, ind in k.priors.items():
        pri[ind] = prior
    mus = [p.mu for p in pri]
    assert np.allclose(mus, [np.log(0.4), np.log(0.1), np.log(0.4), np.log(0.1), np.log(0.05)])   # test_hyperprior.py:41-66
    assert len(k.priors.items()) == 3


def test_randomize_uses_priors():
    hp = workloads.boms_hyperpriors()
    k = workloads.create_1d_kernel('PER', 0, hp)
    np.random.seed(0)
    k.randomize()
    v1 = k.param_array.copy()
    np.random.seed(0)
    z = np.random.normal(size=3)           # consumed first, like paramz
    expect = [np.exp(np.random.randn(1) * p.sigma + p.mu)[0] for p in (hp['PER']['variance'], hp['PER']['period'], hp['PER']['lengthscale'])]
    assert np.allclose(v1, expect) and z.size == 3


def test_brute_force_expansion_counts():
    bases = workloads.get_all_1d_kernels(['SE', 'RQ', 'LIN', 'PER'], 1)
    lvl1 = workloads.expand_full_brute_force(bases, 1, 1000)
    # 4 bases x (4 sums + 4 products) with commutative duplicates removed: 10 sums + 10 products
    assert len(lvl1) == 20
    keys = {workloads.polynomial_key(k) for k in lvl1}
    assert len(keys) == 20
    assert workloads.polynomial_key(from_string('SE_1 + RQ_1')) == workloads.polynomial_key(from_string('RQ_1 + SE_1'))
    assert workloads.polynomial_key(from_string('(SE_1 + RQ_1) * LIN_1')) == workloads.polynomial_key(from_string('SE_1 * LIN_1 + LIN_1 * RQ_1'))
    capped = workloads.expand_full_brute_force(bases, 3, 50)
    assert len(capped) == 50


def test_synthetic_configs_are_standardised_and_seeded():
    X, y, ks = workloads.config_c2(n=64, n_kernels=8)
    assert X.shape == (64, 1) and y.shape == (64, 1) and len(ks) == 8
    assert np.allclose(X.mean(0), 0) and np.allclose(X.std(0), 1) and np.allclose(y.std(0), 1)
    X2, _, ks2 = workloads.config_c2(n=64, n_kernels=8)
    assert np.array_equal(X, X2) and [program.infix_string(a) for a in ks] == [program.infix_string(b) for b in ks2]
    assert all(k.priors.size == k.size for k in ks)
    batch = program.ProgramBatch(ks)
    assert batch.max_params <= 12 and batch.theta_off[-1] == sum(k.size + 1 for k in ks)


def test_program_limits():
    k = RBF(1, active_dims=[0])
    for _ in range(40):
        k = k + RBF(1, active_dims=[0])
    with pytest.raises(ValueError):
        program.compile_kernel(k)
    with pytest.raises(TypeError):
        program.compile_kernel('SE_1')


def test_canonical_shape_table_covers_every_tree_with_up_to_four_leaves():
    """csrc/cprog.cuh lists the term sets the <= 4-leaf element-wise kernels have straight-line code for.  Every +/* tree over
    at most four distinct leaves must expand, up to a relabelling of its leaves, to exactly one of them (the compile kernel
    searches the relabelling; a miss would silently send the program to the slower term-loop kernels)."""
    import itertools
    src = open(os.path.join(ROOT, 'ms_project_b200', 'csrc', 'cprog.cuh')).read()
    body = re.search(r'keys\[CP_NSHAPES\]\s*=\s*\{([^}]*)\}', src).group(1)
    keys = [int(t, 16) for t in re.findall(r'0x([0-9A-Fa-f]+)', body)]
    n_shapes = int(re.search(r'constexpr int CP_NSHAPES = (\d+);', src).group(1))
    assert len(keys) == n_shapes == len(set(keys))

    def leaves_of(s):           # the header's cp_shape_leaves
        return 1 if s == 0 else (2 if s <= 2 else (3 if s <= 6 else 4))

    def polys(leaves):          # every polynomial (frozenset of term masks) a tree over exactly these leaves expands to
        if len(leaves) == 1:
            return {frozenset([1 << leaves[0]])}
        out = set()
        for r in range(1, len(leaves)):
            for sub in itertools.combinations(leaves, r):
                if leaves[0] not in sub:
                    continue
                rest = tuple(x for x in leaves if x not in sub)
                for a in polys(tuple(sub)):
                    for b in polys(rest):
                        out.add(frozenset(a | b))
                        out.add(frozenset(x | y for x in a for y in b))
        return out

    def key(masks):
        k = 0
        for i, m in enumerate(sorted(masks)):
            k |= m << (4 * i)
        return k

    table = {(leaves_of(s), k) for s, k in enumerate(keys)}
    seen = set()
    for n in range(1, 5):
        for p in polys(tuple(range(n))):
            assert len(p) <= 4
            hits = set()
            for perm in itertools.permutations(range(n)):
                k = key([sum(1 << perm[i] for i in range(n) if m >> i & 1) for m in p])
                if (n, k) in table:
                    hits.add(k)
            assert len(hits) == 1, (n, sorted(p), hits)
            seen |= {(n, h) for h in hits}
    assert seen == table          # and no entry of the table is unreachable
