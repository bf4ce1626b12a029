"""GPU parity: LML / gradients / kernel matrices of the CUDA path against the NumPy oracle.
Tolerances are the north star's: NLML rel <= 1e-8, gradients rel <= 1e-6."""
import json
import os

import numpy as np
import pytest

from conftest import to_oracle_expr, from_string, make_data

pytestmark = pytest.mark.gpu

KERNELS = ['SE_1', 'RQ_1', 'PER_1', 'LIN_1', 'SE_1 + RQ_2', 'SE_1 * PER_1 + RQ_2', '(SE_1 + RQ_2) * LIN_2',
           'SE_1 * SE_2 * RQ_1 + PER_2 * LIN_1 + CONST_1', 'LIN_1 * LIN_1 * PER_1', '(SE_1 + PER_2) * (RQ_1 + LIN_2)',
           # 5-8 leaves: register-resident polynomial form with 8 leaf slots
           '(SE_1 + RQ_2) * (PER_1 + LIN_2 + SE_2) * RQ_1', 'SE_1 * RQ_2 + PER_1 * LIN_2 + SE_2 * PER_2 + RQ_1 + LIN_1',
           # 9 leaves / 27 product terms: falls back to the postfix interpreter
           '(SE_1 + RQ_1 + PER_1) * (SE_2 + RQ_2 + LIN_2) * (PER_2 + LIN_1 + SE_2)']


def random_theta(batch, rng, lo=-0.7, hi=0.7):
    ks = [np.exp(rng.uniform(lo, hi, size=int(p))) for p in batch.n_params]
    nz = np.exp(rng.uniform(-3, -0.5, size=batch.n_items))
    return batch.pack_theta(ks, nz), ks, nz


@pytest.mark.parametrize('n,d', [(5, 2), (100, 2), (128, 2), (300, 3), (513, 2)])
def test_lml_grad_matches_oracle(engine, n, d):
    from ms_project_b200.program import ProgramBatch
    from oracle import gp
    X, y = make_data(n, d, seed=n)
    kernels = [from_string(s) for s in KERNELS]
    batch = ProgramBatch(kernels)
    engine.bind(X, y, slots=4, max_params=batch.max_params)     # 4 slots < 13 items: exercises chunking
    engine.upload_programs(batch)
    theta, ks, nz = random_theta(batch, np.random.RandomState(1))
    lml, dlml, status, tries = engine.lml_grad(batch, theta)
    for i, k in enumerate(kernels):
        ref_lml, ref_g, _ = gp.lml_and_grad(to_oracle_expr(k), ks[i], nz[i], X, y)
        a, b = batch.theta_off[i], batch.theta_off[i + 1]
        assert status[i] == 0 and tries[i] == 0
        assert abs(lml[i] - ref_lml) <= 1e-8 * max(1.0, abs(ref_lml)), (KERNELS[i], lml[i], ref_lml)
        scale = max(1.0, np.max(np.abs(ref_g)))
        assert np.max(np.abs(dlml[a:b] - ref_g)) <= 1e-6 * scale, (KERNELS[i], dlml[a:b], ref_g)


# every canonical polynomial shape of the <= 4-leaf kernels (csrc/cprog.cuh), each in two leaf orders / family assignments,
# so that every case of shape_value / shape_adjoints and the compile kernel's relabelling are exercised
SHAPES = ['SE_1', 'PER_2',                                                         # 0  a
          'SE_1 + RQ_2', 'LIN_2 + PER_1',                                          # 1  a + b
          'SE_1 * PER_2', 'LIN_1 * RQ_1',                                          # 2  ab
          'SE_1 + RQ_2 + LIN_1', 'PER_2 + LIN_2 + SE_2',                           # 3  a + b + c
          'SE_1 + RQ_2 * PER_1', 'LIN_1 * SE_2 + PER_2',                           # 4  a + bc
          'SE_1 * (RQ_2 + PER_1)', '(LIN_2 + SE_1) * RQ_1',                        # 5  a (b + c)
          'SE_1 * RQ_2 * LIN_1', 'PER_1 * LIN_2 * RQ_2',                           # 6  abc
          'SE_1 + RQ_2 + PER_1 + LIN_2', 'LIN_1 + PER_2 + RQ_1 + SE_2',            # 7  a + b + c + d
          'SE_1 + LIN_2 + RQ_1 * PER_2', 'PER_1 * SE_2 + RQ_2 + LIN_1',            # 8  a + b + cd
          'SE_1 + RQ_2 * (PER_1 + LIN_2)', '(SE_2 + LIN_1) * PER_2 + RQ_1',        # 9  a + b (c + d)
          'LIN_1 + SE_2 * RQ_1 * PER_2', 'PER_1 * LIN_2 * SE_1 + RQ_2',            # 10 a + bcd
          'SE_1 * (RQ_2 + PER_1 + LIN_2)', '(LIN_1 + SE_2 + PER_2) * RQ_1',        # 11 a (b + c + d)
          '(SE_1 + LIN_2) * (RQ_1 + PER_2)', '(PER_1 + RQ_2) * (LIN_1 + SE_2)',    # 12 (a + d)(b + c)
          'SE_1 * RQ_2 + PER_1 * LIN_2', 'LIN_1 * PER_2 + RQ_1 * SE_2',            # 13 ab + cd
          'SE_1 * (RQ_2 + PER_1 * LIN_2)', '(LIN_1 * SE_2 + PER_2) * RQ_1',        # 14 a (b + cd)
          'SE_1 * RQ_2 * (PER_1 + LIN_2)', '(LIN_1 + SE_2) * PER_2 * RQ_1',        # 15 ab (c + d)
          'SE_1 * RQ_2 * PER_1 * LIN_2', 'LIN_1 * PER_2 * RQ_1 * SE_2']            # 16 abcd


def test_every_polynomial_shape_matches_oracle(engine):
    from ms_project_b200.program import ProgramBatch
    from oracle import gp
    X, y = make_data(150, 2, seed=11)                 # two tile rows: diagonal tiles and one below the diagonal
    kernels = [from_string(s) for s in SHAPES]
    batch = ProgramBatch(kernels)
    engine.bind(X, y, slots=len(kernels), max_params=batch.max_params)
    engine.upload_programs(batch)
    theta, ks, nz = random_theta(batch, np.random.RandomState(5))
    lml, dlml, status, tries = engine.lml_grad(batch, theta)
    K = engine.build_K(batch, theta, add_diag=False)
    from oracle import kernels as ok
    for i, k in enumerate(kernels):
        expr = to_oracle_expr(k)
        assert np.allclose(K[i], ok.K(expr, ks[i], X), rtol=1e-12, atol=1e-12), SHAPES[i]
        ref_lml, ref_g, _ = gp.lml_and_grad(expr, ks[i], nz[i], X, y)
        a, b = batch.theta_off[i], batch.theta_off[i + 1]
        assert status[i] == 0 and tries[i] == 0
        assert abs(lml[i] - ref_lml) <= 1e-8 * max(1.0, abs(ref_lml)), (SHAPES[i], lml[i], ref_lml)
        scale = max(1.0, np.max(np.abs(ref_g)))
        assert np.max(np.abs(dlml[a:b] - ref_g)) <= 1e-6 * scale, (SHAPES[i], dlml[a:b], ref_g)


def test_kernel_matrix_matches_oracle_and_gpml(engine):
    from ms_project_b200.program import ProgramBatch
    from oracle import kernels as ok
    gold = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'gpml_kernels.json')))
    X = np.array(gold['X'], dtype=np.float64).reshape(-1, 1)
    for case in gold['cases']:
        k = from_string(case['kernel'])
        batch = ProgramBatch([k])
        engine.bind(X, np.zeros(5), slots=1, max_params=batch.max_params)
        engine.upload_programs(batch)
        theta = batch.pack_theta([np.array(case['theta'])], [0.0])
        K = engine.build_K(batch, theta)[0]
        assert np.allclose(K, np.array(case['K']), rtol=1e-8, atol=1e-8), case['kernel']
        assert np.allclose(K, ok.K(ok.parse(case['kernel']), case['theta'], X), rtol=1e-13, atol=1e-13)


def test_extreme_hyperparameters(engine):
    """Lengthscales / periods far outside the data scale (line-search excursions of L-BFGS-B reach them): the
    straight-line exp / log1p of csrc/fastmath.cuh must flush, not wrap (exponent arguments down to -1e20)."""
    from ms_project_b200.program import ProgramBatch
    from oracle import kernels as ok
    X, _ = make_data(140, 1, seed=2)
    cases = [('SE_1', [1.3, 1e-10]), ('SE_1', [2.0, 3e-4]), ('SE_1', [0.5, 1e6]), ('RQ_1', [1.1, 1e-9, 0.7]),
             ('RQ_1', [1.1, 2e-3, 40.0]), ('RQ_1', [0.9, 1e5, 1e-3]), ('PER_1', [1.2, 0.37, 1e-9]), ('PER_1', [1.2, 1e-3, 0.4]),
             ('PER_1', [0.8, 50.0, 1e-4]), ('SE_1 * PER_1 + RQ_1', [1.0, 1e-7, 1.0, 2.0, 1e-8, 1.0, 1e-6, 3.0])]
    for name, th in cases:
        k = from_string(name)
        batch = ProgramBatch([k])
        engine.bind(X, np.zeros(X.shape[0]), slots=1, max_params=batch.max_params)
        engine.upload_programs(batch)
        K = engine.build_K(batch, batch.pack_theta([np.array(th)], [0.0]))[0]
        # dist_mode='direct': at lengthscales ~1e-4 GPy's x^2 + x'^2 - 2xx' distance itself carries ~1e-16 x^2 / l^2
        # of cancellation error (SURVEY.md 8c hazard 6); the comparison here is about the transcendental functions
        ref = ok.K(ok.parse(name), th, X, dist_mode='direct')
        assert np.all(np.isfinite(K)), name
        # the angle-difference identity of PER costs ~1e-16 * |pi x / p| absolute in sin; everything else is ulp-level
        assert np.allclose(K, ref, rtol=1e-9 if 'PER' in name else 1e-12, atol=1e-290), (name, th, np.max(np.abs(K - ref)))


def test_subset_ids_and_quirks(engine):
    from ms_project_b200.program import ProgramBatch
    from oracle import gp
    X, y = make_data(200, 2, seed=3)
    kernels = [from_string(s) for s in ['PER_1', 'LIN_2 + SE_1', 'RQ_2']]
    batch = ProgramBatch(kernels)
    engine.bind(X, y, slots=8, max_params=batch.max_params)
    engine.upload_programs(batch)
    theta, ks, nz = random_theta(batch, np.random.RandomState(2))
    engine.set_quirks(3)
    try:
        lml, dlml, status, _ = engine.lml_grad(batch, theta, ids=[1, 0])
        for i in (0, 1):
            ref_lml, ref_g, _ = gp.lml_and_grad(to_oracle_expr(kernels[i]), ks[i], nz[i], X, y, quirks=True)
            a, b = batch.theta_off[i], batch.theta_off[i + 1]
            assert abs(lml[i] - ref_lml) <= 1e-8 * max(1.0, abs(ref_lml))
            assert np.max(np.abs(dlml[a:b] - ref_g)) <= 1e-6 * max(1.0, np.max(np.abs(ref_g)))
    finally:
        engine.set_quirks(0)


def test_jitter_ladder_and_not_pd(engine):
    """GPy's jitchol ladder.  Item 0: K_y = v 11^T exactly (SE with an enormous lengthscale; v = 1e12 swallows the
    noise), so the second pivot is exactly 0 on any IEEE implementation and the first rung (mean(diag) * 1e-6)
    repairs it.  Item 1 is healthy.  Item 2 has a NaN hyperparameter: every rung fails -> NOT_PD / NaN."""
    from ms_project_b200.program import ProgramBatch
    from oracle import gp
    rng = np.random.RandomState(0)
    X = rng.randn(160, 1)
    y = rng.randn(160, 1)
    kernels = [from_string('SE_1'), from_string('SE_1 + RQ_1'), from_string('RQ_1')]
    batch = ProgramBatch(kernels)
    engine.bind(X, y, slots=3, max_params=batch.max_params)
    engine.upload_programs(batch)
    ks = [np.array([1e12, 1e9]), np.array([1.0, 1.0, 1.0, 1.0, 2.0]), np.array([1.0, np.nan, 2.0])]
    nz = [1e-12, 0.1, 0.1]
    theta = batch.pack_theta(ks, nz)
    lml, dlml, status, tries = engine.lml_grad(batch, theta)
    ref_lml, ref_g, inf = gp.lml_and_grad(to_oracle_expr(kernels[0]), ks[0], nz[0], X, y)
    assert inf['tries'] == 1
    assert status[0] == 1 and tries[0] == 1
    assert abs(lml[0] - ref_lml) <= 1e-8 * abs(ref_lml)
    a, b = batch.theta_off[0], batch.theta_off[1]
    assert np.max(np.abs(dlml[a:b] - ref_g)) <= 1e-6 * max(1.0, np.max(np.abs(ref_g)))
    assert status[1] == 0 and tries[1] == 0
    assert status[2] == 2 and np.isnan(lml[2])


def test_fused_kinv_gradient_kernel_matches_oracle(engine):
    """b200gp_set_fused_grad(1): K^-1 tile staged in shared memory and contracted in the same kernel (<= 4 leaves),
    other kinds keep the two-kernel path inside the same batch."""
    from ms_project_b200.program import ProgramBatch
    from oracle import gp
    X, y = make_data(300, 3, seed=11)
    kernels = [from_string(s) for s in KERNELS]
    batch = ProgramBatch(kernels)
    engine.bind(X, y, slots=len(kernels), max_params=batch.max_params)
    engine.upload_programs(batch)
    theta, ks, nz = random_theta(batch, np.random.RandomState(5))
    engine.set_fused_grad(True)
    try:
        lml, dlml, status, tries = engine.lml_grad(batch, theta)
        lml, dlml, status = lml.copy(), dlml.copy(), status.copy()
    finally:
        engine.set_fused_grad(False)
    lml2, dlml2, _, _ = engine.lml_grad(batch, theta)
    for i, k in enumerate(kernels):
        ref_lml, ref_g, _ = gp.lml_and_grad(to_oracle_expr(k), ks[i], nz[i], X, y)
        a, b = batch.theta_off[i], batch.theta_off[i + 1]
        assert status[i] == 0
        assert lml[i] == lml2[i]                                   # the factorisation does not depend on the mode
        scale = max(1.0, np.max(np.abs(ref_g)))
        assert np.max(np.abs(dlml[a:b] - ref_g)) <= 1e-6 * scale, (KERNELS[i], dlml[a:b], ref_g)
        assert np.max(np.abs(dlml[a:b] - dlml2[a:b])) <= 1e-10 * scale


@pytest.mark.parametrize('n', [1, 2, 127, 129, 255, 257])
def test_tile_edge_sizes(engine, n):
    """Orders around the 128-point tile edge (padding rows carry the identity) and the degenerate N = 1, 2."""
    from ms_project_b200.program import ProgramBatch
    from oracle import gp
    rng = np.random.RandomState(n)
    X = rng.randn(n, 2)
    y = rng.randn(n, 1)
    kernels = [from_string(s) for s in ['SE_1 + RQ_2', 'PER_1 * LIN_2 + CONST_1', 'SE_1 * SE_2']]
    batch = ProgramBatch(kernels)
    engine.bind(X, y, slots=2, max_params=batch.max_params)
    engine.upload_programs(batch)
    theta, ks, nz = random_theta(batch, np.random.RandomState(7))
    lml, dlml, status, tries = engine.lml_grad(batch, theta)
    for i, k in enumerate(kernels):
        ref_lml, ref_g, _ = gp.lml_and_grad(to_oracle_expr(k), ks[i], nz[i], X, y)
        a, b = batch.theta_off[i], batch.theta_off[i + 1]
        assert status[i] == 0 and tries[i] == 0
        assert abs(lml[i] - ref_lml) <= 1e-8 * max(1.0, abs(ref_lml)), (n, i, lml[i], ref_lml)
        assert np.max(np.abs(dlml[a:b] - ref_g)) <= 1e-6 * max(1.0, np.max(np.abs(ref_g))), (n, i)


def test_empty_batch_and_duplicate_inputs(engine):
    """No items: nothing is launched, nothing is touched.  Repeated rows of X (collisions) make K singular; with the noise
    term it stays positive definite and matches the oracle, without it the jitter ladder takes over like GPy's jitchol."""
    from ms_project_b200.program import ProgramBatch
    from oracle import gp
    rng = np.random.RandomState(1)
    X = np.repeat(rng.randn(50, 1), 3, axis=0)              # every point three times
    y = rng.randn(150, 1)
    kernels = [from_string('SE_1'), from_string('RQ_1 + LIN_1')]
    batch = ProgramBatch(kernels)
    engine.bind(X, y, slots=2, max_params=batch.max_params)
    engine.upload_programs(batch)
    theta, ks, nz = random_theta(batch, np.random.RandomState(3))
    l0 = engine.launch_count()
    out = engine.lml_grad(batch, theta, ids=np.zeros(0, dtype=np.int32))
    assert engine.launch_count() == l0 and len(out) == 4
    lml, dlml, status, tries = engine.lml_grad(batch, theta)
    for i, k in enumerate(kernels):
        ref_lml, ref_g, _ = gp.lml_and_grad(to_oracle_expr(k), ks[i], nz[i], X, y)
        a, b = batch.theta_off[i], batch.theta_off[i + 1]
        assert status[i] == 0
        assert abs(lml[i] - ref_lml) <= 1e-8 * max(1.0, abs(ref_lml))
        assert np.max(np.abs(dlml[a:b] - ref_g)) <= 1e-6 * max(1.0, np.max(np.abs(ref_g)))
    # (almost) no noise: exactly repeated rows leave only GPy's 1e-8 on the diagonal (condition number ~1e8); both sides must
    # take the same number of jitchol rungs (none, if the 1e-8 is enough) and agree at the accuracy that conditioning allows
    nz0 = np.array([1e-300, 1e-300])
    theta0 = batch.pack_theta(ks, nz0)
    lml, dlml, status, tries = engine.lml_grad(batch, theta0)
    for i, k in enumerate(kernels):
        ref_lml, ref_g, inf = gp.lml_and_grad(to_oracle_expr(k), ks[i], nz0[i], X, y)
        assert tries[i] == inf['tries'] and status[i] == (1 if inf['tries'] else 0), (i, tries[i], inf['tries'])
        assert abs(lml[i] - ref_lml) <= 1e-6 * max(1.0, abs(ref_lml)), (i, lml[i], ref_lml)
