"""GPU parity at BASELINE.json's full sizes.

* C2 size (N = 2048): LML / gradient of a few items against the NumPy oracle (seconds on the CPU), and bit-identical
  results for the same item evaluated alone (look-ahead + pipelined L^-T schedule, one tile column per panel), in a
  batch of 20 (two tile columns per panel) and in a batch of 80 (two slot groups on two streams).
* C5 size (N = 16384): size-independent properties -- the blocked Cholesky reproduces K (K x = L L^T x for random x,
  log-determinant against the sum over a direct slogdet of the leading block is out of reach, so the factor itself is
  checked), and the gradient agrees with central finite differences of the device LML along a random direction.
Tolerances: LML rel 1e-8, gradients rel 1e-6 (north star); finite differences 1e-5 (truncation)."""
import numpy as np
import pytest

from conftest import to_oracle_expr

pytestmark = pytest.mark.gpu


def test_c2_size_parity_and_schedule_independence(engine):
    from ms_project_b200 import workloads
    from ms_project_b200.program import ProgramBatch
    from oracle import gp
    X, y, kernels = workloads.config_c2(n=2048, n_kernels=80)
    rng = np.random.RandomState(3)
    batch = ProgramBatch(kernels)
    engine.bind(X, y, slots=80, max_params=batch.max_params)
    engine.upload_programs(batch)
    ks = [np.exp(rng.uniform(-0.5, 0.5, size=int(p))) for p in batch.n_params]
    nz = np.exp(rng.uniform(-3, -1, size=batch.n_items))
    theta = batch.pack_theta(ks, nz)
    lml80, g80, st80, _ = [a.copy() for a in engine.lml_grad(batch, theta)]                 # split schedule (>= 65 slots)
    assert (st80 == 0).all()
    for i in (0, 7, 33):
        ref, gref, _ = gp.lml_and_grad(to_oracle_expr(kernels[i]), ks[i], nz[i], X, y)
        a, b = batch.theta_off[i], batch.theta_off[i + 1]
        assert abs(lml80[i] - ref) <= 1e-8 * max(1.0, abs(ref)), (i, lml80[i], ref)
        assert np.max(np.abs(g80[a:b] - gref)) <= 1e-6 * max(1.0, np.max(np.abs(gref))), (i, g80[a:b], gref)
    lml1, g1, _, _ = [a.copy() for a in engine.lml_grad(batch, theta, ids=np.array([33], dtype=np.int32))]      # W = 1, pipelined
    lml20, g20, _, _ = [a.copy() for a in engine.lml_grad(batch, theta, ids=np.arange(20, 40, dtype=np.int32))]  # W = 2
    a, b = batch.theta_off[33], batch.theta_off[34]
    assert lml1[33] == lml80[33] and lml20[33] == lml80[33]
    assert np.array_equal(g1[a:b], g80[a:b]) and np.array_equal(g20[a:b], g80[a:b])
    # the 128 x 64 CTA tiles (per-warp skipping of the structural zeros) give the same bits as the default 64 x 64 tiles
    engine.set_tile64(False)
    try:
        lml_w, g_w, st_w, _ = [a_.copy() for a_ in engine.lml_grad(batch, theta)]
        lml_w1, g_w1, _, _ = [a_.copy() for a_ in engine.lml_grad(batch, theta, ids=np.array([33], dtype=np.int32))]
    finally:
        engine.set_tile64(True)
    assert (st_w == 0).all() and np.array_equal(lml_w, lml80) and np.array_equal(g_w, g80)
    assert lml_w1[33] == lml80[33] and np.array_equal(g_w1[a:b], g80[a:b])
    # operand staging of the 64 x 64 tile GEMMs: TMA + mbarrier ring (default, three swizzle modes, 3- and 4-stage rings) and
    # cp.async + block barrier give the same bits (tile_gemm64_tma.cuh: every element sums its products in ascending k)
    try:
        for mode in (0, 1, 2, 4):
            engine.set_tma(mode)
            lml_t, g_t, st_t, _ = [a_.copy() for a_ in engine.lml_grad(batch, theta)]
            assert (st_t == 0).all() and np.array_equal(lml_t, lml80) and np.array_equal(g_t, g80), mode
            lml_t1, g_t1, _, _ = [a_.copy() for a_ in engine.lml_grad(batch, theta, ids=np.array([33], dtype=np.int32))]
            assert lml_t1[33] == lml80[33] and np.array_equal(g_t1[a:b], g80[a:b]), mode
    finally:
        engine.set_tma(3)


def test_c5_size_cholesky_and_gradient_properties(engine):
    import torch
    from ms_project_b200 import workloads
    from ms_project_b200.program import ProgramBatch
    n = 16384
    X, y, k, noise = workloads.config_c5(n=n)
    batch = ProgramBatch([k])
    engine.bind(X, y, slots=1, max_params=batch.max_params)
    engine.upload_programs(batch)
    th0 = batch.pack_theta([np.asarray(k.param_array, dtype=np.float64)], [noise])
    lml, g, st, _ = [a.copy() for a in engine.lml_grad(batch, th0)]
    assert st[0] == 0 and np.isfinite(lml[0])
    # directional derivative by central differences of the device LML
    rng = np.random.RandomState(0)
    d = rng.randn(th0.size)
    d /= np.linalg.norm(d)
    h = 1e-4
    lp = engine.lml_grad(batch, th0 * np.exp(h * d), with_grad=False)[0][0]
    lm = engine.lml_grad(batch, th0 * np.exp(-h * d), with_grad=False)[0][0]
    fd = (lp - lm) / (2 * h)
    an = float(np.dot(g[:th0.size] * th0, d))                 # d/dh of LML(theta exp(h d))
    assert abs(fd - an) <= 1e-5 * max(1.0, abs(an)), (fd, an)
    # the factor reproduces K_y: K_y x = L (L^T x) for random x
    K = torch.empty((n, n), dtype=torch.float64, device=engine.device)
    from ms_project_b200 import _lib
    batch.h_theta.numpy()[:] = th0
    batch.d_theta.copy_(batch.h_theta)
    rc = engine.lib.b200gp_build_K(engine.ctx, batch.d_tokens.data_ptr(), batch.d_prog_off.data_ptr(), batch.d_theta.data_ptr(),
                                   batch.d_theta_off.data_ptr(), batch.d_all_ids.data_ptr(), 1, 1, K.data_ptr())
    _lib.check(engine.ctx, rc, 'b200gp_build_K')
    x = torch.from_numpy(rng.randn(n, 4)).to(engine.device)
    Kx = K @ x
    logdet, status = engine.potrf(K)
    assert status == 0
    L = torch.tril(K)
    del K
    r = L @ (L.T @ x) - Kx
    assert float(r.abs().max()) <= 1e-10 * float(Kx.abs().max())
    # LML assembled from the stand-alone factorisation: -1/2 (N log 2 pi + logdet + y^T K_y^-1 y)
    yd = torch.from_numpy(np.asarray(y, dtype=np.float64).reshape(-1, 1)).to(engine.device)
    z = torch.linalg.solve_triangular(L, yd, upper=False)
    lml_ref = -0.5 * (n * np.log(2 * np.pi) + logdet + float((z * z).sum()))
    assert abs(lml[0] - lml_ref) <= 1e-8 * max(1.0, abs(lml_ref)), (lml[0], lml_ref)


def test_c4_size_parity_every_kernel_family_of_programs(engine):
    """C4 size (N = 4096, D = 8): one program for each device code path -- <= 4 leaves (poly_*<4>), 5-8 leaves
    (poly_*<8>) and > 8 leaves (the interpreter) -- against the oracle at the north-star tolerances."""
    from conftest import from_string
    from ms_project_b200 import workloads
    from ms_project_b200.program import ProgramBatch
    from oracle import gp
    X, y, _ = workloads.config_c4(n=4096, pop=1)
    names = ['SE_1 * RQ_2 + SE_3 * SE_8',                                                          # 4 leaves
             'SE_1 * RQ_2 + SE_3 * RQ_4 * SE_5 + RQ_6 + SE_7 * SE_8',                              # 8 leaves
             'SE_1 + RQ_2 * SE_3 + SE_4 * RQ_5 + RQ_6 * SE_7 + SE_8 * SE_1 + RQ_3 + SE_5']         # 11 leaves: interpreter
    kernels = [from_string(s) for s in names]
    rng = np.random.RandomState(4)
    batch = ProgramBatch(kernels)
    engine.bind(X, y, slots=3, max_params=batch.max_params)
    engine.upload_programs(batch)
    ks = [np.exp(rng.uniform(-0.3, 0.7, size=int(p))) for p in batch.n_params]
    nz = np.array([0.05, 0.1, 0.2])
    lml, g, st, _ = [a.copy() for a in engine.lml_grad(batch, batch.pack_theta(ks, nz))]
    assert (st == 0).all()
    for i, k in enumerate(kernels):
        ref, gref, _ = gp.lml_and_grad(to_oracle_expr(k), ks[i], nz[i], X, y)
        a, b = batch.theta_off[i], batch.theta_off[i + 1]
        assert abs(lml[i] - ref) <= 1e-8 * max(1.0, abs(ref)), (names[i], lml[i], ref)
        assert np.max(np.abs(g[a:b] - gref)) <= 1e-6 * max(1.0, np.max(np.abs(gref))), (names[i], g[a:b], gref)


def test_c5_kernel_against_the_oracle_at_n8192(engine):
    """C5's kernel (SE_1 * PER_1 + RQ_2, D = 2, the look-ahead single-matrix schedule with wide panels: T = 64 > 32) against
    the ORACLE's LML and gradient at N = 8192 (N = 16384 costs the CPU several minutes and 20 GB; the schedule and every
    kernel are the same from T = 33 tile columns on)."""
    from ms_project_b200 import workloads
    from ms_project_b200.program import ProgramBatch
    from oracle import gp
    X, y, k, noise = workloads.config_c5(n=8192)
    batch = ProgramBatch([k])
    engine.bind(X, y, slots=1, max_params=batch.max_params)
    engine.upload_programs(batch)
    th = np.asarray(k.param_array, dtype=np.float64)
    lml, g, st, _ = [a.copy() for a in engine.lml_grad(batch, batch.pack_theta([th], [noise]))]
    assert st[0] == 0
    ref, gref, _ = gp.lml_and_grad(to_oracle_expr(k), th, noise, X, y)
    assert abs(lml[0] - ref) <= 1e-8 * max(1.0, abs(ref)), (lml[0], ref)
    assert np.max(np.abs(g[:th.size + 1] - gref)) <= 1e-6 * max(1.0, np.max(np.abs(gref))), (g[:th.size + 1], gref)
