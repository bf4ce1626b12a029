"""GPU: error behaviour of the C ABI (include/b200gp.h).  API misuse returns a non-zero code with a message from
b200gp_last_error and never touches memory; numerical failure is a per-item status, never a non-zero return."""
import ctypes

import numpy as np
import pytest

from conftest import from_string, make_data

pytestmark = pytest.mark.gpu


def _fresh_ctx():
    import torch
    from ms_project_b200 import _lib
    lib = _lib.lib()
    ctx = ctypes.c_void_p()
    assert lib.b200gp_create(0, None, ctypes.byref(ctx)) == 0
    return lib, ctx, torch


def test_misuse_is_reported_not_executed():
    lib, ctx, torch = _fresh_ctx()
    try:
        dev = torch.device('cuda', 0)
        buf = torch.zeros(64, dtype=torch.float64, device=dev)
        ibuf = torch.zeros(64, dtype=torch.int32, device=dev)
        p, ip = buf.data_ptr(), ibuf.data_ptr()
        # nothing bound yet
        assert lib.b200gp_lml_grad(ctx, ip, ip, p, ip, ip, 1, 1, p, p, ip, ip) != 0
        assert b'b200gp_bind first' in lib.b200gp_last_error(ctx)
        assert lib.b200gp_build_K(ctx, ip, ip, p, ip, ip, 1, 0, p) != 0
        assert lib.b200gp_posterior(ctx, ip, ip, p, ip, ip, 0, p, 4, 1, p, p, p, ip) != 0
        assert lib.b200gp_centered_alignments(ctx, ip, ip, p, ip, ip, 2, p) != 0
        assert lib.b200gp_gram_logdet(ctx, ip, ip, p, ip, ip, 1, 2, p, p, ip) != 0
        assert lib.b200gp_hellinger_pairs_large(ctx, p, p, 2, 4, 40, ip, ip, 1, ip, ip, 1, p, p, ip) != 0
        assert b'b200gp_bind' in lib.b200gp_last_error(ctx)
        # bind: geometry and workspace checks
        X, y = make_data(40, 2, seed=0)
        dX, dy = torch.from_numpy(X).to(dev), torch.from_numpy(y.reshape(-1)).to(dev)
        need = lib.b200gp_workspace_bytes(40, 2, 2, 8)
        assert need > 0 and lib.b200gp_workspace_bytes(0, 2, 2, 8) == 0 and lib.b200gp_workspace_bytes(40, 2, 0, 8) == 0
        ws = torch.empty(need, dtype=torch.uint8, device=dev)
        assert lib.b200gp_bind(ctx, ws.data_ptr(), need - 1024, 2, 8, dX.data_ptr(), 40, 2, dy.data_ptr()) != 0
        assert b'workspace too small' in lib.b200gp_last_error(ctx)
        assert lib.b200gp_bind(ctx, ws.data_ptr(), need, 2, 8, dX.data_ptr(), 40, 17, dy.data_ptr()) != 0      # D > 16
        assert lib.b200gp_bind(ctx, ws.data_ptr(), need, 2, 8, None, 40, 2, dy.data_ptr()) != 0
        assert lib.b200gp_bind(ctx, ws.data_ptr(), need, 2, 200, dX.data_ptr(), 40, 2, dy.data_ptr()) != 0     # > 128 params
        assert lib.b200gp_bind(ctx, ws.data_ptr(), need, 2, 8, dX.data_ptr(), 40, 2, dy.data_ptr()) == 0
        # null pointers after a valid bind
        assert lib.b200gp_lml_grad(ctx, None, ip, p, ip, ip, 1, 1, p, p, ip, ip) != 0
        assert b'null pointer' in lib.b200gp_last_error(ctx)
        assert lib.b200gp_lml_grad(ctx, ip, ip, p, ip, ip, 1, 1, p, None, ip, ip) != 0     # gradient wanted, no buffer
        assert lib.b200gp_posterior(ctx, ip, ip, p, ip, ip, 0, p, 0, 1, p, p, p, ip) != 0   # Ns = 0
        assert lib.b200gp_centered_alignments(ctx, ip, ip, p, ip, ip, 3, p) != 0            # 3 kernels, 2 slots
        assert b'slots' in lib.b200gp_last_error(ctx)
        assert lib.b200gp_gram_logdet(ctx, ip, ip, p, ip, ip, 1, 3, p, p, ip) != 0            # unknown diag_mode
        assert lib.b200gp_gram_logdet(ctx, ip, ip, p, ip, ip, 1, 2, None, p, None) != 0       # log-determinants without a status buffer
        assert lib.b200gp_gram_logdet(ctx, ip, ip, p, ip, ip, 1, 2, None, None, None) == 0    # nothing asked for: no-op
        assert lib.b200gp_hellinger_pairs_large(ctx, p, p, 2, 4, 64, ip, ip, 1, ip, ip, 1, p, p, ip) != 0   # bound to 40 points, not 64
        assert b'bound to 40 points' in lib.b200gp_last_error(ctx)
        assert lib.b200gp_hellinger_pairs_large_work_bytes(0, 4) == 0 and lib.b200gp_hellinger_pairs_large_work_bytes(3, 4) > 0
        # standalone entry points
        assert lib.b200gp_vector_pairs(ctx, p, 2, 4, 3, ip, ip, 1, 7, p) != 0               # unknown metric
        assert lib.b200gp_hellinger_precompute(ctx, p, 129, 2, ip, ip, p, ip, p, 1, 4, p, None, ip) != 0   # subset > 128
        assert lib.b200gp_set_lookup(ctx, p, 0) != 0 and lib.b200gp_set_lookup(ctx, None, 5) != 0
        assert lib.b200gp_set_lookup(ctx, None, 0) == 0
        assert lib.b200gp_potrf_work_bytes(100) == 0                                          # N must be a multiple of 128
        ld, st = ctypes.c_double(), ctypes.c_int32()
        assert lib.b200gp_potrf(ctx, p, 100, p, ctypes.byref(ld), ctypes.byref(st)) != 0
        assert lib.b200gp_set_panel_tiles(ctx, 0) != 0 and lib.b200gp_set_panel_tiles(ctx, 4) == 0
        # n = 0 items is a no-op, not an error
        assert lib.b200gp_lml_grad(ctx, ip, ip, p, ip, ip, 0, 1, p, p, ip, ip) == 0
        assert lib.b200gp_vector_pairs(ctx, p, 2, 4, 3, ip, ip, 0, 0, p) == 0
    finally:
        lib.b200gp_destroy(ctx)
    assert lib.b200gp_last_error(None) == b'null context'


def test_numerical_failure_is_a_status_never_a_return_code(engine):
    """NaN / non-PD inputs come back as status codes and NaN values with rc == 0 (GPModel.fit maps them to failed_fitting,
    src/autoks/core/gp_model.py:67-77)."""
    from ms_project_b200.program import ProgramBatch
    X, y = make_data(50, 1, seed=1)
    ks = [from_string('SE_1'), from_string('PER_1'), from_string('LIN_1 + RQ_1')]
    batch = ProgramBatch(ks)
    engine.bind(X, y, slots=3, max_params=batch.max_params)
    engine.upload_programs(batch)
    theta = batch.pack_theta([np.array([1.0, np.nan]), np.array([1.0, 0.0, 1.0]), np.array([np.inf, 1.0, 1.0, 1.0, 1.0])],
                             [0.1, 0.1, 0.1])
    lml, dlml, status, tries = engine.lml_grad(batch, theta)      # raises only on rc != 0
    assert all(s in (2, 3) for s in status)
    assert all(not np.isfinite(v) for v in lml)
    mean, var, _, st = engine.posterior(batch, theta, 0, X[:5])
    assert st == 2 and np.all(np.isnan(mean)) and np.all(np.isnan(var))
